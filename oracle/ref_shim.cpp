// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points over reference translation units compiled UNMODIFIED from
// /root/reference/src (never copied into this repository): OptimizationBackend/MatrixAccumulators.h,
// OptimizationBackend/ScaleAccumulator.h, util/globalFuncs.h and util/settings.cpp, against the minimal Eigen / Sophus / boost
// stand-ins of oracle/ref_stub/ (see Eigen/Core there for what the stand-in does and does not pin).  Built by
// oracle/ref_build.sh into oracle/_ref/libref_units.so; tests/test_ref_pin.py compares the oracle's restatements
// (orc_accum.h, orc_math.h samplers, orc_config_default, the residual pattern) with these, bit for bit.
#include "OptimizationBackend/MatrixAccumulators.h"
#include "OptimizationBackend/ScaleAccumulator.h"
#include "util/globalFuncs.h"
#include "util/settings.h"

#define REF_API extern "C" __attribute__((visibility("default")))
using namespace dso;

// AccumulatorApprox (MatrixAccumulators.h:744-1170): n residuals; per residual x4,x6,y4,y6 (20 floats), a,b,c, the six
// TopRight and the six BotRight arguments (35 floats per residual).  out: H 13x13 row-major, num.
REF_API void ref_approx_run(int n, const float *in, float *H169, double *num) {
  AccumulatorApprox acc;
  acc.initialize();
  for (int i = 0; i < n; i++) {
    const float *p = in + 35 * (size_t)i;
    acc.update(p, p + 4, p + 10, p + 14, p[20], p[21], p[22]);
    acc.updateTopRight(p, p + 4, p + 10, p + 14, p[23], p[24], p[25], p[26], p[27], p[28]);
    acc.updateBotRight(p[29], p[30], p[31], p[32], p[33], p[34]);
  }
  acc.finish();
  for (int r = 0; r < 13; r++)
    for (int c = 0; c < 13; c++) H169[r * 13 + c] = acc.H(r, c);
  *num = (double)acc.num;
}

// Accumulator9 (MatrixAccumulators.h:1172-1687).  mode 0: updateSSE (groups of 4, J as 9 arrays of n), 1: updateSSE_eighted
// (+ w[n]), 2: updateSingle, 3: updateSingleWeighted.  n must be a multiple of 4 for the SSE modes.  J is [9][n].
REF_API void ref_acc9_run(int mode, int n, const float *J, const float *w, float *H81, double *num) {
  Accumulator9 acc;
  acc.initialize();
  auto col = [&](int k, int i) { return J[(size_t)k * n + i]; };
  if (mode < 2) {
    for (int i = 0; i + 3 < n; i += 4) {
      __m128 v[9];
      for (int k = 0; k < 9; k++) v[k] = _mm_loadu_ps(J + (size_t)k * n + i);
      if (mode == 0) acc.updateSSE(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]);
      else acc.updateSSE_eighted(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], _mm_loadu_ps(w + i));
    }
  } else {
    for (int i = 0; i < n; i++) {
      if (mode == 2) acc.updateSingle(col(0, i), col(1, i), col(2, i), col(3, i), col(4, i), col(5, i), col(6, i), col(7, i), col(8, i));
      else acc.updateSingleWeighted(col(0, i), col(1, i), col(2, i), col(3, i), col(4, i), col(5, i), col(6, i), col(7, i), col(8, i), w[i]);
    }
  }
  acc.finish();
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) H81[r * 9 + c] = acc.H(r, c);
  *num = (double)acc.num;
}

// Accumulator11 (MatrixAccumulators.h:80-150): updateSingle of n values
REF_API float ref_acc11_run(int n, const float *v) {
  Accumulator11 acc;
  acc.initialize();
  for (int i = 0; i < n; i++) acc.updateSingle(v[i]);
  acc.finish();
  return acc.A;
}

// AccumulatorXX<8,8>, AccumulatorXX<8,4>, AccumulatorX<8> (MatrixAccumulators.h:33-78, 152-202) as the Schur accumulators use
// them (AccumulatedSCHessian.h:126-130).  The tier logic (shiftUp) is the header's; the expression `A += w * L * R^T` is
// evaluated by the stand-in Eigen (coefficient-wise (w L_i) R_j).  L: [n][8], R: [n][8|4], w: [n].  out row-major.
REF_API void ref_accxx88_run(int n, const float *L, const float *R, const float *w, float *A64) {
  AccumulatorXX<8, 8> acc;
  acc.initialize();
  for (int i = 0; i < n; i++) {
    Eigen::Matrix<float, 8, 1> l, r;
    for (int k = 0; k < 8; k++) { l[k] = L[8 * (size_t)i + k]; r[k] = R[8 * (size_t)i + k]; }
    acc.update(l, r, w[i]);
  }
  acc.finish();
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < 8; c++) A64[r * 8 + c] = acc.A1m(r, c);
}
REF_API void ref_accxx84_run(int n, const float *L, const float *R, const float *w, float *A32) {
  AccumulatorXX<8, 4> acc;
  acc.initialize();
  for (int i = 0; i < n; i++) {
    Eigen::Matrix<float, 8, 1> l;
    Eigen::Matrix<float, 4, 1> r;
    for (int k = 0; k < 8; k++) l[k] = L[8 * (size_t)i + k];
    for (int k = 0; k < 4; k++) r[k] = R[4 * (size_t)i + k];
    acc.update(l, r, w[i]);
  }
  acc.finish();
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < 4; c++) A32[r * 4 + c] = acc.A1m(r, c);
}
REF_API void ref_accx8_run(int n, const float *L, const float *w, float *A8) {
  AccumulatorX<8> acc;
  acc.initialize();
  for (int i = 0; i < n; i++) {
    Eigen::Matrix<float, 8, 1> l;
    for (int k = 0; k < 8; k++) l[k] = L[8 * (size_t)i + k];
    acc.update(l, w[i]);
  }
  acc.finish();
  for (int k = 0; k < 8; k++) A8[k] = acc.A1m[k];
}

// ScaleAccumulator (ScaleAccumulator.h:27-106): groups of 4
REF_API void ref_scaleacc_run(int n, const float *J0, const float *J1, const float *w, float *H4, double *num) {
  ScaleAccumulator acc;
  acc.initialize();
  for (int i = 0; i + 3 < n; i += 4) acc.updateSSE_oneed(_mm_loadu_ps(J0 + i), _mm_loadu_ps(J1 + i), _mm_loadu_ps(w + i));
  acc.finish();
  for (int r = 0; r < 2; r++)
    for (int c = 0; c < 2; c++) H4[r * 2 + c] = acc.hessian(r, c);
  *num = (double)acc.num;
}

// util/globalFuncs.h:68-182 on an Eigen::Vector3f image (AoS {I, dx, dy}).  which 0: getInterpolatedElement33 -> out[3],
// 1: getInterpolatedElement31 -> out[0], 2: getInterpolatedElement33BiLin -> out[3], 3: getInterpolatedElement (float image
// = channel stride 1 array `img1`) -> out[0]
REF_API void ref_interp(int which, const float *img3, const float *img1, int width, int n, const float *x, const float *y, float *out) {
  const Eigen::Vector3f *m = (const Eigen::Vector3f *)img3;
  for (int i = 0; i < n; i++) {
    float *o = out + 3 * (size_t)i;
    if (which == 0) { const Eigen::Vector3f v = getInterpolatedElement33(m, x[i], y[i], width); o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; }
    else if (which == 1) { o[0] = getInterpolatedElement31(m, x[i], y[i], width); o[1] = o[2] = 0; }
    else if (which == 2) { const Eigen::Vector3f v = getInterpolatedElement33BiLin(m, x[i], y[i], width); o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; }
    else { o[0] = getInterpolatedElement(img1, x[i], y[i], width); o[1] = o[2] = 0; }
  }
}

// util/settings.cpp: the defaults the path reads, in the member order of sosba_config's float block (include/sosba.h), and the
// residual pattern (settings.cpp:307-317 via the patternP macro of settings.h:187-189)
REF_API void ref_settings(float *f18, int *i4) {
  f18[0] = setting_huberTH; f18[1] = setting_outlierTHSumComponent; f18[2] = setting_affineOptModeA; f18[3] = setting_affineOptModeB;
  f18[4] = setting_coarseCutoffTH; f18[5] = setting_idepthFixPrior; f18[6] = setting_idepthFixPriorMargFac;
  f18[7] = setting_frameEnergyTHConstWeight; f18[8] = setting_frameEnergyTHN; f18[9] = setting_frameEnergyTHFacMedian;
  f18[10] = setting_overallEnergyTHWeight; f18[11] = setting_initialCalibHessian; f18[12] = setting_initialRotPrior;
  f18[13] = setting_initialTransPrior; f18[14] = setting_initialAffAPrior; f18[15] = setting_initialAffBPrior;
  f18[16] = setting_margWeightFac; f18[17] = setting_thOptIterations;
  i4[0] = setting_gammaWeightsPixelSelect; i4[1] = setting_minOptIterations; i4[2] = setting_maxOptIterations; i4[3] = patternNum;
}
REF_API void ref_pattern(int *xy16) {
  for (int i = 0; i < patternNum; i++) { xy16[2 * i] = patternP[i][0]; xy16[2 * i + 1] = patternP[i][1]; }
}
