// ORACLE — TEST INFRASTRUCTURE ONLY.  The oracle's restatements of the tiered accumulators (orc_accum.h), the bilinear
// samplers (orc_sample.h), the settings defaults (orc_config_default) and the residual pattern behind the same C signatures
// as oracle/ref_shim.cpp exports for the reference units themselves, so tests/test_ref_pin.py can feed both the same
// streams and compare bit for bit.
#include <cstddef>

#include "orc_accum.h"
#include "orc_core.h"
#include "orc_sample.h"

#define ORC_API extern "C" __attribute__((visibility("default")))
using namespace orc;

ORC_API void orc_pin_approx_run(int n, const float *in, float *H169, double *num) {
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
    for (int c = 0; c < 13; c++) H169[r * 13 + c] = acc.H[r][c];
  *num = (double)acc.num;
}

ORC_API void orc_pin_acc9_run(int mode, int n, const float *J, const float *w, float *H81, double *num) {
  Accumulator9 acc;
  acc.initialize();
  if (mode < 2) {
    for (int i = 0; i + 3 < n; i += 4) {
      float v[9][4], ww[4] = {0, 0, 0, 0};
      for (int k = 0; k < 9; k++)
        for (int l = 0; l < 4; l++) v[k][l] = J[(size_t)k * n + i + l];
      if (mode == 0) acc.updateSSE(v);
      else { for (int l = 0; l < 4; l++) ww[l] = w[i + l]; acc.updateSSE_eighted(v, ww); }
    }
  } else {
    for (int i = 0; i < n; i++) {
      float v[9];
      for (int k = 0; k < 9; k++) v[k] = J[(size_t)k * n + i];
      if (mode == 3) acc.updateSingleWeighted(v, w[i]);   // (mode 2, Accumulator9::updateSingle, is dead on this path: patternNum is a multiple of 4, CoarseInitializer.cpp:595)
    }
  }
  acc.finish();
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) H81[r * 9 + c] = acc.H[r][c];
  *num = (double)acc.num;
}

ORC_API float orc_pin_acc11_run(int n, const float *v) {
  Accumulator11 acc;
  acc.initialize();
  for (int i = 0; i < n; i++) acc.updateSingle(v[i]);
  acc.finish();
  return acc.A;
}

ORC_API void orc_pin_accxx88_run(int n, const float *L, const float *R, const float *w, float *A64) {
  AccumulatorXX<8, 8> acc;
  acc.initialize();
  for (int i = 0; i < n; i++) acc.update(L + 8 * (size_t)i, R + 8 * (size_t)i, w[i]);
  acc.finish();
  for (int k = 0; k < 64; k++) A64[k] = acc.A1m[k];
}
ORC_API void orc_pin_accxx84_run(int n, const float *L, const float *R, const float *w, float *A32) {
  AccumulatorXX<8, 4> acc;
  acc.initialize();
  for (int i = 0; i < n; i++) acc.update(L + 8 * (size_t)i, R + 4 * (size_t)i, w[i]);
  acc.finish();
  for (int k = 0; k < 32; k++) A32[k] = acc.A1m[k];
}
ORC_API void orc_pin_accx8_run(int n, const float *L, const float *w, float *A8) {
  AccumulatorX<8> acc;
  acc.initialize();
  for (int i = 0; i < n; i++) acc.update(L + 8 * (size_t)i, w[i]);
  acc.finish();
  for (int k = 0; k < 8; k++) A8[k] = acc.A1m[k];
}

ORC_API void orc_pin_scaleacc_run(int n, const float *J0, const float *J1, const float *w, float *H4, double *num) {
  ScaleAccumulator acc;
  acc.initialize();
  for (int i = 0; i + 3 < n; i += 4) acc.updateSSE_oneed(J0 + i, J1 + i, w + i);
  acc.finish();
  for (int r = 0; r < 2; r++)
    for (int c = 0; c < 2; c++) H4[r * 2 + c] = acc.hessian[r][c];
  *num = (double)acc.num;
}

ORC_API void orc_pin_interp(int which, const float *img3, const float *img1, int width, int n, const float *x, const float *y, float *out) {
  for (int i = 0; i < n; i++) {
    float *o = out + 3 * (size_t)i;
    if (which == 0) interp33(img3, x[i], y[i], width, o);
    else if (which == 1) { o[0] = interp31(img3, x[i], y[i], width); o[1] = o[2] = 0; }
    else if (which == 2) interp33BiLin(img3, x[i], y[i], width, o);
    else {   // getInterpolatedElement on a float image (globalFuncs.h:36-52): the formula of interp31 with stride 1
      const int ix = (int)x[i], iy = (int)y[i];
      const float dx = x[i] - ix, dy = y[i] - iy, dxdy = dx * dy;
      const float *bp = img1 + ix + iy * width;
      o[0] = dxdy * bp[1 + width] + (dy - dxdy) * bp[width] + (dx - dxdy) * bp[1] + (1 - dx - dy + dxdy) * bp[0];
      o[1] = o[2] = 0;
    }
  }
}

ORC_API void orc_pin_pattern(int *xy16) {
  for (int i = 0; i < patternNum; i++) { xy16[2 * i] = patternP[i][0]; xy16[2 * i + 1] = patternP[i][1]; }
}
