// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).  PARITY UNPINNED (no reference fixtures exist for these loops;
// the accumulators and samplers they call are pinned by tests/test_ref_pin.py).
//
// CPU restatement of the control loops of the direct alignment (SURVEY.md §8 row a16 and the optimizeScale part of a17):
//   CoarseTracker::makeCoarseDepthL0     src/FullSystem/CoarseTracker.cpp:56-230   (as called by setCoarseTrackingRef :232-242)
//   CoarseTracker::scaleCoarseDepthL0    src/FullSystem/CoarseTracker.cpp:244-251
//   CoarseTracker::trackNewestCoarse     src/FullSystem/CoarseTracker.cpp:366-552
//   ScaleOptimizer::optimizeScale        src/FullSystem/ScaleOptimizer.cpp:120-230
// Eigen's `Hl.ldlt().solve(-b)` is orc_math.h ldlt_solve; SE3::exp / operator* follow the vendored Sophus (orc_math.h).
#include <cmath>
#include <cstring>
#include <vector>

#include "orc_core.h"
#include "orc_host.h"

namespace orc {

// ---- makeCoarseDepthL0 -------------------------------------------------------------------------------------
void tracker_makeCoarseDepth(Oracle &o, int ref_slot, int n, const float *cpt, const float *HdiF) {
  const int L = o.levels;
  std::vector<std::vector<float>> idepth(L), wsum(L), wbak(L);
  for (int l = 0; l < L; l++) { idepth[l].assign((size_t)o.wl[l] * o.hl[l], 0.f); wsum[l].assign((size_t)o.wl[l] * o.hl[l], 0.f); wbak[l] = wsum[l]; }
  const int w0 = o.wl[0];
  for (int i = 0; i < n; i++) {   // :62-79
    const int u = cpt[3 * i + 0] + 0.5f;
    const int v = cpt[3 * i + 1] + 0.5f;
    const float new_idepth = cpt[3 * i + 2];
    const float weight = sqrtf(1e-3 / (HdiF[i] + 1e-12));
    idepth[0][u + w0 * v] += new_idepth * weight;
    wsum[0][u + w0 * v] += weight;
  }
  for (int lvl = 1; lvl < L; lvl++) {   // :81-101
    const int lvlm1 = lvl - 1, wl = o.wl[lvl], hl = o.hl[lvl], wlm1 = o.wl[lvlm1];
    float *idepth_l = idepth[lvl].data(), *weight_sums_l = wsum[lvl].data();
    const float *idepth_lm = idepth[lvlm1].data(), *weight_sums_lm = wsum[lvlm1].data();
    for (int y = 0; y < hl; y++)
      for (int x = 0; x < wl; x++) {
        const int bidx = 2 * x + 2 * y * wlm1;
        idepth_l[x + y * wl] = idepth_lm[bidx] + idepth_lm[bidx + 1] + idepth_lm[bidx + wlm1] + idepth_lm[bidx + wlm1 + 1];
        weight_sums_l[x + y * wl] = weight_sums_lm[bidx] + weight_sums_lm[bidx + 1] + weight_sums_lm[bidx + wlm1] + weight_sums_lm[bidx + wlm1 + 1];
      }
  }
  // dilation: levels 0, 1 along the diagonals (:104-146), levels >= 2 along the axes (:149-190).  The reference indexes one
  // element before the map at i = w (i - 1 - wl) and one behind it at i = w*h - w - 1 (i + 1 + wl) on the diagonal variant;
  // those neighbours count as empty here.
  for (int lvl = 0; lvl < L; lvl++) {
    const int wl = o.wl[lvl], N = o.wl[lvl] * o.hl[lvl], wh = N - wl;
    float *weightSumsl = wsum[lvl].data(), *idepthl = idepth[lvl].data();
    wbak[lvl] = wsum[lvl];
    const float *bak = wbak[lvl].data();
    const int off4[2][4] = {{1 + wl, -1 - wl, wl - 1, -wl + 1}, {1, -1, wl, -wl}};
    const int *off = off4[lvl < 2 ? 0 : 1];
    for (int i = wl; i < wh; i++) {
      if (bak[i] <= 0) {
        float sum = 0, num = 0, numn = 0;
        for (int k = 0; k < 4; k++) {
          const int j = i + off[k];
          if (j < 0 || j >= N) continue;
          if (bak[j] > 0) { sum += idepthl[j]; num += bak[j]; numn++; }
        }
        if (numn > 0) { idepthl[i] = sum / numn; weightSumsl[i] = num / numn; }
      }
    }
  }
  // normalisation + point lists (:193-229)
  const Pyramid &ref = o.slots[ref_slot];
  for (int lvl = 0; lvl < L; lvl++) {
    float *weightSumsl = wsum[lvl].data(), *idepthl = idepth[lvl].data();
    const float *dIRefl = ref.lvl[lvl].dI.data();
    const int wl = o.wl[lvl], hl = o.hl[lvl];
    o.pc_u[lvl].clear(); o.pc_v[lvl].clear(); o.pc_idepth[lvl].clear(); o.pc_color[lvl].clear();
    for (int y = 2; y < hl - 2; y++)
      for (int x = 2; x < wl - 2; x++) {
        const int i = x + y * wl;
        if (weightSumsl[i] > 0) {
          idepthl[i] /= weightSumsl[i];
          const float color = dIRefl[3 * i];
          if (!std::isfinite(color) || !(idepthl[i] > 0)) { idepthl[i] = -1; continue; }
          o.pc_u[lvl].push_back((float)x); o.pc_v[lvl].push_back((float)y); o.pc_idepth[lvl].push_back(idepthl[i]); o.pc_color[lvl].push_back(color);
        } else idepthl[i] = -1;
        weightSumsl[i] = 1;
      }
  }
}

void tracker_scaleCoarseDepth(Oracle &o, float scale) {   // :244-251
  for (int lvl = 0; lvl < o.levels; lvl++)
    for (float &id : o.pc_idepth[lvl]) id /= scale;
}

// ---- trackNewestCoarse -----------------------------------------------------------------------------------------
static void affLL_of(float expF, float expT, const double g2F[2], const double g2T[2], float out[2], double *a_out) {   // AffLight::fromToVecExposure, NumType.h:157-168
  if (expF == 0 || expT == 0) expT = expF = 1;
  const double a = exp(g2T[0] - g2F[0]) * expT / expF;
  const double b = g2T[1] - a * g2F[1];
  out[0] = (float)a; out[1] = (float)b;
  if (a_out) *a_out = a;
}

bool tracker_track(Oracle &o, int new_slot, float ref_ab_exposure, float new_ab_exposure, const double ref_aff_g2l[2], int coarsestLvl,
                   sosba_track_hypothesis *hy) {
  for (int i = 0; i < 5; i++) hy->last_residuals[i] = NAN;
  double flow[3] = {1000, 1000, 1000};
  const int maxIterations[] = {10, 20, 50, 50, 50};
  const float lambdaExtrapolationLimit = 0.001;
  SE3 refToNew_current;
  refToNew_current.q = Quat{hy->q[3], hy->q[0], hy->q[1], hy->q[2]};
  refToNew_current.t = V3<double>{{hy->t[0], hy->t[1], hy->t[2]}};
  double aff_current[2] = {hy->aff_g2l[0], hy->aff_g2l[1]};
  bool haveRepeated = false;
  hy->n_passes = 0; hy->ok = 0;
  memset(hy->pass_lvl, 0, sizeof(hy->pass_lvl)); memset(hy->pass_iterations, 0, sizeof(hy->pass_iterations));
  memset(hy->pass_accept, 0, sizeof(hy->pass_accept)); memset(hy->pass_tie, 0, sizeof(hy->pass_tie)); memset(hy->pass_residual, 0, sizeof(hy->pass_residual));
  memset(hy->pass_cutoff_repeat, 0, sizeof(hy->pass_cutoff_repeat));
  const float cutoffTH = o.cfg.coarse_cutoff_th;
  const float modeA = o.cfg.affine_opt_mode_a, modeB = o.cfg.affine_opt_mode_b;

  auto calcRes = [&](int lvl, const SE3 &T, const double aff[2], float cutoff, double out6[6]) {
    double m[12];
    T.to_rowmajor34(m);
    float affLL[2];
    affLL_of(ref_ab_exposure, new_ab_exposure, ref_aff_g2l, aff, affLL, nullptr);
    int32_t counts[3];
    tracker_calcResPose(o, lvl, new_slot, m, affLL, cutoff, out6, counts);
  };
  auto calcGS = [&](int lvl, const double aff[2], double H[64], double b[8]) {
    float affLL[2];
    double a;
    affLL_of(ref_ab_exposure, new_ab_exposure, ref_aff_g2l, aff, affLL, &a);
    tracker_calcGSSSEPose(o, lvl, (float)a, (float)ref_aff_g2l[1], H, b);
  };

  for (int lvl = coarsestLvl; lvl >= 0; lvl--) {
    double H[64], b[8];
    float levelCutoffRepeat = 1;
    double resOld[6];
    calcRes(lvl, refToNew_current, aff_current, cutoffTH * levelCutoffRepeat, resOld);
    while (resOld[5] > 0.6 && levelCutoffRepeat < 50) {
      levelCutoffRepeat *= 2;
      calcRes(lvl, refToNew_current, aff_current, cutoffTH * levelCutoffRepeat, resOld);
    }
    calcGS(lvl, aff_current, H, b);
    float lambda = 0.01;
    const int pass = hy->n_passes < SOSBA_TRACK_MAX_PASSES ? hy->n_passes : SOSBA_TRACK_MAX_PASSES - 1;
    hy->pass_lvl[pass] = lvl; hy->pass_cutoff_repeat[pass] = levelCutoffRepeat;
    int iteration = 0;
    for (; iteration < maxIterations[lvl]; iteration++) {
      double Hl[64], nb[8], inc[8];
      memcpy(Hl, H, sizeof(Hl));
      for (int i = 0; i < 8; i++) Hl[9 * i] *= (1 + lambda);
      for (int i = 0; i < 8; i++) nb[i] = -b[i];
      ldlt_solve(Hl, nb, inc, 8);
      if (modeA < 0 && modeB < 0) {   // fix a, b
        double H6[36], inc6[6];
        for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) H6[6 * r + c] = Hl[8 * r + c];
        ldlt_solve(H6, nb, inc6, 6);
        for (int i = 0; i < 6; i++) inc[i] = inc6[i];
        inc[6] = inc[7] = 0;
      }
      if (!(modeA < 0) && modeB < 0) {   // fix b
        double H7[49], inc7[7];
        for (int r = 0; r < 7; r++) for (int c = 0; c < 7; c++) H7[7 * r + c] = Hl[8 * r + c];
        ldlt_solve(H7, nb, inc7, 7);
        for (int i = 0; i < 7; i++) inc[i] = inc7[i];
        inc[7] = 0;
      }
      if (modeA < 0 && !(modeB < 0)) {   // fix a
        double Hs[64], bs[8], H7[49], nb7[7], inc7[7];
        memcpy(Hs, Hl, sizeof(Hs)); memcpy(bs, b, sizeof(bs));
        for (int r = 0; r < 8; r++) Hs[8 * r + 6] = Hs[8 * r + 7];
        for (int c = 0; c < 8; c++) Hs[8 * 6 + c] = Hs[8 * 7 + c];
        bs[6] = bs[7];
        for (int r = 0; r < 7; r++) { for (int c = 0; c < 7; c++) H7[7 * r + c] = Hs[8 * r + c]; nb7[r] = -bs[r]; }
        ldlt_solve(H7, nb7, inc7, 7);
        for (int i = 0; i < 8; i++) inc[i] = 0;
        for (int i = 0; i < 6; i++) inc[i] = inc7[i];
        inc[6] = 0; inc[7] = inc7[6];
      }
      float extrapFac = 1;
      if (lambda < lambdaExtrapolationLimit) extrapFac = sqrt(sqrt(lambdaExtrapolationLimit / lambda));
      for (int i = 0; i < 8; i++) inc[i] *= extrapFac;
      double incScaled[8];
      for (int i = 0; i < 8; i++) incScaled[i] = inc[i];
      for (int i = 0; i < 3; i++) incScaled[i] *= SCALE_XI_ROT;
      for (int i = 3; i < 6; i++) incScaled[i] *= SCALE_XI_TRANS;
      incScaled[6] *= SCALE_A; incScaled[7] *= SCALE_B;
      double s = 0;
      for (int i = 0; i < 8; i++) s += incScaled[i];
      if (!std::isfinite(s)) for (int i = 0; i < 8; i++) incScaled[i] = 0;
      const SE3 refToNew_new = se3_exp(incScaled) * refToNew_current;
      const double aff_new[2] = {aff_current[0] + incScaled[6], aff_current[1] + incScaled[7]};
      double resNew[6];
      calcRes(lvl, refToNew_new, aff_new, cutoffTH * levelCutoffRepeat, resNew);
      const bool accept = (resNew[0] / resNew[1]) < (resOld[0] / resOld[1]);
      if (iteration < 64 && fabs(resNew[0] / resNew[1] - resOld[0] / resOld[1]) <= 2e-5 * fabs(resOld[0] / resOld[1])) hy->pass_tie[pass] |= 1ull << iteration;   // (reported only)
      if (accept) {
        calcGS(lvl, aff_new, H, b);
        memcpy(resOld, resNew, sizeof(resOld));
        aff_current[0] = aff_new[0]; aff_current[1] = aff_new[1];
        refToNew_current = refToNew_new;
        lambda *= 0.5;
        if (iteration < 64) hy->pass_accept[pass] |= 1ull << iteration;
      } else {
        lambda *= 4;
        if (lambda < lambdaExtrapolationLimit) lambda = lambdaExtrapolationLimit;
      }
      double nrm = 0;
      for (int i = 0; i < 8; i++) nrm += inc[i] * inc[i];
      if (!(sqrt(nrm) > 1e-3)) { iteration++; break; }
    }
    hy->pass_iterations[pass] = iteration;
    hy->last_residuals[lvl] = sqrtf((float)(resOld[0] / resOld[1]));
    hy->pass_residual[pass] = hy->last_residuals[lvl];
    hy->n_passes++;
    flow[0] = resOld[2]; flow[1] = resOld[3]; flow[2] = resOld[4];
    for (int i = 0; i < 3; i++) hy->flow_indicators[i] = flow[i];
    if (hy->last_residuals[lvl] > 1.5 * hy->min_res_for_abort[lvl]) return false;
    if (levelCutoffRepeat > 1 && !haveRepeated) { lvl++; haveRepeated = true; }
  }
  for (int i = 0; i < 3; i++) hy->flow_indicators[i] = flow[i];
  hy->q[0] = refToNew_current.q.x; hy->q[1] = refToNew_current.q.y; hy->q[2] = refToNew_current.q.z; hy->q[3] = refToNew_current.q.w;
  for (int i = 0; i < 3; i++) hy->t[i] = refToNew_current.t[i];
  hy->aff_g2l[0] = aff_current[0]; hy->aff_g2l[1] = aff_current[1];
  if ((modeA != 0 && (fabsf((float)hy->aff_g2l[0]) > 1.2)) || (modeB != 0 && (fabsf((float)hy->aff_g2l[1]) > 200))) return false;
  float relAff[2];
  affLL_of(ref_ab_exposure, new_ab_exposure, ref_aff_g2l, hy->aff_g2l, relAff, nullptr);
  if ((modeA == 0 && (fabsf(logf((float)relAff[0])) > 1.5)) || (modeB == 0 && (fabsf((float)relAff[1]) > 200))) return false;
  if (modeA < 0) hy->aff_g2l[0] = 0;
  if (modeB < 0) hy->aff_g2l[1] = 0;
  hy->ok = 1;
  return true;
}

// ---- optimizeScale ------------------------------------------------------------------------------------------------
void scale_optimize(Oracle &o, int stereo_slot, int coarsestLvl, sosba_scale_hypothesis *hy) {
  for (int i = 0; i < 5; i++) hy->last_residuals[i] = NAN;
  const int maxIterations[] = {10, 20, 50, 50, 50};
  const float lambdaExtrapolationLimit = 0.001;
  float scale_current = hy->scale;
  bool haveRepeated = false;
  hy->n_passes = 0;
  memset(hy->pass_lvl, 0, sizeof(hy->pass_lvl)); memset(hy->pass_iterations, 0, sizeof(hy->pass_iterations)); memset(hy->pass_accept, 0, sizeof(hy->pass_accept));
  memset(hy->pass_tie, 0, sizeof(hy->pass_tie));
  const float cutoffTH = o.cfg.coarse_cutoff_th;
  for (int lvl = coarsestLvl; lvl >= 0; lvl--) {
    float H, b;
    float levelCutoffRepeat = 1;
    double resOld[6], resNew[6];
    int32_t counts[3];
    scale_calcRes(o, lvl, stereo_slot, scale_current, cutoffTH * levelCutoffRepeat, resOld, counts);
    while (resOld[5] > 0.6 && levelCutoffRepeat < 50) {
      levelCutoffRepeat *= 2;
      scale_calcRes(o, lvl, stereo_slot, scale_current, cutoffTH * levelCutoffRepeat, resOld, counts);
    }
    scale_calcGSSSE(o, lvl, scale_current, &H, &b);
    float lambda = 0.01;
    const int pass = hy->n_passes < SOSBA_TRACK_MAX_PASSES ? hy->n_passes : SOSBA_TRACK_MAX_PASSES - 1;
    hy->pass_lvl[pass] = lvl;
    int iteration = 0;
    for (; iteration < maxIterations[lvl]; iteration++) {
      float Hl = H;
      Hl *= (1 + lambda);
      float inc = -b / Hl;
      float extrapFac = 1;
      if (lambda < lambdaExtrapolationLimit) extrapFac = sqrt(sqrt(lambdaExtrapolationLimit / lambda));
      inc *= extrapFac;
      if (!std::isfinite(inc) || fabs(inc) > scale_current) inc = 0.0;
      const float scale_new = scale_current + inc;
      scale_calcRes(o, lvl, stereo_slot, scale_new, cutoffTH * levelCutoffRepeat, resNew, counts);
      const bool accept = (resNew[0] / resNew[1]) < (resOld[0] / resOld[1]);
      if (iteration < 64 && fabs(resNew[0] / resNew[1] - resOld[0] / resOld[1]) <= 2e-5 * fabs(resOld[0] / resOld[1])) hy->pass_tie[pass] |= 1ull << iteration;   // (reported only)
      if (accept) {
        scale_calcGSSSE(o, lvl, scale_new, &H, &b);
        memcpy(resOld, resNew, sizeof(resOld));
        scale_current = scale_new;
        lambda *= 0.5;
        if (iteration < 64) hy->pass_accept[pass] |= 1ull << iteration;
      } else {
        lambda *= 4;
        if (lambda < lambdaExtrapolationLimit) lambda = lambdaExtrapolationLimit;
      }
      if (!(inc > 1e-3)) { iteration++; break; }   // sic: the signed increment (ScaleOptimizer.cpp:200)
    }
    hy->pass_iterations[pass] = iteration;
    hy->last_residuals[lvl] = sqrtf((float)(resOld[0] / resOld[1]));
    hy->n_passes++;
    if (levelCutoffRepeat > 1 && !haveRepeated) { lvl++; haveRepeated = true; }
  }
  hy->scale = scale_current;
  hy->error = (float)hy->last_residuals[0];
}


// ---- CoarseDistanceMap::makeDistanceMap + growDistBFS (CoarseTracker.cpp:789-916) --------------------------------------
void distance_map(Oracle &o, int nhosts, const float *KRKi, const float *Kt, int n, const int32_t *host, const float *u, const float *v,
                  const float *idepth, float *dist) {
  const int w1 = o.wl[1], h1 = o.hl[1], wh1 = w1 * h1;
  for (int i = 0; i < wh1; i++) dist[i] = 1000;
  std::vector<int> l1x, l1y, l2x, l2y;
  for (int i = 0; i < n; i++) {   // :801-821
    const float *M = KRKi + 9 * host[i], *t = Kt + 3 * host[i];
    const float p0 = ((M[0] * u[i] + M[1] * v[i]) + M[2] * 1) + t[0] * idepth[i];
    const float p1 = ((M[3] * u[i] + M[4] * v[i]) + M[5] * 1) + t[1] * idepth[i];
    const float p2 = ((M[6] * u[i] + M[7] * v[i]) + M[8] * 1) + t[2] * idepth[i];
    const int uu = p0 / p2 + 0.5f, vv = p1 / p2 + 0.5f;
    if (!(uu > 0 && vv > 0 && uu < w1 && vv < h1)) continue;
    dist[uu + w1 * vv] = 0;
    l1x.push_back(uu); l1y.push_back(vv);
  }
  for (int k = 1; k < 40; k++) {   // growDistBFS :830-916
    std::swap(l1x, l2x); std::swap(l1y, l2y);
    l1x.clear(); l1y.clear();
    const int nn = (k % 2 == 0) ? 4 : 8;
    const int dx[8] = {1, -1, 0, 0, 1, -1, -1, 1}, dy[8] = {0, 0, 1, -1, 1, 1, -1, -1};
    for (size_t i = 0; i < l2x.size(); i++) {
      const int x = l2x[i], y = l2y[i];
      if (x == 0 || y == 0 || x == w1 - 1 || y == h1 - 1) continue;
      for (int q = 0; q < nn; q++) {
        const int idx = (x + dx[q]) + (y + dy[q]) * w1;
        if (dist[idx] > k) { dist[idx] = k; l1x.push_back(x + dx[q]); l1y.push_back(y + dy[q]); }
      }
    }
  }
}

}  // namespace orc
