// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).  PARITY UNPINNED.
//
// orc_ba.cpp: CPU restatement of the windowed photometric bundle-adjustment path:
//   FrameHessian::makeImages            src/FullSystem/HessianBlocks.cpp:121-176
//   PointFrameResidual::linearize       src/FullSystem/Residuals.cpp:77-271
//   projectPoint x2                     src/FullSystem/ResidualProjections.h:43-73
//   getInterpolatedElement33            src/util/globalFuncs.h:68-82
//   PointFrameResidual::applyRes        src/FullSystem/Residuals.cpp:304-321
//   EFResidual::takeDataF / fixLinearizationF   src/OptimizationBackend/EnergyFunctionalStructs.cpp:36-45, 75-103
//   AccumulatedTopHessianSSE            src/OptimizationBackend/AccumulatedTopHessian.{h,cpp}
//   AccumulatedSCHessianSSE             src/OptimizationBackend/AccumulatedSCHessian.{h,cpp}
//   EnergyFunctional::accumulate*/solveSystemF/resubstituteF_MT/marginalizePointsF
//                                       src/OptimizationBackend/EnergyFunctional.cpp:197-254, 1029-1184, 496-551, 891-936
//   FullSystem::linearizeAll / setNewFrameEnergyTH   src/FullSystem/FullSystemOptimize.cpp:44-182
#include <cassert>
#include <cmath>
#include <cstdio>

#include "orc_core.h"
#include "orc_sample.h"

namespace orc {

// ---------------------------------------------------------------------------------------------
// a1  FrameHessian::makeImages (HessianBlocks.cpp:121-176).  Rows 0 and h-1 keep dx,dy,abs = 0
// here (uninitialised `new[]` memory in the reference).
void make_images(Oracle &o, int slot, const float *color, const float *B) {
  Pyramid &P = o.slots[slot];
  for (int l = 0; l < o.levels; l++) {
    P.lvl[l].w = o.wl[l]; P.lvl[l].h = o.hl[l];
    P.lvl[l].dI.assign((size_t)3 * o.wl[l] * o.hl[l], 0.f);
    P.lvl[l].absg.assign((size_t)o.wl[l] * o.hl[l], 0.f);
  }
  int w = o.wl[0], h = o.hl[0];
  for (int i = 0; i < w * h; i++) P.lvl[0].dI[3 * i] = color[i];
  for (int lvl = 0; lvl < o.levels; lvl++) {
    int wl = o.wl[lvl], hl = o.hl[lvl];
    float *dI_l = P.lvl[lvl].dI.data();
    float *dabs_l = P.lvl[lvl].absg.data();
    if (lvl > 0) {
      int wlm1 = o.wl[lvl - 1];
      const float *dI_lm = P.lvl[lvl - 1].dI.data();
      for (int y = 0; y < hl; y++)
        for (int x = 0; x < wl; x++)
          dI_l[3 * (x + y * wl)] = 0.25f * (dI_lm[3 * (2 * x + 2 * y * wlm1)] + dI_lm[3 * (2 * x + 1 + 2 * y * wlm1)] +
                                            dI_lm[3 * (2 * x + 2 * y * wlm1 + wlm1)] + dI_lm[3 * (2 * x + 1 + 2 * y * wlm1 + wlm1)]);
    }
    for (int idx = wl; idx < wl * (hl - 1); idx++) {
      float dx = 0.5f * (dI_l[3 * (idx + 1)] - dI_l[3 * (idx - 1)]);
      float dy = 0.5f * (dI_l[3 * (idx + wl)] - dI_l[3 * (idx - wl)]);
      if (!std::isfinite(dx)) dx = 0;
      if (!std::isfinite(dy)) dy = 0;
      dI_l[3 * idx + 1] = dx;
      dI_l[3 * idx + 2] = dy;
      dabs_l[idx] = dx * dx + dy * dy;
      if (o.cfg.gamma_weights_pixel_select == 1 && B != nullptr) {
        // CalibHessian::getBGradOnly, HessianBlocks.h:519-526
        int c = (int)(dI_l[3 * idx] + 0.5f);
        if (c < 5) c = 5;
        if (c > 250) c = 250;
        float gw = B[c + 1] - B[c];
        dabs_l[idx] *= gw * gw;
      }
    }
  }
  P.valid = true;
}

// ---------------------------------------------------------------------------------------------
// a3  PointFrameResidual::linearize (Residuals.cpp:77-271)
double linearize(Oracle &o, Res &r) {
  r.state_NewEnergyWithOutlier = -1;
  if (r.state_state == SOSBA_RES_OOB) { r.state_NewState = SOSBA_RES_OOB; return r.state_energy; }

  const Pt &point = o.pts[r.point];
  const Precalc &pc = o.pre[r.host * o.nf + r.target];
  float energyLeft = 0;
  const Level &L0 = o.slots[o.frame_slot[r.target]].lvl[0];
  const float *dIl = L0.dI.data();
  const float *color = point.color, *weights = point.weights;
  const float affLL0 = pc.aff[0], affLL1 = pc.aff[1];
  const float b0 = pc.b0;
  const float fxl = o.fxl, fyl = o.fyl, cxl = o.cxl, cyl = o.cyl, fxli = o.fxli, fyli = o.fyli;
  const float *R0 = pc.RTll_0, *t0 = pc.tTll_0;
  RawJ &J = r.Jdata();

  float d_xi_x[6], d_xi_y[6], d_C_x[4], d_C_y[4], d_d_x, d_d_y;
  {
    // projectPoint (ResidualProjections.h:52-73) with dx=dy=0, idepth_zero_scaled, eval-point pose
    float KliP[3] = {(point.u + 0 - cxl) * fxli, (point.v + 0 - cyl) * fyli, 1};
    float ptp[3];
    for (int i = 0; i < 3; i++)
      ptp[i] = ((R0[3 * i] * KliP[0] + R0[3 * i + 1] * KliP[1]) + R0[3 * i + 2] * KliP[2]) + t0[i] * point.idepth_zero_scaled;
    float drescale = 1.0f / ptp[2];
    float new_idepth = point.idepth_zero_scaled * drescale;
    if (!(drescale > 0)) { r.state_NewState = SOSBA_RES_OOB; return r.state_energy; }
    float u = ptp[0] * drescale, v = ptp[1] * drescale;
    float Ku = u * fxl + cxl, Kv = v * fyl + cyl;
    if (!(Ku > 1.1f && Kv > 1.1f && Ku < o.wM3G && Kv < o.hM3G)) { r.state_NewState = SOSBA_RES_OOB; return r.state_energy; }

    r.centerProjectedTo[0] = Ku; r.centerProjectedTo[1] = Kv; r.centerProjectedTo[2] = new_idepth;

    d_d_x = drescale * (t0[0] - t0[2] * u) * SCALE_IDEPTH * fxl;
    d_d_y = drescale * (t0[1] - t0[2] * v) * SCALE_IDEPTH * fyl;

    d_C_x[2] = drescale * (R0[6] * u - R0[0]);
    d_C_x[3] = fxl * drescale * (R0[7] * u - R0[1]) * fyli;
    d_C_x[0] = KliP[0] * d_C_x[2];
    d_C_x[1] = KliP[1] * d_C_x[3];

    d_C_y[2] = fyl * drescale * (R0[6] * v - R0[3]) * fxli;
    d_C_y[3] = drescale * (R0[7] * v - R0[4]);
    d_C_y[0] = KliP[0] * d_C_y[2];
    d_C_y[1] = KliP[1] * d_C_y[3];

    d_C_x[0] = (d_C_x[0] + u) * SCALE_F;
    d_C_x[1] *= SCALE_F;
    d_C_x[2] = (d_C_x[2] + 1) * SCALE_C;
    d_C_x[3] *= SCALE_C;

    d_C_y[0] *= SCALE_F;
    d_C_y[1] = (d_C_y[1] + v) * SCALE_F;
    d_C_y[2] *= SCALE_C;
    d_C_y[3] = (d_C_y[3] + 1) * SCALE_C;

    d_xi_x[0] = new_idepth * fxl;
    d_xi_x[1] = 0;
    d_xi_x[2] = -new_idepth * u * fxl;
    d_xi_x[3] = -u * v * fxl;
    d_xi_x[4] = (1 + u * u) * fxl;
    d_xi_x[5] = -v * fxl;

    d_xi_y[0] = 0;
    d_xi_y[1] = new_idepth * fyl;
    d_xi_y[2] = -new_idepth * v * fyl;
    d_xi_y[3] = -(1 + v * v) * fyl;
    d_xi_y[4] = u * v * fyl;
    d_xi_y[5] = u * fyl;
  }
  for (int i = 0; i < 6; i++) { J.Jpdxi[0][i] = d_xi_x[i]; J.Jpdxi[1][i] = d_xi_y[i]; }
  for (int i = 0; i < 4; i++) { J.Jpdc[0][i] = d_C_x[i]; J.Jpdc[1][i] = d_C_y[i]; }
  J.Jpdd[0] = d_d_x; J.Jpdd[1] = d_d_y;

  float JIdxJIdx_00 = 0, JIdxJIdx_11 = 0, JIdxJIdx_10 = 0;
  float JabJIdx_00 = 0, JabJIdx_01 = 0, JabJIdx_10 = 0, JabJIdx_11 = 0;
  float JabJab_00 = 0, JabJab_01 = 0, JabJab_11 = 0;
  float wJI2_sum = 0;
  const float *KRKi = pc.KRKi, *Kt = pc.Kt;
  const float huberTH = o.cfg.huber_th, oTH = o.cfg.outlier_th_sum_component;

  for (int idx = 0; idx < patternNum; idx++) {
    // projectPoint (ResidualProjections.h:43-50) with current-state KRKi, Kt and idepth_scaled
    float up = point.u + patternP[idx][0], vp = point.v + patternP[idx][1];
    float ptp[3];
    for (int i = 0; i < 3; i++)
      ptp[i] = ((KRKi[3 * i] * up + KRKi[3 * i + 1] * vp) + KRKi[3 * i + 2] * 1.0f) + Kt[i] * point.idepth_scaled;
    float Ku = ptp[0] / ptp[2], Kv = ptp[1] / ptp[2];
    if (!(Ku > 1.1f && Kv > 1.1f && Ku < o.wM3G && Kv < o.hM3G)) { r.state_NewState = SOSBA_RES_OOB; return r.state_energy; }
    r.projectedTo[idx][0] = Ku; r.projectedTo[idx][1] = Kv;

    float hitColor[3];
    interp33(dIl, Ku, Kv, o.wl[0], hitColor);
    float residual = hitColor[0] - (float)(affLL0 * color[idx] + affLL1);
    float drdA = (color[idx] - b0);
    if (!std::isfinite(hitColor[0])) { r.state_NewState = SOSBA_RES_OOB; return r.state_energy; }

    float w = sqrtf(oTH / (oTH + (hitColor[1] * hitColor[1] + hitColor[2] * hitColor[2])));
    w = 0.5f * (w + weights[idx]);
    float hw = fabsf(residual) < huberTH ? 1 : huberTH / fabsf(residual);
    energyLeft += w * w * hw * residual * residual * (2 - hw);
    {
      if (hw < 1) hw = sqrtf(hw);
      hw = hw * w;
      hitColor[1] *= hw;
      hitColor[2] *= hw;
      J.resF[idx] = residual * hw;
      J.JIdx[0][idx] = hitColor[1];
      J.JIdx[1][idx] = hitColor[2];
      J.JabF[0][idx] = drdA * hw;
      J.JabF[1][idx] = hw;

      JIdxJIdx_00 += hitColor[1] * hitColor[1];
      JIdxJIdx_11 += hitColor[2] * hitColor[2];
      JIdxJIdx_10 += hitColor[1] * hitColor[2];

      JabJIdx_00 += drdA * hw * hitColor[1];
      JabJIdx_01 += drdA * hw * hitColor[2];
      JabJIdx_10 += hw * hitColor[1];
      JabJIdx_11 += hw * hitColor[2];

      JabJab_00 += drdA * drdA * hw * hw;
      JabJab_01 += drdA * hw * hw;
      JabJab_11 += hw * hw;

      wJI2_sum += hw * hw * (hitColor[1] * hitColor[1] + hitColor[2] * hitColor[2]);

      if (o.cfg.affine_opt_mode_a < 0) J.JabF[0][idx] = 0;
      if (o.cfg.affine_opt_mode_b < 0) J.JabF[1][idx] = 0;
    }
  }
  J.JIdx2[0][0] = JIdxJIdx_00; J.JIdx2[0][1] = JIdxJIdx_10; J.JIdx2[1][0] = JIdxJIdx_10; J.JIdx2[1][1] = JIdxJIdx_11;
  J.JabJIdx[0][0] = JabJIdx_00; J.JabJIdx[0][1] = JabJIdx_01; J.JabJIdx[1][0] = JabJIdx_10; J.JabJIdx[1][1] = JabJIdx_11;
  J.Jab2[0][0] = JabJab_00; J.Jab2[0][1] = JabJab_01; J.Jab2[1][0] = JabJab_01; J.Jab2[1][1] = JabJab_11;

  r.state_NewEnergyWithOutlier = energyLeft;
  const float th = std::max<float>(o.frameEnergyTH[r.host], o.frameEnergyTH[r.target]);
  if (energyLeft > th || wJI2_sum < 2) {
    energyLeft = th;
    r.state_NewState = SOSBA_RES_OUTLIER;
  } else {
    r.state_NewState = SOSBA_RES_IN;
  }
  r.state_NewEnergy = energyLeft;
  return energyLeft;
}

// EFResidual::takeDataF (EnergyFunctionalStructs.cpp:36-45)
static void takeDataF(Res &r) {
  r.sel ^= 1;  // std::swap(J, data->J)
  const RawJ &J = r.Jef();
  float JI_JI_Jd[2] = {J.JIdx2[0][0] * J.Jpdd[0] + J.JIdx2[0][1] * J.Jpdd[1], J.JIdx2[1][0] * J.Jpdd[0] + J.JIdx2[1][1] * J.Jpdd[1]};
  for (int i = 0; i < 6; i++) r.JpJdF[i] = J.Jpdxi[0][i] * JI_JI_Jd[0] + J.Jpdxi[1][i] * JI_JI_Jd[1];
  r.JpJdF[6] = J.JabJIdx[0][0] * J.Jpdd[0] + J.JabJIdx[0][1] * J.Jpdd[1];
  r.JpJdF[7] = J.JabJIdx[1][0] * J.Jpdd[0] + J.JabJIdx[1][1] * J.Jpdd[1];
}

// a5  PointFrameResidual::applyRes (Residuals.cpp:304-321)
void applyRes(Res &r, bool copyJacobians) {
  if (copyJacobians) {
    if (r.state_state == SOSBA_RES_OOB) return;  // can never go back from OOB
    if (r.state_NewState == SOSBA_RES_IN) { r.isActive = true; takeDataF(r); }
    else r.isActive = false;
  }
  r.state_state = r.state_NewState;
  r.state_energy = r.state_NewEnergy;
}

// a13  EFResidual::fixLinearizationF (EnergyFunctionalStructs.cpp:75-103)
void fixLinearizationF(Oracle &o, Res &r) {
  const float *dp = &o.adHTdeltaF[8 * (r.host + o.nf * r.target)];
  const RawJ &J = r.Jef();
  const float deltaF = o.pts[r.point].deltaF;
  auto dot6 = [](const float *a, const float *b) { float s = a[0] * b[0]; for (int i = 1; i < 6; i++) s += a[i] * b[i]; return s; };
  auto dot4 = [](const float *a, const float *b) { float s = a[0] * b[0]; for (int i = 1; i < 4; i++) s += a[i] * b[i]; return s; };
  float Jp_delta_x = dot6(J.Jpdxi[0], dp) + dot4(J.Jpdc[0], o.cDeltaF) + J.Jpdd[0] * deltaF;
  float Jp_delta_y = dot6(J.Jpdxi[1], dp) + dot4(J.Jpdc[1], o.cDeltaF) + J.Jpdd[1] * deltaF;
  float delta_a = dp[6], delta_b = dp[7];
  for (int i = 0; i < patternNum; i++) {
    float rtz = J.resF[i];
    rtz = rtz - J.JIdx[0][i] * Jp_delta_x;
    rtz = rtz - J.JIdx[1][i] * Jp_delta_y;
    rtz = rtz - J.JabF[0][i] * delta_a;
    rtz = rtz - J.JabF[1][i] * delta_b;
    r.res_toZeroF[i] = rtz;
  }
  r.isLinearized = true;
}

// ---------------------------------------------------------------------------------------------
// a4  FullSystem::linearizeAll_Reductor (FullSystemOptimize.cpp:44-77)
static void linearizeAll_Reductor(Oracle &o, bool fixLinearization, std::vector<std::vector<int>> *toRemove, int min, int max,
                                  Stats10 *stats, int tid) {
  for (int k = min; k < max; k++) {
    Res &r = o.res[o.activeResiduals[k]];
    stats->v[0] += linearize(o, r);
    if (fixLinearization) {
      applyRes(r, true);
      if (r.isActive) {
        if (r.isNew) {
          Pt &p = o.pts[r.point];
          const Precalc &pc = o.pre[r.host * o.nf + r.target];
          float inf[3], ptp[3];
          for (int i = 0; i < 3; i++) inf[i] = (pc.KRKi[3 * i] * p.u + pc.KRKi[3 * i + 1] * p.v) + pc.KRKi[3 * i + 2] * 1.0f;
          for (int i = 0; i < 3; i++) ptp[i] = inf[i] + pc.Kt[i] * p.idepth_scaled;
          float ex = inf[0] / inf[2] - ptp[0] / ptp[2], ey = inf[1] / inf[2] - ptp[1] / ptp[2];
          float relBS = 0.01 * sqrtf(ex * ex + ey * ey);
          if (relBS > p.maxRelBaseline) p.maxRelBaseline = relBS;
          p.numGoodResiduals++;
        }
      } else {
        (*toRemove)[tid].push_back(o.activeResiduals[k]);
      }
    }
  }
}

// FullSystem::setNewFrameEnergyTH (FullSystemOptimize.cpp:84-124)
static void setNewFrameEnergyTH(Oracle &o) {
  std::vector<float> allResVec;
  allResVec.reserve(o.activeResiduals.size() * 2);
  const int newFrame = o.nf - 1;
  for (int id : o.activeResiduals) {
    const Res &r = o.res[id];
    if (r.state_NewEnergyWithOutlier >= 0 && r.target == newFrame) allResVec.push_back((float)r.state_NewEnergyWithOutlier);
  }
  if (allResVec.size() == 0) { o.frameEnergyTH[newFrame] = 12 * 12 * patternNum; return; }
  int nthIdx = (int)(o.cfg.frame_energy_th_n * allResVec.size());
  std::nth_element(allResVec.begin(), allResVec.begin() + nthIdx, allResVec.end());
  float nthElement = sqrtf(allResVec[nthIdx]);
  float th = nthElement * o.cfg.frame_energy_th_fac_median;
  th = 26.0f * o.cfg.frame_energy_th_const_weight + th * (1 - o.cfg.frame_energy_th_const_weight);
  th = th * th;
  th *= o.cfg.overall_energy_th_weight * o.cfg.overall_energy_th_weight;
  o.frameEnergyTH[newFrame] = th;
}

// FullSystem::linearizeAll (FullSystemOptimize.cpp:125-182)
void linearizeAll(Oracle &o, bool fixLinearization, sosba_linearize_out *out) {
  std::vector<std::vector<int>> toRemove(std::max(o.T, 1));
  double lastEnergyP = 0;
  const int n = (int)o.activeResiduals.size();
  if (o.MT) {
    o.red->reduce([&](int a, int b, Stats10 *s, int tid) { linearizeAll_Reductor(o, fixLinearization, &toRemove, a, b, s, tid); }, 0, n, 0);
    lastEnergyP = o.red->stats.v[0];
  } else {
    Stats10 s; memset(&s, 0, sizeof(s));
    linearizeAll_Reductor(o, fixLinearization, &toRemove, 0, n, &s, 0);
    lastEnergyP = s.v[0];
  }
  setNewFrameEnergyTH(o);
  int nRemoved = 0;
  if (fixLinearization) {
    for (auto &v : toRemove)
      for (int id : v) { o.res[id].dropped = true; nRemoved++; }  // ef->dropResidual + deleteOut (:160-176)
  }
  if (out) {
    out->energy = lastEnergyP;
    out->new_frame_energy_th = o.frameEnergyTH[o.nf - 1];
    out->n_in = out->n_oob = out->n_outlier = 0;
    for (int id : o.activeResiduals) {
      int s = o.res[id].state_NewState;
      if (s == SOSBA_RES_IN) out->n_in++; else if (s == SOSBA_RES_OOB) out->n_oob++; else out->n_outlier++;
    }
    out->n_removed = nRemoved;
  }
}

// ---------------------------------------------------------------------------------------------
// a6  AccumulatedTopHessianSSE::addPoint<mode> (AccumulatedTopHessian.cpp:35-147)
template <int mode> static void top_addPoint(Oracle &o, std::vector<std::vector<AccumulatorApprox>> &acc, std::vector<int> &nres, Pt &p, int tid) {
  const float *dc = o.cDeltaF;
  float dd = p.deltaF;
  float bd_acc = 0, Hdd_acc = 0, Hcd_acc[4] = {0, 0, 0, 0};
  for (int ri = p.res_begin; ri < p.res_end; ri++) {
    Res &r = o.res[ri];
    if (r.dropped) continue;
    if (mode == 0) { if (r.isLinearized || !r.isActive) continue; }
    if (mode == 1) { if (!r.isLinearized || !r.isActive) continue; }
    if (mode == 2) { if (!r.isActive) continue; }
    const RawJ &rJ = r.Jef();
    int htIDX = r.host + r.target * o.nf;
    const float *dp = &o.adHTdeltaF[8 * htIDX];
    float resApprox[8];
    if (mode == 0) for (int i = 0; i < 8; i++) resApprox[i] = rJ.resF[i];
    if (mode == 2) for (int i = 0; i < 8; i++) resApprox[i] = r.res_toZeroF[i];
    if (mode == 1) {
      auto dot6 = [](const float *a, const float *b) { float s = a[0] * b[0]; for (int i = 1; i < 6; i++) s += a[i] * b[i]; return s; };
      auto dot4 = [](const float *a, const float *b) { float s = a[0] * b[0]; for (int i = 1; i < 4; i++) s += a[i] * b[i]; return s; };
      float Jp_delta_x = dot6(rJ.Jpdxi[0], dp) + dot4(rJ.Jpdc[0], dc) + rJ.Jpdd[0] * dd;
      float Jp_delta_y = dot6(rJ.Jpdxi[1], dp) + dot4(rJ.Jpdc[1], dc) + rJ.Jpdd[1] * dd;
      float delta_a = dp[6], delta_b = dp[7];
      for (int i = 0; i < patternNum; i++) {
        float rtz = r.res_toZeroF[i];
        rtz = rtz + rJ.JIdx[0][i] * Jp_delta_x;
        rtz = rtz + rJ.JIdx[1][i] * Jp_delta_y;
        rtz = rtz + rJ.JabF[0][i] * delta_a;
        rtz = rtz + rJ.JabF[1][i] * delta_b;
        resApprox[i] = rtz;
      }
    }
    float JI_r[2] = {0, 0}, Jab_r[2] = {0, 0}, rr = 0;
    for (int i = 0; i < patternNum; i++) {
      JI_r[0] += resApprox[i] * rJ.JIdx[0][i];
      JI_r[1] += resApprox[i] * rJ.JIdx[1][i];
      Jab_r[0] += resApprox[i] * rJ.JabF[0][i];
      Jab_r[1] += resApprox[i] * rJ.JabF[1][i];
      rr += resApprox[i] * resApprox[i];
    }
    AccumulatorApprox &a = acc[tid][htIDX];
    a.update(rJ.Jpdc[0], rJ.Jpdxi[0], rJ.Jpdc[1], rJ.Jpdxi[1], rJ.JIdx2[0][0], rJ.JIdx2[0][1], rJ.JIdx2[1][1]);
    a.updateBotRight(rJ.Jab2[0][0], rJ.Jab2[0][1], Jab_r[0], rJ.Jab2[1][1], Jab_r[1], rr);
    a.updateTopRight(rJ.Jpdc[0], rJ.Jpdxi[0], rJ.Jpdc[1], rJ.Jpdxi[1], rJ.JabJIdx[0][0], rJ.JabJIdx[0][1], rJ.JabJIdx[1][0],
                     rJ.JabJIdx[1][1], JI_r[0], JI_r[1]);
    float Ji2_Jpdd[2] = {rJ.JIdx2[0][0] * rJ.Jpdd[0] + rJ.JIdx2[0][1] * rJ.Jpdd[1], rJ.JIdx2[1][0] * rJ.Jpdd[0] + rJ.JIdx2[1][1] * rJ.Jpdd[1]};
    bd_acc += JI_r[0] * rJ.Jpdd[0] + JI_r[1] * rJ.Jpdd[1];
    Hdd_acc += Ji2_Jpdd[0] * rJ.Jpdd[0] + Ji2_Jpdd[1] * rJ.Jpdd[1];
    for (int i = 0; i < 4; i++) Hcd_acc[i] += rJ.Jpdc[0][i] * Ji2_Jpdd[0] + rJ.Jpdc[1][i] * Ji2_Jpdd[1];
    nres[tid]++;
  }
  if (mode == 0) { p.Hdd_accAF = Hdd_acc; p.bd_accAF = bd_acc; for (int i = 0; i < 4; i++) p.Hcd_accAF[i] = Hcd_acc[i]; }
  if (mode == 1 || mode == 2) { p.Hdd_accLF = Hdd_acc; p.bd_accLF = bd_acc; for (int i = 0; i < 4; i++) p.Hcd_accLF[i] = Hcd_acc[i]; }
  if (mode == 2) { for (int i = 0; i < 4; i++) p.Hcd_accAF[i] = 0; p.Hdd_accAF = 0; p.bd_accAF = 0; }
}

// a8  AccumulatedSCHessianSSE::addPoint (AccumulatedSCHessian.cpp:32-79)
static void sc_addPoint(Oracle &o, Pt &p, bool shiftPriorToZero, int tid) {
  int ngoodres = 0;
  for (int ri = p.res_begin; ri < p.res_end; ri++) if (!o.res[ri].dropped && o.res[ri].isActive) ngoodres++;
  if (ngoodres == 0) { p.HdiF = 0; p.bdSumF = 0; p.idepth_hessian = 0; p.maxRelBaseline = 0; return; }
  float H = p.Hdd_accAF + p.Hdd_accLF + p.priorF;
  if (H < 1e-10) H = 1e-10;
  p.idepth_hessian = H;
  p.HdiF = 1.0 / H;
  p.bdSumF = p.bd_accAF + p.bd_accLF;
  if (shiftPriorToZero) p.bdSumF += p.priorF * p.deltaF;
  float Hcd[4];
  for (int i = 0; i < 4; i++) Hcd[i] = p.Hcd_accAF[i] + p.Hcd_accLF[i];
  o.accHcc[tid].update(Hcd, Hcd, p.HdiF);
  o.accbc[tid].update(Hcd, p.bdSumF * p.HdiF);
  const int nf = o.nf, nFrames2 = nf * nf;
  for (int r1i = p.res_begin; r1i < p.res_end; r1i++) {
    Res &r1 = o.res[r1i];
    if (r1.dropped || !r1.isActive) continue;
    int r1ht = r1.host + r1.target * nf;
    for (int r2i = p.res_begin; r2i < p.res_end; r2i++) {
      Res &r2 = o.res[r2i];
      if (r2.dropped || !r2.isActive) continue;
      o.accD[tid][r1ht + r2.target * nFrames2].update(r1.JpJdF, r2.JpJdF, p.HdiF);
    }
    o.accE[tid][r1ht].update(r1.JpJdF, Hcd, p.HdiF);
    o.accEB[tid][r1ht].update(r1.JpJdF, p.HdiF * p.bdSumF);
  }
}

static inline void mat8_mul(const double *A, const double *B, double *C) {  // C = A*B, 8x8 row-major
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 8; j++) { double s = 0; for (int k = 0; k < 8; k++) s += A[i * 8 + k] * B[k * 8 + j]; C[i * 8 + j] = s; }
}
static inline void mat8_mulT(const double *A, const double *B, double *C) {  // C = A*B^T
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 8; j++) { double s = 0; for (int k = 0; k < 8; k++) s += A[i * 8 + k] * B[j * 8 + k]; C[i * 8 + j] = s; }
}

// a7  AccumulatedTopHessianSSE::stitchDoubleInternal (AccumulatedTopHessian.cpp:231-301); H is D*D row-major.
static void top_stitchInternal(Oracle &o, std::vector<std::vector<AccumulatorApprox>> &acc, double *H, double *b, bool usePrior, int min, int max,
                               int toAggregate, bool doPrior) {
  const int nf = o.nf, D = CPARS + 8 * nf;
#define HH(r, c) H[(size_t)(r) * D + (c)]
  for (int k = min; k < max; k++) {
    int h = k % nf, t = k / nf;
    int hIdx = CPARS + h * 8, tIdx = CPARS + t * 8, aidx = h + nf * t;
    double accH[13][13];
    memset(accH, 0, sizeof(accH));
    for (int tid2 = 0; tid2 < toAggregate; tid2++) {
      acc[tid2][aidx].finish();
      if (acc[tid2][aidx].num == 0) continue;
      for (int i = 0; i < 13; i++) for (int j = 0; j < 13; j++) accH[i][j] += (double)acc[tid2][aidx].H[i][j];
    }
    const double *Ah = &o.adHost[64 * aidx], *At = &o.adTarget[64 * aidx];
    double P[64], tmp[64], out[64];
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) P[i * 8 + j] = accH[CPARS + i][CPARS + j];
    mat8_mul(Ah, P, tmp); mat8_mulT(tmp, Ah, out);
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) HH(hIdx + i, hIdx + j) += out[i * 8 + j];
    mat8_mul(At, P, tmp); mat8_mulT(tmp, At, out);
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) HH(tIdx + i, tIdx + j) += out[i * 8 + j];
    mat8_mul(Ah, P, tmp); mat8_mulT(tmp, At, out);
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) HH(hIdx + i, tIdx + j) += out[i * 8 + j];
    for (int i = 0; i < 8; i++)
      for (int c = 0; c < CPARS; c++) {
        double sh = 0, st = 0;
        for (int k2 = 0; k2 < 8; k2++) { sh += Ah[i * 8 + k2] * accH[CPARS + k2][c]; st += At[i * 8 + k2] * accH[CPARS + k2][c]; }
        HH(hIdx + i, c) += sh; HH(tIdx + i, c) += st;
      }
    for (int i = 0; i < CPARS; i++) for (int j = 0; j < CPARS; j++) HH(i, j) += accH[i][j];
    for (int i = 0; i < 8; i++) {
      double sh = 0, st = 0;
      for (int k2 = 0; k2 < 8; k2++) { sh += Ah[i * 8 + k2] * accH[CPARS + k2][CPARS + 8]; st += At[i * 8 + k2] * accH[CPARS + k2][CPARS + 8]; }
      b[hIdx + i] += sh; b[tIdx + i] += st;
    }
    for (int i = 0; i < CPARS; i++) b[i] += accH[i][CPARS + 8];
  }
  if (doPrior && usePrior) {  // :292-300
    for (int i = 0; i < CPARS; i++) { HH(i, i) += o.cPrior[i]; b[i] += o.cPrior[i] * (double)o.cDeltaF[i]; }
    for (int h = 0; h < nf; h++)
      for (int i = 0; i < 8; i++) { HH(CPARS + h * 8 + i, CPARS + h * 8 + i) += o.fprior[h * 8 + i]; b[CPARS + h * 8 + i] += o.fprior[h * 8 + i] * o.fdelta_prior[h * 8 + i]; }
  }
#undef HH
}

// AccumulatedTopHessianSSE::stitchDoubleMT (AccumulatedTopHessian.h:80-127)
static void top_stitchMT(Oracle &o, std::vector<std::vector<AccumulatorApprox>> &acc, std::vector<int> &nres, double *H, double *b, bool usePrior) {
  const int nf = o.nf, D = CPARS + 8 * nf;
  std::fill(H, H + (size_t)D * D, 0.0); std::fill(b, b + D, 0.0);
  if (o.MT) {
    std::vector<std::vector<double>> Hs(o.T, std::vector<double>((size_t)D * D, 0.0)), bs(o.T, std::vector<double>(D, 0.0));
    o.red->reduce([&](int a, int bb, Stats10 *, int tid) { if (a == bb) return; top_stitchInternal(o, acc, Hs[tid].data(), bs[tid].data(), usePrior, a, bb, o.T, a == 0); }, 0, nf * nf, 0);
    for (int i = 0; i < o.T; i++) {
      for (size_t k = 0; k < (size_t)D * D; k++) H[k] += Hs[i][k];
      for (int k = 0; k < D; k++) b[k] += bs[i][k];
      if (i > 0) nres[0] += nres[i];
    }
  } else {
    top_stitchInternal(o, acc, H, b, usePrior, 0, nf * nf, 1, true);
  }
#define HH(r, c) H[(size_t)(r) * D + (c)]
  for (int h = 0; h < nf; h++) {
    int hIdx = CPARS + h * 8;
    for (int i = 0; i < CPARS; i++) for (int j = 0; j < 8; j++) HH(i, hIdx + j) = HH(hIdx + j, i);
    for (int t = h + 1; t < nf; t++) {
      int tIdx = CPARS + t * 8;
      for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) HH(hIdx + i, tIdx + j) += HH(tIdx + j, hIdx + i);
      for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) HH(tIdx + i, hIdx + j) = HH(hIdx + j, tIdx + i);
    }
  }
#undef HH
}

// a9  AccumulatedSCHessianSSE::stitchDoubleInternal (AccumulatedSCHessian.cpp:80-158)
static void sc_stitchInternal(Oracle &o, double *H, double *b, int min, int max, int toAggregate, bool doCalib) {
  const int nf = o.nf, nframes2 = nf * nf, D = CPARS + 8 * nf;
#define HH(r, c) H[(size_t)(r) * D + (c)]
  for (int k = min; k < max; k++) {
    int i = k % nf, j = k / nf;
    int iIdx = CPARS + i * 8, jIdx = CPARS + j * 8, ijIdx = i + nf * j;
    double Hpc[32], bp[8];
    memset(Hpc, 0, sizeof(Hpc)); memset(bp, 0, sizeof(bp));
    for (int tid2 = 0; tid2 < toAggregate; tid2++) {
      o.accE[tid2][ijIdx].finish(); o.accEB[tid2][ijIdx].finish();
      for (int q = 0; q < 32; q++) Hpc[q] += (double)o.accE[tid2][ijIdx].A1m[q];
      for (int q = 0; q < 8; q++) bp[q] += (double)o.accEB[tid2][ijIdx].A1m[q];
    }
    const double *Ah_ij = &o.adHost[64 * ijIdx], *At_ij = &o.adTarget[64 * ijIdx];
    for (int r = 0; r < 8; r++) {
      for (int c = 0; c < CPARS; c++) {
        double sh = 0, st = 0;
        for (int q = 0; q < 8; q++) { sh += Ah_ij[r * 8 + q] * Hpc[q * 4 + c]; st += At_ij[r * 8 + q] * Hpc[q * 4 + c]; }
        HH(iIdx + r, c) += sh; HH(jIdx + r, c) += st;
      }
      double sh = 0, st = 0;
      for (int q = 0; q < 8; q++) { sh += Ah_ij[r * 8 + q] * bp[q]; st += At_ij[r * 8 + q] * bp[q]; }
      b[iIdx + r] += sh; b[jIdx + r] += st;
    }
    for (int kk = 0; kk < nf; kk++) {
      int kIdx = CPARS + kk * 8, ijkIdx = ijIdx + kk * nframes2, ikIdx = i + nf * kk;
      double accDM[64];
      memset(accDM, 0, sizeof(accDM));
      for (int tid2 = 0; tid2 < toAggregate; tid2++) {
        o.accD[tid2][ijkIdx].finish();
        if (o.accD[tid2][ijkIdx].num == 0) continue;
        for (int q = 0; q < 64; q++) accDM[q] += (double)o.accD[tid2][ijkIdx].A1m[q];
      }
      const double *Ah_ik = &o.adHost[64 * ikIdx], *At_ik = &o.adTarget[64 * ikIdx];
      double tmp[64], out[64];
      mat8_mul(Ah_ij, accDM, tmp); mat8_mulT(tmp, Ah_ik, out);
      for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) HH(iIdx + r, iIdx + c) += out[r * 8 + c];
      mat8_mul(At_ij, accDM, tmp); mat8_mulT(tmp, At_ik, out);
      for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) HH(jIdx + r, kIdx + c) += out[r * 8 + c];
      mat8_mul(At_ij, accDM, tmp); mat8_mulT(tmp, Ah_ik, out);
      for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) HH(jIdx + r, iIdx + c) += out[r * 8 + c];
      mat8_mul(Ah_ij, accDM, tmp); mat8_mulT(tmp, At_ik, out);
      for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) HH(iIdx + r, kIdx + c) += out[r * 8 + c];
    }
  }
  if (doCalib) {
    for (int tid2 = 0; tid2 < toAggregate; tid2++) {
      o.accHcc[tid2].finish(); o.accbc[tid2].finish();
      for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) HH(r, c) += (double)o.accHcc[tid2].A1m[r * 4 + c];
      for (int r = 0; r < 4; r++) b[r] += (double)o.accbc[tid2].A1m[r];
    }
  }
#undef HH
}

// AccumulatedSCHessianSSE::stitchDoubleMT (AccumulatedSCHessian.h:88-124)
static void sc_stitchMT(Oracle &o, double *H, double *b) {
  const int nf = o.nf, D = CPARS + 8 * nf;
  std::fill(H, H + (size_t)D * D, 0.0); std::fill(b, b + D, 0.0);
  if (o.MT) {
    std::vector<std::vector<double>> Hs(o.T, std::vector<double>((size_t)D * D, 0.0)), bs(o.T, std::vector<double>(D, 0.0));
    o.red->reduce([&](int a, int bb, Stats10 *, int tid) { if (a == bb) return; sc_stitchInternal(o, Hs[tid].data(), bs[tid].data(), a, bb, o.T, a == 0); }, 0, nf * nf, 0);
    for (int i = 0; i < o.T; i++) {
      for (size_t k = 0; k < (size_t)D * D; k++) H[k] += Hs[i][k];
      for (int k = 0; k < D; k++) b[k] += bs[i][k];
    }
  } else {
    sc_stitchInternal(o, H, b, 0, nf * nf, 1, true);
  }
  for (int h = 0; h < nf; h++) {
    int hIdx = CPARS + h * 8;
    for (int i = 0; i < CPARS; i++) for (int j = 0; j < 8; j++) H[(size_t)i * D + hIdx + j] = H[(size_t)(hIdx + j) * D + i];
  }
}

static void top_setZero(Oracle &o, std::vector<std::vector<AccumulatorApprox>> &acc, std::vector<int> &nres, int tid) {
  acc[tid].resize((size_t)o.nf * o.nf);
  for (auto &a : acc[tid]) a.initialize();
  nres[tid] = 0;
}
static void sc_setZero(Oracle &o, int tid) {
  const int n = o.nf;
  o.accE[tid].resize((size_t)n * n); o.accEB[tid].resize((size_t)n * n); o.accD[tid].resize((size_t)n * n * n);
  o.accbc[tid].initialize(); o.accHcc[tid].initialize();
  for (auto &a : o.accE[tid]) a.initialize();
  for (auto &a : o.accEB[tid]) a.initialize();
  for (auto &a : o.accD[tid]) a.initialize();
}

// EnergyFunctional::accumulateAF_MT / LF_MT / SCF_MT (EnergyFunctional.cpp:197-254)
void accumulateAF(Oracle &o, double *H, double *b) {
  const int nP = (int)o.pts.size();
  if (o.MT) {
    o.red->reduce([&](int, int, Stats10 *, int tid) { top_setZero(o, o.accA, o.nresA, tid); }, 0, 0, 0);
    o.red->reduce([&](int a, int bb, Stats10 *, int tid) { for (int i = a; i < bb; i++) top_addPoint<0>(o, o.accA, o.nresA, o.pts[i], tid); }, 0, nP, 50);
  } else {
    top_setZero(o, o.accA, o.nresA, 0);
    for (int i = 0; i < nP; i++) top_addPoint<0>(o, o.accA, o.nresA, o.pts[i], 0);
  }
  top_stitchMT(o, o.accA, o.nresA, H, b, false);
  o.resInA = o.nresA[0];
}
void accumulateLF(Oracle &o, double *H, double *b) {
  const int nP = (int)o.pts.size();
  if (o.MT) {
    o.red->reduce([&](int, int, Stats10 *, int tid) { top_setZero(o, o.accL, o.nresL, tid); }, 0, 0, 0);
    o.red->reduce([&](int a, int bb, Stats10 *, int tid) { for (int i = a; i < bb; i++) top_addPoint<1>(o, o.accL, o.nresL, o.pts[i], tid); }, 0, nP, 50);
  } else {
    top_setZero(o, o.accL, o.nresL, 0);
    for (int i = 0; i < nP; i++) top_addPoint<1>(o, o.accL, o.nresL, o.pts[i], 0);
  }
  top_stitchMT(o, o.accL, o.nresL, H, b, true);
  o.resInL = o.nresL[0];
}
void accumulateSCF(Oracle &o, double *H, double *b) {
  const int nP = (int)o.pts.size();
  if (o.MT) {
    o.red->reduce([&](int, int, Stats10 *, int tid) { sc_setZero(o, tid); }, 0, 0, 0);
    o.red->reduce([&](int a, int bb, Stats10 *, int tid) { for (int i = a; i < bb; i++) sc_addPoint(o, o.pts[i], true, tid); }, 0, nP, 50);
  } else {
    sc_setZero(o, 0);
    for (int i = 0; i < nP; i++) sc_addPoint(o, o.pts[i], true, 0);
  }
  sc_stitchMT(o, H, b);
}

// a11  EnergyFunctional::resubstituteF_MT / resubstituteFPt (EnergyFunctional.cpp:496-551)
void resubstituteF(Oracle &o, const double *x) {
  const int nf = o.nf, D = CPARS + 8 * nf;
  std::vector<float> xF(D);
  for (int i = 0; i < D; i++) xF[i] = (float)x[i];
  std::vector<float> xAd((size_t)nf * nf * 8);
  for (int h = 0; h < nf; h++)
    for (int t = 0; t < nf; t++) {
      const float *AhF = &o.adHostF[64 * (h + nf * t)], *AtF = &o.adTargetF[64 * (h + nf * t)];
      for (int c = 0; c < 8; c++) {
        float sh = 0, st = 0;
        for (int k = 0; k < 8; k++) { sh += xF[CPARS + 8 * h + k] * AhF[k * 8 + c]; st += xF[CPARS + 8 * t + k] * AtF[k * 8 + c]; }
        xAd[8 * (nf * h + t) + c] = sh + st;
      }
    }
  const float *xc = xF.data();
  auto fpt = [&](int min, int max) {
    for (int k = min; k < max; k++) {
      Pt &p = o.pts[k];
      int ngoodres = 0;
      for (int ri = p.res_begin; ri < p.res_end; ri++) if (!o.res[ri].dropped && o.res[ri].isActive) ngoodres++;
      if (ngoodres == 0) { p.step = 0; continue; }
      float b = p.bdSumF;
      float dotc = 0;
      for (int i = 0; i < 4; i++) dotc += xc[i] * (p.Hcd_accAF[i] + p.Hcd_accLF[i]);
      b -= dotc;
      for (int ri = p.res_begin; ri < p.res_end; ri++) {
        const Res &r = o.res[ri];
        if (r.dropped || !r.isActive) continue;
        const float *xa = &xAd[8 * (r.host * nf + r.target)];
        float s = 0;
        for (int i = 0; i < 8; i++) s += xa[i] * r.JpJdF[i];
        b -= s;
      }
      p.step = -b * p.HdiF;
    }
  };
  if (o.MT) o.red->reduce([&](int a, int bb, Stats10 *, int) { fpt(a, bb); }, 0, (int)o.pts.size(), 50);
  else fpt(0, (int)o.pts.size());
}

// a10  EnergyFunctional::solveSystemF without IMU (EnergyFunctional.cpp:1029-1184)
void solveSystemF(Oracle &o, const double *HM, const double *bM, double *x_out, double *Hfinal_out, double *bfinal_out) {
  const double lambda = 1e-5;  // :1031
  const int nf = o.nf, D = CPARS + 8 * nf;
  std::vector<double> HL((size_t)D * D), HA((size_t)D * D), Hsc((size_t)D * D), bL(D), bA(D), bsc(D);
  accumulateAF(o, HA.data(), bA.data());
  accumulateLF(o, HL.data(), bL.data());
  accumulateSCF(o, Hsc.data(), bsc.data());
  std::vector<double> HF((size_t)D * D), bF(D);
  for (size_t i = 0; i < (size_t)D * D; i++) HF[i] = HL[i] + HA[i];
  for (int i = 0; i < D; i++) bF[i] = bL[i] + bA[i];
  if (HM && bM) {  // :1070-1091, delta = getStitchedDeltaF()
    std::vector<double> delta(D);
    for (int i = 0; i < CPARS; i++) delta[i] = (double)o.cDeltaF[i];
    for (int h = 0; h < nf; h++) for (int i = 0; i < 8; i++) delta[CPARS + 8 * h + i] = o.fdelta[h * 8 + i];
    for (int r = 0; r < D; r++) {
      double s = 0;
      for (int c = 0; c < D; c++) s += HM[(size_t)r * D + c] * delta[c];
      bF[r] += bM[r] + s;
    }
    for (size_t i = 0; i < (size_t)D * D; i++) HF[i] += HM[i];
  }
  for (int i = 0; i < D; i++) HF[(size_t)i * D + i] *= (1 + lambda);
  const double sc = (double)(1.0f / (1 + lambda));  // float-typed scalar in the reference (:1099)
  for (size_t i = 0; i < (size_t)D * D; i++) HF[i] -= Hsc[i] * sc;
  for (int i = 0; i < D; i++) bF[i] -= bsc[i];
  if (Hfinal_out) memcpy(Hfinal_out, HF.data(), sizeof(double) * D * D);
  if (bfinal_out) memcpy(bfinal_out, bF.data(), sizeof(double) * D);
  std::vector<double> S(D), Hs((size_t)D * D), bs(D), y(D);
  for (int i = 0; i < D; i++) S[i] = 1.0 / std::sqrt(HF[(size_t)i * D + i] + 10.0);
  for (int r = 0; r < D; r++) { for (int c = 0; c < D; c++) Hs[(size_t)r * D + c] = S[r] * HF[(size_t)r * D + c] * S[c]; bs[r] = S[r] * bF[r]; }
  ldlt_solve(Hs.data(), bs.data(), y.data(), D);
  o.lastX.resize(D);
  for (int i = 0; i < D; i++) o.lastX[i] = S[i] * y[i];
  if (x_out) memcpy(x_out, o.lastX.data(), sizeof(double) * D);
  resubstituteF(o, o.lastX.data());
}

// EnergyFunctional::marginalizePointsF (EnergyFunctional.cpp:891-936), serial stitchDouble path.
void marginalizePoints(Oracle &o, const int32_t *ids, int n, double *H, double *b, int *resInM) {
  const int nf = o.nf, D = CPARS + 8 * nf;
  for (int i = 0; i < n; i++) o.pts[ids[i]].priorF *= o.cfg.idepth_fix_prior_marg_fac;
  sc_setZero(o, 0);
  top_setZero(o, o.accA, o.nresA, 0);
  for (int i = 0; i < n; i++) {
    Pt &p = o.pts[ids[i]];
    top_addPoint<2>(o, o.accA, o.nresA, p, 0);
    sc_addPoint(o, p, false, 0);
  }
  std::vector<double> M((size_t)D * D, 0.0), Mb(D, 0.0), Msc((size_t)D * D, 0.0), Mbsc(D, 0.0);
  // stitchDouble (AccumulatedTopHessian.cpp:155-229) == the single-thread internal path + symmetrise
  bool mt = o.MT; o.MT = false;
  top_stitchMT(o, o.accA, o.nresA, M.data(), Mb.data(), false);
  // stitchDouble (AccumulatedSCHessian.cpp:160-223): same sums; the calib block is assigned, H starts at 0
  sc_stitchMT(o, Msc.data(), Mbsc.data());
  o.MT = mt;
  if (resInM) *resInM = o.nresA[0];
  o.resInM += o.nresA[0];
  for (size_t i = 0; i < (size_t)D * D; i++) H[i] = M[i] - Msc[i];
  for (int i = 0; i < D; i++) b[i] = Mb[i] - Mbsc[i];
}

}  // namespace orc
