// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).  PARITY UNPINNED.
//
// orc_tracker.cpp: CPU restatement of the direct-alignment kernels
//   ScaleOptimizer::makeK                 src/FullSystem/ScaleOptimizer.cpp:95-118
//   CoarseTracker::calcResPose            src/FullSystem/CoarseTracker.cpp:612-764
//   CoarseTracker::calcGSSSEPose          src/FullSystem/CoarseTracker.cpp:554-610
//   PoseEstimator::calcRes / calcGSSSE    src/LoopClosure/PoseEstimator.cpp:147-284, 75-145
//   CoarseInitializer::calcResAndGS       src/FullSystem/CoarseInitializer.cpp:450-673
//   ScaleOptimizer::calcResScale          src/FullSystem/ScaleOptimizer.cpp:273-437
//   ScaleOptimizer::calcGSSSEScale        src/FullSystem/ScaleOptimizer.cpp:232-271
//   Accumulator9::updateSSE_eighted       src/OptimizationBackend/MatrixAccumulators.h:1314-1432
//   ScaleAccumulator::updateSSE_oneed     src/OptimizationBackend/ScaleAccumulator.h:60-77
#include <cmath>

#include "orc_core.h"
#include "orc_sample.h"
#include "orc_host.h"

namespace orc {

void tracker_makeK(Oracle &o, const float calib[4]) {
  o.tfx[0] = calib[0]; o.tfy[0] = calib[1]; o.tcx[0] = calib[2]; o.tcy[0] = calib[3];
  for (int level = 1; level < o.levels; ++level) {
    o.tfx[level] = o.tfx[level - 1] * 0.5;
    o.tfy[level] = o.tfy[level - 1] * 0.5;
    o.tcx[level] = (o.tcx[0] + 0.5) / ((int)1 << level) - 0.5;
    o.tcy[level] = (o.tcy[0] + 0.5) / ((int)1 << level) - 0.5;
  }
  for (int level = 0; level < o.levels; ++level) {
    M3<float> Ki;
    for (int i = 0; i < 9; i++) Ki.m[i] = 0;
    Ki(0, 0) = 1.0f / o.tfx[level]; Ki(1, 1) = 1.0f / o.tfy[level];
    Ki(0, 2) = -o.tcx[level] / o.tfx[level]; Ki(1, 2) = -o.tcy[level] / o.tfy[level]; Ki(2, 2) = 1;
    o.tKi[level] = Ki;
  }
}

static void ensure_warp_buffers(Oracle &o) {
  size_t n = (size_t)o.wl[0] * o.hl[0] + 4;
  if (o.bw_u.size() < n) {
    o.bw_idepth.resize(n); o.bw_u.resize(n); o.bw_v.resize(n); o.bw_dx.resize(n); o.bw_dy.resize(n);
    o.bw_residual.resize(n); o.bw_weight.resize(n); o.bw_refColor.resize(n);
    o.sw_rx1.resize(n); o.sw_rx2.resize(n); o.sw_rx3.resize(n); o.sw_dx.resize(n); o.sw_dy.resize(n);
    o.sw_residual.resize(n); o.sw_weight.resize(n); o.sw_ref.resize(n);
  }
}

void tracker_calcResPose(Oracle &o, int lvl, int slot, const double refToNew[12], const float affLL[2], float cutoffTH, double out6[6],
                         int32_t counts[3]) {
  ensure_warp_buffers(o);
  float E = 0;
  int numTermsInE = 0, numTermsInWarped = 0, numSaturated = 0;
  const int wl = o.wl[lvl], hl = o.hl[lvl];
  const float *dINewl = o.slots[slot].lvl[lvl].dI.data();
  const float fxl = o.tfx[lvl], fyl = o.tfy[lvl], cxl = o.tcx[lvl], cyl = o.tcy[lvl];
  M3<float> Rf; V3<float> t;
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) Rf(i, j) = (float)refToNew[i * 4 + j]; t[i] = (float)refToNew[i * 4 + 3]; }
  const M3<float> RKi = mul(Rf, o.tKi[lvl]);
  const M3<float> &Ki = o.tKi[lvl];
  float sumSquaredShiftT = 0, sumSquaredShiftRT = 0, sumSquaredShiftNum = 0;
  const float huberTH = o.cfg.huber_th;
  const float maxEnergy = 2 * huberTH * cutoffTH - huberTH * huberTH;
  const int nl = (int)o.pc_u[lvl].size();
  const float *lpc_u = o.pc_u[lvl].data(), *lpc_v = o.pc_v[lvl].data(), *lpc_idepth = o.pc_idepth[lvl].data(), *lpc_color = o.pc_color[lvl].data();
  for (int i = 0; i < nl; i++) {
    float id = lpc_idepth[i], x = lpc_u[i], y = lpc_v[i];
    V3<float> xy1{{x, y, 1}};
    V3<float> pt = mul(RKi, xy1);
    for (int k = 0; k < 3; k++) pt[k] = pt[k] + t[k] * id;
    float u = pt[0] / pt[2], v = pt[1] / pt[2];
    float Ku = fxl * u + cxl, Kv = fyl * v + cyl;
    float new_idepth = id / pt[2];
    if (lvl == 0 && i % 32 == 0) {
      V3<float> kp = mul(Ki, xy1);
      V3<float> ptT{{kp[0] + t[0] * id, kp[1] + t[1] * id, kp[2] + t[2] * id}};
      float uT = ptT[0] / ptT[2], vT = ptT[1] / ptT[2];
      float KuT = fxl * uT + cxl, KvT = fyl * vT + cyl;
      V3<float> ptT2{{kp[0] - t[0] * id, kp[1] - t[1] * id, kp[2] - t[2] * id}};
      float uT2 = ptT2[0] / ptT2[2], vT2 = ptT2[1] / ptT2[2];
      float KuT2 = fxl * uT2 + cxl, KvT2 = fyl * vT2 + cyl;
      V3<float> rp = mul(RKi, xy1);
      V3<float> pt3{{rp[0] - t[0] * id, rp[1] - t[1] * id, rp[2] - t[2] * id}};
      float u3 = pt3[0] / pt3[2], v3 = pt3[1] / pt3[2];
      float Ku3 = fxl * u3 + cxl, Kv3 = fyl * v3 + cyl;
      sumSquaredShiftT += (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y);
      sumSquaredShiftT += (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
      sumSquaredShiftRT += (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y);
      sumSquaredShiftRT += (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
      sumSquaredShiftNum += 2;
    }
    if (!(Ku > 2 && Kv > 2 && Ku < wl - 3 && Kv < hl - 3 && new_idepth > 0)) continue;
    float refColor = lpc_color[i];
    float hitColor[3];
    interp33(dINewl, Ku, Kv, wl, hitColor);
    if (!std::isfinite(hitColor[0])) continue;
    float residual = hitColor[0] - (float)(affLL[0] * refColor + affLL[1]);
    float hw = fabsf(residual) < huberTH ? 1 : huberTH / fabsf(residual);
    if (fabsf(residual) > cutoffTH) {
      E += maxEnergy; numTermsInE++; numSaturated++;
    } else {
      E += hw * residual * residual * (2 - hw);
      numTermsInE++;
      o.bw_idepth[numTermsInWarped] = new_idepth; o.bw_u[numTermsInWarped] = u; o.bw_v[numTermsInWarped] = v;
      o.bw_dx[numTermsInWarped] = hitColor[1]; o.bw_dy[numTermsInWarped] = hitColor[2];
      o.bw_residual[numTermsInWarped] = residual; o.bw_weight[numTermsInWarped] = hw; o.bw_refColor[numTermsInWarped] = lpc_color[i];
      numTermsInWarped++;
    }
  }
  counts[0] = numTermsInE; counts[1] = numTermsInWarped; counts[2] = numSaturated;
  while (numTermsInWarped % 4 != 0) {
    o.bw_idepth[numTermsInWarped] = 0; o.bw_u[numTermsInWarped] = 0; o.bw_v[numTermsInWarped] = 0; o.bw_dx[numTermsInWarped] = 0;
    o.bw_dy[numTermsInWarped] = 0; o.bw_residual[numTermsInWarped] = 0; o.bw_weight[numTermsInWarped] = 0; o.bw_refColor[numTermsInWarped] = 0;
    numTermsInWarped++;
  }
  o.bw_n = numTermsInWarped;
  out6[0] = E; out6[1] = numTermsInE; out6[2] = sumSquaredShiftT / (sumSquaredShiftNum + 0.1); out6[3] = 0;
  out6[4] = sumSquaredShiftRT / (sumSquaredShiftNum + 0.1); out6[5] = numSaturated / (float)numTermsInE;
}

// PoseEstimator::calcRes (src/LoopClosure/PoseEstimator.cpp:147-284); fills the same warped buffers as calcResPose, so
// PoseEstimator::calcGSSSE (:75-145, identical arithmetic to calcGSSSEPose) is tracker_calcGSSSEPose.
void loop_calcRes(Oracle &o, int lvl, int slot, const double refToNew[12], const float affLL[2], float cutoffTH, double out6[6], int32_t counts[3]) {
  ensure_warp_buffers(o);
  float E = 0;
  int numTermsInE = 0, numTermsInWarped = 0, numSaturated = 0;
  const int wl = o.wl[lvl], hl = o.hl[lvl];
  const float *dINewl = o.slots[slot].lvl[lvl].dI.data();
  const float fxl = o.tfx[lvl], fyl = o.tfy[lvl], cxl = o.tcx[lvl], cyl = o.tcy[lvl];
  M3<float> R; V3<float> t;
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R(i, j) = (float)refToNew[i * 4 + j]; t[i] = (float)refToNew[i * 4 + 3]; }
  float sumSquaredShiftT = 0, sumSquaredShiftRT = 0, sumSquaredShiftNum = 0;
  const float huberTH = o.cfg.huber_th;
  const float maxEnergy = 2 * huberTH * cutoffTH - huberTH * huberTH;
  const size_t n = o.loop_xyz.size() / 3;
  for (size_t i = 0; i < n; i++) {
    float x = o.loop_xyz[3 * i], y = o.loop_xyz[3 * i + 1], z = o.loop_xyz[3 * i + 2];
    float u0 = x / z, v0 = y / z;
    float Ku0 = fxl * u0 + cxl, Kv0 = fyl * v0 + cyl;
    V3<float> p3{{x, y, z}};
    V3<float> pt = mul(R, p3);
    for (int k = 0; k < 3; k++) pt[k] = pt[k] + t[k];
    float u = pt[0] / pt[2], v = pt[1] / pt[2];
    float Ku = fxl * u + cxl, Kv = fyl * v + cyl;
    float new_idepth = 1 / pt[2];
    if (lvl == 0 && i % 32 == 0) {
      V3<float> xy1{{x, y, 1}};   // sic: the un-normalised x, y with z = 1 (PoseEstimator.cpp:208-223)
      V3<float> ptT{{x + t[0], y + t[1], 1 + t[2]}};
      float KuT = fxl * (ptT[0] / ptT[2]) + cxl, KvT = fyl * (ptT[1] / ptT[2]) + cyl;
      V3<float> ptT2{{x - t[0], y - t[1], 1 - t[2]}};
      float KuT2 = fxl * (ptT2[0] / ptT2[2]) + cxl, KvT2 = fyl * (ptT2[1] / ptT2[2]) + cyl;
      V3<float> rp = mul(R, xy1);
      V3<float> pt3{{rp[0] - t[0], rp[1] - t[1], rp[2] - t[2]}};
      float Ku3 = fxl * (pt3[0] / pt3[2]) + cxl, Kv3 = fyl * (pt3[1] / pt3[2]) + cyl;
      sumSquaredShiftT += (KuT - Ku0) * (KuT - Ku0) + (KvT - Kv0) * (KvT - Kv0);
      sumSquaredShiftT += (KuT2 - Ku0) * (KuT2 - Ku0) + (KvT2 - Kv0) * (KvT2 - Kv0);
      sumSquaredShiftRT += (Ku - Ku0) * (Ku - Ku0) + (Kv - Kv0) * (Kv - Kv0);
      sumSquaredShiftRT += (Ku3 - Ku0) * (Ku3 - Ku0) + (Kv3 - Kv0) * (Kv3 - Kv0);
      sumSquaredShiftNum += 2;
    }
    if (!(Ku > 2 && Kv > 2 && Ku < wl - 3 && Kv < hl - 3 && new_idepth > 0)) continue;
    float refColor = o.loop_color[i * o.levels + lvl];
    float hitColor[3];
    interp33(dINewl, Ku, Kv, wl, hitColor);
    if (!std::isfinite(hitColor[0])) continue;
    float residual = hitColor[0] - (float)(affLL[0] * refColor + affLL[1]);
    float hw = fabs(residual) < huberTH ? 1 : huberTH / fabs(residual);
    if (fabs(residual) > cutoffTH) {
      E += maxEnergy; numTermsInE++; numSaturated++;
    } else {
      E += hw * residual * residual * (2 - hw);
      numTermsInE++;
      o.bw_idepth[numTermsInWarped] = new_idepth; o.bw_u[numTermsInWarped] = u; o.bw_v[numTermsInWarped] = v;
      o.bw_dx[numTermsInWarped] = hitColor[1]; o.bw_dy[numTermsInWarped] = hitColor[2];
      o.bw_residual[numTermsInWarped] = residual; o.bw_weight[numTermsInWarped] = hw; o.bw_refColor[numTermsInWarped] = refColor;
      numTermsInWarped++;
    }
  }
  counts[0] = numTermsInE; counts[1] = numTermsInWarped; counts[2] = numSaturated;
  while (numTermsInWarped % 4 != 0) {
    o.bw_idepth[numTermsInWarped] = 0; o.bw_u[numTermsInWarped] = 0; o.bw_v[numTermsInWarped] = 0; o.bw_dx[numTermsInWarped] = 0;
    o.bw_dy[numTermsInWarped] = 0; o.bw_residual[numTermsInWarped] = 0; o.bw_weight[numTermsInWarped] = 0; o.bw_refColor[numTermsInWarped] = 0;
    numTermsInWarped++;
  }
  o.bw_n = numTermsInWarped;
  out6[0] = E; out6[1] = numTermsInE; out6[2] = sumSquaredShiftT / (sumSquaredShiftNum + 0.1); out6[3] = 0;
  out6[4] = sumSquaredShiftRT / (sumSquaredShiftNum + 0.1); out6[5] = numSaturated / (float)numTermsInE;
}

// CoarseInitializer::calcResAndGS (src/FullSystem/CoarseInitializer.cpp:450-673)
void init_calcResAndGS(Oracle &o, int lvl, int ref_slot, int new_slot, const double refToNew[12], const float aff[2], const float tlog[3], float alphaW,
                       float alphaK, float couplingWeight, sosba_init_points *P, float H_out[64], float b_out[8], float H_out_sc[64], float b_out_sc[8],
                       float res3[3]) {
  const int wl = o.wl[lvl], hl = o.hl[lvl];
  const float *colorRef = o.slots[ref_slot].lvl[lvl].dI.data(), *colorNew = o.slots[new_slot].lvl[lvl].dI.data();
  M3<float> Rf; V3<float> t;
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) Rf(i, j) = (float)refToNew[i * 4 + j]; t[i] = (float)refToNew[i * 4 + 3]; }
  const M3<float> RKi = mul(Rf, o.tKi[lvl]);
  const float r2new_aff[2] = {(float)exp((double)aff[0]), aff[1]};   // Eigen::Vector2f(exp(refToNew_aff.a), refToNew_aff.b): a is a double in AffLight
  const float fxl = o.tfx[lvl], fyl = o.tfy[lvl], cxl = o.tcx[lvl], cyl = o.tcy[lvl];
  const float huberTH = o.cfg.huber_th;
  Accumulator11 E;
  Accumulator9 acc9, acc9SC;
  acc9.initialize();
  E.initialize();
  const int npts = P->n;
  for (int i = 0; i < npts; i++) {
    P->maxstep[i] = 1e10;
    float *Jb = P->JbBuffer_new + 10 * (size_t)i;
    if (!P->isGood[i]) {
      E.updateSingle(P->energy[2 * i]);
      P->energy_new[2 * i] = P->energy[2 * i]; P->energy_new[2 * i + 1] = P->energy[2 * i + 1];
      P->isGood_new[i] = 0;
      continue;
    }
    float dp[8][8], dd[8], r[8];
    for (int k = 0; k < 10; k++) Jb[k] = 0;
    bool isGood = true;
    float energy = 0;
    const float pu = P->u[i], pv = P->v[i], idn = P->idepth_new[i];
    for (int idx = 0; idx < 8; idx++) {
      const int dx = patternP[idx][0], dy = patternP[idx][1];
      V3<float> xy1{{pu + dx, pv + dy, 1}};
      V3<float> pt = mul(RKi, xy1);
      for (int k = 0; k < 3; k++) pt[k] = pt[k] + t[k] * idn;
      float u = pt[0] / pt[2], v = pt[1] / pt[2];
      float Ku = fxl * u + cxl, Kv = fyl * v + cyl;
      float new_idepth = idn / pt[2];
      if (!(Ku > 1 && Kv > 1 && Ku < wl - 2 && Kv < hl - 2 && new_idepth > 0)) { isGood = false; break; }
      float hitColor[3];
      interp33(colorNew, Ku, Kv, wl, hitColor);
      float rl3[3];
      interp33(colorRef, pu + dx, pv + dy, wl, rl3);   // getInterpolatedElement31: channel 0 of the same bilinear formula
      float rlR = rl3[0];
      if (!std::isfinite(rlR) || !std::isfinite(hitColor[0])) { isGood = false; break; }
      float residual = hitColor[0] - r2new_aff[0] * rlR - r2new_aff[1];
      float hw = fabs(residual) < huberTH ? 1 : huberTH / fabs(residual);
      energy += hw * residual * residual * (2 - hw);
      float dxdd = (t[0] - t[2] * u) / pt[2];
      float dydd = (t[1] - t[2] * v) / pt[2];
      if (hw < 1) hw = sqrtf(hw);
      float dxInterp = hw * hitColor[1] * fxl;
      float dyInterp = hw * hitColor[2] * fyl;
      dp[0][idx] = new_idepth * dxInterp;
      dp[1][idx] = new_idepth * dyInterp;
      dp[2][idx] = -new_idepth * (u * dxInterp + v * dyInterp);
      dp[3][idx] = -u * v * dxInterp - (1 + v * v) * dyInterp;
      dp[4][idx] = (1 + u * u) * dxInterp + u * v * dyInterp;
      dp[5][idx] = -v * dxInterp + u * dyInterp;
      dp[6][idx] = -hw * r2new_aff[0] * rlR;
      dp[7][idx] = -hw * 1;
      dd[idx] = dxInterp * dxdd + dyInterp * dydd;
      r[idx] = hw * residual;
      float mx = dxdd * fxl, my = dydd * fyl;
      float maxstep = 1.0f / sqrtf(mx * mx + my * my);
      if (maxstep < P->maxstep[i]) P->maxstep[i] = maxstep;
      for (int k = 0; k < 8; k++) Jb[k] += dp[k][idx] * dd[idx];
      Jb[8] += r[idx] * dd[idx];
      Jb[9] += dd[idx] * dd[idx];
    }
    if (!isGood || energy > P->outlierTH[i] * 20) {
      E.updateSingle(P->energy[2 * i]);
      P->isGood_new[i] = 0;
      P->energy_new[2 * i] = P->energy[2 * i]; P->energy_new[2 * i + 1] = P->energy[2 * i + 1];
      continue;
    }
    E.updateSingle(energy);
    P->isGood_new[i] = 1;
    P->energy_new[2 * i] = energy;
    for (int g = 0; g + 3 < 8; g += 4) {
      float J[9][4];
      for (int l = 0; l < 4; l++) { for (int k = 0; k < 8; k++) J[k][l] = dp[k][g + l]; J[8][l] = r[g + l]; }
      acc9.updateSSE(J);
    }
  }
  E.finish();
  acc9.finish();
  // alpha energy: the reference adds these terms to E (after E.finish()) and leaves EAlpha at zero (:606-617)
  for (int i = 0; i < npts; i++) {
    if (!P->isGood_new[i]) E.updateSingle(P->energy[2 * i + 1]);
    else {
      P->energy_new[2 * i + 1] = (P->idepth_new[i] - 1) * (P->idepth_new[i] - 1);
      E.updateSingle(P->energy_new[2 * i + 1]);
    }
  }
  const float EAlphaA = 0;
  const double tsq = refToNew[3] * refToNew[3] + refToNew[7] * refToNew[7] + refToNew[11] * refToNew[11];
  float alphaEnergy = alphaW * (EAlphaA + tsq * npts);
  float alphaOpt;
  if (alphaEnergy > alphaK * npts) { alphaOpt = 0; alphaEnergy = alphaK * npts; }
  else alphaOpt = alphaW;
  acc9SC.initialize();
  for (int i = 0; i < npts; i++) {
    if (!P->isGood_new[i]) continue;
    float *Jb = P->JbBuffer_new + 10 * (size_t)i;
    P->lastHessian_new[i] = Jb[9];
    Jb[8] += alphaOpt * (P->idepth_new[i] - 1);
    Jb[9] += alphaOpt;
    if (alphaOpt == 0) {
      Jb[8] += couplingWeight * (P->idepth_new[i] - P->iR[i]);
      Jb[9] += couplingWeight;
    }
    Jb[9] = 1 / (1 + Jb[9]);
    acc9SC.updateSingleWeighted(Jb, Jb[9]);
  }
  acc9SC.finish();
  for (int rr = 0; rr < 8; rr++) {
    for (int c = 0; c < 8; c++) { H_out[8 * rr + c] = acc9.H[rr][c]; H_out_sc[8 * rr + c] = acc9SC.H[rr][c]; }
    b_out[rr] = acc9.H[rr][8]; b_out_sc[rr] = acc9SC.H[rr][8];
  }
  H_out[0] += alphaOpt * npts; H_out[9] += alphaOpt * npts; H_out[18] += alphaOpt * npts;
  b_out[0] += tlog[0] * alphaOpt * npts; b_out[1] += tlog[1] * alphaOpt * npts; b_out[2] += tlog[2] * alphaOpt * npts;
  res3[0] = E.A; res3[1] = alphaEnergy; res3[2] = (float)E.num;
}

void tracker_calcGSSSEPose(Oracle &o, int lvl, float a, float b0, double H_out[64], double b_out[8]) {
  Accumulator9 acc;
  acc.initialize();
  const float fxl = o.tfx[lvl], fyl = o.tfy[lvl];
  const int n = o.bw_n;
  for (int i = 0; i < n; i += 4) {
    float J[9][4], w[4];
    for (int l = 0; l < 4; l++) {
      float dx = o.bw_dx[i + l] * fxl, dy = o.bw_dy[i + l] * fyl;
      float u = o.bw_u[i + l], v = o.bw_v[i + l], id = o.bw_idepth[i + l];
      J[0][l] = id * dx;
      J[1][l] = id * dy;
      J[2][l] = 0 - id * (u * dx + v * dy);
      J[3][l] = 0 - ((u * v) * dx + dy * (1 + v * v));
      J[4][l] = (u * v) * dy + dx * (1 + u * u);
      J[5][l] = u * dy - v * dx;
      J[6][l] = a * (b0 - o.bw_refColor[i + l]);
      J[7][l] = -1;
      J[8][l] = o.bw_residual[i + l];
      w[l] = o.bw_weight[i + l];
    }
    acc.updateSSE_eighted(J, w);
  }
  acc.finish();
  const float invn = 1.0f / n;
  for (int r = 0; r < 8; r++) { for (int c = 0; c < 8; c++) H_out[r * 8 + c] = (double)acc.H[r][c] * invn; b_out[r] = (double)acc.H[r][8] * invn; }
  const float sc[8] = {SCALE_XI_ROT, SCALE_XI_ROT, SCALE_XI_ROT, SCALE_XI_TRANS, SCALE_XI_TRANS, SCALE_XI_TRANS, SCALE_A, SCALE_B};
  for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) H_out[r * 8 + c] *= sc[c];
  for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) H_out[r * 8 + c] *= sc[r];
  for (int r = 0; r < 8; r++) b_out[r] *= sc[r];
}

void scale_calcRes(Oracle &o, int lvl, int slot, float scale, float cutoffTH, double out6[6], int32_t counts[3]) {
  ensure_warp_buffers(o);
  float E = 0;
  int numTermsInE = 0, numTermsInWarped = 0, numSaturated = 0;
  const int wl = o.wl[lvl], hl = o.hl[lvl];
  const float *dINewl = o.slots[slot].lvl[lvl].dI.data();
  const float fx1l = o.fx1[lvl], fy1l = o.fy1[lvl], cx1l = o.cx1[lvl], cy1l = o.cy1[lvl];
  M3<double> Rd = o.tfmF0ToF1.R();
  M3<float> Rf; V3<float> tsl;
  for (int i = 0; i < 9; i++) Rf.m[i] = (float)Rd.m[i];
  for (int i = 0; i < 3; i++) tsl[i] = (float)o.tfmF0ToF1.t[i];
  const M3<float> RKi = mul(Rf, o.tKi[lvl]);
  const M3<float> &Ki = o.tKi[lvl];
  M3<float> sRKi, sKi;
  for (int i = 0; i < 9; i++) { sRKi.m[i] = scale * RKi.m[i]; sKi.m[i] = scale * Ki.m[i]; }
  float sumSquaredShiftT = 0, sumSquaredShiftRT = 0, sumSquaredShiftNum = 0;
  const float huberTH = o.cfg.huber_th;
  const float maxEnergy = 2 * huberTH * cutoffTH - huberTH * huberTH;
  const int nl = (int)o.pc_u[lvl].size();
  const float *lpc_u = o.pc_u[lvl].data(), *lpc_v = o.pc_v[lvl].data(), *lpc_idepth = o.pc_idepth[lvl].data(), *lpc_color = o.pc_color[lvl].data();
  for (int i = 0; i < nl; i++) {
    float id = lpc_idepth[i], x = lpc_u[i], y = lpc_v[i];
    V3<float> xy1{{x, y, 1}};
    V3<float> pt = mul(sRKi, xy1);
    for (int k = 0; k < 3; k++) pt[k] = pt[k] + tsl[k] * id;
    float u = pt[0] / pt[2], v = pt[1] / pt[2];
    float Ku = fx1l * u + cx1l, Kv = fy1l * v + cy1l;
    float new_idepth = id / pt[2];
    V3<float> rx = mul(RKi, xy1);
    for (int k = 0; k < 3; k++) rx[k] = rx[k] / id;
    if (lvl == 0 && i % 32 == 0) {
      V3<float> kp = mul(sKi, xy1);
      V3<float> ptT{{kp[0] + tsl[0] * id, kp[1] + tsl[1] * id, kp[2] + tsl[2] * id}};
      float uT = ptT[0] / ptT[2], vT = ptT[1] / ptT[2];
      float KuT = fx1l * uT + cx1l, KvT = fy1l * vT + cy1l;
      V3<float> ptT2{{kp[0] - tsl[0] * id, kp[1] - tsl[1] * id, kp[2] - tsl[2] * id}};
      float uT2 = ptT2[0] / ptT2[2], vT2 = ptT2[1] / ptT2[2];
      float KuT2 = fx1l * uT2 + cx1l, KvT2 = fy1l * vT2 + cy1l;
      V3<float> rp = mul(sRKi, xy1);
      V3<float> pt3{{rp[0] - tsl[0] * id, rp[1] - tsl[1] * id, rp[2] - tsl[2] * id}};
      float u3 = pt3[0] / pt3[2], v3 = pt3[1] / pt3[2];
      float Ku3 = fx1l * u3 + cx1l, Kv3 = fy1l * v3 + cy1l;
      sumSquaredShiftT += (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y);
      sumSquaredShiftT += (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
      sumSquaredShiftRT += (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y);
      sumSquaredShiftRT += (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
      sumSquaredShiftNum += 2;
    }
    if (!(Ku > 2 && Kv > 2 && Ku < wl - 3 && Kv < hl - 3 && new_idepth > 0)) continue;
    float refColor = lpc_color[i];
    float hitColor[3];
    interp33(dINewl, Ku, Kv, wl, hitColor);
    if (!std::isfinite(hitColor[0])) continue;
    float residual = hitColor[0] - refColor;
    float hw = fabsf(residual) < huberTH ? 1 : huberTH / fabsf(residual);
    if (fabsf(residual) > cutoffTH) {
      E += maxEnergy; numTermsInE++; numSaturated++;
    } else {
      E += hw * residual * residual * (2 - hw);
      numTermsInE++;
      o.sw_rx1[numTermsInWarped] = rx[0]; o.sw_rx2[numTermsInWarped] = rx[1]; o.sw_rx3[numTermsInWarped] = rx[2];
      o.sw_dx[numTermsInWarped] = hitColor[1]; o.sw_dy[numTermsInWarped] = hitColor[2];
      o.sw_residual[numTermsInWarped] = residual; o.sw_weight[numTermsInWarped] = hw; o.sw_ref[numTermsInWarped] = lpc_color[i];
      numTermsInWarped++;
    }
  }
  counts[0] = numTermsInE; counts[1] = numTermsInWarped; counts[2] = numSaturated;
  while (numTermsInWarped % 4 != 0) {
    o.sw_rx1[numTermsInWarped] = 0; o.sw_rx2[numTermsInWarped] = 0; o.sw_rx3[numTermsInWarped] = 0; o.sw_dx[numTermsInWarped] = 0;
    o.sw_dy[numTermsInWarped] = 0; o.sw_residual[numTermsInWarped] = 0; o.sw_weight[numTermsInWarped] = 0; o.sw_ref[numTermsInWarped] = 0;
    numTermsInWarped++;
  }
  o.sw_n = numTermsInWarped;
  out6[0] = E; out6[1] = numTermsInE; out6[2] = sumSquaredShiftT / (sumSquaredShiftNum + 0.1); out6[3] = 0;
  out6[4] = sumSquaredShiftRT / (sumSquaredShiftNum + 0.1); out6[5] = numSaturated / (float)numTermsInE;
}

void scale_calcGSSSE(Oracle &o, int lvl, float scale, float *H_out, float *b_out) {
  ScaleAccumulator acc;
  acc.initialize();
  const float fx1l = o.fx1[lvl], fy1l = o.fy1[lvl];
  const float s = scale, tx = (float)o.tfmF0ToF1.t[0], ty = (float)o.tfmF0ToF1.t[1], tz = (float)o.tfmF0ToF1.t[2];
  const int n = o.sw_n;
  for (int i = 0; i < n; i += 4) {
    float J0[4], J1[4], w[4];
    for (int l = 0; l < 4; l++) {
      float dxfx = o.sw_dx[i + l] * fx1l, dyfy = o.sw_dy[i + l] * fy1l;
      float rx1 = o.sw_rx1[i + l], rx2 = o.sw_rx2[i + l], rx3 = o.sw_rx3[i + l];
      float deno_sqrt = s * rx3 + tz;
      float deno = 1.0f / (deno_sqrt * deno_sqrt);
      float xno = rx1 * tz - rx3 * tx;
      float yno = rx2 * tz - rx3 * ty;
      J0[l] = dxfx * (deno * xno) + dyfy * (deno * yno);
      J1[l] = o.sw_residual[i + l];
      w[l] = o.sw_weight[i + l];
    }
    acc.updateSSE_oneed(J0, J1, w);
  }
  acc.finish();
  *H_out = acc.hessian[0][0] * (1.0f / n);
  *b_out = acc.hessian[0][1] * (1.0f / n);
}

}  // namespace orc
