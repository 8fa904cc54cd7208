// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).  PARITY UNPINNED.
//
// orc_host.cpp: restatement of the host-side glue the GN loop needs:
//   FrameHessian::setState / setEvalPT / getPrior      src/FullSystem/HessianBlocks.h:217-309
//   CalibHessian::setValue                              src/FullSystem/HessianBlocks.h:487-501
//   FrameFramePrecalc::set                              src/FullSystem/HessianBlocks.cpp:431-461
//   AffLight::fromToVecExposure                         src/util/NumType.h:157-168
//   EnergyFunctional::setAdjointsF / setDeltaF          src/OptimizationBackend/EnergyFunctional.cpp:42-103, 163-194
//   FullSystem::optimize / doStepFromBackup / backupState   src/FullSystem/FullSystemOptimize.cpp:185-271, 305-489
#include <cmath>
#include <cstdio>

#include "orc_core.h"
#include "orc_host.h"

namespace orc {

void FrameH::setState(const double *s) {  // HessianBlocks.h:217-230
  for (int i = 0; i < 10; i++) state[i] = s[i];
  for (int i = 0; i < 3; i++) state_scaled[i] = SCALE_XI_TRANS * state[i];
  for (int i = 3; i < 6; i++) state_scaled[i] = SCALE_XI_ROT * state[i];
  state_scaled[6] = SCALE_A * state[6]; state_scaled[7] = SCALE_B * state[7];
  state_scaled[8] = SCALE_A * state[8]; state_scaled[9] = SCALE_B * state[9];
  PRE_camToWorld = se3_exp(state_scaled) * evalPT;
  PRE_worldToCam = PRE_camToWorld.inverse();
}
void FrameH::setEvalPT(const SE3 &e, const double *s) {  // :245-251
  evalPT = e;
  setState(s);
  for (int i = 0; i < 10; i++) state_zero[i] = state[i];
}
void FrameH::getPrior(const sosba_config &cfg, double p[8]) const {  // :288-309
  for (int i = 0; i < 8; i++) p[i] = 0;
  if (frameID == 0) {
    p[0] = p[1] = p[2] = cfg.initial_trans_prior;
    p[3] = p[4] = p[5] = cfg.initial_rot_prior;
    p[6] = cfg.initial_aff_a_prior; p[7] = cfg.initial_aff_b_prior;
  } else {
    p[6] = cfg.affine_opt_mode_a < 0 ? cfg.initial_aff_a_prior : cfg.affine_opt_mode_a;
    p[7] = cfg.affine_opt_mode_b < 0 ? cfg.initial_aff_b_prior : cfg.affine_opt_mode_b;
  }
}

void CalibH::setValue(const double *v) {  // HessianBlocks.h:487-501
  for (int i = 0; i < 4; i++) value[i] = v[i];
  value_scaled[0] = SCALE_F * value[0]; value_scaled[1] = SCALE_F * value[1];
  value_scaled[2] = SCALE_C * value[2]; value_scaled[3] = SCALE_C * value[3];
  for (int i = 0; i < 4; i++) value_scaledf[i] = (float)value_scaled[i];
  value_scaledi[0] = 1.0f / value_scaledf[0]; value_scaledi[1] = 1.0f / value_scaledf[1];
  value_scaledi[2] = -value_scaledf[2] / value_scaledf[0]; value_scaledi[3] = -value_scaledf[3] / value_scaledf[1];
  for (int i = 0; i < 4; i++) value_minus_value_zero[i] = value[i] - value_zero[i];
}

// NumType.h:157-168
static void fromToVecExposure(float exposureF, float exposureT, double aF, double bF, double aT, double bT, double out[2]) {
  if (exposureF == 0 || exposureT == 0) exposureT = exposureF = 1;
  double a = std::exp(aT - aF) * exposureT / exposureF;
  double b = bT - a * bF;
  out[0] = a; out[1] = b;
}

// FrameFramePrecalc::set (HessianBlocks.cpp:431-461)
void precalc_set(const FrameH &host, const FrameH &target, const CalibH &HCalib, Precalc &pc) {
  SE3 leftToLeft_0 = target.evalPT.inverse() * host.evalPT;
  M3<double> R0 = leftToLeft_0.R();
  for (int i = 0; i < 9; i++) pc.RTll_0[i] = (float)R0.m[i];
  for (int i = 0; i < 3; i++) pc.tTll_0[i] = (float)leftToLeft_0.t[i];
  SE3 leftToLeft = target.PRE_worldToCam * host.PRE_camToWorld;
  M3<double> Rd = leftToLeft.R();
  M3<float> R; V3<float> t;
  for (int i = 0; i < 9; i++) R.m[i] = (float)Rd.m[i];
  for (int i = 0; i < 3; i++) t[i] = (float)leftToLeft.t[i];
  pc.dist = (float)std::sqrt(leftToLeft.t[0] * leftToLeft.t[0] + leftToLeft.t[1] * leftToLeft.t[1] + leftToLeft.t[2] * leftToLeft.t[2]);
  const float fx = HCalib.value_scaledf[0], fy = HCalib.value_scaledf[1], cx = HCalib.value_scaledf[2], cy = HCalib.value_scaledf[3];
  M3<float> K, Ki;
  for (int i = 0; i < 9; i++) K.m[i] = Ki.m[i] = 0;
  K(0, 0) = fx; K(1, 1) = fy; K(0, 2) = cx; K(1, 2) = cy; K(2, 2) = 1;
  // K.inverse() of the upper-triangular pinhole matrix (closed form)
  Ki(0, 0) = 1.0f / fx; Ki(1, 1) = 1.0f / fy; Ki(0, 2) = -cx / fx; Ki(1, 2) = -cy / fy; Ki(2, 2) = 1;
  M3<float> KRKi = mul(mul(K, R), Ki);
  V3<float> Kt = mul(K, t);
  for (int i = 0; i < 9; i++) pc.KRKi[i] = KRKi.m[i];
  for (int i = 0; i < 3; i++) pc.Kt[i] = Kt[i];
  double aff[2];
  fromToVecExposure(host.ab_exposure, target.ab_exposure, host.state_scaled[6], host.state_scaled[7], target.state_scaled[6], target.state_scaled[7], aff);
  pc.aff[0] = (float)aff[0]; pc.aff[1] = (float)aff[1];
  pc.b0 = (float)(host.state_zero[7] * SCALE_B);
}

// EnergyFunctional::setAdjointsF (EnergyFunctional.cpp:42-103)
void set_adjoints(const std::vector<FrameH> &frames, std::vector<double> &adHost, std::vector<double> &adTarget) {
  const int nf = (int)frames.size();
  adHost.assign((size_t)nf * nf * 64, 0.0); adTarget.assign((size_t)nf * nf * 64, 0.0);
  for (int h = 0; h < nf; h++)
    for (int t = 0; t < nf; t++) {
      const FrameH &host = frames[h], &target = frames[t];
      SE3 worldToTarget = target.evalPT.inverse();
      double Adj[36];
      se3_adj(worldToTarget, Adj);
      double AH[64], AT[64];
      for (int i = 0; i < 64; i++) AH[i] = AT[i] = 0;
      for (int i = 0; i < 8; i++) AH[i * 8 + i] = AT[i * 8 + i] = 1;
      for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { AH[i * 8 + j] = Adj[j * 6 + i]; AT[i * 8 + j] = -Adj[j * 6 + i]; }
      double aff[2];
      fromToVecExposure(host.ab_exposure, target.ab_exposure, host.state_zero[6] * SCALE_A, host.state_zero[7] * SCALE_B,
                        target.state_zero[6] * SCALE_A, target.state_zero[7] * SCALE_B, aff);
      float affLL0 = (float)aff[0];
      AT[6 * 8 + 6] = -affLL0; AH[6 * 8 + 6] = affLL0; AT[7 * 8 + 7] = -1; AH[7 * 8 + 7] = affLL0;
      for (int j = 0; j < 8; j++) {
        for (int i = 0; i < 3; i++) { AH[i * 8 + j] *= SCALE_XI_TRANS; AT[i * 8 + j] *= SCALE_XI_TRANS; }
        for (int i = 3; i < 6; i++) { AH[i * 8 + j] *= SCALE_XI_ROT; AT[i * 8 + j] *= SCALE_XI_ROT; }
        AH[6 * 8 + j] *= SCALE_A; AT[6 * 8 + j] *= SCALE_A;
        AH[7 * 8 + j] *= SCALE_B; AT[7 * 8 + j] *= SCALE_B;
      }
      memcpy(&adHost[64 * (h + t * nf)], AH, sizeof(AH));
      memcpy(&adTarget[64 * (h + t * nf)], AT, sizeof(AT));
    }
}

// EnergyFunctional::setDeltaF (EnergyFunctional.cpp:163-194): adHTdeltaF part
void set_delta(const std::vector<FrameH> &frames, const std::vector<float> &adHostF, const std::vector<float> &adTargetF,
               std::vector<float> &adHTdeltaF) {
  const int nf = (int)frames.size();
  adHTdeltaF.assign((size_t)nf * nf * 8, 0.f);
  for (int h = 0; h < nf; h++)
    for (int t = 0; t < nf; t++) {
      int idx = h + t * nf;
      float dh[8], dt[8];
      for (int i = 0; i < 8; i++) { dh[i] = (float)(frames[h].state[i] - frames[h].state_zero[i]); dt[i] = (float)(frames[t].state[i] - frames[t].state_zero[i]); }
      for (int c = 0; c < 8; c++) {
        float sh = 0, st = 0;
        for (int k = 0; k < 8; k++) { sh += dh[k] * adHostF[64 * idx + k * 8 + c]; st += dt[k] * adTargetF[64 * idx + k * 8 + c]; }
        adHTdeltaF[8 * idx + c] = sh + st;
      }
    }
}

// ---------------------------------------------------------------------------------------------
void BAState::load(Oracle &o, const sosba_ba_problem *prob) {
  frames.resize(prob->nf);
  for (int i = 0; i < prob->nf; i++) {
    const sosba_frame_state &fs = prob->frames[i];
    FrameH &f = frames[i];
    f.evalPT = SE3::from_rowmajor34(fs.camToWorld_evalPT);
    f.ab_exposure = fs.ab_exposure; f.frameEnergyTH = fs.frame_energy_th; f.frameID = fs.frame_id; f.slot = fs.slot;
    f.setState(fs.state);
    for (int k = 0; k < 10; k++) f.state_zero[k] = fs.state_zero[k];
    for (int k = 0; k < 10; k++) f.step[k] = 0;
    f.getPrior(o.cfg, f.prior);  // EFFrame::takeData at insertion (EnergyFunctionalStructs.cpp:47-50)
  }
  for (int i = 0; i < 4; i++) calib.value_zero[i] = prob->calib_value_zero[i];
  calib.setValue(prob->calib_value);
  for (int i = 0; i < 4; i++) calib.step[i] = 0;
}

// FullSystem::setPrecalcValues (FullSystem.cpp:1099-1107) + ef->setDeltaF; pushes tables into the Oracle
void BAState::setPrecalcValues(Oracle &o) {
  const int nf = (int)frames.size();
  o.nf = nf;
  o.pre.resize((size_t)nf * nf);
  for (int h = 0; h < nf; h++) for (int t = 0; t < nf; t++) precalc_set(frames[h], frames[t], calib, o.pre[h * nf + t]);
  set_delta(frames, o.adHostF, o.adTargetF, o.adHTdeltaF);
  for (int i = 0; i < 4; i++) o.cDeltaF[i] = (float)calib.value_minus_value_zero[i];
  o.fdelta.resize((size_t)nf * 8); o.fdelta_prior.resize((size_t)nf * 8); o.fprior.resize((size_t)nf * 8);
  for (int h = 0; h < nf; h++)
    for (int i = 0; i < 8; i++) {
      o.fdelta[h * 8 + i] = frames[h].state[i] - frames[h].state_zero[i];
      o.fdelta_prior[h * 8 + i] = frames[h].state[i];  // getPriorZero() == 0
      o.fprior[h * 8 + i] = frames[h].prior[i];
    }
  for (auto &p : o.pts) p.deltaF = p.idepth - p.idepth_zero;
  o.fxl = calib.value_scaledf[0]; o.fyl = calib.value_scaledf[1]; o.cxl = calib.value_scaledf[2]; o.cyl = calib.value_scaledf[3];
  o.fxli = calib.value_scaledi[0]; o.fyli = calib.value_scaledi[1];
  o.frameEnergyTH.resize(nf);
  for (int h = 0; h < nf; h++) o.frameEnergyTH[h] = frames[h].frameEnergyTH;
}
void BAState::setAdjoints(Oracle &o) {
  set_adjoints(frames, o.adHost, o.adTarget);
  o.adHostF.resize(o.adHost.size()); o.adTargetF.resize(o.adTarget.size());
  for (size_t i = 0; i < o.adHost.size(); i++) { o.adHostF[i] = (float)o.adHost[i]; o.adTargetF[i] = (float)o.adTarget[i]; }
  for (int i = 0; i < 4; i++) o.cPrior[i] = o.cfg.initial_calib_hessian;
}

// FullSystem::doStepFromBackup (FullSystemOptimize.cpp:185-257)
bool BAState::doStepFromBackup(Oracle &o, float stepfac, double *sums) {
  double pstepfac[10];
  for (int i = 0; i < 10; i++) pstepfac[i] = stepfac;
  float sumA = 0, sumB = 0, sumT = 0, sumR = 0, sumID = 0, numID = 0, sumNID = 0;
  double v[4];
  for (int i = 0; i < 4; i++) v[i] = calib.value_backup[i] + stepfac * calib.step[i];
  calib.setValue(v);
  const int nf = (int)frames.size();
  for (int fi = 0; fi < nf; fi++) {
    FrameH &fh = frames[fi];
    double s[10];
    for (int i = 0; i < 10; i++) s[i] = fh.state_backup[i] + pstepfac[i] * fh.step[i];
    fh.setState(s);
    sumA += fh.step[6] * fh.step[6];
    sumB += fh.step[7] * fh.step[7];
    sumT += fh.step[0] * fh.step[0] + fh.step[1] * fh.step[1] + fh.step[2] * fh.step[2];
    sumR += fh.step[3] * fh.step[3] + fh.step[4] * fh.step[4] + fh.step[5] * fh.step[5];
  }
  for (auto &ph : o.pts) {  // all pointHessians of all frames
    float nid = ph.idepth_backup + stepfac * ph.step;
    ph.idepth = nid; ph.idepth_scaled = SCALE_IDEPTH * nid;
    sumID += ph.step * ph.step;
    sumNID += fabsf(ph.idepth_backup);
    numID++;
    ph.idepth_zero = nid; ph.idepth_zero_scaled = SCALE_IDEPTH * nid;
  }
  sumA /= nf; sumB /= nf; sumR /= nf; sumT /= nf;
  if (sums) { sums[0] = sumA; sums[1] = sumB; sums[2] = sumT; sums[3] = sumR; sums[4] = sumID; sums[5] = sumNID; sums[6] = numID; }
  sumID /= numID; sumNID /= numID;
  setPrecalcValues(o);
  const float th = o.cfg.th_opt_iterations;
  return sqrtf(sumA) < 0.0005 * th && sqrtf(sumB) < 0.00005 * th && sqrtf(sumR) < 0.00005 * th && sqrtf(sumT) * sumNID < 0.00005 * th;
}

void BAState::backupState(Oracle &o) {  // :260-271
  for (int i = 0; i < 4; i++) calib.value_backup[i] = calib.value[i];
  for (auto &fh : frames) for (int i = 0; i < 10; i++) fh.state_backup[i] = fh.state[i];
  for (auto &ph : o.pts) ph.idepth_backup = ph.idepth;
}

// the loop body of FullSystem::optimize (FullSystemOptimize.cpp:358-413) with setting_forceAceptStep
bool BAState::iterate(Oracle &o, const double *HM, const double *bM, sosba_linearize_out *lo) {
  const int nf = (int)frames.size(), D = CPARS + 8 * nf;
  backupState(o);
  std::vector<double> x(D);
  solveSystemF(o, HM, bM, x.data(), nullptr, nullptr);
  for (int i = 0; i < 4; i++) calib.step[i] = -x[i];  // resubstituteF_MT :500-507
  for (int h = 0; h < nf; h++) { for (int i = 0; i < 8; i++) frames[h].step[i] = -x[CPARS + 8 * h + i]; frames[h].step[8] = frames[h].step[9] = 0; }
  bool canbreak = doStepFromBackup(o, 1.0f);
  linearizeAll(o, false, lo);
  frames.back().frameEnergyTH = o.frameEnergyTH[nf - 1];
  for (int id : o.activeResiduals) applyRes(o.res[id], true);
  return canbreak;
}

// the same body without the solve: the caller provides x (EnergyFunctional::solveSystemF with IMU ends in lastX = x_dso and
// resubstituteF_MT, EnergyFunctional.cpp:1150-1182)
void BAState::step_with_x(Oracle &o, const double *x, sosba_linearize_out *lo, double sums[7]) {
  const int nf = (int)frames.size();
  backupState(o);
  resubstituteF(o, x);
  for (int i = 0; i < 4; i++) calib.step[i] = -x[i];
  for (int h = 0; h < nf; h++) { for (int i = 0; i < 8; i++) frames[h].step[i] = -x[CPARS + 8 * h + i]; frames[h].step[8] = frames[h].step[9] = 0; }
  doStepFromBackup(o, 1.0f, sums);
  linearizeAll(o, false, lo);
  frames.back().frameEnergyTH = o.frameEnergyTH[nf - 1];
  for (int id : o.activeResiduals) applyRes(o.res[id], true);
}

// FullSystem::optimize (FullSystemOptimize.cpp:305-489), IMU off
void BAState::optimize(Oracle &o, const double *HM, const double *bM, int mnumOptIts, sosba_optimize_out *out) {
  const int nf = (int)frames.size();
  memset(out, 0, sizeof(*out));
  if (nf < 2) return;
  if (nf < 3) mnumOptIts = 20;
  if (nf < 4) mnumOptIts = 15;
  // activeResiduals + resetOOB (:316-329)
  o.activeResiduals.clear();
  for (int i = 0; i < (int)o.res.size(); i++) {
    Res &r = o.res[i];
    if (r.dropped) continue;
    if (!r.isLinearized) {
      o.activeResiduals.push_back(i);
      r.state_NewEnergy = r.state_energy = 0; r.state_NewState = SOSBA_RES_OUTLIER; r.state_state = SOSBA_RES_IN;
    }
  }
  setAdjoints(o);
  setPrecalcValues(o);
  sosba_linearize_out lo;
  linearizeAll(o, false, &lo);
  frames.back().frameEnergyTH = o.frameEnergyTH[nf - 1];
  out->energy_initial = lo.energy;
  out->reserved0 = lo.n_in + lo.n_oob + lo.n_outlier;  // residuals linearised per pass (bench bookkeeping)
  for (int id : o.activeResiduals) applyRes(o.res[id], true);
  int it = 0;
  for (int iteration = 0; iteration < mnumOptIts; iteration++) {
    bool canbreak = iterate(o, HM, bM, &lo);
    it++;
    if (canbreak && iteration >= o.cfg.min_opt_iterations) break;
  }
  out->iterations = it;
  double newStateZero[10] = {0, 0, 0, 0, 0, 0, frames.back().state[6], frames.back().state[7], 0, 0};
  frames.back().setEvalPT(frames.back().PRE_camToWorld, newStateZero);
  setAdjoints(o);
  setPrecalcValues(o);
  linearizeAll(o, true, &lo);
  frames.back().frameEnergyTH = o.frameEnergyTH[nf - 1];
  out->energy_final = lo.energy;
  out->n_removed = lo.n_removed;
  out->res_in_a = o.resInA;
  out->rmse = sqrtf((float)(lo.energy / (patternNum * o.resInA)));
  double n2 = 0;
  for (double v : o.lastX) n2 += v * v;
  out->last_x_norm = std::sqrt(n2);
}

void BAState::store(Oracle &o, sosba_ba_problem *prob) const {
  for (int i = 0; i < prob->nf; i++) {
    sosba_frame_state &fs = prob->frames[i];
    const FrameH &f = frames[i];
    f.evalPT.to_rowmajor34(fs.camToWorld_evalPT);
    for (int k = 0; k < 10; k++) { fs.state[k] = f.state[k]; fs.state_zero[k] = f.state_zero[k]; }
    fs.frame_energy_th = f.frameEnergyTH;
  }
  for (int i = 0; i < 4; i++) prob->calib_value[i] = calib.value[i];
  if (prob->idepth_out) for (size_t i = 0; i < o.pts.size(); i++) prob->idepth_out[i] = o.pts[i].idepth;
}

}  // namespace orc
