// host_ba.cpp — see host_ba.h.  Built with -ffp-contract=off: the float tables produced here feed the
// per-residual kernels whose integer outcomes are part of the parity contract.
#include "host_ba.h"

#include <cmath>
#include <cstring>

namespace sosba_host {

using sosba_math::Rigid;

void FrameH::setState(const double *s) {  // HessianBlocks.h:217-230
  for (int i = 0; i < 10; i++) state[i] = s[i];
  for (int i = 0; i < 3; i++) state_scaled[i] = SCALE_XI_TRANS * state[i];
  for (int i = 3; i < 6; i++) state_scaled[i] = SCALE_XI_ROT * state[i];
  state_scaled[6] = SCALE_A * state[6];
  state_scaled[7] = SCALE_B * state[7];
  state_scaled[8] = SCALE_A * state[8];
  state_scaled[9] = SCALE_B * state[9];
  camToWorld = sosba_math::rigid_mul(sosba_math::rigid_exp(state_scaled), evalPT);
  worldToCam = sosba_math::rigid_inverse(camToWorld);
}

void FrameH::setEvalPT(const Rigid &e, const double *s) {  // HessianBlocks.h:245-251
  evalPT = e;
  setState(s);
  for (int i = 0; i < 10; i++) state_zero[i] = state[i];
}

void frame_prior(const sosba_config &cfg, int frameID, double p[8]) {  // FrameHessian::getPrior, HessianBlocks.h:288-309
  for (int i = 0; i < 8; i++) p[i] = 0.0;
  if (frameID == 0) {
    p[0] = p[1] = p[2] = cfg.initial_trans_prior;
    p[3] = p[4] = p[5] = cfg.initial_rot_prior;
    p[6] = cfg.initial_aff_a_prior;
    p[7] = cfg.initial_aff_b_prior;
  } else {
    p[6] = cfg.affine_opt_mode_a < 0 ? cfg.initial_aff_a_prior : cfg.affine_opt_mode_a;
    p[7] = cfg.affine_opt_mode_b < 0 ? cfg.initial_aff_b_prior : cfg.affine_opt_mode_b;
  }
}

void CalibH::setValue(const double *v) {  // HessianBlocks.h:487-501
  for (int i = 0; i < 4; i++) value[i] = v[i];
  value_scaled[0] = SCALE_F * value[0];
  value_scaled[1] = SCALE_F * value[1];
  value_scaled[2] = SCALE_C * value[2];
  value_scaled[3] = SCALE_C * value[3];
  for (int i = 0; i < 4; i++) value_scaledf[i] = (float)value_scaled[i];
  value_scaledi[0] = 1.0f / value_scaledf[0];
  value_scaledi[1] = 1.0f / value_scaledf[1];
  value_scaledi[2] = -value_scaledf[2] / value_scaledf[0];
  value_scaledi[3] = -value_scaledf[3] / value_scaledf[1];
  for (int i = 0; i < 4; i++) value_minus_value_zero[i] = value[i] - value_zero[i];
}

// AffLight::fromToVecExposure, NumType.h:157-168
static void aff_from_to(float expF, float expT, double aF, double bF, double aT, double bT, double out[2]) {
  if (expF == 0 || expT == 0) expT = expF = 1;
  const double a = std::exp(aT - aF) * expT / expF;
  out[0] = a;
  out[1] = bT - a * bF;
}

static inline void mul33f(const float *A, const float *B, float *C) {  // coefficient-wise, left-to-right sums
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[3 * i + j] = (A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j]) + A[3 * i + 2] * B[6 + j];
}

void BAState::load(const sosba_config &cfg, const sosba_ba_problem *prob) {
  frames.resize(prob->nf);
  for (int i = 0; i < prob->nf; i++) {
    const sosba_frame_state &fs = prob->frames[i];
    FrameH &f = frames[i];
    f.evalPT = sosba_math::rigid_from34(fs.camToWorld_evalPT);
    f.ab_exposure = fs.ab_exposure;
    f.frameEnergyTH = fs.frame_energy_th;
    f.frameID = fs.frame_id;
    f.slot = fs.slot;
    f.setState(fs.state);
    for (int k = 0; k < 10; k++) { f.state_zero[k] = fs.state_zero[k]; f.step[k] = 0.0; f.state_backup[k] = fs.state[k]; }
    frame_prior(cfg, f.frameID, f.prior);  // EFFrame::takeData (EnergyFunctionalStructs.cpp:47-50)
  }
  for (int i = 0; i < 4; i++) { calib.value_zero[i] = prob->calib_value_zero[i]; calib.step[i] = 0.0; }
  calib.setValue(prob->calib_value);
  const int D = 4 + 8 * prob->nf;
  if (prob->HM && prob->bM) { HM.assign(prob->HM, prob->HM + (size_t)D * D); bM.assign(prob->bM, prob->bM + D); }
  else { HM.clear(); bM.clear(); }
  loaded = true;
}

void BAState::store(sosba_ba_problem *prob) const {
  for (int i = 0; i < prob->nf; i++) {
    sosba_frame_state &fs = prob->frames[i];
    const FrameH &f = frames[i];
    sosba_math::rigid_to34(f.evalPT, fs.camToWorld_evalPT);
    for (int k = 0; k < 10; k++) { fs.state[k] = f.state[k]; fs.state_zero[k] = f.state_zero[k]; }
    fs.frame_energy_th = f.frameEnergyTH;
  }
  for (int i = 0; i < 4; i++) prob->calib_value[i] = calib.value[i];
}

// EnergyFunctional::setAdjointsF (EnergyFunctional.cpp:42-103)
void BAState::make_adjoints(const sosba_config &cfg, WindowTables &w) const {
  const int nf = (int)frames.size();
  w.nf = nf;
  w.adHost.assign((size_t)nf * nf * 64, 0.0);
  w.adTarget.assign((size_t)nf * nf * 64, 0.0);
  for (int t = 0; t < nf; t++) {
    double Adj[36];
    sosba_math::rigid_adj(sosba_math::rigid_inverse(frames[t].evalPT), Adj);  // worldToTarget at the evaluation point
    for (int h = 0; h < nf; h++) {
      double *AH = &w.adHost[64 * (size_t)(h + t * nf)], *AT = &w.adTarget[64 * (size_t)(h + t * nf)];
      for (int i = 0; i < 8; i++) AH[9 * i] = AT[9 * i] = 1.0;
      for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) { AH[8 * i + j] = Adj[6 * j + i]; AT[8 * i + j] = -Adj[6 * j + i]; }
      double aff[2];
      aff_from_to(frames[h].ab_exposure, frames[t].ab_exposure, frames[h].state_zero[6] * SCALE_A, frames[h].state_zero[7] * SCALE_B,
                  frames[t].state_zero[6] * SCALE_A, frames[t].state_zero[7] * SCALE_B, aff);
      const float a0 = (float)aff[0];
      AT[8 * 6 + 6] = -a0; AH[8 * 6 + 6] = a0; AT[8 * 7 + 7] = -1; AH[8 * 7 + 7] = a0;
      const double rs[8] = {SCALE_XI_TRANS, SCALE_XI_TRANS, SCALE_XI_TRANS, SCALE_XI_ROT, SCALE_XI_ROT, SCALE_XI_ROT, SCALE_A, SCALE_B};
      for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) { AH[8 * i + j] *= rs[i]; AT[8 * i + j] *= rs[i]; }
    }
  }
  w.adHostF.resize(w.adHost.size());
  w.adTargetF.resize(w.adTarget.size());
  for (size_t i = 0; i < w.adHost.size(); i++) { w.adHostF[i] = (float)w.adHost[i]; w.adTargetF[i] = (float)w.adTarget[i]; }
  for (int i = 0; i < 4; i++) w.cPrior[i] = cfg.initial_calib_hessian;
  w.frame_prior.resize((size_t)nf * 8);
  for (int h = 0; h < nf; h++) for (int i = 0; i < 8; i++) w.frame_prior[8 * h + i] = frames[h].prior[i];
  w.frame_slot.resize(nf);
  for (int h = 0; h < nf; h++) w.frame_slot[h] = frames[h].slot;
}

// FullSystem::setPrecalcValues (FullSystem.cpp:1099-1107) -> FrameFramePrecalc::set (HessianBlocks.cpp:431-461)
// + EnergyFunctional::setDeltaF (EnergyFunctional.cpp:163-194)
void BAState::make_precalc(WindowTables &w) const {
  const int nf = (int)frames.size();
  w.precalc.assign((size_t)nf * nf * SOSBA_PRECALC_FLOATS, 0.f);
  const float fx = calib.value_scaledf[0], fy = calib.value_scaledf[1], cx = calib.value_scaledf[2], cy = calib.value_scaledf[3];
  const float K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
  const float Ki[9] = {1.0f / fx, 0, -cx / fx, 0, 1.0f / fy, -cy / fy, 0, 0, 1};
  for (int h = 0; h < nf; h++)
    for (int t = 0; t < nf; t++) {
      float *p = &w.precalc[(size_t)(h * nf + t) * SOSBA_PRECALC_FLOATS];
      const Rigid l0 = sosba_math::rigid_mul(sosba_math::rigid_inverse(frames[t].evalPT), frames[h].evalPT);
      for (int i = 0; i < 9; i++) p[SOSBA_PC_RTLL0 + i] = (float)l0.R[i];
      for (int i = 0; i < 3; i++) p[SOSBA_PC_TTLL0 + i] = (float)l0.t[i];
      const Rigid l = sosba_math::rigid_mul(frames[t].worldToCam, frames[h].camToWorld);
      float R[9], tt[3], KR[9];
      for (int i = 0; i < 9; i++) R[i] = (float)l.R[i];
      for (int i = 0; i < 3; i++) tt[i] = (float)l.t[i];
      mul33f(K, R, KR);
      mul33f(KR, Ki, p + SOSBA_PC_KRKI);
      for (int i = 0; i < 3; i++) p[SOSBA_PC_KT + i] = (K[3 * i] * tt[0] + K[3 * i + 1] * tt[1]) + K[3 * i + 2] * tt[2];
      double aff[2];
      aff_from_to(frames[h].ab_exposure, frames[t].ab_exposure, frames[h].state_scaled[6], frames[h].state_scaled[7],
                  frames[t].state_scaled[6], frames[t].state_scaled[7], aff);
      p[SOSBA_PC_AFF] = (float)aff[0];
      p[SOSBA_PC_AFF + 1] = (float)aff[1];
      p[SOSBA_PC_B0] = (float)(frames[h].state_zero[7] * SCALE_B);
      p[SOSBA_PC_DIST] = (float)std::sqrt(l.t[0] * l.t[0] + l.t[1] * l.t[1] + l.t[2] * l.t[2]);
    }
  // setDeltaF
  w.adHTdeltaF.assign((size_t)nf * nf * 8, 0.f);
  for (int h = 0; h < nf; h++)
    for (int t = 0; t < nf; t++) {
      const int idx = h + t * nf;
      float dh[8], dt[8];
      for (int i = 0; i < 8; i++) {
        dh[i] = (float)(frames[h].state[i] - frames[h].state_zero[i]);
        dt[i] = (float)(frames[t].state[i] - frames[t].state_zero[i]);
      }
      for (int c = 0; c < 8; c++) {
        float sh = 0, st = 0;
        for (int k = 0; k < 8; k++) { sh += dh[k] * w.adHostF[64 * (size_t)idx + 8 * k + c]; st += dt[k] * w.adTargetF[64 * (size_t)idx + 8 * k + c]; }
        w.adHTdeltaF[8 * (size_t)idx + c] = sh + st;
      }
    }
  for (int i = 0; i < 4; i++) { w.calib[i] = calib.value_scaledf[i]; w.cDeltaF[i] = (float)calib.value_minus_value_zero[i]; }
  w.frame_delta.resize((size_t)nf * 8);
  w.frame_delta_prior.resize((size_t)nf * 8);
  w.frameEnergyTH.resize(nf);
  for (int h = 0; h < nf; h++) {
    for (int i = 0; i < 8; i++) {
      w.frame_delta[8 * h + i] = frames[h].state[i] - frames[h].state_zero[i];
      w.frame_delta_prior[8 * h + i] = frames[h].state[i];  // state - getPriorZero(), getPriorZero() == 0
    }
    w.frameEnergyTH[h] = frames[h].frameEnergyTH;
  }
}

void BAState::backup() {  // FullSystem::backupState (FullSystemOptimize.cpp:260-271), frames + calib
  for (int i = 0; i < 4; i++) calib.value_backup[i] = calib.value[i];
  for (auto &f : frames) for (int i = 0; i < 10; i++) f.state_backup[i] = f.state[i];
}

// frames / calib part of EnergyFunctional::resubstituteF_MT (:500-507) + FullSystem::doStepFromBackup (:185-257)
void BAState::step_frames(const double *x, float stepfac, float sums[4]) {
  const int nf = (int)frames.size();
  for (int i = 0; i < 4; i++) calib.step[i] = -x[i];
  double v[4];
  for (int i = 0; i < 4; i++) v[i] = calib.value_backup[i] + stepfac * calib.step[i];
  calib.setValue(v);
  float sumA = 0, sumB = 0, sumT = 0, sumR = 0;
  for (int h = 0; h < nf; h++) {
    FrameH &f = frames[h];
    for (int i = 0; i < 8; i++) f.step[i] = -x[4 + 8 * h + i];
    f.step[8] = f.step[9] = 0.0;
    double s[10];
    for (int i = 0; i < 10; i++) s[i] = f.state_backup[i] + (double)stepfac * f.step[i];
    f.setState(s);
    sumA += f.step[6] * f.step[6];
    sumB += f.step[7] * f.step[7];
    sumT += f.step[0] * f.step[0] + f.step[1] * f.step[1] + f.step[2] * f.step[2];
    sumR += f.step[3] * f.step[3] + f.step[4] * f.step[4] + f.step[5] * f.step[5];
  }
  sums[0] = sumA / nf; sums[1] = sumB / nf; sums[2] = sumT / nf; sums[3] = sumR / nf;
}

}  // namespace sosba_host
