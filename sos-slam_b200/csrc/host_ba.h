// host_ba.h — host-side mirror of the FullSystem state the Gauss-Newton loop touches, in fp64:
//   FrameHessian::setState / setEvalPT / getPrior      src/FullSystem/HessianBlocks.h:217-309
//   CalibHessian::setValue                              src/FullSystem/HessianBlocks.h:487-501
//   FrameFramePrecalc::set                              src/FullSystem/HessianBlocks.cpp:431-461
//   AffLight::fromToVecExposure                         src/util/NumType.h:157-168
//   EnergyFunctional::setAdjointsF / setDeltaF          src/OptimizationBackend/EnergyFunctional.cpp:42-103, 163-194
// These are O(nf^2) scalar computations per Gauss-Newton step; they stay on the host and
// their outputs are the window tables the kernels read.
#pragma once
#include <vector>

#include "../../include/sosba.h"
#include "host_math.h"

namespace sosba_host {

constexpr double SCALE_XI_ROT = 1.0, SCALE_XI_TRANS = 0.5, SCALE_F = 50.0, SCALE_C = 50.0, SCALE_A = 10.0, SCALE_B = 1000.0;

struct FrameH {
  sosba_math::Rigid evalPT, camToWorld, worldToCam;
  double state[10], state_zero[10], state_scaled[10], state_backup[10], step[10], prior[8];
  float ab_exposure, frameEnergyTH;
  int frameID, slot;
  void setState(const double *s);
  void setEvalPT(const sosba_math::Rigid &e, const double *s);
};

struct CalibH {
  double value[4], value_zero[4], value_scaled[4], value_backup[4], step[4], value_minus_value_zero[4];
  float value_scaledf[4], value_scaledi[4];
  void setValue(const double *v);
};

struct WindowTables {   // everything sosba_window carries, host side
  int nf = 0;
  std::vector<int> frame_slot;
  std::vector<float> precalc, adHTdeltaF, frameEnergyTH;
  std::vector<double> adHost, adTarget, frame_prior, frame_delta_prior, frame_delta;
  std::vector<float> adHostF, adTargetF;
  float calib[4], cDeltaF[4];
  double cPrior[4];
};

struct BAState {
  std::vector<FrameH> frames;
  CalibH calib;
  std::vector<double> HM, bM;
  bool loaded = false;
  void load(const sosba_config &cfg, const sosba_ba_problem *prob);
  void store(sosba_ba_problem *prob) const;
  void make_adjoints(const sosba_config &cfg, WindowTables &w) const;   // setAdjointsF
  void make_precalc(WindowTables &w) const;                            // setPrecalcValues + setDeltaF
  void backup();
  // frames/calib part of doStepFromBackup; returns the frame sums {sumA,sumB,sumT,sumR} already divided by nf
  void step_frames(const double *x, float stepfac, float sums[4]);
};

void frame_prior(const sosba_config &cfg, int frameID, double p[8]);

}  // namespace sosba_host
