// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).  PARITY UNPINNED.
//
// orc_core.h: data model of the restated path (PointFrameResidual / EFResidual / EFPoint / EFFrame /
// FrameFramePrecalc / CalibHessian fields that the hot path touches).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

#include "../include/sosba.h"
#include "orc_accum.h"
#include "orc_math.h"
#include "orc_pool.h"

namespace orc {

// HessianBlocks.h:53-60
static const float SCALE_IDEPTH = 1.0f;
static const float SCALE_XI_ROT = 1.0f;
static const float SCALE_XI_TRANS = 0.5f;
static const float SCALE_F = 50.0f;
static const float SCALE_C = 50.0f;
static const float SCALE_A = 10.0f;
static const float SCALE_B = 1000.0f;
static const int CPARS = 4;
static const int patternNum = 8;  // settings.h:187
// staticPattern[8], settings.cpp:307-309
static const int patternP[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};

// RawResidualJacobian.h:29-55
struct RawJ {
  float resF[8];
  float Jpdxi[2][6];
  float Jpdc[2][4];
  float Jpdd[2];
  float JIdx[2][8];
  float JabF[2][8];
  float JIdx2[2][2];
  float JabJIdx[2][2];
  float Jab2[2][2];
};

// HessianBlocks.h:109-134 (the members the path reads)
struct Precalc {
  float RTll_0[9], tTll_0[3], KRKi[9], Kt[3], aff[2], b0, dist;
};

struct Level { int w = 0, h = 0; std::vector<float> dI; std::vector<float> absg; };  // dI: Vector3f AoS
struct Pyramid { Level lvl[SOSBA_MAX_LEVELS]; bool valid = false; };

// PointFrameResidual (Residuals.h:49-93) + EFResidual (EnergyFunctionalStructs.h:43-81)
struct Res {
  int point, host, target;
  int state_state, state_NewState;
  double state_energy, state_NewEnergy, state_NewEnergyWithOutlier;
  bool isNew, isLinearized, isActive, dropped;
  RawJ Jbuf[2];
  int sel;  // PointFrameResidual::J == Jbuf[sel]; EFResidual::J == Jbuf[1-sel]
  float res_toZeroF[8];
  float JpJdF[8];
  float projectedTo[8][2];
  float centerProjectedTo[3];
  RawJ &Jdata() { return Jbuf[sel]; }
  RawJ &Jef() { return Jbuf[1 - sel]; }
};

// PointHessian (HessianBlocks.h:556-649) + EFPoint (EnergyFunctionalStructs.h:85-116)
struct Pt {
  float u, v, idepth, idepth_zero, idepth_scaled, idepth_zero_scaled;
  float color[8], weights[8];
  int host;
  float priorF, deltaF;
  float bdSumF, HdiF, Hdd_accLF, Hcd_accLF[4], bd_accLF, Hdd_accAF, Hcd_accAF[4], bd_accAF;
  float step, idepth_backup, idepth_hessian, maxRelBaseline;
  int numGoodResiduals;
  int res_begin, res_end;  // EFPoint::residualsAll == res[res_begin, res_end) minus dropped
};

struct Oracle {
  sosba_config cfg;
  int levels = 0;
  int wl[SOSBA_MAX_LEVELS], hl[SOSBA_MAX_LEVELS];
  float wM3G, hM3G;  // globalCalib.cpp:63-64
  std::vector<Pyramid> slots;

  // pixel selector (orc_select.cpp)
  struct Selector {
    std::vector<uint8_t> randomPattern;
    int currentPotential = 3, thsStep = 0;
    std::vector<float> ths, thsSmoothed;
  } sel;

  // pre-pyramid image path (orc_trace.cpp: undistort_raw)
  struct Undist {
    bool set = false, passthrough = true, haveG = false, haveV = false;
    int wOrg = 0, hOrg = 0, gDepth = 0;
    std::vector<float> remapX, remapY, G, vignetteInv;
  } und;

  // window
  int nf = 0;
  std::vector<int> frame_slot;
  std::vector<Precalc> pre;  // host*nf+target
  std::vector<double> adHost, adTarget;
  std::vector<float> adHostF, adTargetF, adHTdeltaF;
  std::vector<float> frameEnergyTH;
  float fxl, fyl, cxl, cyl, fxli, fyli;
  float cDeltaF[4];
  double cPrior[4];
  std::vector<double> fprior, fdelta_prior, fdelta;

  std::vector<Pt> pts;
  std::vector<Res> res;
  std::vector<int> activeResiduals;

  // accumulators: [tid][block]
  int T = 1;  // NUM_THREADS of the run (NumType.h:37 has 6)
  bool MT = false;
  std::unique_ptr<ThreadReduce> red;
  std::vector<std::vector<AccumulatorApprox>> accA, accL;
  std::vector<int> nresA, nresL;
  std::vector<std::vector<AccumulatorXX<8, 4>>> accE;
  std::vector<std::vector<AccumulatorX<8>>> accEB;
  std::vector<std::vector<AccumulatorXX<8, 8>>> accD;
  std::vector<AccumulatorXX<4, 4>> accHcc;
  std::vector<AccumulatorX<4>> accbc;
  int resInA = 0, resInL = 0, resInM = 0;
  std::vector<double> lastX;

  // tracker / scale optimizer (ScaleOptimizer.h:60-95, CoarseTracker.h)
  float tfx[SOSBA_MAX_LEVELS], tfy[SOSBA_MAX_LEVELS], tcx[SOSBA_MAX_LEVELS], tcy[SOSBA_MAX_LEVELS];
  M3<float> tKi[SOSBA_MAX_LEVELS];
  std::vector<float> pc_u[SOSBA_MAX_LEVELS], pc_v[SOSBA_MAX_LEVELS], pc_idepth[SOSBA_MAX_LEVELS], pc_color[SOSBA_MAX_LEVELS];
  std::vector<float> bw_idepth, bw_u, bw_v, bw_dx, bw_dy, bw_residual, bw_weight, bw_refColor;  // poseBufWarped_*
  int bw_n = 0;
  std::vector<float> loop_xyz, loop_color;   // PoseEstimator::pts (float-cast xyz [n*3], colour [n*levels])
  SE3 tfmF0ToF1;
  float fx1[SOSBA_MAX_LEVELS], fy1[SOSBA_MAX_LEVELS], cx1[SOSBA_MAX_LEVELS], cy1[SOSBA_MAX_LEVELS];
  std::vector<float> sw_rx1, sw_rx2, sw_rx3, sw_dx, sw_dy, sw_residual, sw_weight, sw_ref;  // scaleBufWarped_*
  int sw_n = 0;
};

}  // namespace orc
