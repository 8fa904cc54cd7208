// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).  PARITY UNPINNED.
#pragma once
#include <vector>

#include "orc_core.h"

namespace orc {

// FrameHessian state the GN loop touches (HessianBlocks.h:136-424)
struct FrameH {
  SE3 evalPT, PRE_camToWorld, PRE_worldToCam;
  double state[10], state_zero[10], state_scaled[10], state_backup[10], step[10];
  double prior[8];
  float ab_exposure, frameEnergyTH;
  int frameID, slot;
  void setState(const double *s);
  void setEvalPT(const SE3 &e, const double *s);
  void getPrior(const sosba_config &cfg, double p[8]) const;
};

// CalibHessian (HessianBlocks.h:426-553)
struct CalibH {
  double value_zero[4], value_scaled[4], value[4], step[4], value_backup[4], value_minus_value_zero[4];
  float value_scaledf[4], value_scaledi[4];
  void setValue(const double *v);
};

void precalc_set(const FrameH &host, const FrameH &target, const CalibH &HCalib, Precalc &pc);
void set_adjoints(const std::vector<FrameH> &frames, std::vector<double> &adHost, std::vector<double> &adTarget);
void set_delta(const std::vector<FrameH> &frames, const std::vector<float> &adHostF, const std::vector<float> &adTargetF,
               std::vector<float> &adHTdeltaF);

struct BAState {
  std::vector<FrameH> frames;
  CalibH calib;
  void load(Oracle &o, const sosba_ba_problem *prob);
  void store(Oracle &o, sosba_ba_problem *prob) const;
  void setPrecalcValues(Oracle &o);
  void setAdjoints(Oracle &o);
  bool doStepFromBackup(Oracle &o, float stepfac, double *sums = nullptr);
  void step_with_x(Oracle &o, const double *x, sosba_linearize_out *lo, double sums[7]);
  void backupState(Oracle &o);
  bool iterate(Oracle &o, const double *HM, const double *bM, sosba_linearize_out *lo);
  void optimize(Oracle &o, const double *HM, const double *bM, int mnumOptIts, sosba_optimize_out *out);
};

// orc_ba.cpp
void make_images(Oracle &o, int slot, const float *color, const float *B);
double linearize(Oracle &o, Res &r);
void applyRes(Res &r, bool copyJacobians);
void fixLinearizationF(Oracle &o, Res &r);
void linearizeAll(Oracle &o, bool fixLinearization, sosba_linearize_out *out);
void accumulateAF(Oracle &o, double *H, double *b);
void accumulateLF(Oracle &o, double *H, double *b);
void accumulateSCF(Oracle &o, double *H, double *b);
void resubstituteF(Oracle &o, const double *x);
void solveSystemF(Oracle &o, const double *HM, const double *bM, double *x_out, double *Hfinal_out, double *bfinal_out);
void marginalizePoints(Oracle &o, const int32_t *ids, int n, double *H, double *b, int *resInM);

// orc_tracker.cpp
void tracker_makeK(Oracle &o, const float calib[4]);
// orc_trace.cpp
int pixel_select(Oracle &o, int slot, float density, int recursionsLeft, float thFactor, float *map_out);
void undistort_raw(const Oracle &o, const void *raw, int raw_bits, float factor, float *out);
void immature_init(Oracle &o, int slot, int n, const int32_t *u, const int32_t *v, float *color, float *weights, float *gradH, float *energyTH);
void trace_immature(Oracle &o, int frame_slot, int nhosts, const float *KRKi, const float *Kt, const float *aff, sosba_immature *pts, int32_t counts[6]);
void optimize_immature(Oracle &o, const sosba_activation_window *win, const sosba_immature *pts, int8_t *result, float *idepth, uint8_t *res_state);
void tracker_calcResPose(Oracle &o, int lvl, int slot, const double refToNew[12], const float affLL[2], float cutoffTH, double out6[6], int32_t counts[3]);
void loop_calcRes(Oracle &o, int lvl, int slot, const double refToNew[12], const float affLL[2], float cutoffTH, double out6[6], int32_t counts[3]);
void init_calcResAndGS(Oracle &o, int lvl, int ref_slot, int new_slot, const double refToNew[12], const float aff[2], const float tlog[3], float alphaW,
                       float alphaK, float couplingWeight, sosba_init_points *pts, float H[64], float b[8], float Hsc[64], float bsc[8], float res3[3]);
void tracker_calcGSSSEPose(Oracle &o, int lvl, float a, float b0, double H[64], double b[8]);
void scale_calcRes(Oracle &o, int lvl, int slot, float scale, float cutoffTH, double out6[6], int32_t counts[3]);
void scale_calcGSSSE(Oracle &o, int lvl, float scale, float *H, float *b);
// orc_lm.cpp
void tracker_makeCoarseDepth(Oracle &o, int ref_slot, int n, const float *center_projected_to, const float *HdiF);
void tracker_scaleCoarseDepth(Oracle &o, float scale);
bool tracker_track(Oracle &o, int new_slot, float ref_ab_exposure, float new_ab_exposure, const double ref_aff_g2l[2], int coarsestLvl, sosba_track_hypothesis *hy);
void scale_optimize(Oracle &o, int stereo_slot, int coarsestLvl, sosba_scale_hypothesis *hy);
void distance_map(Oracle &o, int nhosts, const float *KRKi, const float *Kt, int n, const int32_t *host, const float *u, const float *v, const float *idepth,
                  float *dist);

}  // namespace orc
