/*
 * sosba.h — C ABI of the B200-native photometric bundle-adjustment / direct-alignment path.
 *
 * Drop-in boundary for the hot path of IRVLab/SOS-SLAM.  The reference has no
 * FFI: the "operator surface" is a set of C++ member functions that FullSystem calls directly.
 * Every entry point below names the reference member it replaces (file:line, relative to the
 * reference tree).  Plain pointers and sizes only; all buffers are caller-owned HOST memory
 * unless a name ends in `_dev`.  Every function returns 0 on success, a negative SOSBA_E_* code
 * otherwise; nothing throws, nothing aborts (reference convention: flags + non-finite sentinels,
 * SURVEY.md §8b "Error conventions").  One handle per FullSystem, single caller thread.
 *
 * Index conventions (verbatim from the reference):
 *   block index  = host + target*nf           (AccumulatedTopHessian.cpp:66, EnergyFunctional.cpp:82)
 *   precalc      = host*nf + target           (FrameHessian::targetPrecalc[target->idx])
 *   state order  = [tx ty tz rx ry rz a b]    (8 per frame), calib = [fx fy cx cy] (CPARS=4)
 *   D            = 4 + 8*nf, H row-major D x D double
 */
#ifndef SOSBA_H_
#define SOSBA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOSBA_CPARS 4
#define SOSBA_PATTERN 8            /* patternNum, settings.h:187 */
#define SOSBA_MAX_LEVELS 6         /* PYR_LEVELS, settings.h:34 */
#define SOSBA_PRECALC_FLOATS 32    /* packed FrameFramePrecalc, see sosba_precalc_layout */
#define SOSBA_J_FLOATS 74          /* RawResidualJacobian dump, member order of RawResidualJacobian.h:29-55 */

/* ResState, Residuals.h:43 */
enum { SOSBA_RES_IN = 0, SOSBA_RES_OOB = 1, SOSBA_RES_OUTLIER = 2 };

enum {
  SOSBA_OK = 0,
  SOSBA_E_ARG = -1,      /* bad argument / shape */
  SOSBA_E_CUDA = -2,     /* CUDA runtime error (message via sosba_last_error) */
  SOSBA_E_STATE = -3,    /* call order violated (reference asserts EFDeltaValid/EFIndicesValid) */
  SOSBA_E_NOGPU = -4,    /* no CUDA device: the product has no CPU fallback */
  SOSBA_E_NCCL = -5,
  SOSBA_E_NONFINITE = -6 /* a non-finite value reached the solver (reference: isLost) */
};

/* Offsets inside one packed FrameFramePrecalc record (HessianBlocks.h:109-134), row-major 3x3. */
enum sosba_precalc_layout {
  SOSBA_PC_RTLL0 = 0,   /* PRE_RTll_0   [9] */
  SOSBA_PC_TTLL0 = 9,   /* PRE_tTll_0   [3] */
  SOSBA_PC_KRKI = 12,   /* PRE_KRKiTll  [9] */
  SOSBA_PC_KT = 21,     /* PRE_KtTll    [3] */
  SOSBA_PC_AFF = 24,    /* PRE_aff_mode [2] */
  SOSBA_PC_B0 = 26,     /* PRE_b0_mode      */
  SOSBA_PC_DIST = 27    /* distanceLL       */
};

/* The setting_* globals the path reads (util/settings.cpp:28-204) + compile-time sizes. */
typedef struct sosba_config {
  int32_t w, h;               /* wG[0], hG[0] */
  int32_t pyr_levels;         /* 0 = derive like setGlobalCalib (globalCalib.cpp:39-49) */
  int32_t max_frames;         /* image slots to allocate (live KFs + tracked frame + cam1) */
  int32_t num_threads;        /* CPU oracle only: IndexThreadReduce workers; <=1 = the reference's nomt path */
  int32_t gamma_weights_pixel_select; /* setting_gammaWeightsPixelSelect */
  int32_t min_opt_iterations; /* setting_minOptIterations */
  int32_t reserved0;
  float huber_th;                 /* setting_huberTH = 9 */
  float outlier_th_sum_component; /* setting_outlierTHSumComponent = 2500 */
  float affine_opt_mode_a;        /* setting_affineOptModeA (mode 1 -> 0) */
  float affine_opt_mode_b;
  float coarse_cutoff_th;         /* setting_coarseCutoffTH = 20 */
  float idepth_fix_prior;         /* 2500 */
  float idepth_fix_prior_marg_fac;/* 360000 */
  float frame_energy_th_const_weight; /* 0.5 */
  float frame_energy_th_n;            /* 0.7 */
  float frame_energy_th_fac_median;   /* 1.5 */
  float overall_energy_th_weight;     /* 1 */
  float initial_calib_hessian;        /* 5e9 */
  float initial_rot_prior, initial_trans_prior, initial_aff_a_prior, initial_aff_b_prior;
  float marg_weight_fac;          /* 0.25 */
  float th_opt_iterations;        /* 1.2 */
} sosba_config;

/* Fills the defaults of settings.cpp after settingsDefault(preset=0, mode=1) (main.cpp:27-90). */
void sosba_config_default(sosba_config *cfg, int32_t w, int32_t h);

/* Window tables: what FullSystem::setPrecalcValues (FullSystem.cpp:1099-1107),
 * EnergyFunctional::setAdjointsF (EnergyFunctional.cpp:42-103) and setDeltaF (:163-194) produce. */
typedef struct sosba_window {
  int32_t nf;
  int32_t reserved0;
  const int32_t *frame_slot;       /* [nf] image slot holding frame i's pyramid */
  const float *precalc;            /* [nf*nf*SOSBA_PRECALC_FLOATS], host*nf+target */
  const double *adHost;            /* [nf*nf*64] host+target*nf, row-major 8x8 */
  const double *adTarget;          /* [nf*nf*64] */
  const float *adHTdeltaF;         /* [nf*nf*8]  host+target*nf */
  const float *frame_energy_th;    /* [nf] FrameHessian::frameEnergyTH */
  float calib[4];                  /* fxl fyl cxl cyl = CalibHessian::value_scaledf */
  float cDeltaF[4];                /* EnergyFunctional::cDeltaF */
  double cPrior[4];                /* EnergyFunctional::cPrior */
  const double *frame_prior;       /* [nf*8] EFFrame::prior */
  const double *frame_delta_prior; /* [nf*8] EFFrame::delta_prior */
  const double *frame_delta;       /* [nf*8] EFFrame::delta (state - state_zero) */
} sosba_window;

/* Active points (PointHessian + EFPoint), SoA.  SCALE_IDEPTH == 1 so idepth_scaled == idepth. */
typedef struct sosba_points {
  int32_t n;
  int32_t reserved0;
  const float *u, *v;              /* host pixel */
  const float *idepth;             /* PointHessian::idepth(_scaled) */
  const float *idepth_zero;        /* PointHessian::idepth_zero(_scaled) */
  const float *color;              /* [n*8] */
  const float *weights;            /* [n*8] */
  const int32_t *host;             /* [n] dense frame index (EFFrame::idx) */
  const float *priorF;             /* [n] EFPoint::priorF */
  const float *deltaF;             /* [n] EFPoint::deltaF */
} sosba_points;

/* PointFrameResidual + EFResidual bookkeeping, point-major (residuals of one point contiguous, in
 * EFPoint::residualsAll order; `point` non-decreasing).  Dense ids as after makeIDX()
 * (EnergyFunctional.cpp:1186-1202). */
typedef struct sosba_residuals {
  int32_t n;
  int32_t reserved0;
  const int32_t *point;            /* [n] */
  const int32_t *target;           /* [n] dense frame index */
  const uint8_t *state;            /* [n] state_state (ResState) */
  const uint8_t *is_linearized;    /* [n] EFResidual::isLinearized */
  const uint8_t *is_active;        /* [n] EFResidual::isActiveAndIsGoodNEW */
  const uint8_t *is_new;           /* [n] PointFrameResidual::isNew */
  const float *state_energy;       /* [n] or NULL (=0) */
} sosba_residuals;

typedef struct sosba_linearize_out {
  double energy;            /* stats[0]: sum of linearize() return values (FullSystemOptimize.cpp:49) */
  float new_frame_energy_th;/* frameEnergyTH of the newest frame after setNewFrameEnergyTH (:84-124) */
  int32_t n_in, n_oob, n_outlier; /* histogram of state_NewState over the linearized residuals */
  int32_t n_removed;        /* fixLinearization: residuals that turned inactive (toRemove, :148-179) */
  int32_t reserved0;
} sosba_linearize_out;

typedef struct sosba sosba_t;

/* ---- lifetime ------------------------------------------------------------------------------- */
int sosba_create(const sosba_config *cfg, int32_t device, sosba_t **out);
void sosba_destroy(sosba_t *h);
const char *sosba_last_error(void);
/* Launch everything on this cudaStream_t (0 = the handle's own stream). */
int sosba_set_stream(sosba_t *h, void *cuda_stream);
int sosba_synchronize(sosba_t *h);
/* Number of kernel launches issued by this handle since creation (bench `gpu_launches`). */
int64_t sosba_launch_count(const sosba_t *h);
/* CUDA-event bracket around every PointFrameResidual::linearize kernel launch, on its launch stream:
 * enable, run, then read the summed duration and the number of launches (the bench roofline).
 * on = 1 starts a new measurement, 0 pauses it, 2 resumes it without dropping what was recorded (a bracket breaks the
 * programmatic launch chain around the kernel, so the bench samples every few steps instead of all of them). */
int sosba_profile_enable(sosba_t *h, int32_t on);
int sosba_profile_read(sosba_t *h, double *ms_total, int32_t *launches);
/* Device-side timeline of the kernels of the Gauss-Newton loop (globaltimer stamps taken inside the kernels: first CTA
 * past its dependency wait .. last CTA done), i.e. the duration of a launch as it runs in the programmatic launch chain,
 * which a CUDA-event bracket cannot see (the event record breaks the chain and adds the launch latency).  Enable, run
 * sosba_ba_optimize / sosba_optimize, read the mean of the launches whose name starts with `kernel`
 * ("k_linearize", "k_accumulate_fused", "k_stitch_xchg", "k_solve", "k_step"); process-global; SOSBA_TRACE=1 prints
 * the whole timeline at sosba_destroy. */
int sosba_trace_enable(sosba_t *h, int32_t on);
int sosba_trace_read(sosba_t *h, const char *kernel, int32_t skip_first, double *mean_ns, int32_t *launches);
/* sosba_frame_make_images with the irradiance image (and B, or NULL) already in device memory. */
int sosba_frame_make_images_dev(sosba_t *h, int32_t slot, const float *color_dev, const float *B_dev);
int32_t sosba_pyr_levels(const sosba_t *h);

/* ---- a1: FrameHessian::makeImages (HessianBlocks.cpp:121-176) -------------------------------- */
/* color: w*h irradiance; B: 256-entry response (CalibHessian::B) or NULL (HCalib==0 branch).
 * Host buffers of 64 KB and more that live in page-locked memory (cudaHostAlloc / cudaHostRegister) are read by the DMA in place:
 * keep them unchanged until the next synchronising call (sosba_synchronize, any call that returns results).  Pageable buffers
 * are copied into the library's own pinned staging ring before the call returns. */
int sosba_frame_make_images(sosba_t *h, int32_t slot, const float *color, const float *B);
/* dIp[lvl] as Eigen::Vector3f AoS (w_l*h_l*3) and absSquaredGrad[lvl] (w_l*h_l); either may be NULL.
 * First/last row of dx,dy,absSquaredGrad are uninitialised in the reference; here they are 0. */
int sosba_frame_get_level(sosba_t *h, int32_t slot, int32_t lvl, float *dI3, float *abs_sq_grad);

/* ---- next row: the pre-pyramid image path ------------------------------------------------
 * Undistort::undistort<T> (util/Undistort.cpp:361-458) with PhotometricUndistorter::processFrame<T> (:194-227) fused in
 * front of makeImages: the caller hands over the RAW camera frame (8 or 16 bit), the device applies the inverse response
 * G and the inverse vignette, resamples through the rectification map and builds the pyramid.
 *   w_org, h_org   raw image size (Undistort::wOrg / hOrg); the output size is the handle's w x h
 *   remapX/remapY  [w*h] Undistort::remapX / remapY (source position per output pixel, remapX < 0 = no source: output 0);
 *                  both NULL = passthrough (requires w_org == w, h_org == h)
 *   G              [g_depth] PhotometricUndistorter::G (g_depth 256 or 65536), NULL = no photometric calibration
 *                  (the `!valid || setting_photometricCalibration == 0` branch: factor * raw)
 *   vignette_inv   [w_org*h_org] PhotometricUndistorter::vignetteMapInv, NULL = setting_photometricCalibration < 2
 * benchmark_varNoise / benchmark_varBlurNoise (dataset-corruption experiments, 0 by default) are not supported. */
int sosba_undistort_set(sosba_t *h, int32_t w_org, int32_t h_org, const float *remapX, const float *remapY, const float *G, int32_t g_depth,
                        const float *vignette_inv);
/* raw: w_org*h_org pixels of raw_bits (8 or 16) bits; factor: the `factor` argument of undistort (used when G is NULL);
 * B as in sosba_frame_make_images; image_out: optional [w*h] copy of the undistorted irradiance (ImageAndExposure::image),
 * NULL to keep it on the device. */
int sosba_frame_make_images_raw(sosba_t *h, int32_t slot, const void *raw, int32_t raw_bits, float factor, const float *B, float *image_out);

/* ---- window / point / residual upload ------------------------------------------------------- */
int sosba_window_set(sosba_t *h, const sosba_window *w);
int sosba_points_set(sosba_t *h, const sosba_points *p);
int sosba_residuals_set(sosba_t *h, const sosba_residuals *r);
/* Only the per-iteration part of a window (precalc, adHTdeltaF, cDeltaF, calib, delta, delta_prior,
 * frame_energy_th): FullSystem::setPrecalcValues + setDeltaF after doStepFromBackup. */
int sosba_window_update(sosba_t *h, const sosba_window *w);
/* PointHessian::setIdepth/setIdepthZero + EFPoint::deltaF for all points (doStepFromBackup). */
int sosba_points_update(sosba_t *h, const float *idepth, const float *idepth_zero, const float *deltaF);

/* ---- a3/a4/a5: linearize / applyRes --------------------------------------------------------- */
/* PointFrameResidual::resetOOB over all non-linearized residuals (FullSystemOptimize.cpp:316-329). */
int sosba_reset_oob(sosba_t *h);
/* FullSystem::linearizeAll(fixLinearization) (FullSystemOptimize.cpp:125-182) over activeResiduals
 * (= residuals with !is_linearized), including setNewFrameEnergyTH. */
int sosba_linearize_all(sosba_t *h, int32_t fix_linearization, sosba_linearize_out *out);
/* FullSystem::applyRes_Reductor(true) (FullSystemOptimize.cpp:79-83 -> Residuals.cpp:304-321). */
int sosba_apply_res(sosba_t *h);
/* EFResidual::fixLinearizationF for the listed residuals (EnergyFunctionalStructs.cpp:75-103). */
int sosba_fix_linearization(sosba_t *h, const int32_t *residual_ids, int32_t n);

/* per-residual read-back (any pointer may be NULL) */
int sosba_residuals_get_state(sosba_t *h, uint8_t *state, uint8_t *new_state, float *energy,
                              float *new_energy, float *new_energy_with_outlier, uint8_t *is_active,
                              uint8_t *is_linearized);
/* committed=1: EFResidual::J (after applyRes); 0: PointFrameResidual::J (candidate). [n*74] */
int sosba_residuals_get_jacobians(sosba_t *h, int32_t committed, float *J);
int sosba_residuals_get_aux(sosba_t *h, float *JpJdF /*[n*8]*/, float *res_to_zero /*[n*8]*/,
                            float *projected_to /*[n*16]*/, float *center_projected_to /*[n*3]*/);
/* fixLinearization bookkeeping (FullSystemOptimize.cpp:55-70): per point maxRelBaseline,
 * numGoodResiduals increments; per residual removal flag. */
int sosba_points_get_stats(sosba_t *h, float *max_rel_baseline, int32_t *num_good_residuals);

/* ---- a6-a9: accumulate + stitch ------------------------------------------------------------- */
/* EnergyFunctional::accumulateAF_MT / accumulateLF_MT / accumulateSCF_MT (EnergyFunctional.cpp:197-254)
 * HA,HL,Hsc: D*D row-major; bA,bL,bsc: D.  Any output may be NULL.  With a communicator attached the
 * result is the all-reduced sum over ranks (identical on every rank). */
int sosba_accumulate(sosba_t *h, double *HA, double *bA, double *HL, double *bL, double *Hsc,
                     double *bsc, int32_t *resInA, int32_t *resInL);
/* EFPoint accumulators after sosba_accumulate (any may be NULL): Hdd_accAF,bd_accAF,Hcd_accAF[4],
 * Hdd_accLF,bd_accLF,Hcd_accLF[4],HdiF,bdSumF,idepth_hessian */
int sosba_points_get_acc(sosba_t *h, float *HddA, float *bdA, float *HcdA, float *HddL, float *bdL,
                         float *HcdL, float *HdiF, float *bdSumF);

/* ---- a10/a11: EnergyFunctional::solveSystemF (EnergyFunctional.cpp:1029-1184), IMU off -------- */
/* accumulate A/L/SC, add marginalisation prior (HM,bM of dim D, may be NULL = 0), damp, solve,
 * resubstitute.  x: D doubles (= lastX).  H_final/b_final: optional copies of the solved system. */
int sosba_solve_system(sosba_t *h, const double *HM, const double *bM, double *x, double *H_final,
                       double *b_final);
/* EnergyFunctional::resubstituteF_MT (EnergyFunctional.cpp:496-551) with a caller-provided x. */
int sosba_resubstitute(sosba_t *h, const double *x, float *point_step /*[P] or NULL*/);

/* ---- marginalisation of points: EnergyFunctional::marginalizePointsF (:891-936) --------------- */
/* addPoint<2> + SC addPoint(false) over the listed points; H,b = M - Msc (not yet * margWeightFac). */
int sosba_marginalize_points(sosba_t *h, const int32_t *point_ids, int32_t n, double *H, double *b,
                             int32_t *resInM);

/* ---- a14/a15: CoarseTracker::calcResPose / calcGSSSEPose (CoarseTracker.cpp:612-764, 554-610) -- */
/* makeK (ScaleOptimizer.cpp:95-118): per-level intrinsics from level-0 fx,fy,cx,cy. */
int sosba_tracker_make_k(sosba_t *h, const float calib[4]);
/* pc_u/pc_v/pc_idepth/pc_color of one level (output of makeCoarseDepthL0, CoarseTracker.cpp:56-230) */
int sosba_tracker_set_ref(sosba_t *h, int32_t lvl, int32_t n, const float *pc_u, const float *pc_v,
                          const float *pc_idepth, const float *pc_color);
/* refToNew: row-major 3x4 [R|t] double; affLL = AffLight::fromToVecExposure(...) as float[2].
 * out6 = Vec6 {E, numTermsInE, flowT, 0, flowRT, numSaturated/numTermsInE};
 * counts = {numTermsInE, numTermsInWarped (unpadded), numSaturated}. */
int sosba_tracker_calc_res_pose(sosba_t *h, int32_t lvl, int32_t new_frame_slot,
                                const double refToNew[12], const float affLL[2], float cutoff_th,
                                double out6[6], int32_t counts[3]);
/* Uses the warped buffers of the last calc_res_pose at this level.  a = affLL[0], b0 =
 * lastRef_aff_g2l.b.  H: 8x8 row-major, b: 8 — already divided by n and SCALE_*-scaled. */
int sosba_tracker_calc_gs_pose(sosba_t *h, int32_t lvl, float a, float b0, double H[64], double b[8]);

/* ---- a17: ScaleOptimizer::calcResScale / calcGSSSEScale (ScaleOptimizer.cpp:273-437, 232-271) -- */
/* T10: tfmF0ToF1 row-major 3x4; K1: fx1,fy1,cx1,cy1 of camera 1 (level 0). */
int sosba_scale_set_stereo(sosba_t *h, const double T10[12], const float K1[4]);
int sosba_scale_calc_res(sosba_t *h, int32_t lvl, int32_t stereo_slot, float scale, float cutoff_th,
                         double out6[6], int32_t counts[3]);
int sosba_scale_calc_gs(sosba_t *h, int32_t lvl, float scale, float *H, float *b);

/* ---- a16: the direct-alignment control loops, resident on the device --------------------------------
 * CoarseTracker::makeCoarseDepthL0 (CoarseTracker.cpp:56-230) as called by setCoarseTrackingRef (:232-242): splat the centre
 * projections of the IN residuals on the newest keyframe (PointFrameResidual::centerProjectedTo {u, v, new_idepth}, weight
 * sqrt(1e-3 / (HdiF + 1e-12))), 2x2 sum-pool up the pyramid, dilate, normalise, and emit pc_u / pc_v / pc_idepth / pc_color per
 * level (raster order, 2 <= x < w-2).  ref_slot holds lastRef's pyramid.  The lists replace what sosba_tracker_set_ref uploads.
 * center_projected_to: [n*3], HdiF: [n] (EFPoint::HdiF), in the order FullSystem walks frameHessians / pointHessians.
 * pc_n_out: [levels] or NULL.  (Where the reference's dilation reads one element before / behind the map — index -1 at
 * i = w on levels 0 and 1, index w*h at i = w*h-w-1 — the neighbour counts as empty here.) */
int sosba_tracker_make_coarse_depth(sosba_t *h, int32_t ref_slot, int32_t n, const float *center_projected_to, const float *HdiF,
                                    int32_t *pc_n_out);
/* read back one level of the reference-frame point lists (any pointer may be NULL); returns pc_n[lvl] in *n */
int sosba_tracker_get_ref(sosba_t *h, int32_t lvl, int32_t *n, float *pc_u, float *pc_v, float *pc_idepth, float *pc_color);
/* CoarseTracker::scaleCoarseDepthL0 (CoarseTracker.cpp:244-263) on the resident lists */
int sosba_tracker_scale_coarse_depth(sosba_t *h, float scale);

#define SOSBA_TRACK_MAX_PASSES 8   /* pyramid levels + the one repeated level (CoarseTracker.cpp:516-519) */
/* One pose hypothesis of FullSystem::trackNewCoarse (FullSystem.cpp:175-231) = one call of
 * CoarseTracker::trackNewestCoarse (CoarseTracker.cpp:366-552).  Poses as Sophus::SE3d stores them: unit quaternion
 * (x, y, z, w) + translation. */
typedef struct sosba_track_hypothesis {
  double q[4], t[3];             /* in: lastToNew_out; out: refToNew_current when the loop ran through (else unchanged) */
  double aff_g2l[2];             /* in/out: aff_g2l_out (a, b) */
  double min_res_for_abort[5];   /* in: minResForAbort (NaN = never abort, as in the first try) */
  double last_residuals[5];      /* out: lastResiduals (NaN where a level was not reached) */
  double flow_indicators[3];     /* out: lastFlowIndicators of the last level that ran */
  int32_t ok;                    /* out: the return value */
  int32_t n_passes;              /* out: level passes executed (levels + 1 when a level was repeated) */
  int32_t pass_lvl[SOSBA_TRACK_MAX_PASSES];         /* out */
  int32_t pass_iterations[SOSBA_TRACK_MAX_PASSES];  /* out: LM iterations of the pass */
  uint64_t pass_accept[SOSBA_TRACK_MAX_PASSES];     /* out: bit i = iteration i accepted */
  uint64_t pass_tie[SOSBA_TRACK_MAX_PASSES];        /* out: bit i = the two mean energies of iteration i were closer than 2e-5 relative: the
                                                       decision was taken on energies re-added in point order like the reference's float sum */
  double pass_residual[SOSBA_TRACK_MAX_PASSES];     /* out: sqrt(resOld[0] / resOld[1]) at the end of the pass */
  float pass_cutoff_repeat[SOSBA_TRACK_MAX_PASSES]; /* out: levelCutoffRepeat of the pass */
} sosba_track_hypothesis;
/* n_hyp independent hypotheses against the frame in new_slot: the whole Levenberg-Marquardt loop of every hypothesis (calcResPose,
 * calcGSSSEPose, the damped 8x8 LDL^T solve, SE3::exp update, accept / reject, level schedule, abort and final checks) runs on
 * the device, one thread block per hypothesis, ONE synchronisation per call.  ref_ab_exposure / ref_aff_g2l: lastRef->ab_exposure,
 * lastRef_aff_g2l; new_ab_exposure: newFrame->ab_exposure.  Needs sosba_tracker_make_k and the reference lists
 * (sosba_tracker_make_coarse_depth or sosba_tracker_set_ref on every level <= coarsest_lvl). */
int sosba_tracker_track(sosba_t *h, int32_t new_slot, float ref_ab_exposure, float new_ab_exposure, const double ref_aff_g2l[2],
                        int32_t coarsest_lvl, int32_t n_hyp, sosba_track_hypothesis *hyps);

/* One start value of ScaleOptimizer::optimizeScale (ScaleOptimizer.cpp:120-230; FullSystem::optimizeScale tries 7 of them until
 * the scale is trapped, FullSystem.cpp:1133-1146). */
typedef struct sosba_scale_hypothesis {
  float scale;                   /* in/out */
  float error;                   /* out: the return value last_residuals[0] */
  double last_residuals[5];      /* out */
  int32_t n_passes;
  int32_t pass_lvl[SOSBA_TRACK_MAX_PASSES];
  int32_t pass_iterations[SOSBA_TRACK_MAX_PASSES];
  uint64_t pass_accept[SOSBA_TRACK_MAX_PASSES];
  uint64_t pass_tie[SOSBA_TRACK_MAX_PASSES];
  int32_t reserved0;
} sosba_scale_hypothesis;
/* camera-1 frame in stereo_slot; needs sosba_scale_set_stereo and the reference lists. */
int sosba_scale_optimize(sosba_t *h, int32_t stereo_slot, int32_t coarsest_lvl, int32_t n_hyp, sosba_scale_hypothesis *hyps);

/* ---- a4/a10/a11 composed: the Gauss-Newton loop body of FullSystem::optimize ------------------ */
/* Frame state as FullSystem keeps it (HessianBlocks.h:136-424). */
typedef struct sosba_frame_state {
  double camToWorld_evalPT[12]; /* row-major 3x4 */
  double state[10];             /* unscaled; [8],[9] unused */
  double state_zero[10];
  float ab_exposure;
  float frame_energy_th;
  int32_t frame_id;             /* FrameHessian::frameID (0 => gauge priors) */
  int32_t slot;
} sosba_frame_state;

typedef struct sosba_ba_problem {
  int32_t nf;
  int32_t reserved0;
  sosba_frame_state *frames;    /* [nf] in/out */
  double calib_value[4];        /* CalibHessian::value (unscaled), in/out */
  double calib_value_zero[4];
  sosba_points points;          /* in: initial idepth/idepth_zero; deltaF/priorF as given */
  sosba_residuals residuals;
  const double *HM;             /* [D*D] or NULL */
  const double *bM;             /* [D] or NULL */
  float *idepth_out;            /* [P] optimised idepth (may be NULL) */
} sosba_ba_problem;

typedef struct sosba_optimize_out {
  int32_t iterations;           /* GN iterations actually run */
  int32_t res_in_a;             /* ef->resInA of the last solve */
  double energy_initial;        /* linearizeAll(false) before the loop */
  double energy_final;          /* linearizeAll(true) */
  float rmse;                   /* sqrt(energy_final / (patternNum * resInA)) */
  int32_t n_removed;
  double last_x_norm;
  int32_t reserved0, reserved1;
} sosba_optimize_out;

/* FullSystem::optimize(mnumOptIts) (FullSystemOptimize.cpp:305-489) without IMU: resetOOB,
 * linearizeAll(false), applyRes, then per iteration backupState / solveSystem / doStepFromBackup /
 * linearizeAll(false) / applyRes / break test, then the new evalPT for the newest frame and
 * linearizeAll(true).  Host buffers in, host buffers out. */
int sosba_optimize(sosba_t *h, sosba_ba_problem *prob, int32_t max_iterations, sosba_optimize_out *out);

/* The same loop with the problem already resident on the device: uploads `prob` once ... */
int sosba_ba_upload(sosba_t *h, const sosba_ba_problem *prob);
/* ... FullSystem::optimize on the resident problem (sosba_optimize = ba_upload + ba_optimize + ba_download);
 * out->reserved0 = residuals linearised per linearizeAll pass. */
int sosba_ba_optimize(sosba_t *h, int32_t max_iterations, sosba_optimize_out *out);
/* ... and runs `n` loop bodies (solveSystem + doStepFromBackup + linearizeAll(false) + applyRes)
 * without touching host problem buffers.  n_res_linearized: residuals linearized per body. */
int sosba_ba_iterate(sosba_t *h, int32_t n, int32_t *n_res_linearized);
int sosba_ba_download(sosba_t *h, sosba_ba_problem *prob);

/* ---- the loop body split around a caller-side solve (EnergyFunctional::solveSystemF with IMU) ---------------------------------
 * With setting_enable_imu the reference widens the system between accumulation and solve: expandHbtoFitImu (8 -> 29 states per
 * frame + the scale), getImuHessian, the marginalisation prior in the widened space, KKT rows of the spline constraints and the
 * removal of unconstrained states (EnergyFunctional.cpp:1052-1140), then unpacks the step (:1150-1171).  That algebra works on the
 * FrameHessian spline state and stays with the caller ("IMU spline math host-side"); the two calls below are the
 * device part of the body on the window made resident by sosba_ba_upload:
 *   sosba_ba_system  accumulateAF_MT + accumulateLF_MT (priors included, as usePrior does) + accumulateSCF_MT and their stitches
 *                    (EnergyFunctional.cpp:1040-1047): H_top = HA + HL, b_top = bA + bL, H_sc, b_sc (D x D row-major, D = 4 + 8 nf);
 *                    all-reduced over point shards.  One synchronisation.
 *   sosba_ba_step    the rest of the body for the caller's x_dso (D doubles, = lastX): resubstituteF_MT, backupState,
 *                    doStepFromBackup of points / frames / calibration (stepsize 1), setPrecalcValues, linearizeAll(false), applyRes.
 *                    One synchronisation. */
typedef struct sosba_step_out {
  double energy;               /* linearizeAll(false) after the step */
  float new_frame_energy_th;
  int32_t n_in, n_oob, n_outlier;
  double sum_a, sum_b, sum_t, sum_r;   /* doStepFromBackup: mean squared steps of the frames (FullSystemOptimize.cpp:228-236) */
  double sum_id, sum_nid, num_id;      /* ... and of the points: sum step^2, sum |idepth_backup|, count (over all ranks) */
} sosba_step_out;
int sosba_ba_system(sosba_t *h, double *H_top, double *b_top, double *H_sc, double *b_sc, int32_t *resInA, int32_t *resInL);
int sosba_ba_step(sosba_t *h, const double *x, sosba_step_out *out);

/* ---- next row: immature points ------------------------------------------------ */
/* ImmaturePointStatus, ImmaturePoint.h:40-47 */
enum { SOSBA_IPS_GOOD = 0, SOSBA_IPS_OOB = 1, SOSBA_IPS_OUTLIER = 2, SOSBA_IPS_SKIPPED = 3, SOSBA_IPS_BADCONDITION = 4, SOSBA_IPS_UNINITIALIZED = 5 };

/* The members of ImmaturePoint (ImmaturePoint.h:50-92) that the constructor fills and traceOn reads / updates, SoA. */
typedef struct sosba_immature {
  int32_t n;
  int32_t reserved0;
  const int32_t *host;            /* [n] dense index of the host frame: selects KRKi / Kt / aff of traceNewCoarse */
  const float *u, *v;             /* [n] host pixel (integers stored as float, ImmaturePoint.cpp:30) */
  const float *color;             /* [n*8] */
  const float *weights;           /* [n*8] */
  const float *gradH;             /* [n*4] Mat22f, row-major (symmetric) */
  const float *energy_th;         /* [n] */
  float *idepth_min, *idepth_max; /* [n] in/out; a fresh point has 0 / NaN */
  float *quality;                 /* [n] in/out; a fresh point has 10000 */
  uint8_t *last_trace_status;     /* [n] in/out ImmaturePointStatus */
  float *last_trace_uv;           /* [n*2] out */
  float *last_trace_pixel_interval; /* [n] out */
} sosba_immature;

/* ImmaturePoint::ImmaturePoint (ImmaturePoint.cpp:28-60): color / weights / gradH / energyTH of n candidate pixels of the
 * frame in `host_slot` (getInterpolatedElement33BiLin, globalFuncs.h:162-182).  energy_th is NaN where a colour is not
 * finite (the reference then discards the point).  Outputs are host buffers. */
int sosba_immature_init(sosba_t *h, int32_t host_slot, int32_t n, const int32_t *u, const int32_t *v, float *color /*[n*8]*/,
                        float *weights /*[n*8]*/, float *gradH /*[n*4]*/, float *energy_th /*[n]*/);

/* FullSystem::traceNewCoarse (FullSystem.cpp:311-361): ImmaturePoint::traceOn (ImmaturePoint.cpp:70-415) of every point
 * against the frame in `frame_slot`.  KRKi [nhosts*9] row-major, Kt [nhosts*3], aff [nhosts*2]: hostToFrame_KRKi /
 * hostToFrame_Kt / hostToFrame_affine per host.  counts[6]: points per ImmaturePointStatus after the pass
 * (trace_good, _oob, _out, _skip, _badcondition, _uninitialized). */
int sosba_trace_immature(sosba_t *h, int32_t frame_slot, int32_t nhosts, const float *KRKi, const float *Kt, const float *aff,
                         sosba_immature *pts, int32_t counts[6]);

/* Resident variant: immature points live for many frames and only the traced-into frame changes, so the pool stays in HBM.
 * pool_set uploads every member of `pts` once (after makeNewTraces / whenever the host adds or deletes points),
 * pool_trace runs traceNewCoarse on it in place (per call: 14 floats per host up, 6 counters down), pool_get reads the
 * in/out members (idepth_min/max, quality, last_trace_status / uv / pixel_interval) back into `pts` when the host needs
 * them (activatePointsMT).  pts->n of pool_get must equal the pool size. */
int sosba_immature_pool_set(sosba_t *h, const sosba_immature *pts);
int sosba_immature_pool_trace(sosba_t *h, int32_t frame_slot, int32_t nhosts, const float *KRKi, const float *Kt, const float *aff, int32_t counts[6]);
int sosba_immature_pool_get(sosba_t *h, sosba_immature *pts);

/* The frame pairs of the window as FullSystem::activatePointsMT sees them (FullSystem.cpp:377-505): per (host, target)
 * FrameFramePrecalc::PRE_RTll / PRE_tTll / PRE_aff_mode (HessianBlocks.cpp:184-214), row (host * nf + target). */
typedef struct sosba_activation_window {
  int32_t nf;
  int32_t min_obs;                /* minObs of optimizeImmaturePoint; the reference passes 1 (FullSystem.cpp:371) */
  const int32_t *frame_slot;      /* [nf] image slot of frame i */
  const float *RTll;              /* [nf*nf*9] row-major 3x3 */
  const float *tTll;              /* [nf*nf*3] */
  const float *aff;               /* [nf*nf*2] */
  float calib[4];                 /* fxl fyl cxl cyl (HCalib scaled values) */
  int32_t reserved0;
  int32_t reserved1;
} sosba_activation_window;

enum { SOSBA_ACT_SKIP = 0, SOSBA_ACT_ACTIVATED = 1, SOSBA_ACT_DELETE = -1 };

/* FullSystem::activatePointsMT_Reductor -> optimizeImmaturePoint (FullSystemOptPoint.cpp:47-192) with
 * ImmaturePoint::linearizeResidual (ImmaturePoint.cpp:475-545) for every point of `pts` (reads host, u, v, color, weights,
 * energy_th, idepth_min, idepth_max).  Outputs (host buffers):
 *   result[n]        SOSBA_ACT_ACTIVATED (a PointHessian is created), SOSBA_ACT_SKIP (return 0: not well constrained, the
 *                    point stays immature), SOSBA_ACT_DELETE (return -1: outlier / non-finite, the point is deleted)
 *   idepth[n]        currentIdepth after the Gauss-Newton iterations (setIdepth / setIdepthZero of the new point)
 *   res_state[n*nf]  final state_state of the temporary residual towards frame t (SOSBA_RES_*); 255 at t == host.
 *                    The caller creates a PointFrameResidual for every SOSBA_RES_IN entry of an activated point. */
int sosba_optimize_immature(sosba_t *h, const sosba_activation_window *win, const sosba_immature *pts, int8_t *result, float *idepth,
                            uint8_t *res_state);

/* ---- next row: pixel selection ---------------------------------------------------------
 * PixelSelector (src/FullSystem/PixelSelector2.cpp).  The selector state of the reference object lives in the handle:
 * randomPattern (w*h bytes, `rand() & 0xFF` after srand(3141592), :36-39 -- the caller passes its own array so both sides
 * use the same numbers) and currentPotential (3 after construction, updated by every makeMaps). */
int sosba_pixel_selector_set(sosba_t *h, const uint8_t *random_pattern, int32_t current_potential);
/* PixelSelector::makeMaps(fh, map_out, density, recursionsLeft, plot = false, thFactor) (:146-282) with makeHists (:69-145)
 * and select (:284-422) on the pyramid in `slot`.  Outputs (host buffers, any may be NULL):
 *   n_selected          the return value numHaveSub
 *   u, v, type [cap]    the non-zero entries of the status map in raster order (the order makeNewTraces walks it,
 *                       FullSystem.cpp:1083-1097); type = 1, 2, 4 (the float stored in map_out = ImmaturePoint::my_type)
 *   map_out [w*h]       the full status map of the reference
 *   current_potential   currentPotential after the call
 * Returns SOSBA_E_ARG if cap is smaller than the number of selected pixels. */
int sosba_pixel_select(sosba_t *h, int32_t slot, float density, int32_t recursions_left, float th_factor, int32_t cap, int32_t *n_selected,
                       int32_t *u, int32_t *v, float *type, float *map_out, int32_t *current_potential);

/* CoarseDistanceMap::makeDistanceMap + growDistBFS (src/FullSystem/CoarseTracker.cpp:789-916): the active points of the other
 * keyframes are projected into level 1 of the newest keyframe (KRKi = K[1] R Ki[0], Kt = K[1] t per host frame, the caller's
 * FullSystem.cpp:417-424 expressions), those pixels get distance 0 and a breadth-first flood with alternating 8- / 4-neighbour
 * steps (k = 1 .. 39; sources on the image border do not spread) fills fwdWarpedIDDistFinal; unreached pixels keep 1000.
 *   host [n] index into KRKi / Kt;  u, v, idepth [n]: PointHessian::u, v, idepth_scaled;  dist_out [w1*h1] (level-1 size).
 * The sequential part of activatePointsMT (addIntoDistFinal after every accepted candidate, FullSystem.cpp:471-476) continues
 * on the caller's copy of the map. */
int sosba_distance_map(sosba_t *h, int32_t nhosts, const float *KRKi /*[nhosts*9]*/, const float *Kt /*[nhosts*3]*/, int32_t n, const int32_t *host,
                       const float *u, const float *v, const float *idepth, float *dist_out);

/* ---- next row: loop-closure direct alignment -----------------------------------------
 * PoseEstimator (src/LoopClosure/PoseEstimator.cpp): the LM loop of estimate() (:286-470) stays on the host and calls
 * these two per iteration.  K per level comes from sosba_tracker_make_k (PoseEstimator::makeK :53-73 uses the same
 * recursion as CoarseTracker::makeK). */
/* pts of the matched LoopFrame (LoopHandler.h:85): xyz [n*3] = pts[i].first (camera-frame 3D point, Vector3d),
 * color [n*levels] = pts[i].second[lvl], point-major. */
int sosba_loop_set_points(sosba_t *h, int32_t n, const double *xyz, const float *color);
/* PoseEstimator::calcRes (:147-284); affLL = AffLight::fromToVecExposure(refAbExposure, newFrame->ab_exposure, refAffGToL, aff_g2l).
 * out6 / counts as sosba_tracker_calc_res_pose. */
int sosba_loop_calc_res(sosba_t *h, int32_t lvl, int32_t slot, const double refToNew[12], const float affLL[2], float cutoffTH, double out6[6],
                        int32_t counts[3]);
/* PoseEstimator::calcGSSSE (:75-145); a = fromToVecExposure(...)[0], b0 = refAffGToL.b. */
int sosba_loop_calc_gs(sosba_t *h, int32_t lvl, float a, float b0, double H[64], double b[8]);

/* CoarseInitializer::Pnt (src/FullSystem/CoarseInitializer.h:32-65) of one pyramid level, SoA: what calcResAndGS reads
 * and writes. */
typedef struct sosba_init_points {
  int32_t n;
  int32_t reserved0;
  const float *u, *v;             /* [n] */
  const float *idepth_new;        /* [n] */
  const float *iR;                /* [n] */
  const float *energy;            /* [n*2] (photometric, regulariser) */
  const float *outlierTH;         /* [n] */
  const uint8_t *isGood;          /* [n] */
  float *energy_new;              /* [n*2] out */
  uint8_t *isGood_new;            /* [n] out */
  float *maxstep;                 /* [n] out */
  float *lastHessian_new;         /* [n] out (written for isGood_new points only, like the reference) */
  float *JbBuffer_new;            /* [n*10] out */
} sosba_init_points;

/* CoarseInitializer::calcResAndGS (src/FullSystem/CoarseInitializer.cpp:450-673) on level `lvl`: firstFrame in ref_slot,
 * newFrame in new_slot, K per level from sosba_tracker_make_k (CoarseInitializer::makeK uses the same recursion).
 * aff = (refToNew_aff.a, refToNew_aff.b); tlog = refToNew.log().head<3>(); alphaW / alphaK / couplingWeight as set in
 * setFirst (:236-239).  H, b, Hsc, bsc: 8x8 / 8 floats (row-major); res3 = (E.A, alphaEnergy, E.num).
 * lastHessian_new must be pre-filled by the caller with the points' current values (entries of rejected points are kept). */
int sosba_init_calc_res_and_gs(sosba_t *h, int32_t lvl, int32_t ref_slot, int32_t new_slot, const double refToNew[12], const float aff[2],
                               const float tlog[3], float alphaW, float alphaK, float couplingWeight, sosba_init_points *pts, float H[64], float b[8],
                               float Hsc[64], float bsc[8], float res3[3]);

/* ---- multi-GPU: points shard across ranks, one all-reduce of [H,b] per GN iteration ----------- */
/* 128-byte NCCL unique id (rank 0 creates, caller broadcasts, every rank inits). */
int sosba_comm_unique_id(uint8_t id[128]);
int sosba_comm_init(sosba_t *h, const uint8_t id[128], int32_t rank, int32_t world);
int sosba_comm_destroy(sosba_t *h);
/* 1 when the per-iteration exchange runs over peer memory (experimental, SOSBA_COMM_P2P=1 in the environment at
 * sosba_comm_init: every rank pushes its partial tables into the other ranks' HBM over NVLink, cudaIpc-mapped), 0 when it
 * is the NCCL all-reduce (the default: faster on 2 and 4 B200s). */
int sosba_comm_uses_peer_memory(const sosba_t *h);

#ifdef __cplusplus
}
#endif
#endif /* SOSBA_H_ */
