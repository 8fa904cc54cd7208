// sosba_internal.h — device data layout of the B200-native photometric BA / direct-alignment path.
//
// Everything here is private to libsosba.so; the public boundary is include/sosba.h.
// Layout in HBM:
//   images     float4 {I, dx, dy, absSquaredGrad} row-major per pyramid level, one arena per slot
//              (one 128-bit load per texel; a bilinear tap is 2 x 32 B row segments)
//   points     SoA (u, v, idepth, idepth_zero, color[8], weights[8], host, prior, delta, accumulators)
//   residuals  point-major SoA of flags/energies + AoS Jacobian records:
//                J[2]  80-float RawResidualJacobian records (candidate / committed, `sel` bit per residual)
//                rec   48-float "commit record": everything the block accumulators consume
//   window     precalc[nf*nf*32], adjoints (f32 + f64), adHTdelta, thresholds — read through L1
//   acc        fp64 block accumulators (top: nf*nf x 92, Schur: (D+1)^2), fp64 H/b, fp64 x
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/sosba.h"

#define SOSBA_JREC 80   // floats per RawResidualJacobian record (74 used, 16-byte aligned stride)
#define SOSBA_CREC 48   // floats per commit record
#define SOSBA_TOPB 92   // doubles per (host,target) top block: 91 upper-tri entries of 13x13 + count

// ---- J record (float offsets) --------------------------------------------------------------------
// RawResidualJacobian.h:29-55, regrouped so the 8 pattern lanes of one residual write 32-byte runs.
enum {
  JR_RES = 0,      // resF[8]
  JR_JIDX0 = 8,    // JIdx[0][8]
  JR_JIDX1 = 16,   // JIdx[1][8]
  JR_JAB0 = 24,    // JabF[0][8]
  JR_JAB1 = 32,    // JabF[1][8]
  JR_GEO = 40,     // 34 floats that do not depend on the pattern index:
  JR_JPDXI0 = 40,  //   Jpdxi[0][6]
  JR_JPDXI1 = 46,  //   Jpdxi[1][6]
  JR_JPDC0 = 52,   //   Jpdc[0][4]
  JR_JPDC1 = 56,   //   Jpdc[1][4]
  JR_JPDD = 60,    //   Jpdd[2]
  JR_JIDX2 = 62,   //   JIdx2 (00,01,10,11)
  JR_JABJIDX = 66, //   JabJIdx (00,01,10,11)
  JR_JAB2 = 70     //   Jab2 (00,01,10,11)
};

// ---- commit record (float offsets): what AccumulatedTopHessianSSE::addPoint / SC addPoint read ----
// Written when a linearisation is committed (applyRes -> EFResidual::takeDataF) for active residuals,
// or by the linearised / marginalisation preparation pass (addPoint<1>, addPoint<2>).
enum {
  CR_X = 0,       // x[10] = (Jpdc[0][4], Jpdxi[0][6])
  CR_Y = 10,      // y[10] = (Jpdc[1][4], Jpdxi[1][6])
  CR_A = 20,      // JIdx2(0,0), (0,1), (1,1)
  CR_TR = 23,     // JabJIdx 00,01,10,11, JI_r[0], JI_r[1]        (updateTopRight arguments)
  CR_BR = 29,     // Jab2 00, 01, Jab_r[0], Jab2 11, Jab_r[1], rr  (updateBotRight arguments)
  CR_JPDD = 35,   // Jpdd[2]            (37..39 spare)
  CR_JPJDF = 40   // JpJdF[8]
};

struct DevLevel {
  int w, h;
};

struct sosba {
  sosba_config cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  int64_t launches = 0;
  int sm_count = 148;

  // pyramid geometry (globalCalib.cpp:39-97)
  int levels = 0;
  int wl[SOSBA_MAX_LEVELS], hl[SOSBA_MAX_LEVELS];
  size_t lvl_off[SOSBA_MAX_LEVELS + 1];  // texel offset of each level inside a slot arena
  std::vector<float4 *> slot_img;        // [max_frames] arena base (device)
  std::vector<float *> slot_plane;       // [max_frames] planar channel-0 copy (device), same offsets
  std::vector<char> slot_valid;
  float *d_stage = nullptr;              // w*h input irradiance
  float *d_B = nullptr;                  // 256-entry response
  float *h_pinned = nullptr;             // pinned staging (w*h*4 floats)
  size_t h_pinned_bytes = 0;

  // window tables
  int nf = 0;
  std::vector<int> frame_slot;
  float *d_precalc = nullptr, *d_adHostF = nullptr, *d_adTargetF = nullptr, *d_adHTdeltaF = nullptr, *d_frameEnergyTH = nullptr;
  double *d_adHost = nullptr, *d_adTarget = nullptr;
  float *d_calib = nullptr;              // fxl fyl cxl cyl fxli fyli | cDeltaF[4]  (10 floats)
  double *d_wprior = nullptr;            // cPrior[4] | frame_prior[nf*8] | frame_delta_prior[nf*8] | frame_delta[nf*8]
  const float4 **d_img0 = nullptr;       // [nf] level-0 image of each window frame
  int nf_alloc = 0;
  float h_calib[10];
  std::vector<double> h_wprior;
  std::vector<float> h_frameEnergyTH;
  std::vector<double> h_adHost, h_adTarget;
  std::vector<float> h_adHTdeltaF;

  // points
  int P = 0, P_alloc = 0;
  float *p_u = nullptr, *p_v = nullptr, *p_idepth = nullptr, *p_idepth_zero = nullptr, *p_color = nullptr, *p_weights = nullptr;
  float *p_priorF = nullptr, *p_deltaF = nullptr;
  int *p_host = nullptr, *p_res_begin = nullptr;  // CSR [P+1]
  float *p_HddA = nullptr, *p_bdA = nullptr, *p_HcdA = nullptr, *p_HddL = nullptr, *p_bdL = nullptr, *p_HcdL = nullptr;
  float *p_HdiF = nullptr, *p_bdSumF = nullptr, *p_step = nullptr, *p_idepth_backup = nullptr, *p_idepth_hessian = nullptr;
  float *p_maxRelBaseline = nullptr;
  int *p_numGood = nullptr;

  // residuals
  int R = 0, R_alloc = 0;
  int *r_point = nullptr, *r_target = nullptr, *r_host = nullptr, *r_by_block = nullptr;
  uint8_t *r_state = nullptr, *r_new_state = nullptr, *r_is_lin = nullptr, *r_is_active = nullptr, *r_is_new = nullptr, *r_sel = nullptr,
          *r_dropped = nullptr;
  float *r_energy = nullptr, *r_new_energy = nullptr, *r_new_energy_wo = nullptr;
  float *r_J[2] = {nullptr, nullptr};
  float *r_rec = nullptr, *r_rtz = nullptr, *r_proj = nullptr, *r_center = nullptr;

  // reduction scratch / accumulators
  double *d_stats = nullptr;     // [16]: 0 energy, (ints live in d_counts)
  int *d_counts = nullptr;       // [16]: 0 n_in,1 n_oob,2 n_outlier,3 n_removed,4 n_newframe_energies,5 resIn (top), ...
  float *d_newE = nullptr;       // newest-frame energies (compaction target) [P_alloc... R_alloc]
  float *d_thOut = nullptr;      // [1] new threshold
  double *d_accTop = nullptr;    // [nf*nf*92]
  double *d_accSC = nullptr;     // [(D+1)*(D+1)]
  double *d_H = nullptr;         // [3][D*D + D] : A, L, SC
  double *d_x = nullptr;         // [D]
  float *d_xAd = nullptr;        // [nf*nf*8] + xc[4]
  int D_alloc = 0;

  // tracker / scale optimizer
  float t_K[SOSBA_MAX_LEVELS][4];    // fx fy cx cy per level (ScaleOptimizer::makeK)
  float t_K1[SOSBA_MAX_LEVELS][4];   // camera-1 intrinsics per level
  double t_T10[12];
  bool t_haveK = false, t_haveStereo = false;
  float *t_pc[SOSBA_MAX_LEVELS] = {nullptr};  // u|v|idepth|color, 4*n floats
  int t_n[SOSBA_MAX_LEVELS] = {0}, t_cap[SOSBA_MAX_LEVELS] = {0};
  float *t_warp = nullptr;   // 8 SoA arrays of cap floats (poseBufWarped_* / scaleBufWarped_*)
  int t_warp_cap = 0;
  int t_warp_n[SOSBA_MAX_LEVELS] = {0};  // padded count of the last calcRes per level
  int t_warp_lvl = -1, t_warp_kind = 0;
  double *t_acc = nullptr;   // [64] tracker sums

  // direct-alignment control loops on the device (k_lm.cu): makeCoarseDepthL0 maps and scratch, hypotheses
  float *cd_idepth = nullptr, *cd_wsum = nullptr, *cd_wbak = nullptr;   // per-level maps in one arena (level offsets = lvl_off)
  int *cd_cnt = nullptr;          // level 0: hits per pixel | lowest point index
  uint8_t *cd_mark = nullptr;     // pixels that enter the point list
  int2 *cd_list = nullptr;        // raster-order list per level
  int *cd_scan = nullptr;         // compaction scratch
  void *lm_hyp = nullptr;         // device copy of the hypotheses of one call
  size_t lm_hyp_cap = 0;
  float *lm_terms = nullptr;      // per hypothesis: energy terms of the accepted state and of the candidate
  size_t lm_terms_cap = 0;
  float *dm_buf = nullptr;        // CoarseDistanceMap: staged inputs | level-1 map | seed bytes
  size_t dm_cap = 0;
  std::vector<void *> lm_allocs;

  // composed GN loop (host mirror of FullSystem state)
  struct BA *ba = nullptr;

  // multi-GPU
  void *comm = nullptr;
  int rank = 0, world = 1;
  // what rides on the per-iteration all-reduce besides the block tables (set by sosba_api.cu)
  double *d_rstats_all = nullptr;   // [8] back-substitution sums, both loop-body parities
  int *d_cnt_all = nullptr;         // [2] resInA, resInL
  float *d_newE_all = nullptr;      // [world][newE_cap] newest-frame energies, one segment per rank ...
  int *d_newE_cnt = nullptr;        // ... [world] lengths, stored right behind the segments
  int newE_cap = 0;
  // peer-memory exchange (k_xchg.cu, comm.cu): every rank pushes its stitched values into the other ranks' mailboxes over NVLink
  unsigned char *p2p_mbox = nullptr;          // this rank's mailbox: [2 parities][world slots of {payload, exchange number} words]
  unsigned char *p2p_peer[8] = {nullptr};     // the mailboxes of all ranks as mapped into this process (p2p_peer[rank] = own)
  size_t p2p_slot_bytes = 0;
  int *p2p_epoch_dev = nullptr;               // device: [0] exchange number (advanced by the exchange kernel itself), [1] its CTA ticket
  bool p2p = false;
  int *d_comm_int = nullptr;                  // scratch of sosba_comm_max_int
};

void sosba_set_error(const char *fmt, ...);

#define SOSBA_CUDA(expr)                                                                 \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      sosba_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return SOSBA_E_CUDA;                                                               \
    }                                                                                    \
  } while (0)
