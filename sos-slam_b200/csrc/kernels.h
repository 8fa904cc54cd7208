// kernels.h — launchers of the sm_100a kernels (k_exact.cu: built with -fmad=false so per-residual float
// expressions round exactly like the parity definition; k_accum.cu / k_solve.cu / k_tracker.cu).
#pragma once
#include <utility>

#include "sosba_internal.h"

// Programmatic dependent launch (PDL) for the kernels of the Gauss-Newton loop body: the next launch may be scheduled
// while the current grid drains; every such kernel starts with PDL_ENTER() — let ITS dependents launch early, then wait
// until the grid it depends on has completed and its memory is visible.  Without the launch attribute both are no-ops.
#define PDL_ENTER() do { asm volatile("griddepcontrol.launch_dependents;"); asm volatile("griddepcontrol.wait;" ::: "memory"); } while (0)
// SOSBA_TRACE=1: device-side timeline of the launches of one sosba_ba_optimize (globaltimer, ns).  Per launch 4 words:
// [0] first CTA entered, [1] first CTA past the dependency wait, [2] last CTA done.  Null pointer = off (the default).
__device__ __forceinline__ unsigned long long trace_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define PDL_ENTER_T(tr) do { if ((tr) && threadIdx.x == 0) atomicMin((unsigned long long *)(tr), trace_now()); PDL_ENTER(); \
                             if ((tr) && threadIdx.x == 0) atomicMin((unsigned long long *)(tr) + 1, trace_now()); } while (0)
#define TRACE_EXIT(tr) do { if ((tr) && threadIdx.x == 0) atomicMax((unsigned long long *)(tr) + 2, trace_now()); } while (0)
long long *sosba_trace_slot(const char *name);   // sosba_api.cu: next record of the timeline, or nullptr when tracing is off

template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ---- k_exact.cu ---------------------------------------------------------------------------------
// setNewFrameEnergyTH: k-th smallest of the newest-frame energies -> frameEnergyTH[nf-1], thOut[0].
// The list is `nseg` segments of `seg_stride` floats with `seg_counts[s]` valid entries each: one segment on a single GPU
// (seg_counts = &counts[4]); with point shards one segment per rank, gathered by the per-iteration all-reduce (a sum
// with zeros elsewhere is an exact concatenation).
struct ThArgs {
  const float *newE;
  const int *seg_counts;
  int nseg, seg_stride;
  float *frameEnergyTH;
  int nf;
  float thN, thFacMedian, thConstWeight, overallWeight;
  float *thOut;
};
void launch_energy_th(sosba *h, const ThArgs &a, const int *gate = nullptr);

struct LinArgs {
  int R, nf;
  const int *r_point, *r_target, *r_host;
  uint8_t *r_state, *r_new_state, *r_is_lin, *r_is_active, *r_is_new, *r_sel, *r_dropped;
  float *r_energy, *r_new_energy, *r_new_energy_wo;
  float *J0, *J1, *rec, *rtz, *proj, *center;
  const float *p_u, *p_v, *p_idepth, *p_idepth_zero, *p_color, *p_weights;
  const float *p_deltaF;
  float *p_maxRelBaseline;
  int *p_numGood;
  const float *precalc, *frameEnergyTH, *calib, *adHTdeltaF;  // calib: fxl fyl cxl cyl fxli fyli | cDeltaF[4]
  const float4 *img[16];   // level-0 image of each window frame (kernel parameter space: no pointer-table load)
  int w;              // level-0 width
  float wM3G, hM3G;   // globalCalib.cpp:63-64
  float huberTH, outlierTHSum, affModeA, affModeB;
  double *stats;      // [0] energy
  int *counts;        // 0 in, 1 oob, 2 outlier, 3 removed, 4 n newest-frame energies
  float *newE;        // newest-frame energies of this rank (unordered)
  int *newE_count;    // its length (atomic append)
  // fused variants (launch_linearize_apply)
  ThArgs th;          // setNewFrameEnergyTH in the last CTA
  int *ticket;        // CTA completion counter (self-resetting)
  const int *gate;    // non-null: skip the launch when *gate != 0 (the loop broke on the device)
  double *zero_buf;   // non-null: zero_n double2 to clear (block tables of the next accumulation)
  int zero_n;
  long long *trace;       // SOSBA_TRACE record of this launch, or null
  unsigned opaque_zero;   // always 0; only known at run time (k_linearize_t chains its tap consumers behind the last tap with it)
};

void launch_make_images(sosba *h, int slot, const float *d_color, const float *d_B);
void launch_linearize(sosba *h, const LinArgs &a);
void launch_linearize_fix(sosba *h, const LinArgs &a);   // linearizeAll(true): + applyRes + removal / baseline bookkeeping
// th_inline: the last CTA runs setNewFrameEnergyTH; otherwise the caller schedules it (spare CTA of the next accumulation)
void launch_linearize_apply(sosba *h, const LinArgs &a, bool write_j, bool th_inline);
void launch_apply_res(sosba *h, const LinArgs &a, int fix);
// zero_words != nullptr: the first block also clears n_zero 4-byte words there (linearisation sums) and 4 ints at zero_ctl
// (loop control) - the memsets that would otherwise sit on the stream in front of the first linearisation
void launch_reset_oob(sosba *h, const LinArgs &a, int *zero_words = nullptr, int n_zero = 0, int *zero_ctl = nullptr);
void launch_residual_init(sosba *h, const LinArgs &a, const int *p_host);   // derived residual members after an upload (sosba_residuals_set)
void launch_fix_linearization(sosba *h, const LinArgs &a, const int *d_ids, int n);
// mode 1: linearised residuals (resApprox = res_toZeroF + J*delta), mode 2: marginalisation (res_toZeroF);
// list==nullptr -> all residuals.  Rewrites the commit record of every selected residual.
void launch_prep_records(sosba *h, const LinArgs &a, int mode, const int *d_list, int n);

// frames + calibration part of a Gauss-Newton step, device resident
#define SOSBA_FS 64   // doubles per frame: evalPT[12] state[10]@12 state_zero[10]@22 state_backup[10]@32 step[10]@42 ab_exposure@52
struct StepArgs {
  int nf;
  float stepfac;
  const double *x;
  double *fs, *cs;       // frame states [nf*SOSBA_FS]; calib value[4] | value_zero[4] | value_backup[4] | step[4]
  float *precalc, *adHTdeltaF, *calib;
  const float *adHostF, *adTargetF;
  double *wprior;
  double *iter;          // out: sumA, sumB, sumT, sumR (already / nf)
  double *adHost, *adTarget;   // fp64 adjoints, rewritten by the retarget variant only
  long long *trace;            // SOSBA_TRACE: 4 clock64 stamps of the frame step (start, states staged, frames done, pairs done), or null
};
// th != nullptr: a second CTA runs the pending setNewFrameEnergyTH selection beside the retarget and then clears the n_zero
// 4-byte words at zero_words (the sums of the linearisation that follows); th == nullptr: CTA 0 clears them
void launch_frame_retarget(sosba *h, const StepArgs &a, const ThArgs *th = nullptr, int *zero_words = nullptr, int n_zero = 0);

// tracker / scale optimizer (calcResPose / calcResScale): writes the 8 warped SoA arrays (masked, not
// compacted: invalid points carry weight 0) and the sums
struct TrackResArgs {
  int n, lvl, w, h, cap;
  const float *pc;        // u | v | idepth  (3 arrays of n); loop closure (kind 2): x | y | z of the 3D point
  const float *color;     // reference colour of the level
  const float4 *img;      // level image of the new frame
  float RKi[9], Ki[9], t[3];
  float fx, fy, cx, cy;
  float aff0, aff1;
  float huberTH, cutoffTH, maxEnergy;
  float scale;            // scale variant only
  int kind;               // 0 pose, 1 scale, 2 loop-closure pose (RKi holds R)
  float *warp;            // 8 arrays of cap floats
  double *acc;            // [0] E [1] shiftT [2] shiftRT [3] shiftNum ; ints in icnt
  int *icnt;              // [0] numTermsInE [1] numTermsInWarped [2] numSaturated
};
void launch_track_res(sosba *h, const TrackResArgs &a);

// ---- k_accum.cu ---------------------------------------------------------------------------------
struct AccArgs {
  int R, P, nf, n_list;
  const int *list;   // residual ids ordered by (host + target*nf)
  int mode;          // 0 active, 1 linearised, 2 marginalisation (list = residuals of the chosen points)
  const int *r_point, *r_target, *r_host;
  const uint8_t *r_is_lin, *r_is_active, *r_dropped;
  const float *rec;
  double *accTop;    // [nf*nf*92]
  int *n_acc;        // += residuals accumulated (resInA / resInL / resInM)
};
void launch_top_accumulate(sosba *h, const AccArgs &a);

// per-point sums (A: non-linearised, L: linearised; mode 2: every active residual -> L, A = 0) + Schur term
struct SCArgs {
  int P, nf, D;
  const int *plist;  // nullptr = all points, else the points to process (marginalisation)
  int n_plist;
  int mode;          // 0: accumulateAF/LF/SCF  2: marginalizePointsF
  int shiftPriorToZero;
  const int *res_begin, *r_target, *p_host;
  const uint8_t *r_is_lin, *r_is_active, *r_dropped;
  const float *rec;
  float *HddA, *bdA, *HcdA, *HddL, *bdL, *HcdL;
  const float *priorF, *deltaF;
  float *HdiF, *bdSumF, *idepth_hessian, *maxRelBaseline;
  const float *adHostF, *adTargetF;
  double *accSC;     // [(D+1)*(D+1)] upper triangle
};
void launch_point_sc(sosba *h, const SCArgs &a);

// accumulateAF_MT + accumulateLF_MT + accumulateSCF_MT of all points in one launch (k_accumulate_fused)
struct FusedAccArgs {
  int P, nf, D, R;
  int shiftPriorToZero;
  int do_th;           // the spare CTA runs the pending setNewFrameEnergyTH
  const int4 *tiles;   // per tile: first point, points (<= 32, one host), first residual, residuals
  const int *res_begin, *r_target, *p_host;
  const uint8_t *r_is_lin, *r_is_active, *r_dropped;
  const float *rec;
  double *accTop;      // [2][nf*nf*92]: A | L
  int *n_acc;          // [0] resInA [1] resInL
  float *HddA, *bdA, *HcdA, *HddL, *bdL, *HcdL;
  const float *priorF, *deltaF;
  float *HdiF, *bdSumF, *idepth_hessian, *maxRelBaseline;
  const float *adHostF, *adTargetF;
  double *accSC;       // [(D+1)*(D+1)] upper triangle
  ThArgs th;
  const int *gate;
  long long *dbg;      // optional phase timestamps (SOSBA_SOLVE_DEBUG)
  long long *trace;    // SOSBA_TRACE record of this launch, or null
};
bool launch_accumulate_fused(sosba *h, const FusedAccArgs &a, int max_res_per_tile, int tiles_total);

// top blocks -> H (D*D), b (D): AccumulatedTopHessianSSE::stitchDoubleInternal + stitchDoubleMT epilogue
void launch_stitch_top(sosba *h, const double *accTop, const double *adHost, const double *adTarget, int nf, double *H, double *b,
                       int usePrior, const double *wprior, const float *cDeltaF);

// ---- k_xchg.cu: gather-form stitch of the top blocks (A + L tables) into the final symmetric H, b, fused with the
// point-shard exchange over peer memory (push = 1): H, b, the Schur Gram matrix, the back-substitution sums, the residual
// counters are summed over ranks in rank order, the newest-frame energies concatenated
struct StitchXchgArgs {
  int nf, D;
  const double *accTop;        // [2][nf*nf*92]: A | L
  const double *adHost, *adTarget;
  double *H, *b;               // out: D*D (both triangles), D
  double *accSC;               // (D+1)^2 Schur Gram matrix, upper triangle; summed in place when push
  double *rstats;              // [8] back-substitution sums of both loop-body parities; summed in place when push
  int *cnt;                    // [2] resInA, resInL; summed in place when push
  int rank, world, push;
  unsigned char *peer[8];      // mailboxes of all ranks as mapped into this process
  size_t slot_bytes;
  int *epoch;                  // device: [0] exchange number (starts at 1), [1] CTA ticket
  float *newE_all;             // [world][newE_cap] newest-frame energies, own segment filled by the linearisation
  int *newE_cnt;               // [world]
  int newE_cap, with_newE;
  const int *gate;             // non-null: skip when *gate != 0 (the loop broke on the device -- on every rank alike)
  int *err;                    // set to 2 when a peer's words did not arrive in time
  unsigned backoff_ns;         // nanosleep between polls of a word that has not arrived (0 = spin)
  long long *dbg;              // optional: globaltimer stamps of one launch (SOSBA_XCHG_DEBUG)
  long long *trace;            // SOSBA_TRACE record of this launch, or null
};
#define SOSBA_XCHG_MAX_NF 13        // k_solve's limit
#define SOSBA_XCHG_MAX_NEWE 16384   // newest-frame energies per rank a mailbox slot can carry
size_t stitch_xchg_slot_bytes(int nf_max, int newE_cap);
int launch_stitch_xchg(sosba *h, const StitchXchgArgs &a, int local_points);
void launch_lin_xchg(sosba *h, const StitchXchgArgs &a, double *stats, int *counts, int with_stats, int local_points, double *extra = nullptr, int n_extra = 0);
// accSC -> Hsc (D*D), bsc (D)
void launch_finalize_sc(sosba *h, const double *accSC, int nf, double *H, double *b);
void launch_add_priors(sosba *h, int nf, double *H, double *b, const double *wprior, const float *cDeltaF);

// ---- k_solve.cu ---------------------------------------------------------------------------------
struct SolveArgs {
  int nf, D;
  const double *Htop, *btop;   // stitched A + L top blocks, symmetric (k_stitch_xchg output)
  const double *accSC;         // (D+1)^2 Gram matrix of the Schur term, upper tiles
  const double *HM, *bM;       // may be null
  const double *wprior;        // cPrior[4] | frame_prior | frame_delta_prior | frame_delta
  const float *cDeltaF;
  double *x, *Hfinal, *bfinal; // out (Hfinal/bfinal may be null)
  const float *adHostF, *adTargetF;
  float *xAd;                  // [nf*nf*8] then xc[4]
  int *status;                 // [0] non-finite flag
  long long *dbg;              // optional: clock64() at the phase boundaries (SOSBA_SOLVE_DEBUG)
  int do_step;                 // also run the frames / calibration part of doStepFromBackup (FullSystemOptimize.cpp:185-257)
  StepArgs step;
  int stage_sc, stage_hm;      // set by launch_solve: accSC / HM staged in shared memory
  // device-side loop control of FullSystem::optimize (FullSystemOptimize.cpp:358-413): no host round trip per iteration
  int *ctl;                    // null = always run; [0] latch "loop broke", [1] iterations run
  int iter_index, min_it;      // this is loop body `iter_index`; break after body i if canbreak(i) && i >= min_it
  float th_opt;                // setting_thOptIterations
  const double *prev_rstats;   // back-substitution sums of body iter_index-1: [0] sum step^2 [1] sum |idepth| [2] count
  const int *res_in;           // resInA of the accumulation this solve consumes ...
  int *res_out;                // ... copied where the table clearing of the next linearisation does not reach
  double *zero_rstats;         // non-null: clear the 4 back-substitution sums the following k_resubstitute accumulates into
  const double *stash_src;     // non-null (first body of a loop): the sums of the first linearisation ...
  double *stash_dst;           // ... are kept here (12 doubles) before the step launch clears them
  // point shards: setNewFrameEnergyTH of the previous linearisation (its energies arrived with the all-reduce in front of
  // this launch) runs in a second CTA beside the solve instead of as a launch of its own
  ThArgs th;
  int do_th;
  int smem_words;              // set by launch_solve: 4-byte words of dynamic shared memory (scratch of the spare CTA)
  long long *trace;            // SOSBA_TRACE record of this launch, or null
  int block_pivots;            // set by launch_solve: 1 = 4x4 block pivots (block LDL^T, the default), 0 = scalar pivots (SOSBA_SOLVE_PIVOTS=scalar)
  int pipe;                    // set by launch_solve: 1 = pipelined factorisation (panel warps keep a sliding window of their rows; opt-in with SOSBA_SOLVE_PIPE=1)
  int backsub_rowwise;         // set by launch_solve: 1 = row-by-row back substitution (SOSBA_SOLVE_BACKSUB=row), 0 = blocks of 4 rows
};
int launch_solve(sosba *h, const SolveArgs &a);

struct ResubArgs {
  int P, nf;
  const int *res_begin, *r_target, *p_host;
  const uint8_t *r_is_active, *r_dropped;
  const float *rec, *xAd;
  const float *HcdA, *HcdL, *bdSumF, *HdiF;
  float *step;
  // doStepFromBackup for points (only when do_step): idepth_backup = idepth; idepth = idepth_zero = backup + step
  int do_step;
  float *idepth, *idepth_zero, *idepth_backup, *deltaF;
  double *stats;  // [1] sum step^2  [2] sum |idepth_backup|  [3] count
  const int *gate;     // non-null: skip when *gate != 0
  double *zero_lin;    // non-null: clear the linearisation sums (2 doubles + 5 ints) for the launch that follows
  float *zero_newE;    // non-null (point shards): clear the newest-frame energy counts (NCCL fallback: all ranks' segments too)
  int zero_newE_n;     // floats + ints to clear, as 4-byte words
  long long *trace;    // SOSBA_TRACE record of this launch, or null
};
void launch_resubstitute(sosba *h, const ResubArgs &a);
void launch_step(sosba *h, const ResubArgs &ra, const StepArgs &sa);   // resubstitute + frame step, concurrently, one launch

// ---- pre-pyramid image path (k_exact.cu) ----------------------------------------------------------
struct UndistortArgs {
  int w, h, wOrg, hOrg, bits;     // bits: 8 or 16
  const void *raw;                // wOrg*hOrg raw pixels
  const float2 *remap;            // (remapX, remapY) per output pixel, or nullptr (passthrough)
  const float *G, *vig;           // inverse response / inverse vignette, or nullptr
  float factor;
  float *out;                     // w*h irradiance
};
void launch_undistort(sosba *h, const UndistortArgs &a);

// ---- k_select.cu --------------------------------------------------------------------------------
struct SelectArgs {
  int w, h, w1, w2, w32, h32;
  const float4 *img0, *img1, *img2;   // levels 0..2 of the frame ({I, dx, dy, absSquaredGrad})
  float *ths, *thsSmoothed;           // [w32*h32 + 100]
  const uint8_t *randomPattern;       // [w*h]
  int pot, nbx, nby;                  // block grid of (4*pot)^2 blocks
  float thFactor;
  const int *base;                    // [nb] running level-0 count at the start of each block (exclusive scan of cnt_in)
  const int *cnt_in;                  // [nb] counts the scan was taken from
  int *cnt_out;                       // [nb] counts of this pass
  uint8_t *map;                       // [w*h] status map (0, 1, 2, 4), cleared by the caller
  int *totals;                        // [0..2] n2, n3, n4   [3] flag: counts changed   [4] selected pixels (compaction)
};
void launch_select_hists(sosba *h, const SelectArgs &a);
void launch_select_blocks(sosba *h, const SelectArgs &a, bool write);
void launch_select_serial(sosba *h, const SelectArgs &a);
void launch_select_scan(sosba *h, const int *in, int *out, int n, int *total);
void launch_select_compact(sosba *h, const uint8_t *map, int n, int *chunk_cnt, int *chunk_off, int *total, int cap, int2 *list);

// ---- k_trace.cu ---------------------------------------------------------------------------------
struct TraceArgs {
  int n, w, h;
  const float4 *img;          // level 0 of the traced-into frame (trace_on) / of the host frame (immature_init)
  // immature_init
  const int *iu, *iv;
  float *color_out, *weights_out, *gradH_out, *energyTH_out;
  float outlierTHSum, overallWeight;
  // trace_on
  const int *host;
  const float *u, *v, *color, *weights, *gradH, *energyTH;
  float *idepth_min, *idepth_max, *quality, *uv, *pixint;
  uint8_t *status;
  const float *KRKi, *Kt, *aff;   // per host
  float huberTH;
  int *counts;                // [6] per ImmaturePointStatus
};
#define SOSBA_ACT_MAXF 32
struct ActivateArgs {
  int n, nf, w, h, min_obs;
  const float4 *img[SOSBA_ACT_MAXF];   // level 0 of frame f
  const float *RTll, *tTll, *aff;      // per (host * nf + target)
  float fxl, fyl, cxl, cyl, huberTH;
  const int *host;
  const float *u, *v, *color, *weights, *energyTH, *idepth_min, *idepth_max;
  signed char *result;
  float *idepth;
  uint8_t *res_state;                  // [n * nf]
};
void launch_optimize_immature(sosba *h, const ActivateArgs &a);
struct InitArgs {   // CoarseInitializer::calcResAndGS
  int n, w, h;
  const float4 *imgRef, *imgNew;   // level images of firstFrame / newFrame
  float RKi[9], t[3];
  float fx, fy, cx, cy, aff0 /* exp(a) */, aff1, huberTH, alphaOpt, couplingWeight;
  const float *u, *v, *idepth_new, *iR, *energy, *outlierTH;
  const uint8_t *isGood;
  float *energy_new, *maxstep, *lastHessian_new, *Jb;
  uint8_t *isGood_new;
  double *acc;                      // [0..45) acc9, [45..90) acc9SC, [90] E.A
};
void launch_init_res(sosba *h, const InitArgs &a);
void launch_immature_init(sosba *h, const TraceArgs &a);
void launch_trace_on(sosba *h, const TraceArgs &a);

// ---- k_tracker.cu -------------------------------------------------------------------------------
struct TrackGSArgs {
  int n, cap, kind;
  const float *warp;
  float fx, fy, a, b0;
  float scale, tx, ty, tz;
  double *acc;  // pose: 45 upper-tri sums at [8..53) ; scale: 3 sums at [8..11)
};
void launch_track_gs(sosba *h, const TrackGSArgs &a);
