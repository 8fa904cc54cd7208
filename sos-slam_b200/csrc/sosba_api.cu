// sosba_api.cu — the C ABI of include/sosba.h on top of the sm_100a kernels.  Host C++ only orchestrates:
// uploads, launch order, small read-backs.  There is no CPU implementation of any entry point: without a CUDA
// device sosba_create fails with SOSBA_E_NOGPU.
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <mutex>
#include <new>
#include <vector>

#include "host_ba.h"
#include "kernels.h"

void launch_make_xad(sosba *h, const double *d_x, int nf, const float *adHostF, const float *adTargetF, float *xAd);
void launch_scale_prior(sosba *h, float *priorF, const int *ids, int n, float fac);
void launch_fill_f32(sosba *h, float *dst, int n, float v);

using sosba_host::BAState;
using sosba_host::WindowTables;

int sosba_allreduce_acc(sosba *h, int with_newE);  // comm.cu: no-ops without a communicator
void sosba_xchg_args(sosba *h, StitchXchgArgs *a, int with_newE);
int sosba_allreduce_lin(sosba *h, int with_stats, double *extra = nullptr, int n_extra = 0);
int sosba_comm_max_int(sosba *h, int v, int *out);
static thread_local char g_err[512] = "";
static long long *g_dbg = nullptr;   // SOSBA_SOLVE_DEBUG: k_solve phase timestamps (clock64), 16 per launch
static long g_dbg_n = 0;
// SOSBA_TRACE=1: device-side timeline (globaltimer) of the launches of the last sosba_ba_optimize, printed by sosba_destroy
static const int TRACE_CAP = 256;
static long long *g_trace = nullptr;
static int g_trace_n = 0;
static const char *g_trace_name[TRACE_CAP];
static bool g_trace_enabled = getenv("SOSBA_TRACE") != nullptr;   // also switched by sosba_trace_enable
static bool trace_on() { return g_trace_enabled; }
long long *sosba_trace_slot(const char *name) {
  if (!trace_on() || !g_trace || g_trace_n >= TRACE_CAP) return nullptr;
  g_trace_name[g_trace_n] = name;
  return g_trace + 4 * (g_trace_n++);
}
static void trace_begin(sosba *h) {   // start of a traced sosba_ba_optimize: empty records
  if (!trace_on()) return;
  static long long init[4 * TRACE_CAP];
  if (!g_trace) cudaMalloc(&g_trace, sizeof(init) + 16 * sizeof(long long));   // (+ the phase stamps a linearisation writes behind its own record)
  for (int i = 0; i < TRACE_CAP; i++) { init[4 * i] = init[4 * i + 1] = 0x7fffffffffffffffLL; init[4 * i + 2] = init[4 * i + 3] = 0; }
  cudaMemcpyAsync(g_trace, init, sizeof(init), cudaMemcpyHostToDevice, h->stream);
  g_trace_n = 0;
}
static void trace_print() {
  if (!g_trace || g_trace_n == 0) return;
  std::vector<long long> t(4 * (size_t)TRACE_CAP);
  cudaMemcpy(t.data(), g_trace, t.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  long long t0 = 0x7fffffffffffffffLL;
  for (int i = 0; i < g_trace_n; i++) if (t[4 * i + 2] && g_trace_name[i][0] != '#' && g_trace_name[i][0] != '%') t0 = std::min(t0, t[4 * i]);
  fprintf(stderr, "SOSBA_TRACE: launches of the last sosba_ba_optimize, ns from the first entry: [entered, past the dependency wait, done] (done - past wait)\n");
  long long prev_done = 0;
  for (int i = 0; i < g_trace_n; i++) {
    if (!t[4 * i + 2] && g_trace_name[i][0] != '#' && g_trace_name[i][0] != '%') { fprintf(stderr, "  %-28s (not run)\n", g_trace_name[i]); continue; }
    if (g_trace_name[i][0] == '%') {   // frame step: 4 clock64 stamps
      fprintf(stderr, "  %-28s cycles: stage states %lld | frames %lld | pairs %lld\n", g_trace_name[i], t[4 * i + 1] - t[4 * i], t[4 * i + 2] - t[4 * i + 1], t[4 * i + 3] - t[4 * i + 2]);
      continue;
    }
    if (g_trace_name[i][0] == '#') {   // raw clock64 stamps: cycles between consecutive phases
      if (g_trace_name[i][1]) {
        fprintf(stderr, "  %-28s cycles: ids", g_trace_name[i]);
        static const char *ph[] = {"geometry+point loads", "projections", "taps+interp", "terms+sums", "records", "lists+stats"};
        for (int k = 1; k < 7; k++) fprintf(stderr, " | %s %lld", ph[k - 1], t[4 * i + k] - t[4 * i + k - 1]);
        fprintf(stderr, "\n");
      }
      continue;
    }
    const bool sub = t[4 * i] == 0x7fffffffffffffffLL;   // a sub-record (exit stamp only)
    if (sub) { fprintf(stderr, "  %-28s                  done %7lld\n", g_trace_name[i], t[4 * i + 2] - t0); continue; }
    fprintf(stderr, "  %-28s %7lld %7lld %7lld  (%5lld)  gap after previous done %5lld\n", g_trace_name[i], t[4 * i] - t0, t[4 * i + 1] - t0, t[4 * i + 2] - t0,
            t[4 * i + 2] - t[4 * i + 1], t[4 * i + 1] - t0 - prev_done);
    prev_done = t[4 * i + 2] - t0;
  }
  g_trace_n = 0;
}
static long long *g_xdbg = nullptr;   // SOSBA_XCHG_DEBUG: k_stitch_xchg globaltimer stamps, 8 per launch
static long g_xdbg_n = 0;
void sosba_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

struct BA {
  BAState st;
  WindowTables wt;
  double *d_HM = nullptr, *d_bM = nullptr;
  bool have_HM = false;
  int iterations_done = 0;
  bool mirror_stale = false;   // the device moved the evaluation point of the newest frame (k_frame_retarget)
};

// per-handle host mirrors that the device kernels never read
struct HostSide {
  float *d_imm = nullptr;   // immature-point arena (k_trace.cu): 30 floats + 1 status byte per point, grown on demand
  size_t imm_cap = 0;
  // pre-pyramid image path (sosba_undistort_set)
  bool und_set = false;
  int und_wOrg = 0, und_hOrg = 0, und_gDepth = 0;
  float2 *d_remap = nullptr;
  float *d_G = nullptr, *d_vig = nullptr;
  unsigned char *d_raw = nullptr;
  std::vector<int> seen_tmp, tiles_tmp;   // residuals_set scratch
  std::vector<char> big_stage;        // pageable staging for uploads larger than half the pinned ring
  unsigned char *w_arena = nullptr;   // window tables (win_layout)
  float *p_arena = nullptr;           // uploaded members of the points (carve_points)
  unsigned char *r_arena = nullptr;   // ids / energies / flags of the residuals (carve_residuals)
  float *d_loop = nullptr;      // loop-closure points: x | y | z | colour[level] (3 + levels arrays of loop_n)
  int loop_n = 0, loop_cap = 0, last_res_n = 0;
  // pixel selector (sosba_pixel_selector_set / sosba_pixel_select)
  std::vector<uint8_t> sel_random;
  int sel_pot = 3, sel_nb_cap = 0, sel_list_cap = 0;   // list capacity = w*h: at potential 1 every pixel can be selected
  uint8_t *d_sel_random = nullptr, *d_sel_map = nullptr;
  float *d_sel_ths = nullptr;          // ths | thsSmoothed, (w32*h32 + 100) each
  int *d_sel_int = nullptr;            // totals[8] | chunk_cnt | chunk_off | cnt A | cnt B | base
  int2 *d_sel_list = nullptr;
  int2 *pin_sel_list = nullptr;        // pinned mirror of the first SEL_LIST_FAST entries + totals
  float *d_init = nullptr;      // CoarseInitializer points of one level (sosba_init_calc_res_and_gs) + 91 fp64 sums
  size_t init_cap = 0;
  float *d_pool = nullptr;      // resident immature points (sosba_immature_pool_*): same arena layout as one pass of sosba_trace_immature
  size_t pool_cap = 0;
  int pool_n = 0, pool_hosts = 0;
  // sosba_optimize: the read-back of sosba_ba_download rides on the one synchronisation of sosba_ba_optimize
  bool fold_download = false, download_ready = false;
  float *pin_idepth = nullptr;
  size_t pin_idepth_cap = 0;
  float *d_act = nullptr;       // activation outputs: [idepth n floats][result n bytes][res_state n*nf bytes]
  size_t act_cap = 0;
  float *d_act_win = nullptr;   // PRE_RTll / PRE_tTll / PRE_aff_mode per frame pair
  int act_win_cap = 0;
  float *d_imm_host = nullptr;   // KRKi / Kt / aff per host frame + the 6 status counters
  int imm_host_cap = 0;
  std::vector<int> p_host, res_begin, r_point, r_target, r_host_tmp;
  std::vector<float> delta_tmp;
  bool r_host_copy = false;   // r_point / r_target mirror the device arrays
  bool stash_pending = false; // the next k_solve keeps the sums of the first linearisation of sosba_ba_optimize
  std::vector<void *> allocs;
  int n_lin = 0;
  double *pin_d = nullptr;   // pinned scratch: [4096] doubles
  int *pin_i = nullptr;      // pinned scratch: [64] ints
  float *pin_f = nullptr;    // pinned scratch: [4096] floats
  double *d_HMtmp = nullptr, *d_bMtmp = nullptr;
  int *d_ids = nullptr; int ids_cap = 0;
  int *d_status = nullptr;
  double *d_scratch = nullptr, *d_rstats = nullptr, *d_Hfinal = nullptr;
  double *d_fs = nullptr, *d_cs = nullptr, *d_iter = nullptr;   // device-resident frame / calibration state of the GN loop
  int *d_cnt = nullptr;
  size_t scratch_doubles = 0, scratch_zero_doubles = 0;   // all of it / the part the fused linearize clears (tables + H parts)
  bool tables_clean = false;
  int rstats_par = 0;           // which half of d_rstats the next back-substitution writes
  int *d_ctl = nullptr;         // device loop control: [0] latch, [1] iterations run, [2] non-finite flag, [3] resInA of the last solve
  double *d_stash = nullptr;    // linearisation sums of the first pass of an optimize
  const int *gate = nullptr;    // = d_ctl while a gated optimize loop is being enqueued
  int loop_iter = -1;           // body index while enqueuing a gated loop
  bool th_pending = false;      // setNewFrameEnergyTH of the last fused linearisation still to run
  bool fused_acc_ok = false;    // every (point, target) pair holds at most one residual and a tile fits shared memory
  int max_res_per_tile = 0, n_tiles = 0, tiles_cap = 0;
  int *d_tiles = nullptr;       // int4 per tile of the fused accumulation
  float *p_zero_arena = nullptr;   // point outputs / CSR that start from zero (one memset per points_set)
  size_t p_zero_words = 0;
  char *stage = nullptr;        // pinned staging ring of up()
  size_t stage_cap = 0, stage_off = 0;
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;   // (start, stop) pairs around the linearize kernel
};
static HostSide *HS(sosba *h);
static int solve_flag_error(int flag);

#define API extern "C" __attribute__((visibility("default")))
#define CHECK_H(h) do { if (!(h)) { sosba_set_error("null handle"); return SOSBA_E_ARG; } cudaSetDevice((h)->device); } while (0)

template <class T> static int dalloc(sosba *h, T **p, size_t n) {
  if (n == 0) n = 1;
  void *q = nullptr;
  cudaError_t e = cudaMalloc(&q, n * sizeof(T));
  if (e != cudaSuccess) { sosba_set_error("cudaMalloc(%zu) -> %s", n * sizeof(T), cudaGetErrorString(e)); return SOSBA_E_CUDA; }
  cudaMemsetAsync(q, 0, n * sizeof(T), h->stream);
  *p = (T *)q;
  HS(h)->allocs.push_back(q);
  return SOSBA_OK;
}
template <class T> static void dfree(sosba *h, T *&p) {
  if (!p) return;
  auto &v = HS(h)->allocs;
  v.erase(std::remove(v.begin(), v.end(), (void *)p), v.end());
  cudaFree((void *)p);
  p = nullptr;
}
#define DALLOC(h, p, n) do { int rc__ = dalloc(h, &(p), (size_t)(n)); if (rc__) return rc__; } while (0)
// Host -> device through a pinned staging ring: the source is consumed when the call returns (callers pass pageable
// memory and locals), the copy itself is asynchronous and does not synchronise the stream the way a pageable
// cudaMemcpyAsync does.  The ring is recycled after a stream synchronisation when it runs full.
static int up_bytes(sosba *h, void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return SOSBA_OK;
  HostSide *hs = HS(h);
  if (bytes >= 64 * 1024) {   // a large source in page-locked memory (cudaHostAlloc / cudaHostRegister) needs no staging copy:
                              // the DMA reads it in place; the caller keeps it unchanged until the next synchronising call
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost) {
      SOSBA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
      return SOSBA_OK;
    }
    cudaGetLastError();
  }
  if (!hs->stage || bytes > hs->stage_cap / 2) {   // oversized: plain (synchronising) copy
    SOSBA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return SOSBA_OK;
  }
  if (hs->stage_off + bytes > hs->stage_cap) {
    SOSBA_CUDA(cudaStreamSynchronize(h->stream));
    hs->stage_off = 0;
  }
  memcpy(hs->stage + hs->stage_off, src, bytes);
  SOSBA_CUDA(cudaMemcpyAsync(dst, hs->stage + hs->stage_off, bytes, cudaMemcpyHostToDevice, h->stream));
  hs->stage_off += (bytes + 255) & ~(size_t)255;
  return SOSBA_OK;
}
// A block of the pinned staging ring, valid until the stream is synchronised twice (the ring only wraps behind a sync).
static int stage_reserve(sosba *h, size_t bytes, char **p) {
  HostSide *hs = HS(h);
  if (!hs->stage || bytes > hs->stage_cap) { sosba_set_error("staging block of %zu bytes", bytes); return SOSBA_E_ARG; }
  if (hs->stage_off + bytes > hs->stage_cap) {
    SOSBA_CUDA(cudaStreamSynchronize(h->stream));
    hs->stage_off = 0;
  }
  *p = hs->stage + hs->stage_off;
  hs->stage_off += (bytes + 255) & ~(size_t)255;
  return SOSBA_OK;
}
template <class T> static int up(sosba *h, T *dst, const T *src, size_t n) { return up_bytes(h, dst, src, n * sizeof(T)); }
template <class T> static int down(sosba *h, T *dst, const T *src, size_t n) {
  if (n == 0) return SOSBA_OK;
  SOSBA_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
  return SOSBA_OK;
}
static int sync(sosba *h) {
  SOSBA_CUDA(cudaStreamSynchronize(h->stream));
  SOSBA_CUDA(cudaGetLastError());
  return SOSBA_OK;
}

// the HostSide object is stored behind sosba::ba's neighbour: keep a side table keyed by handle
static std::mutex g_mu;
static std::map<sosba *, HostSide *> g_side;
static HostSide *HS(sosba *h) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_side.find(h);
  if (it != g_side.end()) return it->second;
  HostSide *s = new HostSide();
  g_side[h] = s;
  return s;
}

// ---- config -------------------------------------------------------------------------------------
API void sosba_config_default(sosba_config *c, int32_t w, int32_t h) {
  // util/settings.cpp:28-204 after settingsDefault(preset 0, mode 1) (main.cpp:27-90)
  memset(c, 0, sizeof(*c));
  c->w = w; c->h = h; c->pyr_levels = 0; c->max_frames = 16; c->num_threads = 1;
  c->gamma_weights_pixel_select = 1; c->min_opt_iterations = 1;
  c->huber_th = 9; c->outlier_th_sum_component = 50 * 50;
  c->affine_opt_mode_a = 0; c->affine_opt_mode_b = 0;
  c->coarse_cutoff_th = 20; c->idepth_fix_prior = 50 * 50; c->idepth_fix_prior_marg_fac = 600 * 600;
  c->frame_energy_th_const_weight = 0.5f; c->frame_energy_th_n = 0.7f; c->frame_energy_th_fac_median = 1.5f;
  c->overall_energy_th_weight = 1; c->initial_calib_hessian = 5e9f;
  c->initial_rot_prior = 1e11f; c->initial_trans_prior = 1e10f; c->initial_aff_a_prior = 1e14f; c->initial_aff_b_prior = 1e14f;
  c->marg_weight_fac = 0.5f * 0.5f; c->th_opt_iterations = 1.2f;
}

API const char *sosba_last_error(void) { return g_err; }

// ---- lifetime -----------------------------------------------------------------------------------
static int create_body(sosba *h, const sosba_config *cfg, int device);
API void sosba_destroy(sosba_t *h);
API int sosba_create(const sosba_config *cfg, int32_t device, sosba_t **out) {
  if (!cfg || !out || cfg->w <= 0 || cfg->h <= 0) { sosba_set_error("bad config"); return SOSBA_E_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    sosba_set_error("no CUDA device: libsosba has no CPU fallback");
    return SOSBA_E_NOGPU;
  }
  if (device < 0 || device >= ndev) { sosba_set_error("device %d out of range (%d devices)", device, ndev); return SOSBA_E_ARG; }
  SOSBA_CUDA(cudaSetDevice(device));
  sosba *h = new (std::nothrow) sosba();
  if (!h) return SOSBA_E_ARG;
  h->cfg = *cfg;
  h->device = device;
  const int rc = create_body(h, cfg, device);
  if (rc) { sosba_destroy(h); return rc; }   // a failure half-way must not leak the handle, its stream, or the pinned buffers
  *out = h;
  return SOSBA_OK;
}
static int create_body(sosba *h, const sosba_config *cfg, int device) {
  cudaDeviceProp prop;
  SOSBA_CUDA(cudaGetDeviceProperties(&prop, device));
  h->sm_count = prop.multiProcessorCount;
  SOSBA_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  // setGlobalCalib (globalCalib.cpp:39-49)
  int wlvl = cfg->w, hlvl = cfg->h, lv = 1;
  while (wlvl % 2 == 0 && hlvl % 2 == 0 && wlvl * hlvl > 5000 && lv < SOSBA_MAX_LEVELS) { wlvl /= 2; hlvl /= 2; lv++; }
  h->levels = cfg->pyr_levels > 0 ? std::min<int>(cfg->pyr_levels, SOSBA_MAX_LEVELS) : lv;
  size_t off = 0;
  for (int l = 0; l < h->levels; l++) {
    h->wl[l] = cfg->w >> l; h->hl[l] = cfg->h >> l;
    h->lvl_off[l] = off;
    off += ((size_t)h->wl[l] * h->hl[l] + 31) / 32 * 32;  // keep every level 512-byte aligned
  }
  h->lvl_off[h->levels] = off;
  const int nslots = cfg->max_frames > 0 ? cfg->max_frames : 16;
  h->slot_img.assign(nslots, nullptr);
  h->slot_plane.assign(nslots, nullptr);
  h->slot_valid.assign(nslots, 0);
  HostSide *hs = HS(h);
  for (int s = 0; s < nslots; s++) { DALLOC(h, h->slot_img[s], off); DALLOC(h, h->slot_plane[s], off); }
  DALLOC(h, h->d_stage, (size_t)cfg->w * cfg->h);
  DALLOC(h, h->d_B, 256);
  h->h_pinned_bytes = (size_t)cfg->w * cfg->h * 4 * sizeof(float);
  SOSBA_CUDA(cudaMallocHost((void **)&h->h_pinned, h->h_pinned_bytes));
  SOSBA_CUDA(cudaMallocHost((void **)&hs->pin_d, 4096 * sizeof(double)));
  hs->stage_cap = (size_t)16 << 20;
  SOSBA_CUDA(cudaMallocHost((void **)&hs->stage, hs->stage_cap));
  SOSBA_CUDA(cudaMallocHost((void **)&hs->pin_i, 64 * sizeof(int)));
  SOSBA_CUDA(cudaMallocHost((void **)&hs->pin_f, 4096 * sizeof(float)));
  // linearize statistics, one region so that one memset clears it and one copy reads it back:
  //   double energy | double pad | int counts[16] | float thOut[4]
  DALLOC(h, h->d_stats, 16);
  h->d_counts = (int *)(h->d_stats + 2);
  h->d_thOut = (float *)(h->d_counts + 16);
  DALLOC(h, h->t_acc, 64);
  DALLOC(h, hs->d_status, 4);
  h->ba = new BA();
  return sync(h);
}

int sosba_comm_destroy(sosba_t *h);

API void sosba_destroy(sosba_t *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (g_dbg && g_dbg_n > 8) {
    std::vector<long long> t(64 * 32);
    cudaMemcpy(t.data(), g_dbg, t.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    const int n = (int)std::min<long>(g_dbg_n, 64);
    fprintf(stderr, "k_solve phase cycles (median of last %d launches): ", n);
    for (int ph = 1; ph < 9; ph++) {
      std::vector<long long> v;
      for (int k = 0; k < n; k++) v.push_back(t[32 * k + ph] - t[32 * k + ph - 1]);
      std::sort(v.begin(), v.end());
      fprintf(stderr, " %lld", v[v.size() / 2]);
    }
    std::vector<long long> v;
    for (int k = 0; k < n; k++) v.push_back(t[32 * k + 8] - t[32 * k]);
    std::sort(v.begin(), v.end());
    fprintf(stderr, "  total %lld\n", v[v.size() / 2]);
    {  // panels 4..6 of the last launch: [panel start, panel done, update start, update done] relative to panel 4's start
      const long long *q = t.data() + 32 * ((g_dbg_n - 1) % 64) + 16;
      fprintf(stderr, "k_solve look-ahead timeline, panels 4,5: [panel: data landed, chain done, stored | update: L/Y landed, published, rest done] (cycles):");
      for (int i = 0; i < 16; i++) fprintf(stderr, " %lld%s", q[i] - q[0], (i & 7) == 2 ? " |" : (i & 7) == 5 ? " | through BAR_PUB, published (warp 7):" : (i & 7) == 7 ? " ||" : "");
      fprintf(stderr, "\n");
    }
    {
      long long q[16];
      cudaMemcpy(q, g_dbg + 64 * 32, sizeof(q), cudaMemcpyDeviceToHost);
      fprintf(stderr, "k_accumulate_fused CTA 0 phase cycles (last launch): setup %lld | top (warp 9) %lld [lists %lld, sums %lld, red %lld] || points (warp 1) %lld | barrier %lld schur %lld flush %lld\n",
              q[1] - q[0], q[2] - q[1], q[7] - q[1], q[8] - q[7], q[9] - q[8], q[6] - q[1], q[3] - q[6], q[4] - q[3], q[5] - q[4]);
    }
    g_dbg_n = 0;
  }
  trace_print();
  if (g_xdbg && g_xdbg_n > 8) {   // per launch: [0] misc CTA start, [1] its values ready, [2] summed; [3],[4] diagonal tile 0 ready / summed; [5],[6] energies start / done
    std::vector<long long> t(64 * 8);
    cudaMemcpy(t.data(), g_xdbg, t.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    const int n = (int)std::min<long>(g_xdbg_n, 64);
    auto med = [&](int hi, int lo) { std::vector<long long> v; for (int k = 0; k < n; k++) v.push_back(t[8 * k + hi] - t[8 * k + lo]); std::sort(v.begin(), v.end()); return v[v.size() / 2]; };
    auto mx = [&](int hi, int lo) { long long m = 0; for (int k = 0; k < n; k++) m = std::max(m, t[8 * k + hi] - t[8 * k + lo]); return m; };
    fprintf(stderr, "[rank %d] k_stitch_xchg ns (median / max of last %d launches): misc compute %lld/%lld, misc exchange %lld/%lld, diag0 ready after misc start %lld/%lld, "
            "diag0 exchange %lld/%lld, energies %lld/%lld\n", h->rank, n, med(1, 0), mx(1, 0), med(2, 1), mx(2, 1), med(3, 0), mx(3, 0), med(4, 3), mx(4, 3), med(6, 5), mx(6, 5));
    g_xdbg_n = 0;
  }
  sosba_comm_destroy(h);
  HostSide *hs = HS(h);
  for (void *p : hs->allocs) cudaFree(p);
  for (void *p : h->lm_allocs) cudaFree(p);
  if (h->h_pinned) cudaFreeHost(h->h_pinned);
  if (hs->pin_d) cudaFreeHost(hs->pin_d);
  if (hs->stage) cudaFreeHost(hs->stage);
  if (hs->pin_i) cudaFreeHost(hs->pin_i);
  if (hs->pin_f) cudaFreeHost(hs->pin_f);
  if (hs->pin_sel_list) cudaFreeHost(hs->pin_sel_list);
  if (hs->pin_idepth) cudaFreeHost(hs->pin_idepth);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h->ba;
  { std::lock_guard<std::mutex> lk(g_mu); g_side.erase(h); }
  delete hs;
  delete h;
}

API int sosba_set_stream(sosba_t *h, void *s) {
  CHECK_H(h);
  cudaStream_t ns = s ? (cudaStream_t)s : h->own_stream;
  if (ns != h->stream) {   // copies queued on the old stream still read the staging ring: let them finish before the ring is reused from the new one
    int rc = sync(h);
    if (rc) return rc;
    h->stream = ns;
  }
  return SOSBA_OK;
}
API int sosba_synchronize(sosba_t *h) { CHECK_H(h); return sync(h); }
// CUDA-event timing of the dominant kernel (linearize) on its launch stream, for the bench roofline.
API int sosba_profile_enable(sosba_t *h, int32_t on) {
  CHECK_H(h);
  HostSide *hs = HS(h);
  if (on == 1) {   // 1: start a new measurement; 2: resume (keep what was recorded); 0: pause (sosba_profile_read collects and clears)
    for (cudaEvent_t e : hs->prof_ev) cudaEventDestroy(e);
    hs->prof_ev.clear();
  }
  hs->prof_on = on != 0;
  return SOSBA_OK;
}
API int sosba_profile_read(sosba_t *h, double *ms_total, int32_t *launches) {
  CHECK_H(h);
  HostSide *hs = HS(h);
  int rc = sync(h);
  if (rc) return rc;
  double tot = 0;
  for (size_t i = 0; i + 1 < hs->prof_ev.size(); i += 2) {
    float ms = 0;
    SOSBA_CUDA(cudaEventElapsedTime(&ms, hs->prof_ev[i], hs->prof_ev[i + 1]));
    tot += ms;
  }
  if (ms_total) *ms_total = tot;
  if (launches) *launches = (int32_t)(hs->prof_ev.size() / 2);
  for (cudaEvent_t e : hs->prof_ev) cudaEventDestroy(e);
  hs->prof_ev.clear();
  return SOSBA_OK;
}
// device-side timeline (globaltimer stamps inside the kernels of the Gauss-Newton loop) of the next sosba_ba_optimize calls
API int sosba_trace_enable(sosba_t *h, int32_t on) {
  CHECK_H(h);
  int rc = sync(h);
  if (rc) return rc;
  g_trace_enabled = on != 0;
  if (!on) g_trace_n = 0;
  return SOSBA_OK;
}
// mean of (last CTA done - first CTA past its dependency wait), in ns, over the launches of the last traced
// sosba_ba_optimize whose record name starts with `kernel` ("k_linearize", "k_solve", ...); skip_first leaves out that many
// leading matches (the first linearisation of a step runs on a cold L2)
API int sosba_trace_read(sosba_t *h, const char *kernel, int32_t skip_first, double *mean_ns, int32_t *launches) {
  CHECK_H(h);
  if (!kernel) return SOSBA_E_ARG;
  int rc = sync(h);
  if (rc) return rc;
  double tot = 0;
  int n = 0, seen = 0;
  if (g_trace && g_trace_n > 0) {
    std::vector<long long> t(4 * (size_t)TRACE_CAP);
    SOSBA_CUDA(cudaMemcpy(t.data(), g_trace, t.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    const size_t len = strlen(kernel);
    for (int i = 0; i < g_trace_n; i++) {
      if (strncmp(g_trace_name[i], kernel, len) != 0 || !t[4 * i + 2] || t[4 * i] == 0x7fffffffffffffffLL) continue;
      if (seen++ < skip_first) continue;
      tot += (double)(t[4 * i + 2] - t[4 * i + 1]);
      n++;
    }
  }
  if (mean_ns) *mean_ns = n ? tot / n : 0.0;
  if (launches) *launches = n;
  return SOSBA_OK;
}
API int64_t sosba_launch_count(const sosba_t *h) { return h ? h->launches : 0; }
API int32_t sosba_pyr_levels(const sosba_t *h) { return h ? h->levels : 0; }

// ---- a1 -----------------------------------------------------------------------------------------
API int sosba_frame_make_images(sosba_t *h, int32_t slot, const float *color, const float *B) {
  CHECK_H(h);
  if (slot < 0 || slot >= (int)h->slot_img.size() || !color) { sosba_set_error("bad slot/color"); return SOSBA_E_ARG; }
  const size_t n = (size_t)h->cfg.w * h->cfg.h;
  int rc;
  if ((rc = up(h, h->d_stage, color, n))) return rc;   // staged through the pinned ring, asynchronous
  if (B && (rc = up(h, h->d_B, B, 256))) return rc;
  launch_make_images(h, slot, h->d_stage, B ? h->d_B : nullptr);
  h->slot_valid[slot] = 1;
  SOSBA_CUDA(cudaGetLastError());
  return SOSBA_OK;
}

// ---- 8f rank 2: raw frame -> photometric + geometric undistortion -> pyramid ----------------------
API int sosba_undistort_set(sosba_t *h, int32_t w_org, int32_t h_org, const float *remapX, const float *remapY, const float *G, int32_t g_depth,
                            const float *vignette_inv) {
  CHECK_H(h);
  const int w = h->cfg.w, hh = h->cfg.h;
  if (w_org < 2 || h_org < 2 || (!remapX) != (!remapY) || (G && g_depth != 256 && g_depth != 65536) || (vignette_inv && !G)) {
    sosba_set_error("undistort_set: bad arguments");
    return SOSBA_E_ARG;
  }
  if (!remapX && (w_org != w || h_org != hh)) { sosba_set_error("passthrough needs w_org x h_org == w x h"); return SOSBA_E_ARG; }
  HostSide *hs = HS(h);
  int rc;
  if ((rc = sync(h))) return rc;
  dfree(h, hs->d_remap); dfree(h, hs->d_G); dfree(h, hs->d_vig); dfree(h, hs->d_raw);
  hs->und_set = false;
  const size_t n = (size_t)w * hh, no = (size_t)w_org * h_org;
  if (remapX) {
    std::vector<float2> rm(n);
    for (size_t i = 0; i < n; i++) {
      // the bilinear tap reads (x+1, y+1): the reference's map construction guarantees this margin (Undistort.cpp:862-884)
      if (remapX[i] >= 0 && !(remapX[i] < w_org - 1 && remapY[i] >= 0 && remapY[i] < h_org - 1)) {
        sosba_set_error("remap entry %zu (%g, %g) outside the raw image", i, remapX[i], remapY[i]);
        return SOSBA_E_ARG;
      }
      rm[i] = make_float2(remapX[i], remapY[i]);
    }
    DALLOC(h, hs->d_remap, n);
    SOSBA_CUDA(cudaMemcpyAsync(hs->d_remap, rm.data(), n * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
    if ((rc = sync(h))) return rc;
  }
  if (G) { DALLOC(h, hs->d_G, g_depth); SOSBA_CUDA(cudaMemcpyAsync(hs->d_G, G, g_depth * sizeof(float), cudaMemcpyHostToDevice, h->stream)); }
  if (vignette_inv) { DALLOC(h, hs->d_vig, no); SOSBA_CUDA(cudaMemcpyAsync(hs->d_vig, vignette_inv, no * sizeof(float), cudaMemcpyHostToDevice, h->stream)); }
  DALLOC(h, hs->d_raw, no * 2);
  hs->und_wOrg = w_org; hs->und_hOrg = h_org; hs->und_gDepth = G ? g_depth : 0;
  if ((rc = sync(h))) return rc;
  hs->und_set = true;
  return SOSBA_OK;
}

API int sosba_frame_make_images_raw(sosba_t *h, int32_t slot, const void *raw, int32_t raw_bits, float factor, const float *B, float *image_out) {
  CHECK_H(h);
  HostSide *hs = HS(h);
  if (!hs->und_set) { sosba_set_error("undistort_set first"); return SOSBA_E_STATE; }
  if (slot < 0 || slot >= (int)h->slot_img.size() || !raw || (raw_bits != 8 && raw_bits != 16)) { sosba_set_error("bad slot / raw"); return SOSBA_E_ARG; }
  if (hs->und_gDepth && raw_bits == 16 && hs->und_gDepth != 65536) { sosba_set_error("16-bit frame with a %d-entry response", hs->und_gDepth); return SOSBA_E_ARG; }
  const size_t no = (size_t)hs->und_wOrg * hs->und_hOrg, n = (size_t)h->cfg.w * h->cfg.h;
  int rc;
  if ((rc = up_bytes(h, hs->d_raw, raw, no * (raw_bits / 8)))) return rc;
  if (B && (rc = up(h, h->d_B, B, 256))) return rc;
  UndistortArgs a;
  a.w = h->cfg.w; a.h = h->cfg.h; a.wOrg = hs->und_wOrg; a.hOrg = hs->und_hOrg; a.bits = raw_bits;
  a.raw = hs->d_raw; a.remap = hs->d_remap; a.G = hs->d_G; a.vig = hs->d_vig; a.factor = factor; a.out = h->d_stage;
  launch_undistort(h, a);
  launch_make_images(h, slot, h->d_stage, B ? h->d_B : nullptr);
  h->slot_valid[slot] = 1;
  SOSBA_CUDA(cudaGetLastError());
  if (image_out) {
    if ((rc = down(h, image_out, (const float *)h->d_stage, n))) return rc;
    return sync(h);
  }
  return SOSBA_OK;
}

// device-resident input (bench: inputs already in HBM)
API int sosba_frame_make_images_dev(sosba_t *h, int32_t slot, const float *color_dev, const float *B_dev) {
  CHECK_H(h);
  if (slot < 0 || slot >= (int)h->slot_img.size() || !color_dev) return SOSBA_E_ARG;
  launch_make_images(h, slot, color_dev, B_dev);
  h->slot_valid[slot] = 1;
  SOSBA_CUDA(cudaGetLastError());
  return SOSBA_OK;
}

API int sosba_frame_get_level(sosba_t *h, int32_t slot, int32_t lvl, float *dI3, float *absg) {
  CHECK_H(h);
  if (slot < 0 || slot >= (int)h->slot_img.size() || lvl < 0 || lvl >= h->levels || !h->slot_valid[slot]) { sosba_set_error("bad slot/level"); return SOSBA_E_ARG; }
  const size_t n = (size_t)h->wl[lvl] * h->hl[lvl];
  int rc = sync(h);
  if (rc) return rc;
  if (n * sizeof(float4) > h->h_pinned_bytes) return SOSBA_E_ARG;
  if ((rc = down(h, (float4 *)h->h_pinned, h->slot_img[slot] + h->lvl_off[lvl], n))) return rc;
  if ((rc = sync(h))) return rc;
  const float *p = h->h_pinned;
  for (size_t i = 0; i < n; i++) {
    if (dI3) { dI3[3 * i] = p[4 * i]; dI3[3 * i + 1] = p[4 * i + 1]; dI3[3 * i + 2] = p[4 * i + 2]; }
    if (absg) absg[i] = p[4 * i + 3];
  }
  return SOSBA_OK;
}

// ---- uploads ------------------------------------------------------------------------------------
// The window tables live in ONE arena so that an upload is one staged copy (doubles first, everything 16-byte aligned):
//   head (sosba_window_set only):  img0 [nf pointers, padded] | adHost [n64] | adTarget [n64] | adHostF [n64] | adTargetF [n64]
//   tail (also sosba_window_update): wprior [4 + 24 nf doubles, padded] | precalc [n2*32] | adHTdeltaF [n2*8] | calib [16] | frameEnergyTH [nf, padded]
struct WinLayout { size_t img0, adHost, adTarget, adHostF, adTargetF, tail, wprior, precalc, adHTdelta, calib, th, total; };
static WinLayout win_layout(int nf) {
  const size_t n2 = (size_t)nf * nf, n64 = n2 * 64;
  auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
  WinLayout L;
  size_t o = 0;
  L.img0 = o; o += al(sizeof(void *) * nf);
  L.adHost = o; o += n64 * 8;
  L.adTarget = o; o += n64 * 8;
  L.adHostF = o; o += n64 * 4;
  L.adTargetF = o; o += n64 * 4;
  L.tail = o;
  L.wprior = o; o += al(8 * (4 + 24 * (size_t)nf));
  L.precalc = o; o += n2 * SOSBA_PRECALC_FLOATS * 4;
  L.adHTdelta = o; o += n2 * 8 * 4;
  L.calib = o; o += 16 * 4;
  L.th = o; o += al(4 * (size_t)nf);
  L.total = o;
  return L;
}

static int ensure_window(sosba *h, int nf) {
  if (nf <= h->nf_alloc) return SOSBA_OK;
  HostSide *hs = HS(h);
  h->nf_alloc = 0;   // (see ensure_residuals)
  dfree(h, hs->w_arena);
  dfree(h, h->d_x); dfree(h, h->d_xAd);
  dfree(h, hs->d_scratch); h->d_accTop = h->d_accSC = h->d_H = nullptr;
  dfree(h, hs->d_HMtmp); dfree(h, hs->d_bMtmp);
  const size_t n2 = (size_t)nf * nf;
  const int D = 4 + 8 * nf;
  DALLOC(h, hs->w_arena, win_layout(nf).total);
  {  // one scratch region zeroed by a single memset per API-level solve: top tables (A | L), Schur Gram, counters (ints),
     // H parts A | L | SC, back-substitution sums [4] ; H part 3 (final system) follows and is never cleared.
     // Inside the Gauss-Newton loop only tables + Gram + counters are cleared (by the fused linearisation): the stitch of
     // the loop (k_stitch_xchg) writes every entry of H part 0 and needs no zeroed target.
    const size_t HB = (size_t)D * D + D;
    const size_t scpad = (((size_t)(D + 1) * (D + 1)) + 1) & ~(size_t)1;   // even: H, b behind it stay 16-byte aligned (TMA bulk copies in k_solve)
    hs->scratch_doubles = 2 * n2 * SOSBA_TOPB + scpad + 2 + 3 * HB + 8 + 2;
    DALLOC(h, hs->d_scratch, hs->scratch_doubles + HB);
    h->d_accTop = hs->d_scratch;
    h->d_accSC = h->d_accTop + 2 * n2 * SOSBA_TOPB;
    hs->d_cnt = (int *)(h->d_accSC + scpad);                // [0] resInA [1] resInL (cleared with the tables)
    h->d_H = h->d_accSC + scpad + 2;
    hs->d_rstats = h->d_H + 3 * HB;                         // 2 x 4 doubles: back-substitution sums by loop body parity
    hs->scratch_zero_doubles = (size_t)(h->d_H - hs->d_scratch);
    h->d_rstats_all = hs->d_rstats; h->d_cnt_all = hs->d_cnt;
    hs->d_Hfinal = hs->d_scratch + hs->scratch_doubles;
  }
  DALLOC(h, h->d_x, D);
  DALLOC(h, h->d_xAd, n2 * 8 + 8);
  DALLOC(h, hs->d_HMtmp, (size_t)D * D); DALLOC(h, hs->d_bMtmp, D);
  if (!hs->d_fs) { DALLOC(h, hs->d_fs, 16 * SOSBA_FS); DALLOC(h, hs->d_cs, 16); DALLOC(h, hs->d_iter, 8); DALLOC(h, hs->d_ctl, 4); DALLOC(h, hs->d_stash, 16); }
  h->nf_alloc = nf; h->D_alloc = D;
  return SOSBA_OK;
}

static int window_apply(sosba *h, const sosba_window *w, bool full) {
  const int nf = w->nf;
  if (nf <= 0 || nf > 16) { sosba_set_error("nf=%d out of range", nf); return SOSBA_E_ARG; }
  if (!full && nf != h->nf) { sosba_set_error("window_update before window_set"); return SOSBA_E_STATE; }
  int rc;
  HostSide *hs = HS(h);
  if (full && (rc = ensure_window(h, nf))) return rc;
  // (the arena may be sized for a larger window: the tables of this one are laid out for ITS nf at the front)
  const WinLayout L = win_layout(nf);
  unsigned char *A = hs->w_arena;
  h->d_img0 = (const float4 **)(A + L.img0);
  h->d_adHost = (double *)(A + L.adHost); h->d_adTarget = (double *)(A + L.adTarget);
  h->d_adHostF = (float *)(A + L.adHostF); h->d_adTargetF = (float *)(A + L.adTargetF);
  h->d_wprior = (double *)(A + L.wprior); h->d_precalc = (float *)(A + L.precalc); h->d_adHTdeltaF = (float *)(A + L.adHTdelta);
  h->d_calib = (float *)(A + L.calib); h->d_frameEnergyTH = (float *)(A + L.th);
  const size_t first = full ? 0 : L.tail, bytes = L.total - first;
  char *blk = nullptr;
  if ((rc = stage_reserve(h, bytes, &blk))) return rc;
  char *S = blk - first;   // S + offset = staging address of an arena offset
  const size_t n2 = (size_t)nf * nf, n64 = n2 * 64;
  if (full) {
    h->nf = nf;
    h->frame_slot.assign(w->frame_slot, w->frame_slot + nf);
    const float4 **img = (const float4 **)(S + L.img0);
    for (int i = 0; i < nf; i++) {
      const int sl = w->frame_slot[i];
      if (sl < 0 || sl >= (int)h->slot_img.size()) { sosba_set_error("frame_slot[%d]=%d", i, sl); return SOSBA_E_ARG; }
      img[i] = h->slot_img[sl];
    }
    memcpy(S + L.adHost, w->adHost, n64 * 8);
    memcpy(S + L.adTarget, w->adTarget, n64 * 8);
    float *fh = (float *)(S + L.adHostF), *ft = (float *)(S + L.adTargetF);
    for (size_t i = 0; i < n64; i++) { fh[i] = (float)w->adHost[i]; ft[i] = (float)w->adTarget[i]; }
  }
  memcpy(S + L.precalc, w->precalc, n2 * SOSBA_PRECALC_FLOATS * 4);
  memcpy(S + L.adHTdelta, w->adHTdeltaF, n2 * 8 * 4);
  memcpy(S + L.th, w->frame_energy_th, 4 * (size_t)nf);
  float *c = h->h_calib;
  for (int i = 0; i < 4; i++) c[i] = w->calib[i];
  c[4] = 1.0f / c[0]; c[5] = 1.0f / c[1];  // CalibHessian::setValue, HessianBlocks.h:495-496
  for (int i = 0; i < 4; i++) c[6 + i] = w->cDeltaF[i];
  memcpy(S + L.calib, c, 10 * 4);
  h->h_wprior.resize(4 + 24 * (size_t)nf);
  double *wp = h->h_wprior.data();
  if (full) { for (int i = 0; i < 4; i++) wp[i] = w->cPrior[i]; memcpy(wp + 4, w->frame_prior, sizeof(double) * 8 * nf); }
  memcpy(wp + 4 + 8 * nf, w->frame_delta_prior, sizeof(double) * 8 * nf);
  memcpy(wp + 4 + 16 * nf, w->frame_delta, sizeof(double) * 8 * nf);
  memcpy(S + L.wprior, wp, 8 * (4 + 24 * (size_t)nf));
  SOSBA_CUDA(cudaMemcpyAsync(A + first, blk, bytes, cudaMemcpyHostToDevice, h->stream));
  return SOSBA_OK;
}
API int sosba_window_set(sosba_t *h, const sosba_window *w) { CHECK_H(h); if (!w) return SOSBA_E_ARG; return window_apply(h, w, true); }
API int sosba_window_update(sosba_t *h, const sosba_window *w) { CHECK_H(h); if (!w) return SOSBA_E_ARG; return window_apply(h, w, false); }

static int ensure_points(sosba *h, int P) {
  if (P <= h->P_alloc) return SOSBA_OK;
  HostSide *hs = HS(h);
  h->P_alloc = 0;   // (see ensure_residuals)
  dfree(h, hs->p_arena); dfree(h, hs->p_zero_arena);
  const size_t n = ((size_t)P + (size_t)P / 4 + 64 + 63) & ~(size_t)63;
  // the uploaded members of the points live in ONE arena in the order points_set stages them (carve_points): one H2D
  DALLOC(h, hs->p_arena, 23 * n);
  // everything a new point set starts from zero: one arena, one memset
  hs->p_zero_words = 10 * n + 8 * n + n + (n + 4);
  DALLOC(h, hs->p_zero_arena, hs->p_zero_words);
  float *z = hs->p_zero_arena;
  float **zl[] = {&h->p_HddA, &h->p_bdA, &h->p_HddL, &h->p_bdL, &h->p_HdiF, &h->p_bdSumF, &h->p_step, &h->p_idepth_backup, &h->p_idepth_hessian, &h->p_maxRelBaseline};
  for (auto p : zl) { *p = z; z += n; }
  h->p_HcdA = z; z += 4 * n;
  h->p_HcdL = z; z += 4 * n;
  h->p_numGood = (int *)z; z += n;
  h->p_res_begin = (int *)z;
  h->P_alloc = (int)n;
  return SOSBA_OK;
}

// arena of N = round_up(n, 64) points: [u|v|idepth|idepth_zero|priorF|deltaF] float, color[8], weights[8], host int
static void carve_points(sosba *h, size_t N) {
  float *f = HS(h)->p_arena;
  h->p_u = f; h->p_v = f + N; h->p_idepth = f + 2 * N; h->p_idepth_zero = f + 3 * N; h->p_priorF = f + 4 * N; h->p_deltaF = f + 5 * N;
  h->p_color = f + 6 * N; h->p_weights = f + 14 * N; h->p_host = (int *)(f + 22 * N);
}

API int sosba_points_set(sosba_t *h, const sosba_points *p) {
  CHECK_H(h);
  if (!p || p->n < 0) return SOSBA_E_ARG;
  if (p->n > 0 && (!p->u || !p->v || !p->idepth || !p->idepth_zero || !p->color || !p->weights || !p->host)) {
    sosba_set_error("points_set: u, v, idepth, idepth_zero, color, weights and host are mandatory");
    return SOSBA_E_ARG;
  }
  int rc = ensure_points(h, p->n);
  if (rc) return rc;
  const size_t n = p->n;
  h->P = p->n;
  HostSide *hs = HS(h);
  hs->p_host.assign(p->host, p->host + n);
  {
    const size_t N = (n + 63) & ~(size_t)63, bytes = 23 * N * sizeof(float);
    carve_points(h, N);
    if (n > 0) {
      char *blk = nullptr;
      if (bytes <= hs->stage_cap / 2) { if ((rc = stage_reserve(h, bytes, &blk))) return rc; }
      else { hs->big_stage.resize(bytes); blk = hs->big_stage.data(); }
      float *sf = (float *)blk;
      memcpy(sf, p->u, 4 * n); memcpy(sf + N, p->v, 4 * n); memcpy(sf + 2 * N, p->idepth, 4 * n); memcpy(sf + 3 * N, p->idepth_zero, 4 * n);
      if (p->priorF) memcpy(sf + 4 * N, p->priorF, 4 * n); else memset(sf + 4 * N, 0, 4 * n);
      if (p->deltaF) memcpy(sf + 5 * N, p->deltaF, 4 * n); else memset(sf + 5 * N, 0, 4 * n);
      memcpy(sf + 6 * N, p->color, 32 * n); memcpy(sf + 14 * N, p->weights, 32 * n); memcpy(sf + 22 * N, p->host, 4 * n);
      SOSBA_CUDA(cudaMemcpyAsync(hs->p_arena, blk, bytes, cudaMemcpyHostToDevice, h->stream));
      if (blk == hs->big_stage.data()) { if ((rc = sync(h))) return rc; }
    }
  }
  cudaMemsetAsync(hs->p_zero_arena, 0, hs->p_zero_words * 4, h->stream);   // accumulators, steps, CSR
  hs->res_begin.assign(n + 1, 0);
  h->R = 0;
  return SOSBA_OK;
}

API int sosba_points_update(sosba_t *h, const float *idepth, const float *idepth_zero, const float *deltaF) {
  CHECK_H(h);
  int rc;
  if (idepth && (rc = up(h, h->p_idepth, idepth, h->P))) return rc;
  if (idepth_zero && (rc = up(h, h->p_idepth_zero, idepth_zero, h->P))) return rc;
  if (deltaF && (rc = up(h, h->p_deltaF, deltaF, h->P))) return rc;
  return SOSBA_OK;
}

static int ensure_residuals(sosba *h, int R) {
  if (R <= h->R_alloc) return SOSBA_OK;
  HostSide *hs = HS(h);
  h->R_alloc = 0;   // an allocation failure below must not leave the old capacity standing over freed buffers
  dfree(h, hs->r_arena); dfree(h, h->r_by_block);
  dfree(h, h->r_J[0]); dfree(h, h->r_J[1]); dfree(h, h->r_rec); dfree(h, h->r_rtz); dfree(h, h->r_proj); dfree(h, h->r_center); dfree(h, h->d_newE);
  const size_t n = ((size_t)R + (size_t)R / 4 + 64 + 63) & ~(size_t)63;
  // ids, energies and flags of the residuals live in ONE arena in the order residuals_set stages them: one H2D per upload
  DALLOC(h, hs->r_arena, 31 * n);
  DALLOC(h, h->r_by_block, n);
  DALLOC(h, h->r_J[0], n * SOSBA_JREC); DALLOC(h, h->r_J[1], n * SOSBA_JREC); DALLOC(h, h->r_rec, n * SOSBA_CREC);
  DALLOC(h, h->r_rtz, n * 8); DALLOC(h, h->r_proj, n * 16); DALLOC(h, h->r_center, n * 3); DALLOC(h, h->d_newE, n);
  h->R_alloc = (int)n;
  return SOSBA_OK;
}
// arena of N = round_up(n, 64) residuals.  Uploaded part (16 N bytes, one H2D): [point|target] int, [energy] float, the byte
// planes [state|is_lin|is_active|is_new]; derived part, filled on the device by k_residual_init: [host] int,
// [new_energy|new_energy_wo] float, the byte planes [new_state|sel|dropped]
static void carve_residuals(sosba *h, size_t N) {
  unsigned char *b = HS(h)->r_arena;
  h->r_point = (int *)b; h->r_target = (int *)(b + 4 * N);
  h->r_energy = (float *)(b + 8 * N);
  unsigned char *u = b + 12 * N;
  h->r_state = u; h->r_is_lin = u + N; h->r_is_active = u + 2 * N; h->r_is_new = u + 3 * N;
  unsigned char *d = b + 16 * N;
  h->r_host = (int *)d; h->r_new_energy = (float *)(d + 4 * N); h->r_new_energy_wo = (float *)(d + 8 * N);
  h->r_new_state = d + 12 * N; h->r_sel = d + 13 * N; h->r_dropped = d + 14 * N;
}

static void clear_gathered_energies(sosba *h);
static LinArgs lin_args(sosba *h);
API int sosba_residuals_set(sosba_t *h, const sosba_residuals *r) {
  CHECK_H(h);
  if (!r || r->n < 0 || h->nf <= 0) { sosba_set_error("residuals_set needs window_set/points_set first"); return SOSBA_E_STATE; }
  if (r->n > 0 && (!r->point || !r->target)) { sosba_set_error("residuals_set: point and target are mandatory"); return SOSBA_E_ARG; }
  int rc = ensure_residuals(h, r->n);
  if (rc) return rc;
  HostSide *hs = HS(h);
  const int n = r->n, nf = h->nf, P = h->P;
  h->R = n;
  hs->r_host_copy = false;   // the host mirror of point / target (marginalisation only) is fetched on demand
  std::vector<int> &host = hs->r_host_tmp;   // scratch vectors live in the handle: no allocation per keyframe
  host.resize(n);
  hs->res_begin.assign(P + 1, 0);
  // one pass: validation, host of every residual, CSR counts, and "no point sees a target twice" (what the fused accumulation needs)
  // (run by run: the residuals of a point are consecutive, so the host lookup, the CSR count and the duplicate-target test
  // are per point, and the inner loop only checks the target range)
  bool twice = false;
  {
    const int32_t *pt = r->point, *tg = r->target;
    const int *p_host = hs->p_host.data();
    int *rbeg = hs->res_begin.data();
    int prev = -1;
    std::vector<int> &seen = hs->seen_tmp;   // only for windows of more than 64 frames
    if (nf > 64) seen.assign(nf, -1);
    for (int i = 0; i < n;) {
      const int p = pt[i];
      if (p < 0 || p <= prev || p >= P) {
        sosba_set_error("residual %d: point %d / target %d invalid or not point-major", i, p, tg[i]);
        return SOSBA_E_ARG;
      }
      prev = p;
      const int hst = p_host[p];
      if (hst < 0 || hst >= nf) { sosba_set_error("point %d: host %d invalid", p, hst); return SOSBA_E_ARG; }
      unsigned long long mask = 0ull;
      int j = i;
      for (; j < n && pt[j] == p; j++) {
        const int t = tg[j];
        if (t < 0 || t >= nf) { sosba_set_error("residual %d: point %d / target %d invalid or not point-major", j, p, t); return SOSBA_E_ARG; }
        if (nf <= 64) { const unsigned long long bit = 1ull << t; twice |= (mask & bit) != 0; mask |= bit; }
        else { twice |= seen[t] == p; seen[t] = p; }
        host[j] = hst;
      }
      rbeg[p + 1] = j - i;
      i = j;
    }
    for (int p = 0; p < P; p++) rbeg[p + 1] += rbeg[p];
  }
  {  // the fused accumulation stages the residuals of up to 32 consecutive points of ONE host and lists them per target
    hs->fused_acc_ok = !twice;
    hs->max_res_per_tile = 0;
    std::vector<int> &tiles = hs->tiles_tmp;
    tiles.clear();
    for (int p0 = 0; p0 < P;) {
      int np = 1;
      while (np < 32 && p0 + np < P && hs->p_host[p0 + np] == hs->p_host[p0]) np++;
      const int rb = hs->res_begin[p0], nres = hs->res_begin[p0 + np] - rb;
      tiles.insert(tiles.end(), {p0, np, rb, nres});
      hs->max_res_per_tile = std::max(hs->max_res_per_tile, nres);
      p0 += np;
    }
    hs->n_tiles = (int)tiles.size() / 4;
    if ((int)tiles.size() > hs->tiles_cap) {
      dfree(h, hs->d_tiles);
      DALLOC(h, hs->d_tiles, tiles.size() * 2 + 64);
      hs->tiles_cap = (int)tiles.size() * 2 + 64;
    }
    if (!tiles.empty() && (rc = up(h, hs->d_tiles, tiles.data(), tiles.size()))) return rc;
    hs->th_pending = false;
    hs->tables_clean = false;
  }
  if (h->comm && h->world > 1) {   // point shards: room for every rank's newest-frame energies (one per local point at most)
    int cap = 0;
    if (h->p2p) {   // peer mailboxes: a fixed capacity, so no collective (and no synchronisation) per keyframe
      cap = SOSBA_XCHG_MAX_NEWE;
      if (P > cap) { sosba_set_error("%d points in one shard: the peer exchange carries at most %d (set SOSBA_COMM_NCCL=1)", P, cap); return SOSBA_E_ARG; }
    } else if ((rc = sosba_comm_max_int(h, ((P + 63) / 64 + 1) * 64, &cap))) return rc;
    if (cap > h->newE_cap) {
      dfree(h, h->d_newE_all);
      DALLOC(h, h->d_newE_all, (size_t)h->world * cap + h->world);
      h->newE_cap = cap;
    }
    h->d_newE_cnt = (int *)(h->d_newE_all + (size_t)h->world * h->newE_cap);
    clear_gathered_energies(h);
  }
  hs->n_lin = 0;
  if (n > 0) {
    const size_t N = ((size_t)n + 63) & ~(size_t)63, bytes = 16 * N;   // the uploaded half of the arena (carve_residuals)
    carve_residuals(h, N);
    char *blk = nullptr;
    if (bytes <= hs->stage_cap / 2) { if ((rc = stage_reserve(h, bytes, &blk))) return rc; }
    else { hs->big_stage.resize(bytes); blk = hs->big_stage.data(); }
    int *si = (int *)blk;
    float *sf = (float *)(blk + 8 * N);
    unsigned char *su = (unsigned char *)blk + 12 * N;
    memcpy(si, r->point, 4 * (size_t)n); memcpy(si + N, r->target, 4 * (size_t)n);
    if (r->state_energy) memcpy(sf, r->state_energy, 4 * (size_t)n);
    else memset(sf, 0, 4 * N);
    auto flag = [&](unsigned char *dst, const uint8_t *src, uint8_t dflt) { if (src) memcpy(dst, src, n); else memset(dst, dflt, n); };
    flag(su, r->state, SOSBA_RES_IN); flag(su + N, r->is_linearized, 0); flag(su + 2 * N, r->is_active, 0); flag(su + 3 * N, r->is_new, 1);
    SOSBA_CUDA(cudaMemcpyAsync(hs->r_arena, blk, bytes, cudaMemcpyHostToDevice, h->stream));
    if (blk == hs->big_stage.data()) { if ((rc = sync(h))) return rc; }
    // host of every residual, state_NewEnergy = state_energy, state_NewEnergyWithOutlier = -1 (Residuals.cpp:78), new state
    // OUTLIER, nothing committed, nothing dropped: derived on the device instead of being staged and copied
    launch_residual_init(h, lin_args(h), h->p_host);
  }
  if ((rc = up(h, h->p_res_begin, hs->res_begin.data(), P + 1))) return rc;
  if (!hs->fused_acc_ok) {   // only the un-fused accumulation walks the residuals in (host, target)-block order
    std::vector<int> by_block(n), cnt(nf * nf + 1, 0);
    for (int i = 0; i < n; i++) cnt[host[i] + r->target[i] * nf + 1]++;
    for (int b = 0; b < nf * nf; b++) cnt[b + 1] += cnt[b];
    for (int i = 0; i < n; i++) by_block[cnt[host[i] + r->target[i] * nf]++] = i;   // stable counting sort by block
    if ((rc = up(h, h->r_by_block, by_block.data(), n))) return rc;
  }
  if (r->is_linearized) for (int i = 0; i < n; i++) hs->n_lin += r->is_linearized[i] ? 1 : 0;
  SOSBA_CUDA(cudaGetLastError());
  return SOSBA_OK;
}

static LinArgs lin_args(sosba *h) {
  LinArgs a;
  a.R = h->R; a.nf = h->nf;
  a.r_point = h->r_point; a.r_target = h->r_target; a.r_host = h->r_host;
  a.r_state = h->r_state; a.r_new_state = h->r_new_state; a.r_is_lin = h->r_is_lin; a.r_is_active = h->r_is_active; a.r_is_new = h->r_is_new;
  a.r_sel = h->r_sel; a.r_dropped = h->r_dropped;
  a.r_energy = h->r_energy; a.r_new_energy = h->r_new_energy; a.r_new_energy_wo = h->r_new_energy_wo;
  a.J0 = h->r_J[0]; a.J1 = h->r_J[1]; a.rec = h->r_rec; a.rtz = h->r_rtz; a.proj = h->r_proj; a.center = h->r_center;
  a.p_u = h->p_u; a.p_v = h->p_v; a.p_idepth = h->p_idepth; a.p_idepth_zero = h->p_idepth_zero; a.p_color = h->p_color; a.p_weights = h->p_weights;
  a.p_deltaF = h->p_deltaF; a.p_maxRelBaseline = h->p_maxRelBaseline; a.p_numGood = h->p_numGood;
  a.precalc = h->d_precalc; a.frameEnergyTH = h->d_frameEnergyTH; a.calib = h->d_calib; a.adHTdeltaF = h->d_adHTdeltaF;
  for (int i = 0; i < 16; i++) a.img[i] = i < h->nf ? h->slot_img[h->frame_slot[i]] : nullptr;
  a.w = h->cfg.w; a.wM3G = (float)(h->cfg.w - 3); a.hM3G = (float)(h->cfg.h - 3);
  a.huberTH = h->cfg.huber_th; a.outlierTHSum = h->cfg.outlier_th_sum_component; a.affModeA = h->cfg.affine_opt_mode_a; a.affModeB = h->cfg.affine_opt_mode_b;
  a.stats = h->d_stats; a.counts = h->d_counts;
  if (h->d_newE_all) {   // point shards: this rank's segment of the gathered list
    a.newE = h->d_newE_all + (size_t)h->rank * h->newE_cap; a.newE_count = h->d_newE_cnt + h->rank;
    a.th.newE = h->d_newE_all; a.th.seg_counts = h->d_newE_cnt; a.th.nseg = h->world; a.th.seg_stride = h->newE_cap;
  } else {
    a.newE = h->d_newE; a.newE_count = h->d_counts + 4;
    a.th.newE = h->d_newE; a.th.seg_counts = h->d_counts + 4; a.th.nseg = 1; a.th.seg_stride = 0;
  }
  a.th.frameEnergyTH = h->d_frameEnergyTH; a.th.nf = h->nf;
  a.th.thN = h->cfg.frame_energy_th_n; a.th.thFacMedian = h->cfg.frame_energy_th_fac_median; a.th.thConstWeight = h->cfg.frame_energy_th_const_weight;
  a.th.overallWeight = h->cfg.overall_energy_th_weight; a.th.thOut = h->d_thOut;
  a.ticket = h->d_counts + 12;
  a.gate = nullptr; a.zero_buf = nullptr; a.zero_n = 0; a.opaque_zero = 0u; a.trace = nullptr;
  return a;
}

// ---- a3/a4/a5 -----------------------------------------------------------------------------------
API int sosba_reset_oob(sosba_t *h) {
  CHECK_H(h);
  launch_reset_oob(h, lin_args(h));
  SOSBA_CUDA(cudaGetLastError());
  return SOSBA_OK;
}

// the threshold selection of a fused linearisation that no accumulation picked up yet
static void clear_gathered_energies(sosba *h) {   // point shards: the list must be empty before the next append
  if (!h->d_newE_all) return;
  // peer mailboxes: only the lengths (the exchange overwrites the other ranks' segments); NCCL: a sum with zeros is the
  // concatenation, so every segment has to be zero
  if (h->p2p) cudaMemsetAsync(h->d_newE_cnt, 0, (size_t)h->world * 4, h->stream);
  else cudaMemsetAsync(h->d_newE_all, 0, ((size_t)h->world * h->newE_cap + h->world) * 4, h->stream);
}

// point shards: the sums of a linearizeAll outside the loop (with_stats) and the newest-frame energies over the ranks
static int exchange_lin(sosba *h, int with_stats, double *extra = nullptr, int n_extra = 0) {
  if (!h->comm || h->world <= 1) return SOSBA_OK;
  StitchXchgArgs x;
  memset(&x, 0, sizeof(x));
  sosba_xchg_args(h, &x, 1);
  if (!x.push) return sosba_allreduce_lin(h, with_stats, extra, n_extra);
  HostSide *hs = HS(h);
  x.gate = hs->gate; x.err = hs->d_ctl + 2;
  launch_lin_xchg(h, x, h->d_stats, h->d_counts, with_stats, h->P, extra, n_extra);
  return SOSBA_OK;
}

static void flush_pending_th(sosba *h) {
  HostSide *hs = HS(h);
  if (!hs->th_pending) return;
  exchange_lin(h, 0);
  launch_energy_th(h, lin_args(h).th, hs->gate);
  clear_gathered_energies(h);
  hs->th_pending = false;
}

// enqueue linearizeAll on the stream (no host sync)
static void enqueue_linearize(sosba *h, int fix, bool clear_sums = true) {
  flush_pending_th(h);
  if (clear_sums) cudaMemsetAsync(h->d_stats, 0, 2 * sizeof(double) + 5 * sizeof(int), h->stream);
  clear_gathered_energies(h);
  LinArgs a = lin_args(h);
  a.trace = sosba_trace_slot("k_linearize (API)");
  if (fix) launch_linearize_fix(h, a);   // + applyRes(true) + removal / baseline bookkeeping in the same launch
  else launch_linearize(h, a);          // (the bench roofline brackets the fused launches of the loop, enqueue_linearize_apply)
  exchange_lin(h, 1);   // point shards: global energy, state histogram, removals, newest-frame energies
  launch_energy_th(h, a.th);
  clear_gathered_energies(h);
}

static int read_linearize_out(sosba *h, sosba_linearize_out *out) {
  HostSide *hs = HS(h);
  int rc;
  flush_pending_th(h);
  if ((rc = down(h, hs->pin_d, h->d_stats, 12))) return rc;   // energy | pad | counts[16] | thOut[4]
  const bool sharded = h->comm && h->world > 1 && hs->d_ctl;
  if (sharded && (rc = down(h, hs->pin_i + 8, hs->d_ctl + 2, 1))) return rc;
  if ((rc = sync(h))) return rc;
  if (sharded && (hs->pin_i[8] & 2)) return solve_flag_error(hs->pin_i[8]);
  if (out) {
    const int *ci = (const int *)(hs->pin_d + 2);
    out->energy = hs->pin_d[0];
    out->new_frame_energy_th = ((const float *)(ci + 16))[0];
    out->n_in = ci[0]; out->n_oob = ci[1]; out->n_outlier = ci[2]; out->n_removed = ci[3];
    out->reserved0 = 0;
  }
  return SOSBA_OK;
}

API int sosba_linearize_all(sosba_t *h, int32_t fix, sosba_linearize_out *out) {
  CHECK_H(h);
  if (h->nf <= 0) { sosba_set_error("no window"); return SOSBA_E_STATE; }
  if (h->comm && h->world > 1 && HS(h)->d_ctl) cudaMemsetAsync(HS(h)->d_ctl + 2, 0, sizeof(int), h->stream);   // exchange error word
  enqueue_linearize(h, fix != 0);
  SOSBA_CUDA(cudaGetLastError());
  return read_linearize_out(h, out);
}

API int sosba_apply_res(sosba_t *h) {
  CHECK_H(h);
  launch_apply_res(h, lin_args(h), 0);
  SOSBA_CUDA(cudaGetLastError());
  return SOSBA_OK;
}

static int upload_ids(sosba *h, const int32_t *ids, int n) {
  HostSide *hs = HS(h);
  if (n > hs->ids_cap) {
    dfree(h, hs->d_ids);
    DALLOC(h, hs->d_ids, (size_t)n * 2 + 64);
    hs->ids_cap = n * 2 + 64;
  }
  int rc = up(h, hs->d_ids, ids, n);
  if (rc) return rc;
  return sync(h);
}

API int sosba_fix_linearization(sosba_t *h, const int32_t *ids, int32_t n) {
  CHECK_H(h);
  if (n < 0 || (n > 0 && !ids)) return SOSBA_E_ARG;
  HostSide *hs = HS(h);
  for (int i = 0; i < n; i++) if (ids[i] < 0 || ids[i] >= h->R) { sosba_set_error("residual id %d", ids[i]); return SOSBA_E_ARG; }
  int rc = upload_ids(h, ids, n);
  if (rc) return rc;
  launch_fix_linearization(h, lin_args(h), hs->d_ids, n);
  hs->n_lin += n;
  SOSBA_CUDA(cudaGetLastError());
  return sync(h);
}

// ---- read-back ----------------------------------------------------------------------------------
template <class T> static int fetch(sosba *h, T *dst, const T *src, size_t n) {
  if (!dst) return SOSBA_OK;
  return down(h, dst, src, n);
}

API int sosba_residuals_get_state(sosba_t *h, uint8_t *state, uint8_t *new_state, float *energy, float *new_energy, float *new_energy_wo,
                                  uint8_t *is_active, uint8_t *is_linearized) {
  CHECK_H(h);
  const size_t n = h->R;
  int rc;
  if ((rc = fetch(h, state, h->r_state, n)) || (rc = fetch(h, new_state, h->r_new_state, n)) || (rc = fetch(h, energy, h->r_energy, n)) ||
      (rc = fetch(h, new_energy, h->r_new_energy, n)) || (rc = fetch(h, new_energy_wo, h->r_new_energy_wo, n)) ||
      (rc = fetch(h, is_active, h->r_is_active, n)) || (rc = fetch(h, is_linearized, h->r_is_lin, n)))
    return rc;
  return sync(h);
}

API int sosba_residuals_get_jacobians(sosba_t *h, int32_t committed, float *J) {
  CHECK_H(h);
  if (!J) return SOSBA_E_ARG;
  const size_t n = h->R;
  std::vector<float> j0(n * SOSBA_JREC), j1(n * SOSBA_JREC);
  std::vector<uint8_t> sel(n);
  int rc;
  if ((rc = down(h, j0.data(), h->r_J[0], n * SOSBA_JREC)) || (rc = down(h, j1.data(), h->r_J[1], n * SOSBA_JREC)) || (rc = down(h, sel.data(), h->r_sel, n)))
    return rc;
  if ((rc = sync(h))) return rc;
  for (size_t i = 0; i < n; i++) {
    const int which = committed ? 1 - sel[i] : sel[i];  // PointFrameResidual::J == J[sel], EFResidual::J == J[1-sel]
    const float *s = (which ? j1.data() : j0.data()) + i * SOSBA_JREC;
    float *o = J + i * SOSBA_J_FLOATS;
    int k = 0;
    for (int q = 0; q < 8; q++) o[k++] = s[JR_RES + q];
    for (int q = 0; q < 12; q++) o[k++] = s[JR_JPDXI0 + q];
    for (int q = 0; q < 8; q++) o[k++] = s[JR_JPDC0 + q];
    o[k++] = s[JR_JPDD]; o[k++] = s[JR_JPDD + 1];
    for (int q = 0; q < 16; q++) o[k++] = s[JR_JIDX0 + q];
    for (int q = 0; q < 16; q++) o[k++] = s[JR_JAB0 + q];
    for (int q = 0; q < 12; q++) o[k++] = s[JR_JIDX2 + q];
  }
  return SOSBA_OK;
}

API int sosba_residuals_get_aux(sosba_t *h, float *JpJdF, float *rtz, float *proj, float *center) {
  CHECK_H(h);
  const size_t n = h->R;
  int rc;
  std::vector<float> rec;
  if (JpJdF) { rec.resize(n * SOSBA_CREC); if ((rc = down(h, rec.data(), h->r_rec, n * SOSBA_CREC))) return rc; }
  if ((rc = fetch(h, rtz, h->r_rtz, n * 8)) || (rc = fetch(h, proj, h->r_proj, n * 16)) || (rc = fetch(h, center, h->r_center, n * 3))) return rc;
  if ((rc = sync(h))) return rc;
  if (JpJdF) for (size_t i = 0; i < n; i++) memcpy(JpJdF + 8 * i, rec.data() + i * SOSBA_CREC + CR_JPJDF, 8 * sizeof(float));
  return SOSBA_OK;
}

API int sosba_points_get_stats(sosba_t *h, float *mrb, int32_t *ngr) {
  CHECK_H(h);
  int rc;
  if ((rc = fetch(h, mrb, h->p_maxRelBaseline, h->P)) || (rc = fetch(h, ngr, h->p_numGood, h->P))) return rc;
  return sync(h);
}

API int sosba_points_get_acc(sosba_t *h, float *HddA, float *bdA, float *HcdA, float *HddL, float *bdL, float *HcdL, float *HdiF, float *bdSumF) {
  CHECK_H(h);
  const size_t n = h->P;
  int rc;
  if ((rc = fetch(h, HddA, h->p_HddA, n)) || (rc = fetch(h, bdA, h->p_bdA, n)) || (rc = fetch(h, HcdA, h->p_HcdA, 4 * n)) ||
      (rc = fetch(h, HddL, h->p_HddL, n)) || (rc = fetch(h, bdL, h->p_bdL, n)) || (rc = fetch(h, HcdL, h->p_HcdL, 4 * n)) ||
      (rc = fetch(h, HdiF, h->p_HdiF, n)) || (rc = fetch(h, bdSumF, h->p_bdSumF, n)))
    return rc;
  return sync(h);
}

// ---- a6-a9 --------------------------------------------------------------------------------------
static inline double *Hpart(sosba *h, int which) {
  const int D = 4 + 8 * h->nf;
  if (which == 3) return HS(h)->d_Hfinal;
  return h->d_H + (size_t)which * ((size_t)D * D + D);
}
static inline double *bpart(sosba *h, int which) { const int D = 4 + 8 * h->nf; return Hpart(h, which) + (size_t)D * D; }


static SCArgs sc_args(sosba *h, int mode, const int *plist, int n_plist, int shift) {
  SCArgs s;
  s.P = h->P; s.nf = h->nf; s.D = 4 + 8 * h->nf; s.plist = plist; s.n_plist = n_plist; s.mode = mode; s.shiftPriorToZero = shift;
  s.res_begin = h->p_res_begin; s.r_target = h->r_target; s.p_host = h->p_host;
  s.r_is_lin = h->r_is_lin; s.r_is_active = h->r_is_active; s.r_dropped = h->r_dropped;
  s.rec = h->r_rec; s.HddA = h->p_HddA; s.bdA = h->p_bdA; s.HcdA = h->p_HcdA; s.HddL = h->p_HddL; s.bdL = h->p_bdL; s.HcdL = h->p_HcdL;
  s.priorF = h->p_priorF; s.deltaF = h->p_deltaF; s.HdiF = h->p_HdiF; s.bdSumF = h->p_bdSumF; s.idepth_hessian = h->p_idepth_hessian;
  s.maxRelBaseline = h->p_maxRelBaseline; s.adHostF = h->d_adHostF; s.adTargetF = h->d_adTargetF; s.accSC = h->d_accSC;
  return s;
}

// the block tables of accumulateAF_MT / accumulateLF_MT / accumulateSCF_MT (EnergyFunctional.cpp:197-254):
// one memset, top blocks (A, and L when linearised residuals exist), per-point sums + Schur Gram.
// Point shards: with `nccl_reduce` the un-stitched tables are summed over ranks by NCCL right here (API-level calls, and
// the loop when the peer mailboxes are unavailable) and the pending newest-frame energies ride along; otherwise the caller
// (enqueue_solve) sums the STITCHED system inside k_stitch_xchg.  Either way the pending setNewFrameEnergyTH of a sharded
// window is handed back in `defer_th` so that it runs in the spare CTA of the caller's k_solve launch.
static int enqueue_blocks(sosba *h, bool nccl_reduce, ThArgs *defer_th = nullptr, int *deferred = nullptr) {
  if (deferred) *deferred = 0;
  HostSide *hs = HS(h);
  const int nf = h->nf;
  const size_t n2 = (size_t)nf * nf;
  const bool sharded = h->comm && h->world > 1;
  if (!hs->tables_clean) cudaMemsetAsync(hs->d_scratch, 0, sizeof(double) * hs->scratch_doubles, h->stream);
  // (clean tables: the back-substitution sums of this body are cleared by k_solve)
  hs->tables_clean = false;
  bool fused = false;
  if (hs->fused_acc_ok) {
    if (hs->n_lin > 0) launch_prep_records(h, lin_args(h), 1, nullptr, h->R);
    FusedAccArgs f;
    f.P = h->P; f.nf = nf; f.D = 4 + 8 * nf; f.R = h->R; f.shiftPriorToZero = 1; f.do_th = (hs->th_pending && !sharded) ? 1 : 0;
    f.tiles = (const int4 *)hs->d_tiles;
    f.res_begin = h->p_res_begin; f.r_target = h->r_target; f.p_host = h->p_host;
    f.r_is_lin = h->r_is_lin; f.r_is_active = h->r_is_active; f.r_dropped = h->r_dropped; f.rec = h->r_rec;
    f.accTop = h->d_accTop; f.n_acc = hs->d_cnt;
    f.HddA = h->p_HddA; f.bdA = h->p_bdA; f.HcdA = h->p_HcdA; f.HddL = h->p_HddL; f.bdL = h->p_bdL; f.HcdL = h->p_HcdL;
    f.priorF = h->p_priorF; f.deltaF = h->p_deltaF; f.HdiF = h->p_HdiF; f.bdSumF = h->p_bdSumF; f.idepth_hessian = h->p_idepth_hessian;
    f.maxRelBaseline = h->p_maxRelBaseline; f.adHostF = h->d_adHostF; f.adTargetF = h->d_adTargetF; f.accSC = h->d_accSC;
    f.th = lin_args(h).th; f.gate = hs->gate;
    f.dbg = (g_dbg && getenv("SOSBA_SOLVE_DEBUG")) ? g_dbg + 64 * 32 : nullptr;
    f.trace = sosba_trace_slot("k_accumulate_fused");
    fused = launch_accumulate_fused(h, f, hs->max_res_per_tile, hs->n_tiles);
    if (fused && f.do_th) hs->th_pending = false;
  }
  if (!fused) {
    if (!sharded) flush_pending_th(h);
    AccArgs a;
    a.R = h->R; a.P = h->P; a.nf = nf; a.n_list = h->R; a.list = h->r_by_block; a.mode = 0;
    a.r_point = h->r_point; a.r_target = h->r_target; a.r_host = h->r_host;
    a.r_is_lin = h->r_is_lin; a.r_is_active = h->r_is_active; a.r_dropped = h->r_dropped;
    a.rec = h->r_rec; a.accTop = h->d_accTop; a.n_acc = hs->d_cnt;
    launch_top_accumulate(h, a);
    if (hs->n_lin > 0) {
      launch_prep_records(h, lin_args(h), 1, nullptr, h->R);
      AccArgs l = a;
      l.mode = 1; l.accTop = h->d_accTop + n2 * SOSBA_TOPB; l.n_acc = hs->d_cnt + 1;
      launch_top_accumulate(h, l);
    }
    launch_point_sc(h, sc_args(h, 0, nullptr, 0, 1));
  }
  const bool shard_th = sharded && hs->th_pending;
  if (nccl_reduce) {
    int rc = sosba_allreduce_acc(h, shard_th ? 1 : 0);
    if (rc) return rc;
  }
  if (shard_th) {
    if (defer_th) { *defer_th = lin_args(h).th; *deferred = 1; }
    else if (nccl_reduce) { launch_energy_th(h, lin_args(h).th, hs->gate); hs->th_pending = false; }
  }
  return SOSBA_OK;
}

// API path: the three stitched systems separately (d_H parts 0..2)
static int enqueue_accumulate(sosba *h) {
  const int nf = h->nf;
  const size_t n2 = (size_t)nf * nf;
  const bool only_tables_clean = HS(h)->tables_clean;   // a loop body cleared the tables, not the H parts the stitch adds into
  int rc = enqueue_blocks(h, true);
  if (rc) return rc;
  if (only_tables_clean) cudaMemsetAsync(h->d_H, 0, sizeof(double) * 3 * ((size_t)(4 + 8 * nf) * (4 + 8 * nf) + 4 + 8 * nf), h->stream);
  launch_stitch_top(h, h->d_accTop, h->d_adHost, h->d_adTarget, nf, Hpart(h, 0), bpart(h, 0), 0, h->d_wprior, h->d_calib + 6);
  launch_stitch_top(h, h->d_accTop + n2 * SOSBA_TOPB, h->d_adHost, h->d_adTarget, nf, Hpart(h, 1), bpart(h, 1), 1, h->d_wprior, h->d_calib + 6);
  launch_finalize_sc(h, h->d_accSC, nf, Hpart(h, 2), bpart(h, 2));
  SOSBA_CUDA(cudaGetLastError());
  return SOSBA_OK;
}

API int sosba_accumulate(sosba_t *h, double *HA, double *bA, double *HL, double *bL, double *Hsc, double *bsc, int32_t *resInA, int32_t *resInL) {
  CHECK_H(h);
  if (h->nf <= 0) return SOSBA_E_STATE;
  int rc = enqueue_accumulate(h);
  if (rc) return rc;
  const int D = 4 + 8 * h->nf;
  HostSide *hs = HS(h);
  if ((rc = fetch(h, HA, Hpart(h, 0), (size_t)D * D)) || (rc = fetch(h, bA, bpart(h, 0), D)) || (rc = fetch(h, HL, Hpart(h, 1), (size_t)D * D)) ||
      (rc = fetch(h, bL, bpart(h, 1), D)) || (rc = fetch(h, Hsc, Hpart(h, 2), (size_t)D * D)) || (rc = fetch(h, bsc, bpart(h, 2), D)) ||
      (rc = down(h, hs->pin_i, hs->d_cnt, 2)))
    return rc;
  if ((rc = sync(h))) return rc;
  if (resInA) *resInA = hs->pin_i[0];
  if (resInL) *resInL = hs->pin_i[1];
  return SOSBA_OK;
}

// ---- a10/a11 ------------------------------------------------------------------------------------
// the device error word of a solve (d_ctl[2]): k_solve raises 1 for a non-finite step (the reference: isLost), the peer
// exchange raises bit 1 when a rank's words did not arrive in time (its results are then not stored)
static int solve_flag_error(int flag) {
  if (flag & 2) { sosba_set_error("point-shard exchange timed out: a peer rank did not deliver its part of the system"); return SOSBA_E_NCCL; }
  sosba_set_error("non-finite solution");
  return SOSBA_E_NONFINITE;
}
static ResubArgs resub_args(sosba *h, int do_step) {
  ResubArgs r;
  r.P = h->P; r.nf = h->nf; r.res_begin = h->p_res_begin; r.r_target = h->r_target; r.p_host = h->p_host;
  r.r_is_active = h->r_is_active; r.r_dropped = h->r_dropped; r.rec = h->r_rec; r.xAd = h->d_xAd;
  r.HcdA = h->p_HcdA; r.HcdL = h->p_HcdL; r.bdSumF = h->p_bdSumF; r.HdiF = h->p_HdiF; r.step = h->p_step;
  r.do_step = do_step; r.idepth = h->p_idepth; r.idepth_zero = h->p_idepth_zero; r.idepth_backup = h->p_idepth_backup; r.deltaF = h->p_deltaF;
  r.stats = HS(h)->d_rstats + 4 * HS(h)->rstats_par - 1;   // the kernel writes stats[1..3]
  r.gate = HS(h)->gate; r.zero_lin = nullptr; r.zero_newE = nullptr; r.zero_newE_n = 0; r.trace = nullptr;
  return r;
}

static StepArgs step_args(sosba *h) {
  HostSide *hs = HS(h);
  StepArgs st;
  st.nf = h->nf; st.stepfac = 1.0f; st.x = h->d_x; st.fs = hs->d_fs; st.cs = hs->d_cs;
  st.precalc = h->d_precalc; st.adHTdeltaF = h->d_adHTdeltaF; st.calib = h->d_calib;
  st.adHostF = h->d_adHostF; st.adTargetF = h->d_adTargetF; st.wprior = h->d_wprior; st.iter = hs->d_iter;
  st.adHost = h->d_adHost; st.adTarget = h->d_adTarget; st.trace = nullptr;
  return st;
}

// accumulate + stitch + solve + back-substitution; with do_step also backupState / doStepFromBackup of the points (in
// k_resubstitute) and of the frames, calibration, precalc and deltas (fused into the tail of k_solve)
static int enqueue_solve(sosba *h, const double *d_HM, const double *d_bM, int do_step, bool want_final = false) {
  HostSide *hs = HS(h);
  const int nf = h->nf, D = 4 + 8 * nf;
  ThArgs th_def = {};
  int th_deferred = 0;
  StitchXchgArgs x;
  memset(&x, 0, sizeof(x));
  sosba_xchg_args(h, &x, hs->th_pending ? 1 : 0);   // push = 1: the ranks' stitched systems are summed over peer memory
  const bool sharded = h->comm && h->world > 1;
  int rc = enqueue_blocks(h, sharded && !x.push, &th_def, &th_deferred);
  if (rc) return rc;
  x.nf = nf; x.D = D; x.accTop = h->d_accTop; x.adHost = h->d_adHost; x.adTarget = h->d_adTarget;
  x.H = Hpart(h, 0); x.b = bpart(h, 0); x.accSC = h->d_accSC; x.rstats = h->d_rstats_all; x.cnt = h->d_cnt_all;
  x.gate = hs->gate; x.err = hs->d_ctl + 2;
  static const bool xchg_debug = getenv("SOSBA_XCHG_DEBUG") != nullptr;
  if (xchg_debug) {   // globaltimer stamps of the last 64 launches, printed by sosba_destroy
    if (!g_xdbg) { cudaMalloc(&g_xdbg, 64 * 8 * sizeof(long long)); cudaMemset(g_xdbg, 0, 64 * 8 * sizeof(long long)); }
    x.dbg = g_xdbg + 8 * (g_xdbg_n++ % 64);
  }
  x.trace = sosba_trace_slot("k_stitch_xchg");
  if ((rc = launch_stitch_xchg(h, x, h->P))) return rc;
  if (th_deferred) hs->th_pending = false;   // runs in the spare CTA of the solve launch below
  SolveArgs s;
  s.th = th_def; s.do_th = th_deferred;
  s.nf = nf; s.D = D;
  s.Htop = Hpart(h, 0); s.btop = bpart(h, 0); s.accSC = h->d_accSC;
  s.HM = d_HM; s.bM = d_bM; s.wprior = h->d_wprior; s.cDeltaF = h->d_calib + 6;
  s.x = h->d_x; s.Hfinal = want_final ? Hpart(h, 3) : nullptr; s.bfinal = want_final ? bpart(h, 3) : nullptr;
  s.adHostF = h->d_adHostF; s.adTargetF = h->d_adTargetF; s.xAd = h->d_xAd; s.status = hs->d_ctl + 2;
  s.dbg = nullptr; s.trace = sosba_trace_slot("k_solve");
  if (do_step) s.xAd = nullptr;   // the step launch builds xAd per CTA
  s.do_step = 0;            // the frame step runs in the spare CTA of the step launch, beside the back-substitution
  s.step = step_args(h);    // (k_solve still reads step.iter: the step norms of the previous body, for the loop latch)
  s.stage_sc = s.stage_hm = 0;
  s.ctl = hs->gate ? hs->d_ctl : nullptr; s.iter_index = hs->loop_iter; s.min_it = h->cfg.min_opt_iterations; s.th_opt = h->cfg.th_opt_iterations;
  s.prev_rstats = hs->d_rstats + 4 * (hs->rstats_par ^ 1);
  s.res_in = hs->d_cnt; s.res_out = hs->d_ctl + 3;
  s.zero_rstats = hs->d_rstats + 4 * hs->rstats_par;
  s.stash_src = nullptr; s.stash_dst = nullptr;
  if (hs->stash_pending) { s.stash_src = h->d_stats; s.stash_dst = hs->d_stash; hs->stash_pending = false; }
  static const bool solve_debug = getenv("SOSBA_SOLVE_DEBUG") != nullptr;
  if (solve_debug) {   // phase timestamps of the last 64 launches, no host sync: read back and printed by sosba_destroy
    if (!g_dbg) cudaMalloc(&g_dbg, (64 * 32 + 16) * sizeof(long long));
    s.dbg = g_dbg + 32 * (g_dbg_n++ % 64);
  }
  if ((rc = launch_solve(h, s))) return rc;
  {
    ResubArgs ra = resub_args(h, do_step);
    if (do_step) {   // the fused linearisation follows: no memsets on the stream
      ra.zero_lin = h->d_stats;
      if (h->d_newE_all) {
        if (h->p2p) { ra.zero_newE = (float *)h->d_newE_cnt; ra.zero_newE_n = h->world; }
        else { ra.zero_newE = h->d_newE_all; ra.zero_newE_n = h->world * h->newE_cap + h->world; }
      }
    }
    if (do_step) {
      StepArgs sa = step_args(h);
      ra.trace = sosba_trace_slot("k_step (points)");
      if (ra.trace) { sosba_trace_slot("k_step (frame CTA)"); sa.trace = sosba_trace_slot("%  its phases"); }
      launch_step(h, ra, sa);
    }
    else launch_resubstitute(h, ra);
  }
  SOSBA_CUDA(cudaGetLastError());
  return SOSBA_OK;
}

API int sosba_solve_system(sosba_t *h, const double *HM, const double *bM, double *x, double *Hf, double *bf) {
  CHECK_H(h);
  if (h->nf <= 0) return SOSBA_E_STATE;
  HostSide *hs = HS(h);
  const int D = 4 + 8 * h->nf;
  int rc;
  const bool prior = HM && bM;
  if (prior) {
    if ((rc = up(h, hs->d_HMtmp, HM, (size_t)D * D)) || (rc = up(h, hs->d_bMtmp, bM, D))) return rc;
    if ((rc = sync(h))) return rc;
  }
  cudaMemsetAsync(hs->d_ctl + 2, 0, sizeof(int), h->stream);
  if ((rc = enqueue_solve(h, prior ? hs->d_HMtmp : nullptr, prior ? hs->d_bMtmp : nullptr, 0, Hf || bf))) return rc;
  if ((rc = fetch(h, x, h->d_x, D)) || (rc = fetch(h, Hf, Hpart(h, 3), (size_t)D * D)) || (rc = fetch(h, bf, bpart(h, 3), D)) ||
      (rc = down(h, hs->pin_i, hs->d_ctl + 2, 1)))
    return rc;
  if ((rc = sync(h))) return rc;
  if (hs->pin_i[0]) return solve_flag_error(hs->pin_i[0]);
  return SOSBA_OK;
}

API int sosba_resubstitute(sosba_t *h, const double *x, float *step) {
  CHECK_H(h);
  if (!x || h->nf <= 0) return SOSBA_E_ARG;
  const int D = 4 + 8 * h->nf;
  int rc;
  if ((rc = up(h, h->d_x, x, D))) return rc;
  if ((rc = sync(h))) return rc;
  launch_make_xad(h, h->d_x, h->nf, h->d_adHostF, h->d_adTargetF, h->d_xAd);
  launch_resubstitute(h, resub_args(h, 0));
  SOSBA_CUDA(cudaGetLastError());
  if ((rc = fetch(h, step, h->p_step, h->P))) return rc;
  return sync(h);
}

// ---- marginalizePointsF ---------------------------------------------------------------------------
API int sosba_marginalize_points(sosba_t *h, const int32_t *ids, int32_t n, double *H, double *b, int32_t *resInM) {
  CHECK_H(h);
  if (n < 0 || (n > 0 && !ids) || !H || !b || h->nf <= 0) return SOSBA_E_ARG;
  HostSide *hs = HS(h);
  const int nf = h->nf, D = 4 + 8 * nf;
  const size_t HB = (size_t)D * D + D;
  if (!hs->r_host_copy && h->R > 0) {   // host mirror of point / target for the block ordering below (not kept per upload)
    hs->r_point.resize(h->R); hs->r_target.resize(h->R);
    int rc0;
    if ((rc0 = fetch(h, hs->r_point.data(), h->r_point, h->R)) || (rc0 = fetch(h, hs->r_target.data(), h->r_target, h->R)) || (rc0 = sync(h))) return rc0;
    hs->r_host_copy = true;
  }
  // residual list of the chosen points, ordered by block
  std::vector<int> rl;
  for (int i = 0; i < n; i++) {
    if (ids[i] < 0 || ids[i] >= h->P) { sosba_set_error("point id %d", ids[i]); return SOSBA_E_ARG; }
    for (int r = hs->res_begin[ids[i]]; r < hs->res_begin[ids[i] + 1]; r++) rl.push_back(r);
  }
  std::stable_sort(rl.begin(), rl.end(), [&](int a, int c) {
    return hs->p_host[hs->r_point[a]] + hs->r_target[a] * nf < hs->p_host[hs->r_point[c]] + hs->r_target[c] * nf;
  });
  std::vector<int> all(ids, ids + n);
  all.insert(all.end(), rl.begin(), rl.end());
  int rc = upload_ids(h, all.data(), (int)all.size());
  if (rc) return rc;
  const int *d_pl = hs->d_ids, *d_rl = hs->d_ids + n;
  launch_scale_prior(h, h->p_priorF, d_pl, n, h->cfg.idepth_fix_prior_marg_fac);  // EnergyFunctional.cpp:901
  cudaMemsetAsync(hs->d_scratch, 0, sizeof(double) * hs->scratch_doubles, h->stream);
  LinArgs la = lin_args(h);
  launch_prep_records(h, la, 2, d_rl, (int)rl.size());
  AccArgs a;
  a.R = h->R; a.P = h->P; a.nf = nf; a.n_list = (int)rl.size(); a.list = d_rl; a.mode = 2;
  a.r_point = h->r_point; a.r_target = h->r_target; a.r_host = h->r_host;
  a.r_is_lin = h->r_is_lin; a.r_is_active = h->r_is_active; a.r_dropped = h->r_dropped;
  a.rec = h->r_rec; a.accTop = h->d_accTop; a.n_acc = hs->d_cnt;
  launch_top_accumulate(h, a);
  launch_point_sc(h, sc_args(h, 2, d_pl, n, 0));
  if ((rc = sosba_allreduce_acc(h, 0))) return rc;
  launch_stitch_top(h, h->d_accTop, h->d_adHost, h->d_adTarget, nf, Hpart(h, 0), bpart(h, 0), 0, h->d_wprior, h->d_calib + 6);
  launch_finalize_sc(h, h->d_accSC, nf, Hpart(h, 2), bpart(h, 2));
  SOSBA_CUDA(cudaGetLastError());
  std::vector<double> M(HB), Msc(HB);
  if ((rc = down(h, M.data(), Hpart(h, 0), HB)) || (rc = down(h, Msc.data(), Hpart(h, 2), HB)) || (rc = down(h, hs->pin_i, hs->d_cnt, 1))) return rc;
  if ((rc = sync(h))) return rc;
  for (size_t i = 0; i < (size_t)D * D; i++) H[i] = M[i] - Msc[i];
  for (int i = 0; i < D; i++) b[i] = M[(size_t)D * D + i] - Msc[(size_t)D * D + i];
  if (resInM) *resInM = hs->pin_i[0];
  return SOSBA_OK;
}

// ---- tracker / scale optimizer ----------------------------------------------------------------------
API int sosba_tracker_make_k(sosba_t *h, const float calib[4]) {  // ScaleOptimizer::makeK (ScaleOptimizer.cpp:95-118)
  CHECK_H(h);
  float (*K)[4] = h->t_K;
  for (int i = 0; i < 4; i++) K[0][i] = calib[i];
  for (int l = 1; l < h->levels; l++) {
    K[l][0] = K[l - 1][0] * 0.5;
    K[l][1] = K[l - 1][1] * 0.5;
    K[l][2] = (K[0][2] + 0.5) / ((int)1 << l) - 0.5;
    K[l][3] = (K[0][3] + 0.5) / ((int)1 << l) - 0.5;
  }
  h->t_haveK = true;
  return SOSBA_OK;
}

static void make_Ki(const float K[4], float Ki[9]) {
  for (int i = 0; i < 9; i++) Ki[i] = 0.f;
  Ki[0] = 1.0f / K[0]; Ki[4] = 1.0f / K[1]; Ki[2] = -K[2] / K[0]; Ki[5] = -K[3] / K[1]; Ki[8] = 1.f;
}
static void mul33f(const float *A, const float *B, float *C) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[3 * i + j] = (A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j]) + A[3 * i + 2] * B[6 + j];
}

// room for n reference points on level lvl (pc_u | pc_v | pc_idepth | pc_color) and for the warped buffers of the per-call
// calcRes / calcGSSSE entry points; shared by sosba_tracker_set_ref and sosba_tracker_make_coarse_depth (k_lm.cu)
int sosba_tracker_reserve(sosba *h, int lvl, int n) {
  if (n > h->t_cap[lvl]) {
    dfree(h, h->t_pc[lvl]);
    h->t_cap[lvl] = n + n / 4 + 64;
    DALLOC(h, h->t_pc[lvl], 4 * (size_t)h->t_cap[lvl]);
  }
  if (n > h->t_warp_cap) {
    dfree(h, h->t_warp);
    h->t_warp_cap = n + n / 4 + 64;
    DALLOC(h, h->t_warp, 8 * (size_t)h->t_warp_cap);
  }
  return SOSBA_OK;
}

API int sosba_tracker_set_ref(sosba_t *h, int32_t lvl, int32_t n, const float *u, const float *v, const float *id, const float *c) {
  CHECK_H(h);
  if (lvl < 0 || lvl >= h->levels || n < 0) return SOSBA_E_ARG;
  { int rc0 = sosba_tracker_reserve(h, lvl, n); if (rc0) return rc0; }
  h->t_n[lvl] = n;
  int rc;
  if ((rc = up(h, h->t_pc[lvl], u, n)) || (rc = up(h, h->t_pc[lvl] + n, v, n)) || (rc = up(h, h->t_pc[lvl] + 2 * (size_t)n, id, n)) ||
      (rc = up(h, h->t_pc[lvl] + 3 * (size_t)n, c, n)))
    return rc;
  return sync(h);
}

static int track_res_common(sosba *h, int kind, int lvl, int slot, const float R[9], const float t[3], const float K[4], float aff0, float aff1,
                            float scale, float cutoff, double out6[6], int32_t counts[3]) {
  HostSide *hs = HS(h);
  if (kind == 2 && !hs->d_loop) { sosba_set_error("loop_set_points first"); return SOSBA_E_STATE; }
  if (!h->t_haveK) { sosba_set_error("tracker_make_k first"); return SOSBA_E_STATE; }
  if (lvl < 0 || lvl >= h->levels || slot < 0 || slot >= (int)h->slot_img.size() || !h->slot_valid[slot]) { sosba_set_error("bad level/slot"); return SOSBA_E_ARG; }
  TrackResArgs a;
  a.lvl = lvl; a.w = h->wl[lvl]; a.h = h->hl[lvl]; a.cap = h->t_warp_cap;
  a.img = h->slot_img[slot] + h->lvl_off[lvl];
  make_Ki(h->t_K[lvl], a.Ki);
  if (kind == 2) {   // loop closure: 3D points x | y | z, then one colour array per level; the rotation is used as it is
    a.n = hs->loop_n; a.pc = hs->d_loop; a.color = hs->d_loop + (size_t)(3 + lvl) * hs->loop_n;
    for (int i = 0; i < 9; i++) a.RKi[i] = R[i];
  } else {
    a.n = h->t_n[lvl]; a.pc = h->t_pc[lvl]; a.color = h->t_pc[lvl] + 3 * (size_t)a.n;
    mul33f(R, a.Ki, a.RKi);
  }
  hs->last_res_n = a.n;
  for (int i = 0; i < 3; i++) a.t[i] = t[i];
  a.fx = K[0]; a.fy = K[1]; a.cx = K[2]; a.cy = K[3];
  a.aff0 = aff0; a.aff1 = aff1;
  a.huberTH = h->cfg.huber_th; a.cutoffTH = cutoff; a.maxEnergy = 2 * a.huberTH * cutoff - a.huberTH * a.huberTH;
  a.scale = scale; a.kind = kind; a.warp = h->t_warp; a.acc = h->t_acc; a.icnt = h->d_counts + 8;
  cudaMemsetAsync(h->t_acc, 0, sizeof(double) * 8, h->stream);
  cudaMemsetAsync(h->d_counts + 8, 0, sizeof(int) * 3, h->stream);
  launch_track_res(h, a);
  SOSBA_CUDA(cudaGetLastError());
  int rc;
  if ((rc = down(h, hs->pin_d, h->t_acc, 4)) || (rc = down(h, hs->pin_i, h->d_counts + 8, 3))) return rc;
  if ((rc = sync(h))) return rc;
  const float E = (float)hs->pin_d[0], sT = (float)hs->pin_d[1], sRT = (float)hs->pin_d[2], sN = (float)hs->pin_d[3];
  const int nE = hs->pin_i[0], nW = hs->pin_i[1], nS = hs->pin_i[2];
  if (counts) { counts[0] = nE; counts[1] = nW; counts[2] = nS; }
  h->t_warp_n[lvl] = (nW + 3) / 4 * 4;  // padded to the SSE width (CoarseTracker.cpp:736-746)
  h->t_warp_lvl = lvl; h->t_warp_kind = kind;
  if (out6) {
    out6[0] = E; out6[1] = nE; out6[2] = sT / (sN + 0.1); out6[3] = 0; out6[4] = sRT / (sN + 0.1); out6[5] = nS / (float)nE;
  }
  return SOSBA_OK;
}

API int sosba_tracker_calc_res_pose(sosba_t *h, int32_t lvl, int32_t slot, const double refToNew[12], const float affLL[2], float cutoff,
                                    double out6[6], int32_t counts[3]) {
  CHECK_H(h);
  if (!refToNew || !affLL) return SOSBA_E_ARG;
  float R[9], t[3];
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R[3 * i + j] = (float)refToNew[4 * i + j]; t[i] = (float)refToNew[4 * i + 3]; }
  if (lvl < 0 || lvl >= h->levels) return SOSBA_E_ARG;
  return track_res_common(h, 0, lvl, slot, R, t, h->t_K[lvl], affLL[0], affLL[1], 1.f, cutoff, out6, counts);
}

static int gs_pose_common(sosba *h, int kind, int32_t lvl, float a, float b0, double H[64], double b[8]) {
  HostSide *hs = HS(h);
  if (lvl < 0 || lvl >= h->levels || lvl != h->t_warp_lvl || h->t_warp_kind != kind) {
    sosba_set_error("calc_gs needs the matching calc_res at the same level first");
    return SOSBA_E_STATE;
  }
  TrackGSArgs g;
  g.n = hs->last_res_n; g.cap = h->t_warp_cap; g.kind = 0; g.warp = h->t_warp; g.fx = h->t_K[lvl][0]; g.fy = h->t_K[lvl][1]; g.a = a; g.b0 = b0;
  g.scale = 1.f; g.tx = g.ty = g.tz = 0.f; g.acc = h->t_acc;
  cudaMemsetAsync(h->t_acc + 8, 0, sizeof(double) * 48, h->stream);
  launch_track_gs(h, g);
  SOSBA_CUDA(cudaGetLastError());
  int rc;
  if ((rc = down(h, hs->pin_d, h->t_acc + 8, 45))) return rc;
  if ((rc = sync(h))) return rc;
  float A[9][9];
  int q = 0;
  for (int r = 0; r < 9; r++) for (int c = r; c < 9; c++) { A[r][c] = A[c][r] = (float)hs->pin_d[q++]; }
  const int n = h->t_warp_n[lvl];
  const float invn = 1.0f / n;   // CoarseTracker.cpp:595-596: divided by the padded count
  const float sc[8] = {1.0f, 1.0f, 1.0f, 0.5f, 0.5f, 0.5f, 10.0f, 1000.0f};  // SCALE_XI_ROT x3, SCALE_XI_TRANS x3, SCALE_A, SCALE_B (:598-609)
  for (int r = 0; r < 8; r++) {
    for (int c = 0; c < 8; c++) H[8 * r + c] = (double)A[r][c] * invn;
    b[r] = (double)A[r][8] * invn;
  }
  for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) H[8 * r + c] *= sc[c];
  for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) H[8 * r + c] *= sc[r];
  for (int r = 0; r < 8; r++) b[r] *= sc[r];
  return SOSBA_OK;
}

API int sosba_tracker_calc_gs_pose(sosba_t *h, int32_t lvl, float a, float b0, double H[64], double b[8]) {
  CHECK_H(h);
  return gs_pose_common(h, 0, lvl, a, b0, H, b);
}

// ---- 8f rank 4: PoseEstimator (LoopClosure/PoseEstimator.cpp) -----------------------------------
API int sosba_loop_set_points(sosba_t *h, int32_t n, const double *xyz, const float *color) {
  CHECK_H(h);
  if (n < 0 || (n > 0 && (!xyz || !color))) { sosba_set_error("loop_set_points: bad arguments"); return SOSBA_E_ARG; }
  HostSide *hs = HS(h);
  const int L = h->levels;
  if (n > hs->loop_cap) {
    dfree(h, hs->d_loop);
    hs->loop_cap = n + n / 4 + 64;
    DALLOC(h, hs->d_loop, (size_t)(3 + L) * hs->loop_cap);
  }
  if (n > h->t_warp_cap) {
    dfree(h, h->t_warp);
    h->t_warp_cap = n + n / 4 + 64;
    DALLOC(h, h->t_warp, 8 * (size_t)h->t_warp_cap);
  }
  hs->loop_n = n;
  if (n == 0) return SOSBA_OK;
  std::vector<float> soa((size_t)(3 + L) * n);
  for (int i = 0; i < n; i++) {
    for (int k = 0; k < 3; k++) soa[(size_t)k * n + i] = (float)xyz[3 * (size_t)i + k];   // float x = pts[i].first(0) (PoseEstimator.cpp:188-190)
    for (int l = 0; l < L; l++) soa[(size_t)(3 + l) * n + i] = color[(size_t)i * L + l];
  }
  int rc;
  if ((rc = up(h, hs->d_loop, soa.data(), soa.size()))) return rc;
  return sync(h);
}

API int sosba_loop_calc_res(sosba_t *h, int32_t lvl, int32_t slot, const double refToNew[12], const float affLL[2], float cutoff, double out6[6],
                            int32_t counts[3]) {
  CHECK_H(h);
  if (!refToNew || !affLL) return SOSBA_E_ARG;
  if (lvl < 0 || lvl >= h->levels) return SOSBA_E_ARG;
  float R[9], t[3];
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R[3 * i + j] = (float)refToNew[4 * i + j]; t[i] = (float)refToNew[4 * i + 3]; }
  return track_res_common(h, 2, lvl, slot, R, t, h->t_K[lvl], affLL[0], affLL[1], 1.f, cutoff, out6, counts);
}

API int sosba_loop_calc_gs(sosba_t *h, int32_t lvl, float a, float b0, double H[64], double b[8]) {
  CHECK_H(h);
  return gs_pose_common(h, 2, lvl, a, b0, H, b);
}

API int sosba_scale_set_stereo(sosba_t *h, const double T10[12], const float K1[4]) {
  CHECK_H(h);
  if (!T10 || !K1) return SOSBA_E_ARG;
  memcpy(h->t_T10, T10, sizeof(double) * 12);
  float (*K)[4] = h->t_K1;
  for (int i = 0; i < 4; i++) K[0][i] = K1[i];
  for (int l = 1; l < h->levels; l++) {  // ScaleOptimizer.cpp:72-77
    K[l][0] = K[l - 1][0] * 0.5;
    K[l][1] = K[l - 1][1] * 0.5;
    K[l][2] = (K[0][2] + 0.5) / ((int)1 << l) - 0.5;
    K[l][3] = (K[0][3] + 0.5) / ((int)1 << l) - 0.5;
  }
  h->t_haveStereo = true;
  return SOSBA_OK;
}

API int sosba_scale_calc_res(sosba_t *h, int32_t lvl, int32_t slot, float scale, float cutoff, double out6[6], int32_t counts[3]) {
  CHECK_H(h);
  if (!h->t_haveStereo) { sosba_set_error("scale_set_stereo first"); return SOSBA_E_STATE; }
  if (lvl < 0 || lvl >= h->levels) return SOSBA_E_ARG;
  float R[9], t[3];
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R[3 * i + j] = (float)h->t_T10[4 * i + j]; t[i] = (float)h->t_T10[4 * i + 3]; }
  return track_res_common(h, 1, lvl, slot, R, t, h->t_K1[lvl], 1.f, 0.f, scale, cutoff, out6, counts);
}

API int sosba_scale_calc_gs(sosba_t *h, int32_t lvl, float scale, float *H, float *b) {
  CHECK_H(h);
  HostSide *hs = HS(h);
  if (lvl != h->t_warp_lvl || h->t_warp_kind != 1) { sosba_set_error("scale_calc_gs needs scale_calc_res at the same level first"); return SOSBA_E_STATE; }
  TrackGSArgs g;
  g.n = h->t_n[lvl]; g.cap = h->t_warp_cap; g.kind = 1; g.warp = h->t_warp; g.fx = h->t_K1[lvl][0]; g.fy = h->t_K1[lvl][1]; g.a = 0.f; g.b0 = 0.f;
  g.scale = scale; g.tx = (float)h->t_T10[3]; g.ty = (float)h->t_T10[7]; g.tz = (float)h->t_T10[11]; g.acc = h->t_acc;
  cudaMemsetAsync(h->t_acc + 8, 0, sizeof(double) * 48, h->stream);
  launch_track_gs(h, g);
  SOSBA_CUDA(cudaGetLastError());
  int rc;
  if ((rc = down(h, hs->pin_d, h->t_acc + 8, 3))) return rc;
  if ((rc = sync(h))) return rc;
  const int n = h->t_warp_n[lvl];
  if (H) *H = (float)hs->pin_d[0] * (1.0f / n);
  if (b) *b = (float)hs->pin_d[1] * (1.0f / n);
  return SOSBA_OK;
}

// ---- composed Gauss-Newton loop: FullSystem::optimize (FullSystemOptimize.cpp:305-489), IMU off -----
static int upload_tables(sosba *h, BA *ba, bool full) {
  WindowTables &w = ba->wt;
  sosba_window win;
  memset(&win, 0, sizeof(win));
  win.nf = w.nf;
  win.frame_slot = w.frame_slot.data();
  win.precalc = w.precalc.data();
  win.adHost = w.adHost.data(); win.adTarget = w.adTarget.data();
  win.adHTdeltaF = w.adHTdeltaF.data();
  win.frame_energy_th = w.frameEnergyTH.data();
  for (int i = 0; i < 4; i++) { win.calib[i] = w.calib[i]; win.cDeltaF[i] = w.cDeltaF[i]; win.cPrior[i] = w.cPrior[i]; }
  win.frame_prior = w.frame_prior.data(); win.frame_delta_prior = w.frame_delta_prior.data(); win.frame_delta = w.frame_delta.data();
  return window_apply(h, &win, full);
}

static int upload_frame_state(sosba *h);

API int sosba_ba_upload(sosba_t *h, const sosba_ba_problem *prob) {
  CHECK_H(h);
  if (!prob || prob->nf <= 0 || !prob->frames) return SOSBA_E_ARG;
  BA *ba = h->ba;
  static const bool timing = getenv("SOSBA_TIMING") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto us = [](auto a, auto b) { return (long)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count(); };
  const auto t0 = now();
  ba->st.load(h->cfg, prob);
  ba->st.make_adjoints(h->cfg, ba->wt);
  ba->st.make_precalc(ba->wt);
  int rc;
  const auto t1 = now();
  if ((rc = upload_tables(h, ba, true))) return rc;
  const auto t2 = now();
  {  // EnergyFunctional::setDeltaF: p->deltaF = idepth - idepth_zero (EnergyFunctional.cpp:187-191), staged with the other point arrays
    sosba_points pp = prob->points;
    std::vector<float> &d = HS(h)->delta_tmp;
    if (pp.n > 0 && pp.idepth && pp.idepth_zero) {
      d.resize(pp.n);
      for (int i = 0; i < pp.n; i++) d[i] = pp.idepth[i] - pp.idepth_zero[i];
      pp.deltaF = d.data();
    }
    if ((rc = sosba_points_set(h, &pp))) return rc;
  }
  const auto t3 = now();
  if ((rc = sosba_residuals_set(h, &prob->residuals))) return rc;
  if (timing) fprintf(stderr, "ba_upload: host tables %ld us, window %ld us, points %ld us, delta+residuals %ld us\n", us(t0, t1), us(t1, t2), us(t2, t3), us(t3, now()));
  const int D = 4 + 8 * prob->nf;
  ba->have_HM = !ba->st.HM.empty();
  if (ba->have_HM) {
    HostSide *hs = HS(h);
    if ((rc = up(h, hs->d_HMtmp, ba->st.HM.data(), (size_t)D * D)) || (rc = up(h, hs->d_bMtmp, ba->st.bM.data(), D))) return rc;
  }
  ba->iterations_done = 0;
  return upload_frame_state(h);
}

// host mirror -> device frame / calibration state
static int upload_frame_state(sosba *h) {
  BA *ba = h->ba;
  HostSide *hs = HS(h);
  const int nf = (int)ba->st.frames.size();
  int rc;
  std::vector<double> buf((size_t)nf * SOSBA_FS + 16, 0.0);   // staged by up(): no synchronisation
  double *p = buf.data();
  for (int f = 0; f < nf; f++) {
    const sosba_host::FrameH &F = ba->st.frames[f];
    double *q = p + SOSBA_FS * f;
    sosba_math::rigid_to34(F.evalPT, q);
    for (int i = 0; i < 10; i++) { q[12 + i] = F.state[i]; q[22 + i] = F.state_zero[i]; q[32 + i] = F.state_backup[i]; q[42 + i] = F.step[i]; }
    q[52] = F.ab_exposure;
  }
  double *c = p + nf * SOSBA_FS;
  for (int i = 0; i < 4; i++) { c[i] = ba->st.calib.value[i]; c[4 + i] = ba->st.calib.value_zero[i]; c[8 + i] = ba->st.calib.value_backup[i]; c[12 + i] = ba->st.calib.step[i]; }
  if ((rc = up(h, hs->d_fs, p, (size_t)nf * SOSBA_FS)) || (rc = up(h, hs->d_cs, c, 16))) return rc;
  return SOSBA_OK;
}

// device frame / calibration state -> host mirror
static const int PIN_FS_OFF = 1024;   // doubles: where a prefetched copy of the frame states sits in pin_d
static int enqueue_frame_state_download(sosba *h, double *p) {
  BA *ba = h->ba;
  HostSide *hs = HS(h);
  const int nf = (int)ba->st.frames.size();
  int rc;
  if ((rc = down(h, p, hs->d_fs, (size_t)nf * SOSBA_FS)) || (rc = down(h, p + nf * SOSBA_FS, hs->d_cs, 16)) || (rc = down(h, hs->pin_f, h->d_frameEnergyTH, nf))) return rc;
  return SOSBA_OK;
}
static int download_frame_state(sosba *h, bool prefetched = false) {
  BA *ba = h->ba;
  HostSide *hs = HS(h);
  const int nf = (int)ba->st.frames.size();
  int rc;
  double *p = prefetched ? hs->pin_d + PIN_FS_OFF : hs->pin_d;
  if (!prefetched) {
    if ((rc = enqueue_frame_state_download(h, p))) return rc;
    if ((rc = sync(h))) return rc;
  }
  for (int f = 0; f < nf; f++) {
    sosba_host::FrameH &F = ba->st.frames[f];
    const double *q = p + SOSBA_FS * f;
    for (int i = 0; i < 10; i++) { F.state_backup[i] = q[32 + i]; F.step[i] = q[42 + i]; F.state_zero[i] = q[22 + i]; }
    F.evalPT = sosba_math::rigid_from34(q);   // k_frame_retarget re-anchors the newest frame on the device
    F.setState(q + 12);
    F.frameEnergyTH = hs->pin_f[f];
  }
  const double *c = p + nf * SOSBA_FS;
  for (int i = 0; i < 4; i++) { ba->st.calib.value_backup[i] = c[8 + i]; ba->st.calib.step[i] = c[12 + i]; }
  ba->st.calib.setValue(c);
  ba->mirror_stale = false;
  return SOSBA_OK;
}

// one loop body of FullSystem::optimize (FullSystemOptimize.cpp:358-413), enqueued on the stream with no host
// round trip: backupState + solveSystemF + doStepFromBackup (points in k_resubstitute, frames/calib/precalc in
// k_frame_step) + linearizeAll(false) + applyRes
static void enqueue_linearize_apply(sosba *h, bool zero_tables);
static int enqueue_iteration(sosba *h) {
  BA *ba = h->ba;
  HostSide *hs = HS(h);
  int rc = enqueue_solve(h, ba->have_HM ? hs->d_HMtmp : nullptr, ba->have_HM ? hs->d_bMtmp : nullptr, 1);
  if (rc) return rc;
  enqueue_linearize_apply(h, true);
  hs->rstats_par ^= 1;
  SOSBA_CUDA(cudaGetLastError());
  ba->iterations_done++;
  return SOSBA_OK;
}

// linearizeAll(false) + applyRes in ONE launch (threshold in its last CTA); also clears the block tables for the
// accumulation of the next loop body
static void enqueue_linearize_apply(sosba *h, bool zero_tables) {
  HostSide *hs = HS(h);
  flush_pending_th(h);
  LinArgs a = lin_args(h);   // the linearisation sums were cleared by the back-substitution launch of this body
  a.gate = hs->gate;
  if (zero_tables) { a.zero_buf = hs->d_scratch; a.zero_n = (int)(hs->scratch_zero_doubles / 2); }
  a.trace = sosba_trace_slot("k_linearize (fused apply)");
  if (a.trace) { sosba_trace_slot("#  phases of its middle CTA"); sosba_trace_slot("#"); }   // 7 clock64 stamps (k_linearize_t: LIN_TS)
  // the selection runs in the spare CTA of the next accumulation, or (point shards) behind the next all-reduce
  const bool th_inline = !hs->fused_acc_ok && !(h->comm && h->world > 1);
  if (hs->prof_on) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, h->stream);
    launch_linearize_apply(h, a, false, th_inline);
    cudaEventRecord(e1, h->stream);
    hs->prof_ev.push_back(e0); hs->prof_ev.push_back(e1);
  } else launch_linearize_apply(h, a, false, th_inline);
  hs->th_pending = !th_inline;
  hs->tables_clean = zero_tables;
}

API int sosba_ba_iterate(sosba_t *h, int32_t n, int32_t *n_res) {
  CHECK_H(h);
  if (!h->ba->st.loaded) { sosba_set_error("ba_upload first"); return SOSBA_E_STATE; }
  HostSide *hs = HS(h);
  int rc;
  cudaMemsetAsync(hs->d_ctl + 2, 0, sizeof(int), h->stream);
  for (int i = 0; i < n; i++)
    if ((rc = enqueue_iteration(h))) return rc;
  flush_pending_th(h);
  if ((rc = down(h, hs->pin_i, hs->d_ctl + 2, 1))) return rc;
  if ((rc = sync(h))) return rc;
  if (hs->pin_i[0]) return solve_flag_error(hs->pin_i[0]);
  if (n_res) *n_res = h->R - hs->n_lin;
  return SOSBA_OK;
}

// ---- the loop body split around a caller-side solve (IMU configurations; include/sosba.h) ---------------------------------
API int sosba_ba_system(sosba_t *h, double *H_top, double *b_top, double *H_sc, double *b_sc, int32_t *resInA, int32_t *resInL) {
  CHECK_H(h);
  if (!h->ba->st.loaded) { sosba_set_error("ba_upload first"); return SOSBA_E_STATE; }
  HostSide *hs = HS(h);
  const int nf = h->nf, D = 4 + 8 * nf;
  int rc;
  cudaMemsetAsync(hs->d_ctl + 2, 0, sizeof(int), h->stream);
  StitchXchgArgs x;
  memset(&x, 0, sizeof(x));
  sosba_xchg_args(h, &x, hs->th_pending ? 1 : 0);
  const bool sharded = h->comm && h->world > 1;
  ThArgs th_def = {};
  int th_deferred = 0;
  if ((rc = enqueue_blocks(h, sharded && !x.push, &th_def, &th_deferred))) return rc;
  x.nf = nf; x.D = D; x.accTop = h->d_accTop; x.adHost = h->d_adHost; x.adTarget = h->d_adTarget;
  x.H = Hpart(h, 0); x.b = bpart(h, 0); x.accSC = h->d_accSC; x.rstats = h->d_rstats_all; x.cnt = h->d_cnt_all;
  x.gate = nullptr; x.err = hs->d_ctl + 2;
  if ((rc = launch_stitch_xchg(h, x, h->P))) return rc;
  if (th_deferred) { launch_energy_th(h, th_def); hs->th_pending = false; }
  launch_add_priors(h, nf, Hpart(h, 0), bpart(h, 0), h->d_wprior, h->d_calib + 6);
  launch_finalize_sc(h, h->d_accSC, nf, Hpart(h, 2), bpart(h, 2));
  SOSBA_CUDA(cudaGetLastError());
  if ((rc = fetch(h, H_top, Hpart(h, 0), (size_t)D * D)) || (rc = fetch(h, b_top, bpart(h, 0), D)) || (rc = fetch(h, H_sc, Hpart(h, 2), (size_t)D * D)) ||
      (rc = fetch(h, b_sc, bpart(h, 2), D)) || (rc = down(h, hs->pin_i, hs->d_cnt, 2)) || (rc = down(h, hs->pin_i + 8, hs->d_ctl + 2, 1)))
    return rc;
  if ((rc = sync(h))) return rc;
  if (hs->pin_i[8] & 2) return solve_flag_error(hs->pin_i[8]);
  if (resInA) *resInA = hs->pin_i[0];
  if (resInL) *resInL = hs->pin_i[1];
  return SOSBA_OK;
}

API int sosba_ba_step(sosba_t *h, const double *x, sosba_step_out *out) {
  CHECK_H(h);
  if (!x || !out) return SOSBA_E_ARG;
  if (!h->ba->st.loaded) { sosba_set_error("ba_upload first"); return SOSBA_E_STATE; }
  HostSide *hs = HS(h);
  const int nf = h->nf, D = 4 + 8 * nf;
  int rc;
  if ((rc = up(h, h->d_x, x, D))) return rc;
  double *rst = hs->d_rstats + 4 * hs->rstats_par;   // [0] sum step^2 [1] sum |idepth_backup| [2] count of this body
  cudaMemsetAsync(rst, 0, 4 * sizeof(double), h->stream);
  cudaMemsetAsync(hs->d_ctl + 2, 0, sizeof(int), h->stream);
  flush_pending_th(h);
  {
    ResubArgs ra = resub_args(h, 1);
    ra.zero_lin = h->d_stats;
    if (h->d_newE_all) {
      if (h->p2p) { ra.zero_newE = (float *)h->d_newE_cnt; ra.zero_newE_n = h->world; }
      else { ra.zero_newE = h->d_newE_all; ra.zero_newE_n = h->world * h->newE_cap + h->world; }
    }
    launch_step(h, ra, step_args(h));
  }
  enqueue_linearize_apply(h, true);
  // sums over the point shards: energy, state histogram, the step sums of the points; then setNewFrameEnergyTH
  if ((rc = exchange_lin(h, 1, rst, 3))) return rc;
  if (hs->th_pending) { launch_energy_th(h, lin_args(h).th); clear_gathered_energies(h); hs->th_pending = false; }
  h->ba->mirror_stale = true;
  h->ba->iterations_done++;
  SOSBA_CUDA(cudaGetLastError());
  if ((rc = down(h, hs->pin_d, h->d_stats, 12)) || (rc = down(h, hs->pin_d + 16, hs->d_iter, 4)) || (rc = down(h, hs->pin_d + 24, rst, 3)) ||
      (rc = down(h, hs->pin_i + 8, hs->d_ctl + 2, 1)))
    return rc;
  if ((rc = sync(h))) return rc;
  if (hs->pin_i[8] & 2) return solve_flag_error(hs->pin_i[8]);
  const int *ci = (const int *)(hs->pin_d + 2);
  out->energy = hs->pin_d[0];
  out->new_frame_energy_th = ((const float *)(ci + 16))[0];
  out->n_in = ci[0]; out->n_oob = ci[1]; out->n_outlier = ci[2];
  out->sum_a = hs->pin_d[16]; out->sum_b = hs->pin_d[17]; out->sum_t = hs->pin_d[18]; out->sum_r = hs->pin_d[19];
  out->sum_id = hs->pin_d[24]; out->sum_nid = hs->pin_d[25]; out->num_id = hs->pin_d[26];
  return SOSBA_OK;
}

API int sosba_ba_download(sosba_t *h, sosba_ba_problem *prob) {
  CHECK_H(h);
  if (!prob || !h->ba->st.loaded) return SOSBA_E_STATE;
  HostSide *hs = HS(h);
  const bool ready = hs->download_ready;
  hs->download_ready = false;
  int rc0 = download_frame_state(h, ready);
  if (rc0) return rc0;
  h->ba->st.store(prob);
  if (prob->idepth_out) {
    if (ready && hs->pin_idepth) { memcpy(prob->idepth_out, hs->pin_idepth, sizeof(float) * (size_t)h->P); return SOSBA_OK; }
    int rc = down(h, prob->idepth_out, h->p_idepth, h->P);
    if (rc) return rc;
    return sync(h);
  }
  return SOSBA_OK;
}

API int sosba_ba_optimize(sosba_t *h, int32_t mnumOptIts, sosba_optimize_out *out) {
  CHECK_H(h);
  if (!out) return SOSBA_E_ARG;
  memset(out, 0, sizeof(*out));
  BA *ba = h->ba;
  if (!ba->st.loaded) { sosba_set_error("ba_upload first"); return SOSBA_E_STATE; }
  HostSide *hs = HS(h);
  const int nf = h->nf;
  int rc;
  if (nf < 2) return SOSBA_OK;
  if (nf < 3) mnumOptIts = 20;
  if (nf < 4) mnumOptIts = 15;
  // Everything below goes onto the stream without a host round trip; ONE synchronisation at the end reads the results.
  trace_begin(h);
  flush_pending_th(h);
  // resetOOB; the same launch clears the linearisation sums (2 doubles + 5 ints) and the loop control words, so no memset
  // and no copy sits between the launches below (each would break the programmatic launch chain)
  launch_reset_oob(h, lin_args(h), (int *)h->d_stats, 9, hs->d_ctl);
  clear_gathered_energies(h);
  enqueue_linearize_apply(h, false);   // linearizeAll(false) + applyRes, fused
  // energy | pad | counts[16] | thOut of the first linearisation: stashed on the device by the first k_solve, read at the end.
  // The whole loop: k_solve of body i latches "converged" from the step norms of body i-1 (doStepFromBackup's canbreak,
  // iteration >= setting_minOptIterations) and every later launch returns immediately
  hs->gate = hs->d_ctl;
  hs->stash_pending = true;
  for (int iteration = 0; iteration < mnumOptIts; iteration++) {
    hs->loop_iter = iteration;
    if ((rc = enqueue_iteration(h))) { hs->gate = nullptr; hs->loop_iter = -1; return rc; }
  }
  const bool sharded_loop = h->comm && h->world > 1;
  if (hs->stash_pending) {   // no loop body ran: keep the sums of the first linearisation with a plain copy
    cudaMemcpyAsync(hs->d_stash, h->d_stats, 12 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
    hs->stash_pending = false;
  }
  hs->gate = nullptr; hs->loop_iter = -1;
  hs->tables_clean = false;   // a loop that broke early leaves partial block tables behind
  // new evaluation point for the newest frame (FullSystemOptimize.cpp:415-423) with its adjoints / precalc / deltas, on
  // the device; then linearizeAll(true).  One GPU: the pending threshold selection of the loop's last linearisation runs in
  // a second CTA of the retarget launch, which also clears the sums of the linearisation that follows.
  if (!sharded_loop && mnumOptIts > 0) {
    const ThArgs th = lin_args(h).th;
    launch_frame_retarget(h, step_args(h), hs->th_pending ? &th : nullptr, (int *)h->d_stats, 9);
    hs->th_pending = false;
    ba->mirror_stale = true;
    enqueue_linearize(h, 1, false);
  } else {
    flush_pending_th(h);
    launch_frame_retarget(h, step_args(h));
    ba->mirror_stale = true;    // the host mirror (ba->st) is refreshed by download_frame_state
    enqueue_linearize(h, 1);
  }
  SOSBA_CUDA(cudaGetLastError());
  const int D = 4 + 8 * nf;
  if ((rc = down(h, hs->pin_i, hs->d_ctl, 4)) || (rc = down(h, hs->pin_d + 16, hs->d_stash, 12)) || (rc = down(h, hs->pin_d + 32, h->d_x, D))) return rc;
  hs->download_ready = false;
  if (hs->fold_download && (size_t)nf * SOSBA_FS + 16 + PIN_FS_OFF <= 4096) {   // what sosba_ba_download will want, on the same synchronisation
    if ((size_t)h->P > hs->pin_idepth_cap) {
      if (hs->pin_idepth) cudaFreeHost(hs->pin_idepth);
      hs->pin_idepth_cap = (size_t)h->P + h->P / 4 + 256;
      SOSBA_CUDA(cudaMallocHost((void **)&hs->pin_idepth, hs->pin_idepth_cap * sizeof(float)));
    }
    if ((rc = enqueue_frame_state_download(h, hs->pin_d + PIN_FS_OFF)) || (rc = down(h, hs->pin_idepth, (const float *)h->p_idepth, h->P))) return rc;
    hs->download_ready = true;
  }
  sosba_linearize_out lo;
  if ((rc = read_linearize_out(h, &lo))) { hs->download_ready = false; return rc; }   // the one synchronisation
  if (hs->pin_i[2]) return solve_flag_error(hs->pin_i[2]);
  {
    const int *c0 = (const int *)(hs->pin_d + 16 + 2);
    out->reserved0 = c0[0] + c0[1] + c0[2];   // residuals linearised per pass (bench bookkeeping)
    out->energy_initial = hs->pin_d[16];
  }
  out->iterations = hs->pin_i[1];
  out->energy_final = lo.energy;
  out->n_removed = lo.n_removed;
  out->res_in_a = hs->pin_i[3];               // resInA of the last solve
  out->rmse = sqrtf((float)(lo.energy / (SOSBA_PATTERN * out->res_in_a)));
  double n2 = 0;
  for (int i = 0; i < D; i++) n2 += hs->pin_d[32 + i] * hs->pin_d[32 + i];
  out->last_x_norm = sqrt(n2);
  return SOSBA_OK;
}

API int sosba_optimize(sosba_t *h, sosba_ba_problem *prob, int32_t mnumOptIts, sosba_optimize_out *out) {
  CHECK_H(h);
  if (!prob || !out) return SOSBA_E_ARG;
  static const bool timing = getenv("SOSBA_TIMING") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  const auto t0 = now();
  int rc = sosba_ba_upload(h, prob);
  if (rc) return rc;
  const auto t1 = now();
  HS(h)->fold_download = true;
  rc = sosba_ba_optimize(h, mnumOptIts, out);
  HS(h)->fold_download = false;
  if (rc) { HS(h)->download_ready = false; return rc; }
  const auto t2 = now();
  rc = sosba_ba_download(h, prob);
  if (timing) {
    const auto t3 = now();
    auto us = [](auto a, auto b) { return (long)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count(); };
    fprintf(stderr, "sosba_optimize host wall: upload %ld us, optimize %ld us, download %ld us\n", us(t0, t1), us(t1, t2), us(t2, t3));
  }
  return rc;
}

// ---- next row: immature points --------------------------------------------
// arena layout for n points (floats): [u n][v n][color 8n][weights 8n][gradH 4n][energyTH n][idmin n][idmax n][quality n][uv 2n][pixint n][host n (int)][status n bytes]
static int ensure_immature(sosba *h, size_t n, int nhosts) {
  HostSide *hs = HS(h);
  if (n > hs->imm_cap) {
    dfree(h, hs->d_imm);
    hs->imm_cap = n + n / 4 + 256;
    DALLOC(h, hs->d_imm, 31 * hs->imm_cap);
  }
  if (nhosts > hs->imm_host_cap) {
    dfree(h, hs->d_imm_host);
    hs->imm_host_cap = nhosts + 8;
    DALLOC(h, hs->d_imm_host, 14 * (size_t)hs->imm_host_cap + 8);
  }
  return SOSBA_OK;
}

API int sosba_immature_init(sosba_t *h, int32_t host_slot, int32_t n, const int32_t *u, const int32_t *v, float *color, float *weights, float *gradH,
                            float *energy_th) {
  CHECK_H(h);
  if (host_slot < 0 || host_slot >= (int)h->slot_img.size() || !h->slot_valid[host_slot] || n < 0) { sosba_set_error("bad slot / n"); return SOSBA_E_ARG; }
  if (n == 0) return SOSBA_OK;
  if (!u || !v || !color || !weights || !gradH || !energy_th) { sosba_set_error("null buffer"); return SOSBA_E_ARG; }
  for (int i = 0; i < n; i++)   // the pattern reaches 2 px and the bilinear tap one more: the reference's selector keeps candidates inside this margin
    if (u[i] < 2 || v[i] < 2 || u[i] > h->wl[0] - 4 || v[i] > h->hl[0] - 4) { sosba_set_error("candidate %d outside the image margin", i); return SOSBA_E_ARG; }
  int rc;
  if ((rc = ensure_immature(h, (size_t)n, 1))) return rc;
  HostSide *hs = HS(h);
  const size_t N = (size_t)n;
  float *d = hs->d_imm;
  TraceArgs a = {};
  a.n = n; a.w = h->wl[0]; a.h = h->hl[0]; a.img = h->slot_img[host_slot] + h->lvl_off[0];
  a.iu = (const int *)d; a.iv = (const int *)(d + N);
  a.color_out = d + 2 * N; a.weights_out = d + 10 * N; a.gradH_out = d + 18 * N; a.energyTH_out = d + 22 * N;
  a.outlierTHSum = h->cfg.outlier_th_sum_component; a.overallWeight = h->cfg.overall_energy_th_weight;
  if ((rc = up(h, (int *)d, (const int *)u, N)) || (rc = up(h, (int *)(d + N), (const int *)v, N))) return rc;
  // a colour that is not finite ends the pattern loop early (ImmaturePoint.cpp:43): the entries behind it are left as 0
  SOSBA_CUDA(cudaMemsetAsync(d + 2 * N, 0, 16 * N * sizeof(float), h->stream));
  launch_immature_init(h, a);
  SOSBA_CUDA(cudaGetLastError());
  if ((rc = down(h, color, a.color_out, 8 * N)) || (rc = down(h, weights, a.weights_out, 8 * N)) || (rc = down(h, gradH, a.gradH_out, 4 * N)) ||
      (rc = down(h, energy_th, a.energyTH_out, N)))
    return rc;
  return sync(h);
}

static const int IMM_CHUNK = 48 * 1024;   // points per pass: 48k x 121 B = 5.8 MB of the 16 MB ring

// One pass = one H2D of the packed SoA, one launch, one D2H of the in/out tail, one synchronisation.
API int sosba_trace_immature(sosba_t *h, int32_t frame_slot, int32_t nhosts, const float *KRKi, const float *Kt, const float *aff, sosba_immature *pts,
                             int32_t counts[6]) {
  CHECK_H(h);
  if (frame_slot < 0 || frame_slot >= (int)h->slot_img.size() || !h->slot_valid[frame_slot] || !pts || pts->n < 0 || nhosts < 0) {
    sosba_set_error("bad slot / points");
    return SOSBA_E_ARG;
  }
  if (counts) for (int i = 0; i < 6; i++) counts[i] = 0;
  if (pts->n == 0) return SOSBA_OK;
  if (!KRKi || !Kt || !aff || !pts->host || !pts->u || !pts->v || !pts->color || !pts->weights || !pts->gradH || !pts->energy_th || !pts->idepth_min ||
      !pts->idepth_max || !pts->quality || !pts->last_trace_status || !pts->last_trace_uv || !pts->last_trace_pixel_interval) {
    sosba_set_error("null buffer");
    return SOSBA_E_ARG;
  }
  for (int i = 0; i < pts->n; i++)
    if (pts->host[i] < 0 || pts->host[i] >= nhosts) { sosba_set_error("point %d: host %d outside [0,%d)", i, pts->host[i], nhosts); return SOSBA_E_ARG; }
  int rc;
  if ((rc = ensure_immature(h, (size_t)std::min(pts->n, IMM_CHUNK), nhosts))) return rc;
  HostSide *hs = HS(h);
  float *dh = hs->d_imm_host;
  int *d_counts = (int *)(dh + 14 * (size_t)hs->imm_host_cap);
  if ((rc = up(h, dh, KRKi, 9 * (size_t)nhosts)) || (rc = up(h, dh + 9 * (size_t)hs->imm_host_cap, Kt, 3 * (size_t)nhosts)) ||
      (rc = up(h, dh + 12 * (size_t)hs->imm_host_cap, aff, 2 * (size_t)nhosts)))
    return rc;
  SOSBA_CUDA(cudaMemsetAsync(d_counts, 0, 6 * sizeof(int), h->stream));
  for (int off = 0; off < pts->n; off += IMM_CHUNK) {
    const int n = std::min(IMM_CHUNK, pts->n - off);
    const size_t N = (size_t)n, bytes = 30 * N * sizeof(float) + N;
    char *blk;
    if ((rc = stage_reserve(h, bytes, &blk))) return rc;
    float *s = (float *)blk, *d = hs->d_imm;
    // arena (floats): [u][v][color 8][weights 8][gradH 4][energyTH] | [idmin][idmax][quality][uv 2][pixint] [host (int)] [status (bytes)]
    memcpy(s, pts->u + off, 4 * N); memcpy(s + N, pts->v + off, 4 * N);
    memcpy(s + 2 * N, pts->color + 8 * (size_t)off, 32 * N); memcpy(s + 10 * N, pts->weights + 8 * (size_t)off, 32 * N);
    memcpy(s + 18 * N, pts->gradH + 4 * (size_t)off, 16 * N); memcpy(s + 22 * N, pts->energy_th + off, 4 * N);
    memcpy(s + 23 * N, pts->idepth_min + off, 4 * N); memcpy(s + 24 * N, pts->idepth_max + off, 4 * N); memcpy(s + 25 * N, pts->quality + off, 4 * N);
    memcpy(s + 26 * N, pts->last_trace_uv + 2 * (size_t)off, 8 * N); memcpy(s + 28 * N, pts->last_trace_pixel_interval + off, 4 * N);
    memcpy(s + 29 * N, pts->host + off, 4 * N); memcpy(s + 30 * N, pts->last_trace_status + off, N);
    SOSBA_CUDA(cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, h->stream));
    TraceArgs a = {};
    a.n = n; a.w = h->wl[0]; a.h = h->hl[0]; a.img = h->slot_img[frame_slot] + h->lvl_off[0];
    a.u = d; a.v = d + N; a.color = d + 2 * N; a.weights = d + 10 * N; a.gradH = d + 18 * N; a.energyTH = d + 22 * N;
    a.idepth_min = d + 23 * N; a.idepth_max = d + 24 * N; a.quality = d + 25 * N; a.uv = d + 26 * N; a.pixint = d + 28 * N;
    a.host = (const int *)(d + 29 * N); a.status = (uint8_t *)(d + 30 * N);
    a.KRKi = dh; a.Kt = dh + 9 * (size_t)hs->imm_host_cap; a.aff = dh + 12 * (size_t)hs->imm_host_cap;
    a.huberTH = h->cfg.huber_th; a.counts = d_counts;
    launch_trace_on(h, a);
    SOSBA_CUDA(cudaGetLastError());
    SOSBA_CUDA(cudaMemcpyAsync(s + 23 * N, d + 23 * N, 7 * N * sizeof(float) + N, cudaMemcpyDeviceToHost, h->stream));
    if (off + n >= pts->n) { if ((rc = down(h, hs->pin_i, (const int *)d_counts, 6))) return rc; }
    if ((rc = sync(h))) return rc;
    memcpy(pts->idepth_min + off, s + 23 * N, 4 * N); memcpy(pts->idepth_max + off, s + 24 * N, 4 * N); memcpy(pts->quality + off, s + 25 * N, 4 * N);
    memcpy(pts->last_trace_uv + 2 * (size_t)off, s + 26 * N, 8 * N); memcpy(pts->last_trace_pixel_interval + off, s + 28 * N, 4 * N);
    memcpy(pts->last_trace_status + off, s + 30 * N, N);
  }
  if (counts) for (int i = 0; i < 6; i++) counts[i] = hs->pin_i[i];
  return SOSBA_OK;
}

API int sosba_optimize_immature(sosba_t *h, const sosba_activation_window *win, const sosba_immature *pts, int8_t *result, float *idepth, uint8_t *res_state) {
  CHECK_H(h);
  if (!win || !pts || win->nf < 1 || win->nf > SOSBA_ACT_MAXF || pts->n < 0 || !win->frame_slot || !win->RTll || !win->tTll || !win->aff) {
    sosba_set_error("bad activation window (1 <= nf <= %d)", SOSBA_ACT_MAXF);
    return SOSBA_E_ARG;
  }
  const int nf = win->nf;
  if (pts->n == 0) return SOSBA_OK;
  if (!pts->host || !pts->u || !pts->v || !pts->color || !pts->weights || !pts->energy_th || !pts->idepth_min || !pts->idepth_max || !result || !idepth ||
      !res_state) {
    sosba_set_error("null buffer");
    return SOSBA_E_ARG;
  }
  ActivateArgs a = {};
  for (int f = 0; f < nf; f++) {
    const int slot = win->frame_slot[f];
    if (slot < 0 || slot >= (int)h->slot_img.size() || !h->slot_valid[slot]) { sosba_set_error("frame %d: bad slot %d", f, slot); return SOSBA_E_ARG; }
    a.img[f] = h->slot_img[slot] + h->lvl_off[0];
  }
  for (int i = 0; i < pts->n; i++)
    if (pts->host[i] < 0 || pts->host[i] >= nf) { sosba_set_error("point %d: host %d outside [0,%d)", i, pts->host[i], nf); return SOSBA_E_ARG; }
  int rc;
  const int chunk = std::min(pts->n, IMM_CHUNK);
  if ((rc = ensure_immature(h, (size_t)chunk, 1))) return rc;
  HostSide *hs = HS(h);
  const size_t out_bytes = (size_t)chunk * (4 + 1 + nf);
  if (out_bytes > hs->act_cap) {
    dfree(h, hs->d_act);
    hs->act_cap = out_bytes + out_bytes / 4 + 1024;
    DALLOC(h, hs->d_act, (hs->act_cap + 3) / 4);
  }
  if (nf * nf > hs->act_win_cap) {
    dfree(h, hs->d_act_win);
    hs->act_win_cap = nf * nf;
    DALLOC(h, hs->d_act_win, 14 * (size_t)hs->act_win_cap);
  }
  float *dw = hs->d_act_win;
  const size_t NP = (size_t)nf * nf;
  if ((rc = up(h, dw, win->RTll, 9 * NP)) || (rc = up(h, dw + 9 * NP, win->tTll, 3 * NP)) || (rc = up(h, dw + 12 * NP, win->aff, 2 * NP))) return rc;
  a.nf = nf; a.w = h->wl[0]; a.h = h->hl[0]; a.min_obs = win->min_obs;
  a.RTll = dw; a.tTll = dw + 9 * NP; a.aff = dw + 12 * NP;
  a.fxl = win->calib[0]; a.fyl = win->calib[1]; a.cxl = win->calib[2]; a.cyl = win->calib[3]; a.huberTH = h->cfg.huber_th;
  for (int off = 0; off < pts->n; off += IMM_CHUNK) {
    const int n = std::min(IMM_CHUNK, pts->n - off);
    const size_t N = (size_t)n, in_bytes = 22 * N * sizeof(float), ob = N * (4 + 1 + nf);
    char *blk;
    if ((rc = stage_reserve(h, std::max(in_bytes, ob), &blk))) return rc;
    float *s = (float *)blk, *d = hs->d_imm;
    // arena (floats): [u][v][color 8][weights 8][energyTH][idmin][idmax][host (int)]
    memcpy(s, pts->u + off, 4 * N); memcpy(s + N, pts->v + off, 4 * N);
    memcpy(s + 2 * N, pts->color + 8 * (size_t)off, 32 * N); memcpy(s + 10 * N, pts->weights + 8 * (size_t)off, 32 * N);
    memcpy(s + 18 * N, pts->energy_th + off, 4 * N); memcpy(s + 19 * N, pts->idepth_min + off, 4 * N); memcpy(s + 20 * N, pts->idepth_max + off, 4 * N);
    memcpy(s + 21 * N, pts->host + off, 4 * N);
    SOSBA_CUDA(cudaMemcpyAsync(d, s, in_bytes, cudaMemcpyHostToDevice, h->stream));
    a.n = n;
    a.u = d; a.v = d + N; a.color = d + 2 * N; a.weights = d + 10 * N; a.energyTH = d + 18 * N; a.idepth_min = d + 19 * N; a.idepth_max = d + 20 * N;
    a.host = (const int *)(d + 21 * N);
    a.idepth = hs->d_act; a.result = (signed char *)(hs->d_act + N); a.res_state = (uint8_t *)(hs->d_act + N) + N;
    launch_optimize_immature(h, a);
    SOSBA_CUDA(cudaGetLastError());
    // the input block has been consumed by the copy engine once the kernel ran; reuse it for the results
    SOSBA_CUDA(cudaMemcpyAsync(blk, hs->d_act, ob, cudaMemcpyDeviceToHost, h->stream));
    if ((rc = sync(h))) return rc;
    memcpy(idepth + off, blk, 4 * N); memcpy(result + off, blk + 4 * N, N); memcpy(res_state + (size_t)off * nf, blk + 5 * N, N * nf);
  }
  return SOSBA_OK;
}

// ---- next row: pixel selection --------------------------------------------

API int sosba_pixel_selector_set(sosba_t *h, const uint8_t *random_pattern, int32_t current_potential) {
  CHECK_H(h);
  if (!random_pattern || current_potential < 1) { sosba_set_error("pixel_selector_set: bad arguments"); return SOSBA_E_ARG; }
  if (h->levels < 3) { sosba_set_error("the selector reads pyramid levels 0..2"); return SOSBA_E_STATE; }
  HostSide *hs = HS(h);
  const size_t n = (size_t)h->cfg.w * h->cfg.h;
  const int w32 = h->cfg.w / 32, h32 = h->cfg.h / 32, chunks = (int)((n + 1023) / 1024);
  const int nb_max = ((h->cfg.w + 3) / 4) * ((h->cfg.h + 3) / 4);   // pot = 1
  int rc;
  if (!hs->d_sel_random) {
    DALLOC(h, hs->d_sel_random, n);
    DALLOC(h, hs->d_sel_map, n + 16);
    DALLOC(h, hs->d_sel_ths, 2 * ((size_t)w32 * h32 + 100));
    DALLOC(h, hs->d_sel_int, 8 + 2 * (size_t)chunks + 3 * (size_t)nb_max);
    hs->sel_list_cap = (int)n;
    DALLOC(h, hs->d_sel_list, n);
    SOSBA_CUDA(cudaMallocHost((void **)&hs->pin_sel_list, sizeof(int2) * (n + 8192 + 8)));
    hs->sel_nb_cap = nb_max;
    SOSBA_CUDA(cudaMemsetAsync(hs->d_sel_ths, 0, 2 * ((size_t)w32 * h32 + 100) * sizeof(float), h->stream));
  }
  hs->sel_random.assign(random_pattern, random_pattern + n);
  hs->sel_pot = current_potential;
  if ((rc = up_bytes(h, hs->d_sel_random, random_pattern, n))) return rc;
  return sync(h);
}

// PixelSelector::select on the device: fills d_sel_map, returns n2 / n3 / n4, the selected count and the raster-order list
static int select_device(sosba *h, int slot, int pot, float thFactor, int n234[3], int *n_list) {
  HostSide *hs = HS(h);
  const int w = h->cfg.w, hh = h->cfg.h, n = w * hh, chunks = (n + 1023) / 1024;
  SelectArgs a = {};
  a.w = w; a.h = hh; a.w1 = h->wl[1]; a.w2 = h->wl[2]; a.w32 = w / 32; a.h32 = hh / 32;
  a.img0 = h->slot_img[slot] + h->lvl_off[0]; a.img1 = h->slot_img[slot] + h->lvl_off[1]; a.img2 = h->slot_img[slot] + h->lvl_off[2];
  a.ths = hs->d_sel_ths; a.thsSmoothed = hs->d_sel_ths + ((size_t)a.w32 * a.h32 + 100);
  a.randomPattern = hs->d_sel_random;
  a.pot = pot; a.nbx = (w + 4 * pot - 1) / (4 * pot); a.nby = (hh + 4 * pot - 1) / (4 * pot);
  a.thFactor = thFactor;
  const int nb = a.nbx * a.nby;
  int *totals = hs->d_sel_int, *chunk_cnt = totals + 8, *chunk_off = chunk_cnt + chunks, *cntA = chunk_off + chunks, *cntB = cntA + hs->sel_nb_cap,
      *base = cntB + hs->sel_nb_cap;
  a.map = hs->d_sel_map; a.totals = totals; a.base = base;
  // pass 0: counts with tentative directions
  a.cnt_out = cntA; a.cnt_in = nullptr;
  launch_select_blocks(h, a, false);
  int rc;
  int *cin = cntA, *cout = cntB;
  for (int round = 0;; round++) {
    launch_select_scan(h, cin, base, nb, nullptr);
    SOSBA_CUDA(cudaMemsetAsync(totals, 0, 8 * sizeof(int), h->stream));
    SOSBA_CUDA(cudaMemsetAsync(hs->d_sel_map, 0, (size_t)n, h->stream));
    a.cnt_in = cin; a.cnt_out = cout;
    if (round < 4) launch_select_blocks(h, a, true);
    else launch_select_serial(h, a);   // counts that do not settle: sequential fallback (k_select.cu)
    launch_select_compact(h, hs->d_sel_map, n, chunk_cnt, chunk_off, totals + 4, hs->sel_list_cap, hs->d_sel_list);
    SOSBA_CUDA(cudaGetLastError());
    // totals and the head of the list come back together; a longer list needs a second copy
    const int fast = std::min(8192, hs->sel_list_cap);
    if ((rc = down(h, (int *)hs->pin_sel_list, (const int *)totals, 8)) || (rc = down(h, hs->pin_sel_list + 4, (const int2 *)hs->d_sel_list, fast))) return rc;
    if ((rc = sync(h))) return rc;
    const int *t = (const int *)hs->pin_sel_list;
    if (t[3] == 0 || round > 64) {
      if (t[3]) { sosba_set_error("pixel_select: direction counts did not reach a fixed point"); return SOSBA_E_STATE; }
      n234[0] = t[0]; n234[1] = t[1]; n234[2] = t[2];
      *n_list = t[4];
      if (t[4] > fast) {
        if ((rc = down(h, hs->pin_sel_list + 4 + fast, (const int2 *)hs->d_sel_list + fast, (size_t)t[4] - fast))) return rc;
        if ((rc = sync(h))) return rc;
      }
      return SOSBA_OK;
    }
    std::swap(cin, cout);   // the counts of this pass become the scan input of the next
  }
}

API int sosba_pixel_select(sosba_t *h, int32_t slot, float density, int32_t recursions_left, float th_factor, int32_t cap, int32_t *n_selected,
                           int32_t *u, int32_t *v, float *type, float *map_out, int32_t *current_potential) {
  CHECK_H(h);
  HostSide *hs = HS(h);
  if (hs->sel_random.empty()) { sosba_set_error("pixel_selector_set first"); return SOSBA_E_STATE; }
  if (slot < 0 || slot >= (int)h->slot_img.size() || !h->slot_valid[slot] || !(density > 0)) { sosba_set_error("bad slot / density"); return SOSBA_E_ARG; }
  const int w = h->cfg.w, hh = h->cfg.h;
  {  // makeHists (once per frame; the recursion of makeMaps reuses it)
    SelectArgs a = {};
    a.w = w; a.h = hh; a.w32 = w / 32; a.h32 = hh / 32;
    a.img0 = h->slot_img[slot] + h->lvl_off[0];
    a.ths = hs->d_sel_ths; a.thsSmoothed = hs->d_sel_ths + ((size_t)a.w32 * a.h32 + 100);
    if (a.w32 * a.h32 > 0) launch_select_hists(h, a);
  }
  // PixelSelector::makeMaps (PixelSelector2.cpp:146-282): the control flow stays on the host
  int rc, n234[3], n_list = 0;
  float numHave, quotia;
  const float numWant = density;
  int idealPotential;
  for (;;) {
    if ((rc = select_device(h, slot, hs->sel_pot, th_factor, n234, &n_list))) return rc;
    numHave = n234[0] + n234[1] + n234[2];
    quotia = numWant / numHave;
    const float K = numHave * (hs->sel_pot + 1) * (hs->sel_pot + 1);
    idealPotential = sqrtf(K / numWant) - 1;
    if (idealPotential < 1) idealPotential = 1;
    if (recursions_left > 0 && quotia > 1.25 && hs->sel_pot > 1) {
      if (idealPotential >= hs->sel_pot) idealPotential = hs->sel_pot - 1;
      hs->sel_pot = idealPotential; recursions_left--;
      continue;
    } else if (recursions_left > 0 && quotia < 0.25) {
      if (idealPotential <= hs->sel_pot) idealPotential = hs->sel_pot + 1;
      hs->sel_pot = idealPotential; recursions_left--;
      continue;
    }
    break;
  }
  // sub-sampling by the random pattern, indexed by the rank of the pixel among the selected ones in raster order (:226-238)
  const int2 *list = hs->pin_sel_list + 4;
  int numHaveSub = (int)numHave;
  const bool sub = quotia < 0.95;
  const unsigned char charTH = sub ? (unsigned char)(255 * quotia) : 255;
  if (map_out) memset(map_out, 0, sizeof(float) * (size_t)w * hh);
  int k = 0;
  for (int rn = 0; rn < n_list; rn++) {
    if (sub && hs->sel_random[rn] > charTH) { numHaveSub--; continue; }
    const int idx = list[rn].x;
    if (u || v || type) {
      if (k >= cap) { sosba_set_error("pixel_select: more than cap = %d selected pixels", cap); return SOSBA_E_ARG; }
      if (u) u[k] = idx % w;
      if (v) v[k] = idx / w;
      if (type) type[k] = (float)list[rn].y;
    }
    if (map_out) map_out[idx] = (float)list[rn].y;
    k++;
  }
  hs->sel_pot = idealPotential;
  if (n_selected) *n_selected = numHaveSub;
  if (current_potential) *current_potential = hs->sel_pot;
  return SOSBA_OK;
}

// ---- 8f rank 4, second half: CoarseInitializer::calcResAndGS (CoarseInitializer.cpp:450-673) ----------------------
API int sosba_init_calc_res_and_gs(sosba_t *h, int32_t lvl, int32_t ref_slot, int32_t new_slot, const double refToNew[12], const float aff[2],
                                   const float tlog[3], float alphaW, float alphaK, float couplingWeight, sosba_init_points *pts, float H[64], float b[8],
                                   float Hsc[64], float bsc[8], float res3[3]) {
  CHECK_H(h);
  HostSide *hs = HS(h);
  if (!h->t_haveK) { sosba_set_error("tracker_make_k first"); return SOSBA_E_STATE; }
  auto bad = [&](int s) { return s < 0 || s >= (int)h->slot_img.size() || !h->slot_valid[s]; };
  if (lvl < 0 || lvl >= h->levels || bad(ref_slot) || bad(new_slot) || !refToNew || !aff || !tlog || !pts || pts->n < 0 || !H || !b || !Hsc || !bsc || !res3) {
    sosba_set_error("init_calc_res_and_gs: bad arguments");
    return SOSBA_E_ARG;
  }
  const int n = pts->n;
  if (n > 0 && (!pts->u || !pts->v || !pts->idepth_new || !pts->iR || !pts->energy || !pts->outlierTH || !pts->isGood || !pts->energy_new ||
                !pts->isGood_new || !pts->maxstep || !pts->lastHessian_new || !pts->JbBuffer_new)) {
    sosba_set_error("null buffer");
    return SOSBA_E_ARG;
  }
  for (int i = 0; i < n; i++)   // the reference tap of firstFrame reads (u + dx, v + dy) .. +1 without a bounds test (:518-519)
    if (!(pts->u[i] >= 2 && pts->v[i] >= 2 && pts->u[i] < h->wl[lvl] - 3 && pts->v[i] < h->hl[lvl] - 3)) { sosba_set_error("point %d outside the pattern margin", i); return SOSBA_E_ARG; }
  int rc;
  const size_t N = ((size_t)n + 63) & ~(size_t)63;
  // arena (floats): in [u][v][idepth_new][iR][energy 2][outlierTH][isGood (bytes, N)] | out [energy_new 2][maxstep][lastHessian_new][Jb 10][isGood_new (bytes)] | 91 doubles
  const size_t in_f = 7 * N + N / 4, out_f = 14 * N + N / 4, total_f = in_f + out_f + 2 * 92;
  if (total_f > hs->init_cap) {
    dfree(h, hs->d_init);
    hs->init_cap = total_f + total_f / 4;
    DALLOC(h, hs->d_init, hs->init_cap);
  }
  float *d = hs->d_init;
  float *d_out = d + in_f;
  double *d_acc = (double *)(d + in_f + out_f);
  InitArgs a = {};
  a.n = n; a.w = h->wl[lvl]; a.h = h->hl[lvl];
  a.imgRef = h->slot_img[ref_slot] + h->lvl_off[lvl]; a.imgNew = h->slot_img[new_slot] + h->lvl_off[lvl];
  float R[9], Ki[9];
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R[3 * i + j] = (float)refToNew[4 * i + j]; a.t[i] = (float)refToNew[4 * i + 3]; }
  make_Ki(h->t_K[lvl], Ki);
  mul33f(R, Ki, a.RKi);
  a.fx = h->t_K[lvl][0]; a.fy = h->t_K[lvl][1]; a.cx = h->t_K[lvl][2]; a.cy = h->t_K[lvl][3];
  a.aff0 = (float)exp((double)aff[0]); a.aff1 = aff[1]; a.huberTH = h->cfg.huber_th; a.couplingWeight = couplingWeight;
  // alpha energy (:618-630): EAlpha stays empty in the reference, so the decision depends on the translation only
  const double tsq = refToNew[3] * refToNew[3] + refToNew[7] * refToNew[7] + refToNew[11] * refToNew[11];
  float alphaEnergy = alphaW * (0.f + tsq * n);
  float alphaOpt;
  if (alphaEnergy > alphaK * n) { alphaOpt = 0; alphaEnergy = alphaK * n; }
  else alphaOpt = alphaW;
  a.alphaOpt = alphaOpt;
  a.u = d; a.v = d + N; a.idepth_new = d + 2 * N; a.iR = d + 3 * N; a.energy = d + 4 * N; a.outlierTH = d + 6 * N; a.isGood = (const uint8_t *)(d + 7 * N);
  a.energy_new = d_out; a.maxstep = d_out + 2 * N; a.lastHessian_new = d_out + 3 * N; a.Jb = d_out + 4 * N; a.isGood_new = (uint8_t *)(d_out + 14 * N);
  a.acc = d_acc;
  if (n > 0) {
    const size_t in_bytes = in_f * 4, out_bytes = out_f * 4;
    if (in_bytes + out_bytes > hs->stage_cap / 2) { sosba_set_error("too many initializer points"); return SOSBA_E_ARG; }
    char *blk;
    if ((rc = stage_reserve(h, in_bytes + out_bytes, &blk))) return rc;
    float *sf = (float *)blk;
    memcpy(sf, pts->u, 4 * (size_t)n); memcpy(sf + N, pts->v, 4 * (size_t)n); memcpy(sf + 2 * N, pts->idepth_new, 4 * (size_t)n);
    memcpy(sf + 3 * N, pts->iR, 4 * (size_t)n); memcpy(sf + 4 * N, pts->energy, 8 * (size_t)n); memcpy(sf + 6 * N, pts->outlierTH, 4 * (size_t)n);
    memcpy(sf + 7 * N, pts->isGood, (size_t)n);
    // in/out members the kernel only partly rewrites: lastHessian_new and JbBuffer_new keep the caller's values for rejected points
    float *so = sf + in_f;
    memcpy(so + 3 * N, pts->lastHessian_new, 4 * (size_t)n); memcpy(so + 4 * N, pts->JbBuffer_new, 40 * (size_t)n);
    SOSBA_CUDA(cudaMemcpyAsync(d, sf, in_bytes + out_bytes, cudaMemcpyHostToDevice, h->stream));
    SOSBA_CUDA(cudaMemsetAsync(d_acc, 0, 92 * sizeof(double), h->stream));
    launch_init_res(h, a);
    SOSBA_CUDA(cudaGetLastError());
    SOSBA_CUDA(cudaMemcpyAsync(so, d_out, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    if ((rc = down(h, hs->pin_d, (const double *)d_acc, 91))) return rc;
    if ((rc = sync(h))) return rc;
    memcpy(pts->energy_new, so, 8 * (size_t)n); memcpy(pts->maxstep, so + 2 * N, 4 * (size_t)n); memcpy(pts->lastHessian_new, so + 3 * N, 4 * (size_t)n);
    memcpy(pts->JbBuffer_new, so + 4 * N, 40 * (size_t)n); memcpy(pts->isGood_new, so + 14 * N, (size_t)n);
  } else {
    for (int i = 0; i < 91; i++) hs->pin_d[i] = 0;
  }
  float A[2][9][9];
  for (int m = 0; m < 2; m++) {
    int q = 45 * m;
    for (int r = 0; r < 9; r++) for (int c = r; c < 9; c++) { A[m][r][c] = A[m][c][r] = (float)hs->pin_d[q++]; }
  }
  for (int r = 0; r < 8; r++) {
    for (int c = 0; c < 8; c++) { H[8 * r + c] = A[0][r][c]; Hsc[8 * r + c] = A[1][r][c]; }
    b[r] = A[0][r][8]; bsc[r] = A[1][r][8];
  }
  H[0] += alphaOpt * n; H[9] += alphaOpt * n; H[18] += alphaOpt * n;
  b[0] += tlog[0] * alphaOpt * n; b[1] += tlog[1] * alphaOpt * n; b[2] += tlog[2] * alphaOpt * n;
  res3[0] = (float)hs->pin_d[90]; res3[1] = alphaEnergy; res3[2] = (float)(2 * (size_t)n);   // E.num: both loops update E (:606-617)
  return SOSBA_OK;
}

// ---- resident immature-point pool ---------------------------------------------------------------
// arena of N = pool_n points (floats): [u][v][color 8][weights 8][gradH 4][energyTH] | [idmin][idmax][quality][uv 2][pixint] [host (int)] [status (bytes)]
static TraceArgs pool_trace_args(sosba *h, float *d, size_t N, int frame_slot) {
  TraceArgs a = {};
  a.n = (int)N; a.w = h->wl[0]; a.h = h->hl[0]; a.img = h->slot_img[frame_slot] + h->lvl_off[0];
  a.u = d; a.v = d + N; a.color = d + 2 * N; a.weights = d + 10 * N; a.gradH = d + 18 * N; a.energyTH = d + 22 * N;
  a.idepth_min = d + 23 * N; a.idepth_max = d + 24 * N; a.quality = d + 25 * N; a.uv = d + 26 * N; a.pixint = d + 28 * N;
  a.host = (const int *)(d + 29 * N); a.status = (uint8_t *)(d + 30 * N);
  a.huberTH = h->cfg.huber_th;
  return a;
}

API int sosba_immature_pool_set(sosba_t *h, const sosba_immature *pts) {
  CHECK_H(h);
  if (!pts || pts->n < 0) { sosba_set_error("pool_set: bad arguments"); return SOSBA_E_ARG; }
  HostSide *hs = HS(h);
  const int n = pts->n;
  hs->pool_n = 0;
  if (n == 0) return SOSBA_OK;
  if (!pts->host || !pts->u || !pts->v || !pts->color || !pts->weights || !pts->gradH || !pts->energy_th || !pts->idepth_min || !pts->idepth_max ||
      !pts->quality || !pts->last_trace_status || !pts->last_trace_uv || !pts->last_trace_pixel_interval) {
    sosba_set_error("null buffer");
    return SOSBA_E_ARG;
  }
  const size_t N = (size_t)n, bytes = 30 * N * sizeof(float) + N;
  int rc;
  if (31 * N > hs->pool_cap) {
    dfree(h, hs->d_pool);
    hs->pool_cap = 31 * (N + N / 4 + 256);
    DALLOC(h, hs->d_pool, hs->pool_cap);
  }
  std::vector<char> tmp;
  char *blk;
  if (bytes <= hs->stage_cap / 2) { if ((rc = stage_reserve(h, bytes, &blk))) return rc; }
  else { tmp.resize(bytes); blk = tmp.data(); }
  float *s = (float *)blk;
  memcpy(s, pts->u, 4 * N); memcpy(s + N, pts->v, 4 * N); memcpy(s + 2 * N, pts->color, 32 * N); memcpy(s + 10 * N, pts->weights, 32 * N);
  memcpy(s + 18 * N, pts->gradH, 16 * N); memcpy(s + 22 * N, pts->energy_th, 4 * N);
  memcpy(s + 23 * N, pts->idepth_min, 4 * N); memcpy(s + 24 * N, pts->idepth_max, 4 * N); memcpy(s + 25 * N, pts->quality, 4 * N);
  memcpy(s + 26 * N, pts->last_trace_uv, 8 * N); memcpy(s + 28 * N, pts->last_trace_pixel_interval, 4 * N);
  memcpy(s + 29 * N, pts->host, 4 * N); memcpy(s + 30 * N, pts->last_trace_status, N);
  SOSBA_CUDA(cudaMemcpyAsync(hs->d_pool, s, bytes, cudaMemcpyHostToDevice, h->stream));
  if (blk == tmp.data()) { if ((rc = sync(h))) return rc; }
  hs->pool_n = n;
  hs->pool_hosts = 0;
  for (int i = 0; i < n; i++) hs->pool_hosts = std::max(hs->pool_hosts, pts->host[i] + 1);
  for (int i = 0; i < n; i++) if (pts->host[i] < 0) { hs->pool_n = 0; sosba_set_error("point %d: negative host", i); return SOSBA_E_ARG; }
  return SOSBA_OK;
}

API int sosba_immature_pool_trace(sosba_t *h, int32_t frame_slot, int32_t nhosts, const float *KRKi, const float *Kt, const float *aff, int32_t counts[6]) {
  CHECK_H(h);
  HostSide *hs = HS(h);
  if (frame_slot < 0 || frame_slot >= (int)h->slot_img.size() || !h->slot_valid[frame_slot] || nhosts < 0) { sosba_set_error("bad slot"); return SOSBA_E_ARG; }
  if (counts) for (int i = 0; i < 6; i++) counts[i] = 0;
  if (hs->pool_n == 0) return SOSBA_OK;
  if (!KRKi || !Kt || !aff || nhosts < hs->pool_hosts) { sosba_set_error("pool_trace: %d hosts given, the pool refers to %d", nhosts, hs->pool_hosts); return SOSBA_E_ARG; }
  int rc;
  if ((rc = ensure_immature(h, 1, nhosts))) return rc;
  float *dh = hs->d_imm_host;
  int *d_counts = (int *)(dh + 14 * (size_t)hs->imm_host_cap);
  // the per-host tables go up as ONE block: [KRKi 9 x cap][Kt 3 x cap][aff 2 x cap] is how the kernel addresses them
  {
    char *blk;
    const size_t cap = hs->imm_host_cap, bytes = 14 * cap * sizeof(float);
    if ((rc = stage_reserve(h, bytes, &blk))) return rc;
    float *s = (float *)blk;
    memcpy(s, KRKi, 36 * (size_t)nhosts); memcpy(s + 9 * cap, Kt, 12 * (size_t)nhosts); memcpy(s + 12 * cap, aff, 8 * (size_t)nhosts);
    SOSBA_CUDA(cudaMemcpyAsync(dh, s, bytes, cudaMemcpyHostToDevice, h->stream));
  }
  SOSBA_CUDA(cudaMemsetAsync(d_counts, 0, 6 * sizeof(int), h->stream));
  TraceArgs a = pool_trace_args(h, hs->d_pool, (size_t)hs->pool_n, frame_slot);
  a.KRKi = dh; a.Kt = dh + 9 * (size_t)hs->imm_host_cap; a.aff = dh + 12 * (size_t)hs->imm_host_cap; a.counts = d_counts;
  launch_trace_on(h, a);
  SOSBA_CUDA(cudaGetLastError());
  if (counts) {
    if ((rc = down(h, hs->pin_i, (const int *)d_counts, 6))) return rc;
    if ((rc = sync(h))) return rc;
    for (int i = 0; i < 6; i++) counts[i] = hs->pin_i[i];
  }
  return SOSBA_OK;
}

API int sosba_immature_pool_get(sosba_t *h, sosba_immature *pts) {
  CHECK_H(h);
  HostSide *hs = HS(h);
  if (!pts || pts->n != hs->pool_n) { sosba_set_error("pool_get: pts->n must equal the pool size %d", hs->pool_n); return SOSBA_E_ARG; }
  const size_t N = (size_t)hs->pool_n;
  if (N == 0) return SOSBA_OK;
  if (!pts->idepth_min || !pts->idepth_max || !pts->quality || !pts->last_trace_status || !pts->last_trace_uv || !pts->last_trace_pixel_interval) {
    sosba_set_error("null buffer");
    return SOSBA_E_ARG;
  }
  int rc;
  const size_t bytes = 7 * N * sizeof(float) + N;   // [idmin .. host] + status
  std::vector<char> tmp;
  char *blk;
  if (bytes <= hs->stage_cap / 2) { if ((rc = stage_reserve(h, bytes, &blk))) return rc; }
  else { tmp.resize(bytes); blk = tmp.data(); }
  SOSBA_CUDA(cudaMemcpyAsync(blk, hs->d_pool + 23 * N, bytes, cudaMemcpyDeviceToHost, h->stream));
  if ((rc = sync(h))) return rc;
  const float *s = (const float *)blk;
  memcpy(pts->idepth_min, s, 4 * N); memcpy(pts->idepth_max, s + N, 4 * N); memcpy(pts->quality, s + 2 * N, 4 * N);
  memcpy(pts->last_trace_uv, s + 3 * N, 8 * N); memcpy(pts->last_trace_pixel_interval, s + 5 * N, 4 * N);
  memcpy(pts->last_trace_status, s + 7 * N, N);
  return SOSBA_OK;
}
