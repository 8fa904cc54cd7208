// comm.cu — multi-GPU: active points shard across the ranks of one box; the only data-path
// exchange per Gauss-Newton iteration is ONE sum all-reduce of the fp64 block tables ([top A | top L] + Schur
// Gram matrix + residual counts), issued on the compute stream so it is ordered between the accumulation
// kernels and the stitch.  NCCL is bound lazily with dlopen so that single-GPU use needs only libcudart
// (torch's bundled libnccl.so.2 is picked up when the caller already loaded it).
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "kernels.h"
#include "energy_th.cuh"

namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt32 = 2, ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };
struct Nccl {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

int load_nccl() {
  if (g_nccl.lib) return SOSBA_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) { g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
  if (!g_nccl.lib) { sosba_set_error("dlopen(libnccl.so.2) failed: %s", dlerror()); return SOSBA_E_NCCL; }
#define BIND(f) *(void **)(&g_nccl.f) = dlsym(g_nccl.lib, "nccl" #f); if (!g_nccl.f) { sosba_set_error("nccl" #f " missing"); return SOSBA_E_NCCL; }
  BIND(GetUniqueId) BIND(CommInitRank) BIND(CommDestroy) BIND(AllReduce) BIND(AllGather) BIND(GroupStart) BIND(GroupEnd) BIND(GetErrorString)
#undef BIND
  return SOSBA_OK;
}
#define NCCLCHK(expr) do { ncclResult_t r__ = (expr); if (r__ != 0) { sosba_set_error("%s -> %s", #expr, g_nccl.GetErrorString(r__)); return SOSBA_E_NCCL; } } while (0)
}  // namespace

// ---- peer-memory exchange -------------------------------------------------------------------------
// (Experimental, enabled with SOSBA_COMM_P2P=1; results identical to the NCCL path to 1e-14, but slower -- see DESIGN.md §6.)
// The per-iteration exchange is ~0.14 MB per rank: latency, not bandwidth.  Instead of an NCCL all-reduce every rank PUSHES
// its partial block tables (+ back-substitution sums, residual counters, its segment of newest-frame energies) straight
// into a mailbox slot in every peer's HBM with plain stores over NVLink (the mailboxes are cudaIpc-mapped at
// sosba_comm_init) and then sums the slots of its own mailbox in rank order -- local reads, and the same summation order on
// every rank, so all ranks hold bit-identical tables.  Two parities of slots: a slot of parity p is rewritten two exchanges
// later, which a peer can only reach after it has received this rank's words of the exchange in between, i.e. after this
// rank finished reading.  Every word carries the exchange number; a spin that does not see it within ~2 s raises the
// solver's non-finite flag instead of hanging the stream.
#define P2P_FLAG_BYTES 256   // mailbox header (unused by the in-band-flag protocol; keeps the slots 256-byte aligned)
#define P2P_MAX_NEWE 16384   // floats of newest-frame energies per rank a slot can carry

struct XchgArgs {
  int rank, world, epoch;
  size_t slot_bytes;
  unsigned char *peer[8];     // mailboxes of all ranks
  // payload (device pointers of this rank)
  double *tables; size_t nd;  // top A | L | Schur Gram
  double *rstats;             // [8]
  int *cnt;                   // [2]
  float *newE_all; int *newE_cnt; int newE_cap; int with_newE;
  int *ticket;
  const int *gate;
  int *err;                   // loop-control word that reports a failed exchange (non-finite flag)
};

__device__ __forceinline__ uint2 *slot_of(const XchgArgs &a, int owner, int parity, int from) {
  return (uint2 *)(a.peer[owner] + P2P_FLAG_BYTES + ((size_t)parity * a.world + from) * a.slot_bytes);
}
// Slot layout, in 8-byte words {payload 32 bits, exchange number}: the flag travels inside every word (an aligned 8-byte
// store is atomic over NVLink), so there is no fence and no separate "data ready" flag -- the receiver spins on the word
// it needs (the low-latency protocol of the collective libraries).  Words: [2 per double: tables nd, rstats 8]
// [cnt 2, newE count 1, pad 1] [newE: 1 per float].
__device__ __forceinline__ void ll_store(uint2 *p, unsigned v, unsigned epoch) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v), "r"(epoch) : "memory");
}
__device__ __forceinline__ bool ll_load(const uint2 *p, unsigned epoch, unsigned &v, long long t0) {
  unsigned f;
  int spins = 0;
  for (;;) {
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(f) : "l"(p) : "memory");
    if (f == epoch) return true;
    if ((++spins & 1023) == 0 && clock64() - t0 > 4000000000LL) return false;   // ~2 s at 1.9 GHz: give up instead of hanging the stream
  }
}

// One launch: push -> (the peers' words arrive) -> sum in rank order -> optionally setNewFrameEnergyTH on the gathered
// energies in the CTA that finishes last.  A thread pushes and later sums the same indices, so the in-place result never
// overwrites a value that still has to go out; no CTA waits for another CTA of the same launch, so the grid can be as
// wide as the tables (one entry per thread: the exchange is a latency problem).
__global__ void __launch_bounds__(256) k_xchg(XchgArgs a, ThArgs th, int do_th) {
  if (a.gate && *a.gate) return;
  const int parity = a.epoch & 1;
  const unsigned ep = (unsigned)a.epoch;
  const size_t nd = a.nd + 8;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const long long t0 = clock64();
  __shared__ int s_last;
  // ---- push into every mailbox (this rank's own included)
  const int my_n = a.with_newE ? a.newE_cnt[a.rank] : 0;
  for (size_t i = tid; i < nd; i += nth) {
    const double d = i < a.nd ? a.tables[i] : a.rstats[i - a.nd];
    const unsigned lo = (unsigned)__double2loint(d), hi = (unsigned)__double2hiint(d);
    for (int p = 0; p < a.world; p++) {   // one 16-byte store per double: each 8-byte half validates itself, so tearing between the halves is harmless
      uint2 *dst = slot_of(a, p, parity, a.rank) + 2 * i;
      asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %2};" ::"l"(dst), "r"(lo), "r"(ep), "r"(hi) : "memory");
    }
  }
  if (tid < 3) {
    const unsigned v = tid < 2 ? (unsigned)a.cnt[tid] : (unsigned)my_n;
    for (int p = 0; p < a.world; p++) ll_store(slot_of(a, p, parity, a.rank) + 2 * nd + tid, v, ep);
  }
  if (a.with_newE) {
    const float *src = a.newE_all + (size_t)a.rank * a.newE_cap;
    for (int i = tid; i < my_n; i += nth) {
      const unsigned v = __float_as_uint(src[i]);
      for (int p = 0; p < a.world; p++) ll_store(slot_of(a, p, parity, a.rank) + 2 * nd + 4 + i, v, ep);
    }
  }
  // ---- sum in rank order (identical on every rank); the words of all ranks are requested before any is checked, so a
  // thread pays one L2 round trip, not one per rank
  bool ok = true;
  for (size_t i = tid; i < nd && ok; i += nth) {
    unsigned lo[8], hi[8], flo[8], fhi[8];
#pragma unroll
    for (int r = 0; r < 8; r++)
      if (r < a.world) {
        const uint2 *src = slot_of(a, a.rank, parity, r) + 2 * i;
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo[r]), "=r"(flo[r]), "=r"(hi[r]), "=r"(fhi[r]) : "l"(src) : "memory");
      }
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < 8; r++)
      if (r < a.world) {
        if (flo[r] != ep || fhi[r] != ep) {   // not there yet: wait for this rank's words
          const uint2 *src = slot_of(a, a.rank, parity, r) + 2 * i;
          ok = ok && ll_load(src, ep, lo[r], t0) && ll_load(src + 1, ep, hi[r], t0);
        }
        s += __hiloint2double((int)hi[r], (int)lo[r]);
      }
    if (i < a.nd) a.tables[i] = s; else a.rstats[i - a.nd] = s;
  }
  if (tid < 2 && ok) {
    int s = 0;
    for (int r = 0; r < a.world; r++) { unsigned v; ok = ok && ll_load(slot_of(a, a.rank, parity, r) + 2 * nd + tid, ep, v, t0); s += (int)v; }
    a.cnt[tid] = s;
  }
  if (a.with_newE) {
    for (int r = 0; r < a.world && ok; r++) {
      const uint2 *src = slot_of(a, a.rank, parity, r) + 2 * nd;
      unsigned n;
      ok = ok && ll_load(src + 2, ep, n, t0);
      if (!ok) break;
      if (tid == 0) a.newE_cnt[r] = (int)n;
      float *dst = a.newE_all + (size_t)r * a.newE_cap;
      for (int i = tid; i < (int)n && ok; i += nth) { unsigned v; ok = ll_load(src + 4 + i, ep, v, t0); dst[i] = __uint_as_float(v); }
    }
  }
  if (!ok && a.err) *a.err = 1;
  if (!do_th) return;
  // ---- the CTA that finishes last selects the 70th percentile of the gathered energies (setNewFrameEnergyTH)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(a.ticket + 1, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) a.ticket[1] = 0;
  __threadfence();
  energy_th_body(th);
}

#define API extern "C" __attribute__((visibility("default")))

// map every rank's mailbox into this process; any failure leaves the NCCL path in place
static int p2p_setup(sosba *h) {
  h->p2p = false;
  // Opt-in: measured on 2 and 4 B200s the NCCL all-reduce of the same payload is faster, so it stays the default.
  if (h->world < 2 || h->world > 8 || !getenv("SOSBA_COMM_P2P")) return SOSBA_OK;
  const size_t max_nd = 2 * (size_t)13 * 13 * SOSBA_TOPB + (size_t)109 * 109 + 8;   // k_solve's limit: nf <= 13, D + 1 <= 109
  h->p2p_slot_bytes = (((2 * max_nd + 4 + (size_t)P2P_MAX_NEWE) * 8) + 255) & ~(size_t)255;   // 8-byte words {payload, exchange number}
  const size_t bytes = P2P_FLAG_BYTES + 2 * (size_t)h->world * h->p2p_slot_bytes;
  if (cudaMalloc((void **)&h->p2p_mbox, bytes) != cudaSuccess) { cudaGetLastError(); return SOSBA_OK; }
  cudaMemset(h->p2p_mbox, 0, bytes);
  if (cudaMalloc((void **)&h->p2p_ticket, 2 * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return SOSBA_OK; }
  cudaMemset(h->p2p_ticket, 0, 2 * sizeof(int));
  cudaIpcMemHandle_t mine;
  if (cudaIpcGetMemHandle(&mine, h->p2p_mbox) != cudaSuccess) { cudaGetLastError(); return SOSBA_OK; }
  // all-gather the handles through the communicator that exists already
  unsigned char *d_h = nullptr;
  if (cudaMalloc((void **)&d_h, (size_t)(h->world + 1) * sizeof(mine)) != cudaSuccess) { cudaGetLastError(); return SOSBA_OK; }
  cudaMemcpy(d_h, &mine, sizeof(mine), cudaMemcpyHostToDevice);
  ncclResult_t r = g_nccl.AllGather(d_h, d_h + sizeof(mine), sizeof(mine), /*ncclInt8*/ 0, (ncclComm_t)h->comm, h->stream);
  cudaStreamSynchronize(h->stream);
  std::vector<cudaIpcMemHandle_t> all(h->world);
  cudaMemcpy(all.data(), d_h + sizeof(mine), (size_t)h->world * sizeof(mine), cudaMemcpyDeviceToHost);
  cudaFree(d_h);
  if (r != 0) return SOSBA_OK;
  int ok = 1;
  for (int p = 0; p < h->world && ok; p++) {
    if (p == h->rank) { h->p2p_peer[p] = h->p2p_mbox; continue; }
    void *q = nullptr;
    if (cudaIpcOpenMemHandle(&q, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
    h->p2p_peer[p] = (unsigned char *)q;
  }
  // every rank must take the same path: agree on the minimum
  int agreed = ok;
  {
    int *d = nullptr;
    cudaMalloc(&d, sizeof(int));
    cudaMemcpy(d, &ok, sizeof(int), cudaMemcpyHostToDevice);
    g_nccl.AllReduce(d, d, 1, ncclInt32, /*ncclMin*/ 3, (ncclComm_t)h->comm, h->stream);
    cudaStreamSynchronize(h->stream);
    cudaMemcpy(&agreed, d, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d);
  }
  h->p2p = agreed == 1;
  h->p2p_epoch = 0;
  return SOSBA_OK;
}

API int sosba_comm_unique_id(uint8_t id[128]) {
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId u;
  NCCLCHK(g_nccl.GetUniqueId(&u));
  memcpy(id, u.internal, 128);
  return SOSBA_OK;
}

API int sosba_comm_init(sosba_t *h, const uint8_t id[128], int32_t rank, int32_t world) {
  if (!h || !id || world < 1 || rank < 0 || rank >= world) return SOSBA_E_ARG;
  int rc = load_nccl();
  if (rc) return rc;
  cudaSetDevice(h->device);
  ncclUniqueId u;
  memcpy(u.internal, id, 128);
  ncclComm_t c = nullptr;
  NCCLCHK(g_nccl.CommInitRank(&c, world, u, rank));
  h->comm = c; h->rank = rank; h->world = world;
  return p2p_setup(h);
}

// 1 = the per-iteration exchange goes through peer memory, 0 = NCCL all-reduce
API int sosba_comm_uses_peer_memory(const sosba_t *h) { return h && h->p2p ? 1 : 0; }

API int sosba_comm_destroy(sosba_t *h) {
  if (h && h->comm) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (int p = 0; p < h->world && p < 8; p++) {
      if (p != h->rank && h->p2p_peer[p]) cudaIpcCloseMemHandle(h->p2p_peer[p]);
      h->p2p_peer[p] = nullptr;
    }
    g_nccl.CommDestroy((ncclComm_t)h->comm);   // (a collective: every peer has stopped writing into this rank's mailbox)
    if (h->p2p_mbox) cudaFree(h->p2p_mbox);
    if (h->p2p_ticket) cudaFree(h->p2p_ticket);
    if (h->d_comm_int) cudaFree(h->d_comm_int);
    h->d_comm_int = nullptr;
    h->p2p_mbox = nullptr; h->p2p_ticket = nullptr; h->p2p = false;
    h->comm = nullptr; h->world = 1; h->rank = 0;
  }
  return SOSBA_OK;
}

// sum of the block tables over ranks, in place, on the compute stream; the same (aggregated) launch carries the
// back-substitution sums of the previous loop body (so that every rank takes the same break decision), the residual
// counters and — when `with_newE` — the newest-frame energies of the last linearisation (a sum with zeros elsewhere =
// exact concatenation, so every rank selects the same 70th percentile as a single GPU would)
// *th_done = 1 when the exchange kernel also ran setNewFrameEnergyTH (`th` non-null, peer-memory path)
int sosba_allreduce_acc(sosba *h, int with_newE, const int *gate, int *err, const ThArgs *th, int *th_done) {
  if (th_done) *th_done = 0;
  if (!h->comm || h->world <= 1) return SOSBA_OK;
  const int nf = h->nf, D = 4 + 8 * nf;
  if (h->p2p && (!with_newE || (h->d_newE_all && h->newE_cap <= P2P_MAX_NEWE))) {
    XchgArgs a = {};
    a.rank = h->rank; a.world = h->world; a.epoch = ++h->p2p_epoch; a.slot_bytes = h->p2p_slot_bytes;
    for (int p = 0; p < h->world; p++) a.peer[p] = h->p2p_peer[p];
    a.tables = h->d_accTop; a.nd = 2 * (size_t)nf * nf * SOSBA_TOPB + (size_t)(D + 1) * (D + 1);
    a.rstats = h->d_rstats_all; a.cnt = h->d_cnt_all;
    a.newE_all = h->d_newE_all; a.newE_cnt = h->d_newE_cnt; a.newE_cap = h->newE_cap; a.with_newE = with_newE && h->d_newE_all;
    a.ticket = h->p2p_ticket; a.gate = gate; a.err = err;
    const int do_th = with_newE && th != nullptr;
    const int blocks = (int)((a.nd + 8 + 255) / 256);   // one table entry per thread
    k_xchg<<<blocks, 256, 0, h->stream>>>(a, do_th ? *th : ThArgs{}, do_th);
    h->launches += 1;
    if (th_done) *th_done = do_th;
    return SOSBA_OK;
  }
  ncclComm_t c = (ncclComm_t)h->comm;
  NCCLCHK(g_nccl.GroupStart());
  // d_accTop (A | L) and d_accSC are adjacent in the scratch region: one fp64 sum
  const size_t nd = 2 * (size_t)nf * nf * SOSBA_TOPB + (size_t)(D + 1) * (D + 1);
  NCCLCHK(g_nccl.AllReduce(h->d_accTop, h->d_accTop, nd, ncclFloat64, ncclSum, c, h->stream));
  NCCLCHK(g_nccl.AllReduce(h->d_rstats_all, h->d_rstats_all, 8, ncclFloat64, ncclSum, c, h->stream));
  NCCLCHK(g_nccl.AllReduce(h->d_cnt_all, h->d_cnt_all, 2, ncclInt32, ncclSum, c, h->stream));
  if (with_newE && h->d_newE_all) {
    NCCLCHK(g_nccl.AllReduce(h->d_newE_all, h->d_newE_all, (size_t)h->world * h->newE_cap, ncclFloat32, ncclSum, c, h->stream));
    NCCLCHK(g_nccl.AllReduce(h->d_newE_cnt, h->d_newE_cnt, h->world, ncclInt32, ncclSum, c, h->stream));
  }
  NCCLCHK(g_nccl.GroupEnd());
  h->launches += 1;
  return SOSBA_OK;
}

// the linearisation sums of an API-level linearizeAll: energy (1 double), state histogram + removals (4 ints), energies
int sosba_allreduce_lin(sosba *h, int with_stats) {
  if (!h->comm || h->world <= 1) return SOSBA_OK;
  ncclComm_t c = (ncclComm_t)h->comm;
  NCCLCHK(g_nccl.GroupStart());
  if (with_stats) {
    NCCLCHK(g_nccl.AllReduce(h->d_stats, h->d_stats, 1, ncclFloat64, ncclSum, c, h->stream));
    NCCLCHK(g_nccl.AllReduce(h->d_counts, h->d_counts, 4, ncclInt32, ncclSum, c, h->stream));
  }
  if (h->d_newE_all) {
    NCCLCHK(g_nccl.AllReduce(h->d_newE_all, h->d_newE_all, (size_t)h->world * h->newE_cap, ncclFloat32, ncclSum, c, h->stream));
    NCCLCHK(g_nccl.AllReduce(h->d_newE_cnt, h->d_newE_cnt, h->world, ncclInt32, ncclSum, c, h->stream));
  }
  NCCLCHK(g_nccl.GroupEnd());
  h->launches += 1;
  return SOSBA_OK;
}

// max of one host integer over ranks (setup only: synchronises)
int sosba_comm_max_int(sosba *h, int v, int *out) {
  *out = v;
  if (!h->comm || h->world <= 1) return SOSBA_OK;
  if (!h->d_comm_int && cudaMalloc(&h->d_comm_int, sizeof(int)) != cudaSuccess) return SOSBA_E_CUDA;   // kept: a per-call cudaMalloc / cudaFree costs more than the reduction
  int *d = h->d_comm_int;
  cudaMemcpyAsync(d, &v, sizeof(int), cudaMemcpyHostToDevice, h->stream);
  ncclResult_t r = g_nccl.AllReduce(d, d, 1, ncclInt32, ncclMax, (ncclComm_t)h->comm, h->stream);
  cudaMemcpyAsync(out, d, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  cudaStreamSynchronize(h->stream);
  if (r != 0) { sosba_set_error("ncclAllReduce(max) -> %s", g_nccl.GetErrorString(r)); return SOSBA_E_NCCL; }
  return SOSBA_OK;
}
