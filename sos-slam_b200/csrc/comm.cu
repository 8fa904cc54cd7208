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
// The per-iteration exchange of the Gauss-Newton loop does not go through NCCL: k_stitch_xchg (k_xchg.cu) writes every
// stitched value straight into the peers' mailboxes over NVLink.  Here: the mailboxes (cudaIpc-mapped at sosba_comm_init,
// handles all-gathered through the NCCL communicator that exists anyway) and the NCCL collectives of the API-level calls
// outside the loop (sosba_accumulate, sosba_marginalize_points, sosba_linearize_all) and of the fallback when peer
// mapping is unavailable (SOSBA_COMM_NCCL=1 forces it).

#define API extern "C" __attribute__((visibility("default")))

// map every rank's mailbox into this process; any failure leaves the NCCL path in place
static int p2p_setup(sosba *h) {
  h->p2p = false;
  if (h->world < 2 || h->world > 8 || getenv("SOSBA_COMM_NCCL")) return SOSBA_OK;
  h->p2p_slot_bytes = stitch_xchg_slot_bytes(SOSBA_XCHG_MAX_NF, SOSBA_XCHG_MAX_NEWE);
  const size_t bytes = 2 * (size_t)h->world * h->p2p_slot_bytes;
  int ok = 1;
  if (cudaMalloc((void **)&h->p2p_mbox, bytes) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  if (ok && cudaMalloc((void **)&h->p2p_epoch_dev, 2 * sizeof(int)) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) {
    cudaMemset(h->p2p_mbox, 0, bytes);   // exchange numbers start at 1: a zeroed word is "not there yet"
    const int init[2] = {1, 0};
    cudaMemcpy(h->p2p_epoch_dev, init, sizeof(init), cudaMemcpyHostToDevice);
    if (cudaIpcGetMemHandle(&mine, h->p2p_mbox) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  }
  // all-gather the handles through the communicator that exists already (every rank takes part, whatever its own state)
  unsigned char *d_h = nullptr;
  if (cudaMalloc((void **)&d_h, (size_t)(h->world + 1) * sizeof(mine)) != cudaSuccess) { cudaGetLastError(); return SOSBA_E_CUDA; }
  cudaMemcpy(d_h, &mine, sizeof(mine), cudaMemcpyHostToDevice);
  ncclResult_t r = g_nccl.AllGather(d_h, d_h + sizeof(mine), sizeof(mine), /*ncclInt8*/ 0, (ncclComm_t)h->comm, h->stream);
  cudaStreamSynchronize(h->stream);
  std::vector<cudaIpcMemHandle_t> all(h->world);
  cudaMemcpy(all.data(), d_h + sizeof(mine), (size_t)h->world * sizeof(mine), cudaMemcpyDeviceToHost);
  cudaFree(d_h);
  if (r != 0) ok = 0;
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
    if (h->p2p_epoch_dev) cudaFree(h->p2p_epoch_dev);
    if (h->d_comm_int) cudaFree(h->d_comm_int);
    h->d_comm_int = nullptr;
    h->p2p_mbox = nullptr; h->p2p_epoch_dev = nullptr; h->p2p = false;
    h->comm = nullptr; h->world = 1; h->rank = 0;
  }
  return SOSBA_OK;
}

// NCCL sum of the un-stitched block tables over ranks, in place, on the compute stream (API-level calls outside the
// Gauss-Newton loop, and the loop itself when the peer mailboxes are unavailable); the same aggregated launch carries the
// back-substitution sums of the previous loop body (so that every rank takes the same break decision), the residual
// counters and — when `with_newE` — the newest-frame energies of the last linearisation (a sum with zeros elsewhere =
// exact concatenation, so every rank selects the same 70th percentile as a single GPU would)
int sosba_allreduce_acc(sosba *h, int with_newE) {
  if (!h->comm || h->world <= 1) return SOSBA_OK;
  const int nf = h->nf, D = 4 + 8 * nf;
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

// arguments of the fused stitch + exchange for this handle; push = 0 when the mailboxes are not in use
void sosba_xchg_args(sosba *h, StitchXchgArgs *a, int with_newE) {
  a->rank = h->rank; a->world = h->world; a->slot_bytes = h->p2p_slot_bytes;
  for (int p = 0; p < 8; p++) a->peer[p] = p < h->world ? h->p2p_peer[p] : nullptr;
  a->epoch = h->p2p_epoch_dev;
  a->newE_all = h->d_newE_all; a->newE_cnt = h->d_newE_cnt; a->newE_cap = h->newE_cap;
  a->with_newE = with_newE && h->d_newE_all;
  static const char *bo = getenv("SOSBA_XCHG_BACKOFF");
  a->backoff_ns = bo ? (unsigned)atoi(bo) : 0u;
  a->dbg = nullptr;
  a->push = (h->comm && h->world > 1 && h->p2p && h->nf <= SOSBA_XCHG_MAX_NF && h->newE_cap <= SOSBA_XCHG_MAX_NEWE) ? 1 : 0;
}

// the linearisation sums of an API-level linearizeAll: energy (1 double), state histogram + removals (4 ints), energies
int sosba_allreduce_lin(sosba *h, int with_stats, double *extra, int n_extra) {
  if (!h->comm || h->world <= 1) return SOSBA_OK;
  ncclComm_t c = (ncclComm_t)h->comm;
  NCCLCHK(g_nccl.GroupStart());
  if (extra && n_extra > 0) NCCLCHK(g_nccl.AllReduce(extra, extra, n_extra, ncclFloat64, ncclSum, c, h->stream));
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
