// comm.cu — multi-GPU: active points shard across the ranks of one box; the only data-path
// exchange per Gauss-Newton iteration is ONE sum all-reduce of the fp64 block tables ([top A | top L] + Schur
// Gram matrix + residual counts), issued on the compute stream so it is ordered between the accumulation
// kernels and the stitch.  NCCL is bound lazily with dlopen so that single-GPU use needs only libcudart
// (torch's bundled libnccl.so.2 is picked up when the caller already loaded it).
#include <dlfcn.h>
#include <string.h>

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

#define API extern "C" __attribute__((visibility("default")))

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
  return SOSBA_OK;
}

API int sosba_comm_destroy(sosba_t *h) {
  if (h && h->comm) { g_nccl.CommDestroy((ncclComm_t)h->comm); h->comm = nullptr; h->world = 1; h->rank = 0; }
  return SOSBA_OK;
}

// sum of the block tables over ranks, in place, on the compute stream; the same (aggregated) launch carries the
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
  int *d = nullptr;
  if (cudaMalloc(&d, sizeof(int)) != cudaSuccess) return SOSBA_E_CUDA;
  cudaMemcpyAsync(d, &v, sizeof(int), cudaMemcpyHostToDevice, h->stream);
  ncclResult_t r = g_nccl.AllReduce(d, d, 1, ncclInt32, ncclMax, (ncclComm_t)h->comm, h->stream);
  cudaMemcpyAsync(out, d, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  cudaStreamSynchronize(h->stream);
  cudaFree(d);
  if (r != 0) { sosba_set_error("ncclAllReduce(max) -> %s", g_nccl.GetErrorString(r)); return SOSBA_E_NCCL; }
  return SOSBA_OK;
}
