// k_xchg.cu — the stitch of the top blocks FUSED with the point-shard exchange: one launch per Gauss-Newton iteration.
//
//   compute   AccumulatedTopHessianSSE::stitchDoubleInternal + the symmetrisation epilogue of stitchDoubleMT
//             (AccumulatedTopHessian.cpp:231-301, .h:80-127) in GATHER form: one CTA per output tile of H
//             (frame a x frame b, a <= b), summing the adjoint sandwiches of the (host,target) blocks that reach it in a
//             fixed order — no atomics, no zeroed H, the same bits on every run;
//   exchange  (point shards) every finished value goes straight from the producing thread into every
//             peer's mailbox over NVLink (cudaIpc-mapped HBM) as a 16-byte {lo, n, hi, n} word, n = exchange number: the
//             flag travels inside the data, so there is no fence, no separate "ready" flag and no collective launch; the
//             thread then reads the same word of every peer from its own mailbox and adds them in rank order, which
//             makes H, b, the Schur Gram matrix, the back-substitution sums and the residual counters bit-identical on
//             every rank (so every rank takes the same loop-break decision, FullSystemOptimize.cpp:411).  The
//             newest-frame energies of setNewFrameEnergyTH (:84-124) travel the same way as 8-byte {float, n} words and
//             are concatenated, not summed.
// Payload per rank at nf = 8: 1 684 stitched words + 2 415 live Schur words + 31 = 66 KB of 16-byte words per peer (the
// un-stitched tables of round 1 were 264 KB), written tile by tile while the other tiles are still being computed.
// Mailbox slots are double-buffered by exchange parity: a slot is rewritten two exchanges later, which a peer can only
// reach after it consumed this rank's words of the exchange in between, i.e. after this rank finished reading.
#include "kernels.h"
#include "top_entries.cuh"

namespace {

constexpr int XT = 256;                 // threads per CTA
constexpr int XW_DIAG = 104;            // words of a diagonal tile: 64 H[a,a] + 32 H[a,calib] + 8 b[a]
constexpr int XW_MISC = 32;             // 16 H[calib,calib] + 4 b[calib] + 8 back-substitution sums + 2 residual counters (+2 spare)
constexpr long long X_TIMEOUT = 20000000000LL;   // ~10 s of clock64: give up instead of hanging the stream

struct Layout {
  int nf, D, DP, npairs, base_pair, base_misc, base_sc, nv;
  int cta_pair, cta_misc, cta_sc, cta_e, n_cta;
};
__host__ __device__ inline Layout make_layout(int nf, int n_energy_ctas) {
  Layout L;
  L.nf = nf; L.D = 4 + 8 * nf; L.DP = L.D + 1;
  L.npairs = nf * (nf - 1) / 2;
  L.base_pair = nf * XW_DIAG;
  L.base_misc = L.base_pair + L.npairs * 64;
  L.base_sc = L.base_misc + XW_MISC;
  L.nv = L.base_sc + L.DP * L.DP;
  L.cta_pair = nf;
  L.cta_misc = L.cta_pair + (L.npairs + 3) / 4;
  L.cta_sc = L.cta_misc + 1;
  L.cta_e = L.cta_sc + (L.DP * L.DP + XT - 1) / XT;
  L.n_cta = L.cta_e + n_energy_ctas;
  return L;
}

__device__ __forceinline__ unsigned char *slot_of(const StitchXchgArgs &a, int owner, int parity, int from) {
  return a.peer[owner] + ((size_t)parity * a.world + from) * a.slot_bytes;
}

// one value: push to every peer, then the rank-ordered sum over all ranks (own value in its place)
__device__ __forceinline__ double xchg_sum(const StitchXchgArgs &a, unsigned ep, int parity, int widx, double v, long long t0, bool &ok) {
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  for (int p = 0; p < a.world; p++) {
    if (p == a.rank) continue;
    unsigned char *dst = slot_of(a, p, parity, a.rank) + (size_t)widx * 16;
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %2};" ::"l"(dst), "r"(lo), "r"(ep), "r"(hi) : "memory");
  }
  unsigned rl[8], rh[8], fl[8], fh[8];
#pragma unroll
  for (int r = 0; r < 8; r++)
    if (r < a.world && r != a.rank) {   // all peers' words are requested before any is checked: one round trip, not one per rank
      const unsigned char *src = slot_of(a, a.rank, parity, r) + (size_t)widx * 16;
      asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rl[r]), "=r"(fl[r]), "=r"(rh[r]), "=r"(fh[r]) : "l"(src) : "memory");
    }
  double s = 0.0;
#pragma unroll
  for (int r = 0; r < 8; r++)
    if (r < a.world) {
      if (r == a.rank) { s += v; continue; }
      int spins = 0;
      while (fl[r] != ep || fh[r] != ep) {   // each 8-byte half validates itself, so tearing between the halves is harmless
        const unsigned char *src = slot_of(a, a.rank, parity, r) + (size_t)widx * 16;
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rl[r]), "=r"(fl[r]), "=r"(rh[r]), "=r"(fh[r]) : "l"(src) : "memory");
        if (a.backoff_ns) __nanosleep(a.backoff_ns);
        if ((++spins & 255) == 0 && clock64() - t0 > X_TIMEOUT) { ok = false; break; }
      }
      s += __hiloint2double((int)rh[r], (int)rl[r]);
    }
  return s;
}

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define X_TS(slot) do { if (a.dbg && threadIdx.x == 0) a.dbg[slot] = gtime(); } while (0)

__device__ __forceinline__ void ll8_store(unsigned char *p, unsigned v, unsigned ep) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v), "r"(ep) : "memory");
}
__device__ __forceinline__ bool ll8_load(const unsigned char *p, unsigned ep, unsigned &v, long long t0) {
  unsigned f;
  int spins = 0;
  for (;;) {
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(f) : "l"(p) : "memory");
    if (f == ep) return true;
    __nanosleep(40);
    if ((++spins & 255) == 0 && clock64() - t0 > X_TIMEOUT) return false;
  }
}

// newest-frame energies: every rank's segment is copied into every other rank's list (8-byte {float, exchange number}
// words behind `ebase` of the slot, word 0 = the segment length); CTA `cta` of `ncta`
__device__ __forceinline__ void exchange_energies(const StitchXchgArgs &a, unsigned ep, int parity, size_t ebase, int cta, int ncta, long long t0, bool &ok) {
  const int first = cta * XT + (int)threadIdx.x, stride = ncta * XT;
  const int my_n = min(a.newE_cnt[a.rank], a.newE_cap);
  const float *src = a.newE_all + (size_t)a.rank * a.newE_cap;
  for (int p = 0; p < a.world; p++) {
    if (p == a.rank) continue;
    unsigned char *dst = slot_of(a, p, parity, a.rank) + ebase;
    if (first == 0) ll8_store(dst, (unsigned)my_n, ep);
    for (int k = first; k < my_n; k += stride) ll8_store(dst + 8 * (size_t)(1 + k), __float_as_uint(src[k]), ep);
  }
  // receive: the words of all peers are requested before any is checked (one round trip, not one per rank)
  unsigned cn[8], cf[8];
#pragma unroll
  for (int r = 0; r < 8; r++) {
    cn[r] = 0; cf[r] = ep;
    if (r < a.world && r != a.rank) {
      const unsigned char *s8 = slot_of(a, a.rank, parity, r) + ebase;
      asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(cn[r]), "=r"(cf[r]) : "l"(s8) : "memory");
    }
  }
  int maxn = 0;
#pragma unroll
  for (int r = 0; r < 8; r++)
    if (r < a.world && r != a.rank) {
      if (cf[r] != ep) ok = ok && ll8_load(slot_of(a, a.rank, parity, r) + ebase, ep, cn[r], t0);
      if (!ok) cn[r] = 0;
      if ((int)cn[r] > a.newE_cap) cn[r] = (unsigned)a.newE_cap;
      if (first == 0 && ok) a.newE_cnt[r] = (int)cn[r];
      maxn = max(maxn, (int)cn[r]);
    }
  for (int k = first; k < maxn && ok; k += stride) {
    unsigned v[8], f[8];
#pragma unroll
    for (int r = 0; r < 8; r++)
      if (r < a.world && r != a.rank && k < (int)cn[r]) {
        const unsigned char *s8 = slot_of(a, a.rank, parity, r) + ebase + 8 * (size_t)(1 + k);
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v[r]), "=r"(f[r]) : "l"(s8) : "memory");
      }
#pragma unroll
    for (int r = 0; r < 8; r++)
      if (r < a.world && r != a.rank && k < (int)cn[r]) {
        if (f[r] != ep) ok = ok && ll8_load(slot_of(a, a.rank, parity, r) + ebase + 8 * (size_t)(1 + k), ep, v[r], t0);
        if (ok) a.newE_all[(size_t)r * a.newE_cap + k] = __uint_as_float(v[r]);
      }
  }
}

// the CTA that finishes last advances the exchange number (every CTA read it before it could finish)
__device__ __forceinline__ void finish_exchange(const StitchXchgArgs &a, unsigned ep) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(a.epoch + 1, 1) == (int)gridDim.x - 1) {
      a.epoch[1] = 0;
      a.epoch[0] = ep + 1u == 0u ? 1 : (int)(ep + 1u);
    }
  }
}

// Shared-memory staging of the (host,target) blocks one CTA needs ("terms"): every term's 92-double table row (A pass and L
// pass) and its 8x8 adjoint arrive by TMA bulk copies (cp.async.bulk + mbarrier) issued by one thread each, all in flight at
// once; the code that consumes them is a few compact loops (a fully unrolled gather is instruction-fetch bound here: every
// instruction of this kernel runs once).
constexpr int X_MAXT = 24;              // terms per CTA: 2 (nf - 1) for a diagonal tile, nf <= 13
constexpr int XS_A = 0, XS_L = XS_A + X_MAXT * SOSBA_TOPB, XS_M = XS_L + X_MAXT * SOSBA_TOPB, XS_TOTAL = XS_M + X_MAXT * 64;
constexpr int XS_X = XS_L;              // the L rows are consumed once they are added to the A rows: X = M P lives there

__device__ __forceinline__ unsigned xs_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void xs_bulk(void *dst, const void *src, unsigned bytes, unsigned long long *mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(xs_u32(dst)), "l"(src), "r"(bytes),
               "r"(xs_u32(mbar))
               : "memory");
}
__device__ __forceinline__ void xs_wait(unsigned long long *mbar) {
  unsigned done = 0;
  while (!done)
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(xs_u32(mbar)), "r"(0u) : "memory");
}
// index of entry (r, c), r <= c, of the symmetric 13x13 block inside a table row (inverse of entry_rc)
__device__ __forceinline__ int entry_index(int r, int c) {
  if (c < 10) return r * 10 - r * (r - 1) / 2 + (c - r);
  if (r < 10) return 55 + 3 * r + (c - 10);
  return r == 10 ? 85 + (c - 10) : r == 11 ? 88 + (c - 11) : 90;
}

__global__ void __launch_bounds__(XT, 1) k_stitch_xchg(StitchXchgArgs a, Layout L) {
  __shared__ __align__(16) double sm[XS_TOTAL];
  __shared__ __align__(8) unsigned long long mbar;
  PDL_ENTER_T(a.trace);
  if (a.gate && *a.gate) return;
  const int tid = threadIdx.x, g = tid >> 6, l = tid & 63, i = l >> 3, j = l & 7;
  const int nf = a.nf, D = a.D, bid = blockIdx.x, n2 = nf * nf;
  const bool push = a.push != 0;
  const unsigned ep = push ? (unsigned)a.epoch[0] : 0u;
  const int parity = (int)(ep & 1u);
  const long long t0 = clock64();
  bool ok = true;

  if (bid < L.cta_misc) {
    // ---- tiles of H: a diagonal tile (CTA af < nf: H[af,af], H[af,calib], b[af]) or four off-diagonal tiles ------------
    const bool diag = bid < L.cta_pair;
    const int af = bid;
    // diagonal: terms k < nf-1 are the host terms Ah P Ah^T of blocks (af, t), t != af; the others the target terms
    //           At P At^T of blocks (h, af), h != af; dealt round-robin to the four 64-thread groups
    // off-diagonal: group g owns tile q = (fa, fb), fa < fb: H[fa,fb] = Ah P At^T of block (fa,fb) + (Ah P At^T of (fb,fa))^T,
    //           terms 2g, 2g+1 with left / right factors (adHost, adTarget) of (fa,fb), then (adTarget, adHost) of (fb,fa)
    const int nterms = diag ? 2 * (nf - 1) : 8;
    const int q = 4 * (bid - L.cta_pair) + g;
    const bool act = diag || q < L.npairs;
    int fa = 0, fb = 1;
    if (!diag && q < L.npairs) { int rem = q; while (rem >= nf - 1 - fa) { rem -= nf - 1 - fa; fa++; } fb = fa + 1 + rem; }
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xs_u32(&mbar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const unsigned total = (unsigned)nterms * (2u * SOSBA_TOPB * 8u + (diag ? 512u : 1024u));
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xs_u32(&mbar)), "r"(total) : "memory");
    }
    __syncthreads();
    if (tid < nterms) {   // one thread per term: its A row, its L row, its adjoint(s)
      const int k = tid;
      int blk, left_is_host;
      if (diag) {
        left_is_host = k < nf - 1;
        if (left_is_host) { const int t = k < af ? k : k + 1; blk = af + nf * t; }
        else { const int kk = k - (nf - 1), hh = kk < af ? kk : kk + 1; blk = hh + nf * af; }
      } else {   // (an idle group stages the blocks of tile (0,1) and drops the result)
        int pa = 0, pb = 1, rem = 4 * (bid - L.cta_pair) + (k >> 1);
        if (rem < L.npairs) { while (rem >= nf - 1 - pa) { rem -= nf - 1 - pa; pa++; } pb = pa + 1 + rem; }
        left_is_host = (k & 1) == 0;
        blk = left_is_host ? pa + nf * pb : pb + nf * pa;
      }
      xs_bulk(sm + XS_A + k * SOSBA_TOPB, a.accTop + (size_t)blk * SOSBA_TOPB, SOSBA_TOPB * 8, &mbar);
      xs_bulk(sm + XS_L + k * SOSBA_TOPB, a.accTop + ((size_t)n2 + blk) * SOSBA_TOPB, SOSBA_TOPB * 8, &mbar);
      xs_bulk(sm + XS_M + k * 64, (left_is_host ? a.adHost : a.adTarget) + 64 * (size_t)blk, 512, &mbar);
      if (!diag) xs_bulk(sm + XS_M + (8 + k) * 64, (left_is_host ? a.adTarget : a.adHost) + 64 * (size_t)blk, 512, &mbar);   // right factors behind the 8 left ones
    }
    // where this thread's operands sit inside a table row: P[q][j] = block(4+q, 4+j), C[q][j] = block(j, 4+q), r[q] = block(4+q, 12)
    int pe[8], ce[8];
#pragma unroll
    for (int qq = 0; qq < 8; qq++) {
      pe[qq] = entry_index(min(4 + qq, 4 + j), max(4 + qq, 4 + j));
      ce[qq] = j < 4 ? entry_index(j, 4 + qq) : entry_index(4 + qq, 12);
    }
    xs_wait(&mbar);
    for (int idx = tid; idx < nterms * SOSBA_TOPB; idx += XT) sm[XS_A + idx] += sm[XS_L + idx];
    __syncthreads();
    const int kstep = diag ? 4 : 1, kbeg = diag ? g : 2 * g, kend = diag ? nterms : 2 * g + 2;
    for (int k = kbeg; k < kend; k += kstep) {
      const double *M = sm + XS_M + k * 64, *B = sm + XS_A + k * SOSBA_TOPB;
      double x = 0.0;
#pragma unroll
      for (int qq = 0; qq < 8; qq++) x += M[i * 8 + qq] * B[pe[qq]];
      sm[XS_X + k * 64 + l] = x;
    }
    __syncthreads();
    double o = 0.0, fc = 0.0;
    const int ii = diag ? max(i, j) : i, jj = diag ? min(i, j) : j;   // diagonal tile: both triangles from the same expression
    for (int k = kbeg; k < kend; k += kstep) {
      const double *M = sm + XS_M + k * 64, *R = sm + XS_M + (diag ? k : 8 + k) * 64, *X = sm + XS_X + k * 64, *B = sm + XS_A + k * SOSBA_TOPB;
#pragma unroll
      for (int qq = 0; qq < 8; qq++) o += X[ii * 8 + qq] * R[jj * 8 + qq];
      if (diag && j <= 4) {
#pragma unroll
        for (int qq = 0; qq < 8; qq++) fc += M[i * 8 + qq] * B[ce[qq]];
      }
    }
    if (diag) {
      __syncthreads();
      double *part = sm + XS_A;   // [4][XW_DIAG], the staged rows are consumed
      part[g * XW_DIAG + l] = o;
      if (j < 4) part[g * XW_DIAG + 64 + i * 4 + j] = fc;
      else if (j == 4) part[g * XW_DIAG + 96 + i] = fc;
      __syncthreads();
      if (tid < XW_DIAG) {
        double v = ((part[tid] + part[XW_DIAG + tid]) + part[2 * XW_DIAG + tid]) + part[3 * XW_DIAG + tid];
        if (af == 0) X_TS(3);
        if (push) v = xchg_sum(a, ep, parity, af * XW_DIAG + tid, v, t0, ok);
        if (af == 0) X_TS(4);
        if (ok) {
          const int r0 = 4 + 8 * af;
          if (tid < 64) a.H[(size_t)(r0 + i) * D + r0 + j] = v;
          else if (tid < 96) { const int e = tid - 64, fi = e >> 2, cj = e & 3; a.H[(size_t)(r0 + fi) * D + cj] = v; a.H[(size_t)cj * D + r0 + fi] = v; }
          else a.b[r0 + tid - 96] = v;
        }
      }
    } else if (act) {
      double v = o;
      if (push) v = xchg_sum(a, ep, parity, L.base_pair + q * 64 + l, v, t0, ok);
      if (ok) {
        a.H[(size_t)(4 + 8 * fa + i) * D + 4 + 8 * fb + j] = v;
        a.H[(size_t)(4 + 8 * fb + j) * D + 4 + 8 * fa + i] = v;
      }
    }
  } else if (bid < L.cta_sc) {
    // ---- calibration block, b[calib], back-substitution sums, residual counters ------------------------------------
    const int nblk = 2 * nf * nf;
    X_TS(0);
    if (tid < 240) {   // 20 entries x 12 slices of the block list
      const int ent = tid % 20, slice = tid / 20;
      int e;
      if (ent < 16) { const int r = min(ent >> 2, ent & 3), c = max(ent >> 2, ent & 3); e = r * 10 - r * (r - 1) / 2 + (c - r); }
      else e = 55 + 3 * (ent - 16) + 2;
      constexpr int MI = (2 * (X_MAXT / 2 + 1) * (X_MAXT / 2 + 1) + 11) / 12;   // all loads in flight before the first add
      double v[MI];
#pragma unroll
      for (int it = 0; it < MI; it++) { const int blk = slice + 12 * it; v[it] = blk < nblk ? a.accTop[(size_t)blk * SOSBA_TOPB + e] : 0.0; }
      double s = 0.0;
#pragma unroll
      for (int it = 0; it < MI; it++) s += v[it];
      sm[slice * 20 + ent] = s;
    }
    __syncthreads();
    if (tid < 30) {
      double v;
      if (tid < 20) { v = 0.0; for (int s = 0; s < 12; s++) v += sm[s * 20 + tid]; }
      else if (tid < 28) v = a.rstats[tid - 20];
      else v = (double)a.cnt[tid - 28];
      X_TS(1);
      if (push) v = xchg_sum(a, ep, parity, L.base_misc + tid, v, t0, ok);
      X_TS(2);
      if (ok) {
        if (tid < 16) a.H[(size_t)(tid >> 2) * D + (tid & 3)] = v;
        else if (tid < 20) a.b[tid - 16] = v;
        else if (tid < 28) a.rstats[tid - 20] = v;
        else a.cnt[tid - 28] = (int)v;
      }
    }
  } else if (bid < L.cta_e) {
    // ---- Schur Gram matrix (already in stitched space): summed in place, upper triangle ---------------------------
    if (push) {
      const int e = (bid - L.cta_sc) * XT + tid;
      if (e < L.DP * L.DP && e / L.DP <= e % L.DP) {
        const double v = xchg_sum(a, ep, parity, L.base_sc + e, a.accSC[e], t0, ok);
        if (ok) a.accSC[e] = v;
      }
    }
  } else if (push && a.with_newE) {
    if (bid == L.cta_e) X_TS(5);
    exchange_energies(a, ep, parity, (size_t)L.nv * 16, bid - L.cta_e, gridDim.x - L.cta_e, t0, ok);
    if (bid == L.cta_e) X_TS(6);
  }
  if (!ok && a.err) atomicOr(a.err, 2);   // reported by the caller as SOSBA_E_NCCL (peer exchange timed out)
  if (push) finish_exchange(a, ep);
  TRACE_EXIT(a.trace);
}

// The sums of an API-level linearizeAll outside the loop over the ranks (FullSystemOptimize.cpp:125-182: energy, state
// histogram, removals) and the newest-frame energies for setNewFrameEnergyTH, through the same mailboxes.
// CTA 0: sums; CTAs 1..: energies.  with_stats = 0: energies only (the pending selection of a fused linearisation).
__global__ void __launch_bounds__(XT) k_lin_xchg(StitchXchgArgs a, double *stats, int *counts, int with_stats, double *extra, int n_extra) {
  PDL_ENTER();
  if (a.gate && *a.gate) return;
  const unsigned ep = (unsigned)a.epoch[0];
  const int parity = (int)(ep & 1u), tid = threadIdx.x;
  const long long t0 = clock64();
  bool ok = true;
  if (blockIdx.x == 0) {
    if (with_stats && tid < 5) {
      double v = tid == 0 ? stats[0] : (double)counts[tid - 1];
      v = xchg_sum(a, ep, parity, tid, v, t0, ok);
      if (ok) { if (tid == 0) stats[0] = v; else counts[tid - 1] = (int)v; }
    }
    if (tid >= 5 && tid < 5 + n_extra) {   // further sums of the caller (the step sums of sosba_ba_step)
      const double v = xchg_sum(a, ep, parity, tid, extra[tid - 5], t0, ok);
      if (ok) extra[tid - 5] = v;
    }
  } else if (a.with_newE) {
    exchange_energies(a, ep, parity, 16 * 16, blockIdx.x - 1, gridDim.x - 1, t0, ok);
  }
  if (!ok && a.err) atomicOr(a.err, 2);
  finish_exchange(a, ep);
}

}  // namespace

size_t stitch_xchg_slot_bytes(int nf_max, int newE_cap) {
  const Layout L = make_layout(nf_max, 0);
  return (((size_t)L.nv * 16 + 8 * (size_t)(1 + newE_cap)) + 255) & ~(size_t)255;
}

int launch_stitch_xchg(sosba *h, const StitchXchgArgs &a0, int local_points) {
  StitchXchgArgs a = a0;
  if (2 * (a.nf - 1) > X_MAXT) { sosba_set_error("the fused stitch supports nf <= %d, got %d", X_MAXT / 2 + 1, a.nf); return SOSBA_E_ARG; }
  const int ne = (a.push && a.with_newE) ? (local_points + XT - 1) / XT + 1 : 0;
  const Layout L = make_layout(a.nf, ne);
  launch_pdl(k_stitch_xchg, a.push ? L.n_cta : L.cta_sc, XT, 0, h->stream, a, L);   // without peers the Schur matrix stays as it is
  h->launches++;
  return SOSBA_OK;
}

// point shards, outside the loop: see k_lin_xchg.  Only with the peer mailboxes (a.push); the caller falls back to NCCL.
void launch_lin_xchg(sosba *h, const StitchXchgArgs &a, double *stats, int *counts, int with_stats, int local_points, double *extra, int n_extra) {
  const int ne = a.with_newE ? (local_points + XT - 1) / XT + 1 : 0;
  launch_pdl(k_lin_xchg, 1 + ne, XT, 0, h->stream, a, stats, counts, with_stats, extra, n_extra > 10 ? 10 : n_extra);
  h->launches++;
}
