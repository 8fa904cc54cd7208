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
        if ((++spins & 255) == 0 && clock64() - t0 > X_TIMEOUT) { ok = false; break; }
      }
      s += __hiloint2double((int)rh[r], (int)rl[r]);
    }
  return s;
}

__device__ __forceinline__ void ll8_store(unsigned char *p, unsigned v, unsigned ep) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v), "r"(ep) : "memory");
}
__device__ __forceinline__ bool ll8_load(const unsigned char *p, unsigned ep, unsigned &v, long long t0) {
  unsigned f;
  int spins = 0;
  for (;;) {
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(f) : "l"(p) : "memory");
    if (f == ep) return true;
    if ((++spins & 255) == 0 && clock64() - t0 > X_TIMEOUT) return false;
  }
}

// the 13x13 block of one (host,target) pair, A and L passes summed, into shared memory (symmetric); 64 threads
__device__ __forceinline__ void load_block(const StitchXchgArgs &a, int blk, int l, double (*accH)[13]) {
  const int n2 = a.nf * a.nf;
  for (int e = l; e < 91; e += 64) {
    const double v = a.accTop[(size_t)blk * SOSBA_TOPB + e] + a.accTop[((size_t)n2 + blk) * SOSBA_TOPB + e];
    int r, c;
    entry_rc(e, r, c);
    accH[r][c] = v; accH[c][r] = v;
  }
}

__global__ void __launch_bounds__(XT) k_stitch_xchg(StitchXchgArgs a, Layout L) {
  __shared__ double s_acc[4][13][13];
  __shared__ double s_M[4][2][64];
  __shared__ double s_X[4][64];
  __shared__ double s_part[12][64 + 32 + 8];
  PDL_ENTER();
  if (a.gate && *a.gate) return;
  const int tid = threadIdx.x, g = tid >> 6, l = tid & 63, i = l >> 3, j = l & 7;
  const int nf = a.nf, D = a.D, bid = blockIdx.x;
  const bool push = a.push != 0;
  const unsigned ep = push ? (unsigned)a.epoch[0] : 0u;
  const int parity = (int)(ep & 1u);
  const long long t0 = clock64();
  bool ok = true;

  if (bid < L.cta_pair) {
    // ---- diagonal tile of frame af: H[af,af], H[af,calib], b[af] ------------------------------------------------
    // = sum over targets t != af of the host terms Ah P Ah^T (block af + nf t) + sum over hosts h != af of the target terms
    //   At P At^T (block h + nf af), four terms at a time (one per 64-thread group), partial sums added in group order
    const int af = bid, nterms = 2 * (nf - 1);
    double o = 0.0, fc = 0.0, bb = 0.0;
    for (int k0 = 0; k0 < nterms; k0 += 4) {
      const int k = k0 + g;
      const bool act = k < nterms;
      if (act) {
        int blk;
        const double *Msrc;
        if (k < nf - 1) { const int t = k < af ? k : k + 1; blk = af + nf * t; Msrc = a.adHost + 64 * (size_t)blk; }
        else { const int kk = k - (nf - 1), hh = kk < af ? kk : kk + 1; blk = hh + nf * af; Msrc = a.adTarget + 64 * (size_t)blk; }
        load_block(a, blk, l, s_acc[g]);
        s_M[g][0][l] = Msrc[l];
      }
      __syncthreads();
      if (act) {
        double x = 0.0;
#pragma unroll
        for (int q = 0; q < 8; q++) x += s_M[g][0][i * 8 + q] * s_acc[g][4 + q][4 + j];
        s_X[g][l] = x;
      }
      __syncthreads();
      if (act) {
        const int ii = max(i, j), jj = min(i, j);   // both triangles from the same expression: exactly symmetric
#pragma unroll
        for (int q = 0; q < 8; q++) o += s_X[g][ii * 8 + q] * s_M[g][0][jj * 8 + q];
        if (j < 4) {
#pragma unroll
          for (int q = 0; q < 8; q++) fc += s_M[g][0][i * 8 + q] * s_acc[g][4 + q][j];
        } else if (j == 4) {
#pragma unroll
          for (int q = 0; q < 8; q++) bb += s_M[g][0][i * 8 + q] * s_acc[g][4 + q][12];
        }
      }
      __syncthreads();
    }
    s_part[g][l] = o;
    if (j < 4) s_part[g][64 + i * 4 + j] = fc;
    else if (j == 4) s_part[g][96 + i] = bb;
    __syncthreads();
    if (tid < XW_DIAG) {
      double v = ((s_part[0][tid] + s_part[1][tid]) + s_part[2][tid]) + s_part[3][tid];
      if (push) v = xchg_sum(a, ep, parity, af * XW_DIAG + tid, v, t0, ok);
      if (ok) {
        const int r0 = 4 + 8 * af;
        if (tid < 64) a.H[(size_t)(r0 + i) * D + r0 + j] = v;
        else if (tid < 96) { const int e = tid - 64, fi = e >> 2, cj = e & 3; a.H[(size_t)(r0 + fi) * D + cj] = v; a.H[(size_t)cj * D + r0 + fi] = v; }
        else a.b[r0 + tid - 96] = v;
      }
    }
  } else if (bid < L.cta_misc) {
    // ---- four off-diagonal tiles, one per 64-thread group: H[fa,fb] = Ah P At^T of block (fa,fb) + (Ah P At^T of block (fb,fa))^T
    const int q = 4 * (bid - L.cta_pair) + g;
    const bool act = q < L.npairs;
    int fa = 0, fb = 1;
    if (act) { int rem = q; while (rem >= nf - 1 - fa) { rem -= nf - 1 - fa; fa++; } fb = fa + 1 + rem; }
    double o = 0.0;
    for (int term = 0; term < 2; term++) {
      const int blk = term == 0 ? fa + nf * fb : fb + nf * fa;
      if (act) {
        load_block(a, blk, l, s_acc[g]);
        // left factor, right factor: (adHost, adTarget) of block (fa,fb), then (adTarget, adHost) of block (fb,fa)
        s_M[g][0][l] = (term == 0 ? a.adHost : a.adTarget)[64 * (size_t)blk + l];
        s_M[g][1][l] = (term == 0 ? a.adTarget : a.adHost)[64 * (size_t)blk + l];
      }
      __syncthreads();
      if (act) {
        double x = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) x += s_M[g][0][i * 8 + k] * s_acc[g][4 + k][4 + j];
        s_X[g][l] = x;
      }
      __syncthreads();
      if (act) {
#pragma unroll
        for (int k = 0; k < 8; k++) o += s_X[g][i * 8 + k] * s_M[g][1][j * 8 + k];
      }
      __syncthreads();
    }
    if (act) {
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
    if (tid < 240) {   // 20 entries x 12 slices of the block list
      const int ent = tid % 20, slice = tid / 20;
      int e;
      if (ent < 16) { const int r = min(ent >> 2, ent & 3), c = max(ent >> 2, ent & 3); e = r * 10 - r * (r - 1) / 2 + (c - r); }
      else e = 55 + 3 * (ent - 16) + 2;
      double s = 0.0;
      for (int blk = slice; blk < nblk; blk += 12) s += a.accTop[(size_t)blk * SOSBA_TOPB + e];
      s_part[slice][ent] = s;
    }
    __syncthreads();
    if (tid < 30) {
      double v;
      if (tid < 20) { v = 0.0; for (int s = 0; s < 12; s++) v += s_part[s][tid]; }
      else if (tid < 28) v = a.rstats[tid - 20];
      else v = (double)a.cnt[tid - 28];
      if (push) v = xchg_sum(a, ep, parity, L.base_misc + tid, v, t0, ok);
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
    // ---- newest-frame energies: every rank's segment is copied into every other rank's list ------------------------
    const int ne = gridDim.x - L.cta_e, first = (bid - L.cta_e) * XT + tid, stride = ne * XT;
    const size_t ebase = (size_t)L.nv * 16;
    const int my_n = a.newE_cnt[a.rank];
    const float *src = a.newE_all + (size_t)a.rank * a.newE_cap;
    for (int p = 0; p < a.world; p++) {
      if (p == a.rank) continue;
      unsigned char *dst = slot_of(a, p, parity, a.rank) + ebase;
      if (first == 0) ll8_store(dst, (unsigned)my_n, ep);
      for (int k = first; k < my_n; k += stride) ll8_store(dst + 8 * (size_t)(1 + k), __float_as_uint(src[k]), ep);
    }
    for (int r = 0; r < a.world && ok; r++) {
      if (r == a.rank) continue;
      const unsigned char *s8 = slot_of(a, a.rank, parity, r) + ebase;
      unsigned n = 0;
      ok = ll8_load(s8, ep, n, t0);
      if (!ok) break;
      if ((int)n > a.newE_cap) n = (unsigned)a.newE_cap;
      if (first == 0) a.newE_cnt[r] = (int)n;
      float *dst = a.newE_all + (size_t)r * a.newE_cap;
      for (int k = first; k < (int)n && ok; k += stride) {
        unsigned v;
        ok = ll8_load(s8 + 8 * (size_t)(1 + k), ep, v, t0);
        if (ok) dst[k] = __uint_as_float(v);
      }
    }
  }
  if (!ok && a.err) atomicOr(a.err, 2);   // reported by the caller as SOSBA_E_NCCL (peer exchange timed out)
  if (!push) return;
  // ---- the CTA that finishes last advances the exchange number (every CTA read it before it could finish) ------------
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(a.epoch + 1, 1) == (int)gridDim.x - 1) {
      a.epoch[1] = 0;
      a.epoch[0] = ep + 1u == 0u ? 1 : (int)(ep + 1u);
    }
  }
}

}  // namespace

size_t stitch_xchg_slot_bytes(int nf_max, int newE_cap) {
  const Layout L = make_layout(nf_max, 0);
  return (((size_t)L.nv * 16 + 8 * (size_t)(1 + newE_cap)) + 255) & ~(size_t)255;
}

void launch_stitch_xchg(sosba *h, const StitchXchgArgs &a0, int local_points) {
  StitchXchgArgs a = a0;
  const int ne = (a.push && a.with_newE) ? (local_points + XT - 1) / XT + 1 : 0;
  const Layout L = make_layout(a.nf, ne);
  launch_pdl(k_stitch_xchg, a.push ? L.n_cta : L.cta_sc, XT, 0, h->stream, a, L);   // without peers the Schur matrix stays as it is
  h->launches++;
}
