// k_solve.cu — EnergyFunctional::solveSystemF without IMU (EnergyFunctional.cpp:1029-1184) as ONE single-CTA
// fp64 kernel.  Input is the stitch of the top blocks (A and L passes summed, symmetrised: k_stitch_xchg), the
// stitched-space Schur Gram matrix, the priors and (optionally) the marginalisation prior HM, bM:
//   HFinal = Htop + priors (+HM), b = btop + prior*delta_prior (+bM + HM*delta); diag *= (1+1e-5);
//   HFinal -= H_sc/(1+1e-5); b -= b_sc; Jacobi scaling 1/sqrt(diag+10); LDL^T with the transposition order of
//   Eigen::LDLT (the solver called at :1147-1148), factorised with 4x4 pivot blocks (panel_rows<true>; scalar pivots with
//   SOSBA_SOLVE_PIVOTS=scalar); back-substitution in blocks of 4 rows; x = S * y; then
//   xAd[h*nf+t] = x_h^T adHostF[h+nf*t] + x_t^T adTargetF[h+nf*t] for resubstituteF_MT (:496-524).
// D = 4 + 8*nf <= 132; the (D+1)^2 augmented matrix lives in shared memory.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "host_math.h"
#include "kernels.h"
#include "energy_th.cuh"
#include "resub.cuh"

namespace {

// 256 threads as a 16x16 grid.  Thread (ty,tx) keeps the entries {(ty+16(kb+a), tx+16(kb+b))} of the trailing part of
// the augmented, permuted, preconditioned lower triangle in registers (T = ceil((D+1)/16) tile rows/columns).
// Pivot order: Eigen::LDLT is left-looking, so when it searches the largest |diagonal| of the trailing block none
// of those entries has been updated yet — the transposition sequence depends only on the input diagonal and equals
// a descending sort of |diag| (ties broken by index).  The factorisation itself then needs no pivot search and runs
// right-looking in panels of 4 columns with ONE barrier per panel: the owners publish the panel, every thread
// factorises the 4x4 diagonal block redundantly in registers, forms the L21 rows of its own tile rows / columns and
// applies the rank-4 update.  After every 16 columns the register tiles shift by one, so the loop body indexes
// registers statically yet stays small enough to live in the instruction cache (a fully unrolled factorisation is
// instruction-fetch bound: every instruction would run exactly once).  The right-hand side rides along as row D,
// ending as D^-1 L^-1 P b.
// Inputs arrive by TMA bulk copies (cp.async.bulk + mbarrier) into shared memory, so the permuted gathers of the
// assembly never touch L2.
// This file is compiled with -fmad=false (the fused frame step mirrors host_ba.cpp operation by operation), so the
// factorisation spells its fused multiply-adds out with fma().
// 1/v: MUFU.RCP64H seed (about 20 bits) + two Newton steps, no slow path — the pivots of the Jacobi-scaled system are O(1).
// v == 0 gives inf -> nan -> 0: the column is left alone (Eigen: pivot_is_valid).
__device__ __forceinline__ double safe_rcp(double v) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
  double e = fma(-v, r, 1.0);
  r = fma(r, e, r);
  e = fma(-v, r, 1.0);
  r = fma(r, e, r);
  return isfinite(r) ? r : 0.0;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(mbar))
               : "memory");
}

constexpr int SOLVE_UPD = 256;      // update warps: 16x16 thread grid over the trailing tiles
constexpr int SOLVE_PAN = 128;      // panel warps: one thread per remaining row
constexpr int SOLVE_THREADS = SOLVE_UPD + SOLVE_PAN;
constexpr int BAR_PUB = 1, BAR_LY = 2, BAR_PL = 3;   // named barriers (0 is __syncthreads)
constexpr int BAR_PN = 3;   // pipelined factorisation: the panel warps among themselves (BAR_PL is not used there)

__device__ __forceinline__ void nb_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(SOLVE_THREADS) : "memory"); }
// bar.arrive orders the arriving thread's earlier shared-memory writes before the barrier completes (the PTX ISA's
// producer/consumer pattern: st.shared; bar.arrive  ||  bar.sync; ld.shared), so no fence is needed.
__device__ __forceinline__ void nb_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(SOLVE_THREADS) : "memory"); }

// Update warps, panel p (k0 = 16 kb + 4 q) with NA live tile rows.  L (l = y / d) of the panel sits in the persistent
// factor storage Lp (row-major [row][4]), Y in the double-buffered Y4.  The next panel lives in tile column JP: the
// lanes that own it (8 per warp) update and publish it FIRST, with only their own shared-memory loads in the way, so
// the panel warps can start on panel p+1 while the bulk of the rank-4 update runs (look-ahead).
// LAST (q == 3): tile column 0 is finished, the next panel sits in column 1.
template <int T, int NA, bool LAST>
__device__ __forceinline__ void update_step(double (&reg)[T][T], const double *__restrict__ Lp, const double *__restrict__ Y4, double *__restrict__ pbn,
                                            const int kb, const int q, const int ty, const int tx, long long *dbg) {
  constexpr int JP = LAST ? 1 : 0;
  const int k0n = 16 * kb + 4 * q + 4;
  const int cn = tx - 4 * ((q + 1) & 3);   // my column inside the next panel, if 0 <= cn < 4
  const bool owner = cn >= 0 && cn < 4;
  const double2 *pl = (const double2 *)(Lp + (size_t)(ty + 16 * (kb + JP)) * 4);   // + 32 double2 per tile row
  const double2 *py = (const double2 *)(Y4 + (size_t)(tx + 16 * (kb + JP)) * 4);
  double *pub = pbn + (size_t)(ty + 16 * (kb + JP)) * 4 + cn;
  nb_sync(BAR_LY);
  double li[NA][4];
#pragma unroll
  for (int ia = JP; ia < NA; ia++) {   // two rows per warp: one wavefront per load
    const double2 l01 = pl[32 * (ia - JP)], l23 = pl[32 * (ia - JP) + 1];
    li[ia][0] = l01.x; li[ia][1] = l01.y; li[ia][2] = l23.x; li[ia][3] = l23.y;
  }
  if (NA > JP) {
    const double2 y01 = py[0], y23 = py[1];
    if (dbg && threadIdx.x == 0) dbg[0] = clock64() + (long long)(y23.y == 123.0);
#pragma unroll
    for (int ia = JP; ia < NA; ia++) {
      const double v = fma(-li[ia][3], y23.y, fma(-li[ia][2], y23.x, fma(-li[ia][1], y01.y, fma(-li[ia][0], y01.x, reg[ia][JP]))));
      reg[ia][JP] = v;
      if (owner && ty + 16 * (kb + ia) >= k0n + cn) pub[64 * (ia - JP)] = v;   // rows past D are zero tiles landing in the padding
    }
  }
  if (dbg) dbg[threadIdx.x == 0 ? 1 : 4] = clock64() + (long long)(reg[JP][JP] == 123.0);   // [7]: the same moment seen by warp 7
  nb_arrive(BAR_PUB);
  // the bulk of the rank-4 update runs while the panel warps work on panel p+1; its shared-memory loads wait until the
  // panel warps have fetched the new panel (BAR_PL), otherwise those few loads queue behind 64 of ours
  nb_sync(BAR_PL);
#pragma unroll
  for (int jb = JP + 1; jb < NA; jb++) {
    const double2 y01 = py[32 * (jb - JP)], y23 = py[32 * (jb - JP) + 1];
#pragma unroll
    for (int ia = jb; ia < NA; ia++)
      reg[ia][jb] = fma(-li[ia][3], y23.y, fma(-li[ia][2], y23.x, fma(-li[ia][1], y01.y, fma(-li[ia][0], y01.x, reg[ia][jb]))));
  }
}

template <int T, int NA>
struct UpdateDispatch {
  static __device__ __forceinline__ void run(int nact, bool last, double (&reg)[T][T], const double *Lp, const double *Y4, double *pbn, int kb, int q,
                                             int ty, int tx, long long *dbg) {
    if (nact == NA) {
      if (last) update_step<T, NA, true>(reg, Lp, Y4, pbn, kb, q, ty, tx, dbg);
      else update_step<T, NA, false>(reg, Lp, Y4, pbn, kb, q, ty, tx, dbg);
    } else UpdateDispatch<T, NA - 1>::run(nact, last, reg, Lp, Y4, pbn, kb, q, ty, tx, dbg);
  }
};
template <int T>
struct UpdateDispatch<T, 0> {
  static __device__ __forceinline__ void run(int, bool, double (&)[T][T], const double *, const double *, double *, int, int, int, int, long long *) {}
};

// Panel warps, panel k0: the 4x4 diagonal block (redundantly per thread), then one row of L21 per thread: threads 0..7
// retire the rows of this and the previous pivot block from Y4 (they must read as zero columns from now on), thread
// 8 + r owns row k0 + 4 + r.  L goes to the persistent factor storage Lp ([row][4]; row D = the right-hand side,
// ending as D^-1 L^-1 P b), Y to Y4; thread 0 also stores L11 into the rows of the pivot block.
// BLOCK (the default): the 4x4 diagonal block is the pivot as a whole (block LDL^T).  Its inverse comes from the adjugate
// (2x2 minors -> cofactors -> one reciprocal of the determinant), and the row of the factor is W_i = a_i A11^-1, so the
// dependent chain of a panel is ~12 fp64 operations and ONE reciprocal instead of four reciprocals in sequence; the
// rank-4 update is S -= W A21^T with A21 = the published panel itself (no Y rows to store).  The right-hand side row ends
// as A11^-1 (L^-1 P b)_block, the back substitution sees an identity diagonal block.  A singular block (determinant 0)
// leaves its four columns alone, like Eigen's pivot_is_valid.
template <bool BLOCK>
__device__ __forceinline__ void panel_rows(const double *__restrict__ pb, double *__restrict__ Lp, double *__restrict__ Y4, const int D, const int k0,
                                           const int pt, long long *dbg) {
  const double2 *pd = (const double2 *)(pb + (size_t)k0 * 4);
  const int i = k0 - 4 + pt;
  const bool live = pt >= 8 && i <= D;
  const double2 *pp = (const double2 *)(pb + (size_t)(live ? i : k0) * 4);
  double2 *pl = (double2 *)(Lp + (size_t)i * 4), *py = (double2 *)(Y4 + (size_t)i * 4);
  nb_sync(BAR_PUB);
  if (dbg) dbg[6] = clock64();   // through the barrier
  const double a00 = pd[0].x;
  const double2 q1 = pd[2], q2a = pd[4], q2b = pd[5], q3a = pd[6], q3b = pd[7];
  const double2 pa = pp[0], pc = pp[1];
  const double a10 = q1.x, a11 = q1.y, a20 = q2a.x, a21 = q2a.y, a22 = q2b.x, a30 = q3a.x, a31 = q3a.y, a32 = q3b.x, a33 = q3b.y;
  if (k0 > 0) nb_arrive(BAR_PL);   // our loads are queued: the update warps may start theirs (pairs with the sync in update_step of the panel before)
  if (dbg) dbg[0] = clock64() + (long long)(a00 == 123.0) + (long long)(pc.y == 123.0);   // panel data landed
  if (BLOCK) {
    const double s0 = fma(a00, a11, -(a10 * a10)), s1 = fma(a00, a21, -(a10 * a20)), s2 = fma(a00, a31, -(a10 * a30));
    const double s3 = fma(a10, a21, -(a11 * a20)), s4 = fma(a10, a31, -(a11 * a30)), s5 = fma(a20, a31, -(a21 * a30));
    const double c5 = fma(a22, a33, -(a32 * a32)), c4 = fma(a21, a33, -(a31 * a32)), c3 = fma(a21, a32, -(a31 * a22));
    const double c2 = fma(a20, a33, -(a30 * a32)), c1 = fma(a20, a32, -(a30 * a22));
    const double det = fma(s5, s5, fma(-s4, c1, fma(s3, c2, fma(s2, c3, fma(-s1, c4, s0 * c5)))));   // c0 == s5
    const double rdet = safe_rcp(det);
    // adjugate (symmetric): i_rc
    const double i00 = fma(a31, c3, fma(-a21, c4, a11 * c5)), i01 = fma(-a30, c3, fma(a20, c4, -(a10 * c5)));
    const double i02 = fma(a33, s3, fma(-a32, s4, a31 * s5)), i03 = fma(-a32, s3, fma(a22, s4, -(a21 * s5)));
    const double i11 = fma(a30, c1, fma(-a20, c2, a00 * c5)), i12 = fma(-a33, s1, fma(a32, s2, -(a30 * s5)));
    const double i13 = fma(a32, s1, fma(-a22, s2, a20 * s5)), i22 = fma(a33, s0, fma(-a31, s2, a30 * s4));
    const double i23 = fma(-a32, s0, fma(a21, s2, -(a20 * s4))), i33 = fma(a22, s0, fma(-a21, s1, a20 * s3));
    // W_i = (a_i adj) / det
    const double w0 = fma(pc.y, i03, fma(pc.x, i02, fma(pa.y, i01, pa.x * i00))) * rdet;
    const double w1 = fma(pc.y, i13, fma(pc.x, i12, fma(pa.y, i11, pa.x * i01))) * rdet;
    const double w2 = fma(pc.y, i23, fma(pc.x, i22, fma(pa.y, i12, pa.x * i02))) * rdet;
    const double w3 = fma(pc.y, i33, fma(pc.x, i23, fma(pa.y, i13, pa.x * i03))) * rdet;
    if (dbg) dbg[1] = clock64() + (long long)(w3 == 123.0);   // chain done
    if (live) {
      pl[0] = make_double2(w0, w1);
      pl[1] = make_double2(w2, w3);
    }
    return;
  }
  const double r0 = safe_rcp(a00);
  const double l10 = a10 * r0, l20 = a20 * r0, l30 = a30 * r0, m0 = pa.x * r0;
  const double d1 = fma(-l10, a10, a11), r1 = safe_rcp(d1);
  const double y21 = fma(-l20, a10, a21), y31 = fma(-l30, a10, a31), y1 = fma(-m0, a10, pa.y);
  const double l21 = y21 * r1, l31 = y31 * r1, m1 = y1 * r1;
  const double d2 = fma(-l21, y21, fma(-l20, a20, a22)), r2 = safe_rcp(d2);
  const double y32 = fma(-l31, y21, fma(-l30, a20, a32)), y2 = fma(-m1, y21, fma(-m0, a20, pc.x));
  const double l32 = y32 * r2, m2 = y2 * r2;
  const double d3 = fma(-l32, y32, fma(-l31, y31, fma(-l30, a30, a33))), r3 = safe_rcp(d3);
  const double y3 = fma(-m2, y32, fma(-m1, y31, fma(-m0, a30, pc.y))), m3 = y3 * r3;
  if (dbg) dbg[1] = clock64() + (long long)(m3 == 123.0);   // chain done
  if (live) {
    pl[0] = make_double2(m0, m1);
    pl[1] = make_double2(m2, m3);
    if (i < D) { py[0] = make_double2(pa.x, y1); py[1] = make_double2(y2, y3); }   // the rhs row is nobody's column
  } else if (pt < 8 && i >= 0) {
    py[0] = make_double2(0.0, 0.0);
    py[1] = make_double2(0.0, 0.0);
  }
  if (pt == 0) {   // L11 (unit lower) into the pivot rows of this panel's factor block
    Lp[(size_t)(k0 + 1) * 4] = l10;
    Lp[(size_t)(k0 + 2) * 4] = l20; Lp[(size_t)(k0 + 2) * 4 + 1] = l21;
    Lp[(size_t)(k0 + 3) * 4] = l30; Lp[(size_t)(k0 + 3) * 4 + 1] = l31; Lp[(size_t)(k0 + 3) * 4 + 2] = l32;
  }
}

// ---- pipelined factorisation (experimental, SOSBA_SOLVE_PIPE=1; parity-green, not faster: see launch_solve) ------------------
// The panel warps keep a sliding window of their row in registers: the current pivot group of 4 columns, the next one, and the
// one after that as it arrives.  A panel step is then entirely theirs: exchange the current columns of all rows through shared
// memory (a barrier among the 4 panel warps only), invert the 4x4 pivot block (adjugate, one reciprocal), form W_i, apply it to
// their own next two groups, go on.  The update warps follow one step behind: with W_p they update the register tiles of the
// trailing matrix and hand over column group p + 3 — which the panel warps need only one step later, so the hand-off of the old
// scheme (publish -> barrier -> factor -> barrier -> update, 1.7 k cycles per panel around a 0.4 k chain) is off the critical path.
//   pan4[p & 3][row][4]   current columns of every row at step p (written by the panel warps; the Y operand of the update warps)
//   hand[p & 1][row][4]   column group p + 3, updated through panel p (written by the update warps in their step p)
//   BAR_PN  panel warps only;  BAR_LY  W_p and pan4[p & 3] ready (panel -> update);  BAR_PUB  hand[p & 1] ready (update -> panel;
//           the first one says "factor storage zeroed")
template <int T, int NA, bool LATE>
__device__ __forceinline__ void update_step_pipe(double (&reg)[T][T], const double *__restrict__ Lp, const double *__restrict__ Y4, double *__restrict__ hand,
                                                 const int kb, const int q, const int ty, const int tx) {
  constexpr int JP = LATE ? 1 : 0;          // tile column of group p + 3 (q = 0: columns 12..15 of tile column 0; q >= 1: tile column 1)
  const int off = (4 * q + 12) & 15;        // first column of the group inside its tile column
  const int k0n = 16 * kb + 4 * q + 12;     // ... and in the matrix
  const int cn = tx - off;
  const bool owner = cn >= 0 && cn < 4;
  const double2 *pl = (const double2 *)(Lp + (size_t)(ty + 16 * (kb + JP)) * 4);
  const double2 *py = (const double2 *)(Y4 + (size_t)(tx + 16 * (kb + JP)) * 4);
  double *pub = hand + (size_t)(ty + 16 * (kb + JP)) * 4 + cn;
  nb_sync(BAR_LY);
  double li[NA][4];
#pragma unroll
  for (int ia = JP; ia < NA; ia++) {
    const double2 l01 = pl[32 * (ia - JP)], l23 = pl[32 * (ia - JP) + 1];
    li[ia][0] = l01.x; li[ia][1] = l01.y; li[ia][2] = l23.x; li[ia][3] = l23.y;
  }
  if (NA > JP) {
    const double2 y01 = py[0], y23 = py[1];
#pragma unroll
    for (int ia = JP; ia < NA; ia++) {
      const double v = fma(-li[ia][3], y23.y, fma(-li[ia][2], y23.x, fma(-li[ia][1], y01.y, fma(-li[ia][0], y01.x, reg[ia][JP]))));
      reg[ia][JP] = v;
      if (owner && ty + 16 * (kb + ia) >= k0n + cn) pub[64 * (ia - JP)] = v;
    }
  }
  nb_arrive(BAR_PUB);
#pragma unroll
  for (int jb = JP + 1; jb < NA; jb++) {
    const double2 y01 = py[32 * (jb - JP)], y23 = py[32 * (jb - JP) + 1];
#pragma unroll
    for (int ia = jb; ia < NA; ia++)
      reg[ia][jb] = fma(-li[ia][3], y23.y, fma(-li[ia][2], y23.x, fma(-li[ia][1], y01.y, fma(-li[ia][0], y01.x, reg[ia][jb]))));
  }
}
template <int T, int NA>
struct UpdatePipeDispatch {
  static __device__ __forceinline__ void run(int nact, bool late, double (&reg)[T][T], const double *Lp, const double *Y4, double *hand, int kb, int q, int ty, int tx) {
    if (nact == NA) {
      if (late) update_step_pipe<T, NA, true>(reg, Lp, Y4, hand, kb, q, ty, tx);
      else update_step_pipe<T, NA, false>(reg, Lp, Y4, hand, kb, q, ty, tx);
    } else UpdatePipeDispatch<T, NA - 1>::run(nact, late, reg, Lp, Y4, hand, kb, q, ty, tx);
  }
};
template <int T>
struct UpdatePipeDispatch<T, 0> {
  static __device__ __forceinline__ void run(int, bool, double (&)[T][T], const double *, const double *, double *, int, int, int, int) {}
};

// the panel warps of the pipelined factorisation: thread pt owns row pt for the whole factorisation (row D = right-hand side)
template <int T>
__device__ __forceinline__ void panel_pipeline(const double *__restrict__ As, double *__restrict__ M, double *__restrict__ pan4, const double *__restrict__ hand,
                                               const int D, const int PST, const int pt, const int npanels, long long *dbg_base) {
  const int i = pt;
  const bool mine = i <= D;
  double cur[4], nxt[4], nx2[4];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    cur[c] = (mine && c <= i && c < D) ? As[(size_t)i * D + c] : 0.0;
    nxt[c] = (mine && 4 + c <= i && 4 + c < D) ? As[(size_t)i * D + 4 + c] : 0.0;
    nx2[c] = (mine && 8 + c <= i && 8 + c < D) ? As[(size_t)i * D + 8 + c] : 0.0;
  }
#pragma unroll 1
  for (int p = 0; p < npanels; p++) {
    const int k0 = 4 * p;
    long long *dbg = (dbg_base && pt == 8 && p >= 4 && p < 6) ? dbg_base + 16 + 8 * (p - 4) : nullptr;
    double *pb = pan4 + (size_t)(p & 3) * (16 * T) * 4;
    if (mine) {
      *(double2 *)(pb + (size_t)i * 4) = make_double2(cur[0], cur[1]);
      *(double2 *)(pb + (size_t)i * 4 + 2) = make_double2(cur[2], cur[3]);
    }
    asm volatile("bar.sync %0, %1;" ::"r"(BAR_PN), "r"(SOLVE_PAN) : "memory");
    const double2 *pd = (const double2 *)(pb + (size_t)k0 * 4);
    const double a00 = pd[0].x;
    const double2 q1 = pd[2], q2a = pd[4], q2b = pd[5], q3a = pd[6], q3b = pd[7];
    const double a10 = q1.x, a11 = q1.y, a20 = q2a.x, a21 = q2a.y, a22 = q2b.x, a30 = q3a.x, a31 = q3a.y, a32 = q3b.x, a33 = q3b.y;
    if (dbg) dbg[0] = clock64() + (long long)(a00 == 123.0) + (long long)(a33 == 123.0);   // pivot block landed
    const double s0 = fma(a00, a11, -(a10 * a10)), s1 = fma(a00, a21, -(a10 * a20)), s2 = fma(a00, a31, -(a10 * a30));
    const double s3 = fma(a10, a21, -(a11 * a20)), s4 = fma(a10, a31, -(a11 * a30)), s5 = fma(a20, a31, -(a21 * a30));
    const double c5 = fma(a22, a33, -(a32 * a32)), c4 = fma(a21, a33, -(a31 * a32)), c3 = fma(a21, a32, -(a31 * a22));
    const double c2 = fma(a20, a33, -(a30 * a32)), c1 = fma(a20, a32, -(a30 * a22));
    const double det = fma(s5, s5, fma(-s4, c1, fma(s3, c2, fma(s2, c3, fma(-s1, c4, s0 * c5)))));   // c0 == s5
    const double rdet = (mine && i >= k0 + 4) ? safe_rcp(det) : 0.0;   // pivot rows and finished rows take no part
    const double i00 = fma(a31, c3, fma(-a21, c4, a11 * c5)), i01 = fma(-a30, c3, fma(a20, c4, -(a10 * c5)));
    const double i02 = fma(a33, s3, fma(-a32, s4, a31 * s5)), i03 = fma(-a32, s3, fma(a22, s4, -(a21 * s5)));
    const double i11 = fma(a30, c1, fma(-a20, c2, a00 * c5)), i12 = fma(-a33, s1, fma(a32, s2, -(a30 * s5)));
    const double i13 = fma(a32, s1, fma(-a22, s2, a20 * s5)), i22 = fma(a33, s0, fma(-a31, s2, a30 * s4));
    const double i23 = fma(-a32, s0, fma(a21, s2, -(a20 * s4))), i33 = fma(a22, s0, fma(-a21, s1, a20 * s3));
    double w0 = fma(cur[3], i03, fma(cur[2], i02, fma(cur[1], i01, cur[0] * i00))) * rdet;
    double w1 = fma(cur[3], i13, fma(cur[2], i12, fma(cur[1], i11, cur[0] * i01))) * rdet;
    double w2 = fma(cur[3], i23, fma(cur[2], i22, fma(cur[1], i12, cur[0] * i02))) * rdet;
    double w3 = fma(cur[3], i33, fma(cur[2], i23, fma(cur[1], i13, cur[0] * i03))) * rdet;
    if (!(mine && i >= k0 + 4)) w0 = w1 = w2 = w3 = 0.0;   // (0 * inf of a dead row would be NaN)
    if (dbg) dbg[1] = clock64() + (long long)(w3 == 123.0);   // chain done
    // p = 0: the update warps have zeroed the factor storage; p >= 1: their step p - 1 has handed over column group p + 2.
    // Taken BEFORE this step's BAR_LY arrival on purpose: it proves that every update warp has passed BAR_LY of step p - 1, so
    // two arrivals of ours can never pile up on that barrier (a named barrier completes as soon as its count is reached).
    nb_sync(BAR_PUB);
    if (mine && i >= k0 + 4) {
      double2 *pl = (double2 *)(M + (size_t)p * PST * 4 + (size_t)i * 4);
      pl[0] = make_double2(w0, w1);
      pl[1] = make_double2(w2, w3);
    }
    if (dbg) dbg[2] = clock64();
    nb_arrive(BAR_LY);
    if (p + 1 < npanels) {   // my next group: columns k0+4 .. k0+7, rows of those columns = A21 of this panel
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const double2 r01 = *(const double2 *)(pb + (size_t)(k0 + 4 + c) * 4), r23 = *(const double2 *)(pb + (size_t)(k0 + 4 + c) * 4 + 2);
        nxt[c] = fma(-w3, r23.y, fma(-w2, r23.x, fma(-w1, r01.y, fma(-w0, r01.x, nxt[c]))));
      }
    }
    if (p >= 1) {   // column group p + 2 as the update warps left it after panel p - 1
      if (mine) {
        const double2 h01 = *(const double2 *)(hand + (size_t)((p - 1) & 1) * (16 * T) * 4 + (size_t)i * 4),
                      h23 = *(const double2 *)(hand + (size_t)((p - 1) & 1) * (16 * T) * 4 + (size_t)i * 4 + 2);
        nx2[0] = h01.x; nx2[1] = h01.y; nx2[2] = h23.x; nx2[3] = h23.y;
      }
    }
    if (p + 2 < npanels) {
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const double2 r01 = *(const double2 *)(pb + (size_t)(k0 + 8 + c) * 4), r23 = *(const double2 *)(pb + (size_t)(k0 + 8 + c) * 4 + 2);
        nx2[c] = fma(-w3, r23.y, fma(-w2, r23.x, fma(-w1, r01.y, fma(-w0, r01.x, nx2[c]))));
      }
    }
#pragma unroll
    for (int c = 0; c < 4; c++) { cur[c] = nxt[c]; nxt[c] = nx2[c]; }
  }
}

// Back substitution L^T w = z, rows jhi..jlo (descending), all inside lane-register block MT (j >> 5 == MT): the step
// chain is shuffle (w_j) -> fma; the factor row of the next step is always in flight.  L[j][i] = M4[off_i + 4 j].
template <int W, int MT>
__device__ __forceinline__ void backsub_segment(double (&w)[W], const double *__restrict__ M4, const int (&off)[W], const int jhi, const int jlo, const int lane) {
  if (jhi < jlo) return;
  double Ln[MT + 1];
#pragma unroll
  for (int m = 0; m <= MT; m++) Ln[m] = (m < MT || lane + 32 * m < jhi) ? M4[off[m] + jhi * 4] : 0.0;
#pragma unroll 2
  for (int j = jhi; j >= jlo; j--) {
    double Lc[MT + 1];
#pragma unroll
    for (int m = 0; m <= MT; m++) {
      Lc[m] = Ln[m];
      Ln[m] = (j > jlo && (m < MT || lane + 32 * m < j - 1)) ? M4[off[m] + (j - 1) * 4] : 0.0;
    }
    const double wj = __shfl_sync(0xffffffffu, w[MT], j & 31);
#pragma unroll
    for (int m = 0; m <= MT; m++) w[m] = fma(-Lc[m], wj, w[m]);   // Lc is 0 for columns i >= j
  }
}

// Back substitution L^T w = z in blocks of 4 rows (one warp; lane l keeps the partially updated rows l, l+32, ...):
// per block the four finished rows are broadcast, the unit 4x4 triangle is solved redundantly in every lane and the
// remaining rows take one rank-4 update, so the dependent chain is 4 shuffles + ~4 fma per FOUR rows instead of
// shuffle + fma per row.  The factor entries of the next block are always in flight.  L[j][i] = M4[off_i + 4 j].
template <int W>
__device__ __forceinline__ void backsub_blocked(const double *__restrict__ M4, double *__restrict__ z, const int D, const int PST, const int lane) {
  double w[W];
  const double *col[W];   // column (lane + 32 m) of the factor: L[j][i] = col[m][4 j]
#pragma unroll
  for (int m = 0; m < W; m++) {
    const int i = lane + 32 * m;
    col[m] = M4 + (size_t)(i >> 2) * PST * 4 + (i & 3);
    w[m] = i < D ? col[m][D * 4] : 0.0;
  }
#pragma unroll 1
  for (int b = (D >> 2) - 1; b >= 0; b--) {
    const int j0 = 4 * b, mb = j0 >> 5, l0 = j0 & 31;
    double ws = w[0];
#pragma unroll
    for (int m = 1; m < W; m++) ws = mb == m ? w[m] : ws;
    // the chain: the four finished rows of this block, broadcast ...
    const double v0 = __shfl_sync(0xffffffffu, ws, l0), v1 = __shfl_sync(0xffffffffu, ws, l0 + 1), v2 = __shfl_sync(0xffffffffu, ws, l0 + 2),
                 v3 = __shfl_sync(0xffffffffu, ws, l0 + 3);
    // ... while the factor entries of the block arrive (they do not depend on the chain)
    const double *d = M4 + ((size_t)b * PST + j0) * 4;   // rows j0..j0+3 of this panel's block: the unit lower 4x4 triangle
    const double l10 = d[4], l20 = d[8], l21 = d[9], l30 = d[12], l31 = d[13], l32 = d[14];
    double L[W][4];
#pragma unroll
    for (int m = 0; m < W; m++) {
      const bool on = lane + 32 * m < j0;
      const double *c = on ? col[m] + j0 * 4 : M4;   // rows at or past the block: any valid address, the value is dropped
#pragma unroll
      for (int k = 0; k < 4; k++) { const double t = c[4 * k]; L[m][k] = on ? t : 0.0; }
    }
    const double x3 = v3;
    const double x2 = fma(-l32, x3, v2);
    const double x1 = fma(-l21, x2, fma(-l31, x3, v1));
    const double x0 = fma(-l10, x1, fma(-l20, x2, fma(-l30, x3, v0)));
#pragma unroll
    for (int m = 0; m < W; m++) w[m] = fma(-L[m][0], x0, fma(-L[m][1], x1, fma(-L[m][2], x2, fma(-L[m][3], x3, w[m]))));
    if (lane == 0) { *(double2 *)(z + j0) = make_double2(x0, x1); *(double2 *)(z + j0 + 2) = make_double2(x2, x3); }
  }
}

template <bool RETARGET>
__device__ void frame_step_body(const StepArgs &a);

template <int T>
__global__ void __launch_bounds__(SOLVE_THREADS) k_solve(SolveArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int D = a.D, LD = D + 1, tid = threadIdx.x, nf = a.nf;
  const int DP = D + 1;
  constexpr int PST = 16 * T + 1;          // rows per panel block of the factor storage (+1: de-phases the banks of consecutive blocks)
  double *M = sm;                          // [D][D] staged raw top H; later M4: [D/4][PST][4] unit lower factor, panel by panel, row D = rhs
  double *As = M + (size_t)D * PST;        // [D+1][D] assembled lower triangle, row D = rhs
  double *S = As + (size_t)DP * D;         // [D] Jacobi scaling
  double *bb = S + D;                      // [D] unscaled rhs
  double *dtmp = bb + D;                   // [D] damped diagonal
  double *delta = dtmp + D;                // [D]
  double *key = delta + D;                 // [D]
  double *z = key + D;                     // [D]
  double *pan = z + D;                     // [2][16T*4] panel, double buffered, rows past D stay zero
  double *Yb = pan + 2 * (size_t)(16 * T) * 4;   // [2][16T*4] Y rows of the current panel, double buffered (pipelined scheme: pan[2..3])
  double *hand = Yb + 2 * (size_t)(16 * T) * 4;  // [2][16T*4] pipelined scheme: column group p + 3 on its way to the panel warps
  double *stage = hand + 2 * (size_t)(16 * T) * 4; // optional: accSC [(D+1)^2 (+1)], then HM [D*D]
  __shared__ int perm[144];
  __shared__ __align__(8) unsigned long long mbar;
  const double lambda = 1e-5;                               // EnergyFunctional.cpp:1031
  const double sc = (double)(1.0f / (float)(1 + lambda));   // float-typed scalar (:1099)
  const double *cPrior = a.wprior, *fprior = a.wprior + 4, *fdp = a.wprior + 4 + 8 * nf, *fdelta = a.wprior + 4 + 16 * nf;
  const int warp = tid >> 5, lane = tid & 31;
  const int ty = tid >> 4, tx = tid & 15;
#define SOLVE_TS(n) do { if (a.dbg && tid == 0) a.dbg[n] = clock64(); } while (0)
  PDL_ENTER_T(a.trace);
  if (blockIdx.x == 1) {
    // spare CTA (point shards).  The selection belongs to the linearisation of the previous body, so it runs whenever the
    // loop had not broken BEFORE this launch: body i-1's solve left ctl[1] = i unless the latch was already set (CTA 0 of
    // this launch may be raising it to i+1 right now -- either value says "not broken before").
    if (a.do_th && (!a.ctl || a.ctl[1] >= a.iter_index)) energy_th_body(a.th, (unsigned *)sm, a.smem_words);   // this CTA's share of the dynamic shared memory is free
    return;
  }
  SOLVE_TS(0);
  // ---- stage the inputs in shared memory: one thread issues the bulk copies (first thing, so they overlap the loop-control
  // check below), everyone waits on the mbarrier -----------------------------------------------------------------
  const double *Hsrc = M, *Ssrc = a.accSC, *HMsrc = a.HM;
  const unsigned bytesH = (unsigned)((size_t)D * D * 8), bytesS = (unsigned)((((size_t)DP * DP + 1) & ~(size_t)1) * 8);
  double *Ss = stage, *Hm = stage + (bytesS >> 3);
  if (a.stage_sc) Ssrc = Ss;
  if (a.HM && a.stage_hm) HMsrc = Hm;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const unsigned total = bytesH + (a.stage_sc ? bytesS : 0u) + ((a.HM && a.stage_hm) ? bytesH : 0u);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(total) : "memory");
    bulk_g2s(M, a.Htop, bytesH, &mbar);
    if (a.stage_sc) bulk_g2s(Ss, a.accSC, bytesS, &mbar);
    if (a.HM && a.stage_hm) bulk_g2s(Hm, a.HM, bytesH, &mbar);
  }
  if (a.ctl) {   // did the previous loop body converge?  (doStepFromBackup's return value, FullSystemOptimize.cpp:238-256)
    __shared__ int s_stop;
    if (tid == 0) {
      int stop = a.ctl[0];
      if (!stop && a.iter_index >= 1 && a.iter_index - 1 >= a.min_it) {
        const double *it = a.step.iter;   // sumA, sumB, sumT, sumR of the previous body (already / nf)
        const float sumA = (float)it[0], sumB = (float)it[1], sumT = (float)it[2], sumR = (float)it[3];
        const float numID = (float)a.prev_rstats[2];
        const float sumNID = (float)a.prev_rstats[1] / numID;   // no points: 0/0 = NaN and the loop never breaks, like the reference (FullSystemOptimize.cpp:228-256)
        const float th = a.th_opt;
        stop = sqrtf(sumA) < 0.0005 * th && sqrtf(sumB) < 0.00005 * th && sqrtf(sumR) < 0.00005 * th && sqrtf(sumT) * sumNID < 0.00005 * th;
        if (stop) a.ctl[0] = 1;
      }
      if (!stop) a.ctl[1] = a.iter_index + 1;
      s_stop = stop;
    }
    __syncthreads();
    if (s_stop) {   // the copies into this CTA's shared memory must land before it exits
      unsigned done = 0;
      while (!done)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
      return;
    }
  }
  if (tid == 0 && a.res_out) a.res_out[0] = a.res_in[0];
  if (a.stash_dst && tid >= 32 && tid < 44) a.stash_dst[tid - 32] = a.stash_src[tid - 32];   // sums of the first linearisation (the step launch clears them)
  if (a.zero_rstats && tid < 4) a.zero_rstats[tid] = 0.0;   // the back-substitution sums of this body (k_resubstitute follows)

  for (int i = tid; i < 6 * (16 * T) * 4; i += SOLVE_THREADS) pan[i] = 0.0;   // panel + Y + hand-over buffers: rows past D must read as zero
  for (int i = tid; i < D; i += SOLVE_THREADS) delta[i] = i < 4 ? (double)a.cDeltaF[i] : fdelta[i - 4];
  double pr = 0.0, dpr = 0.0, bt = 0.0, bm = 0.0;   // per-row inputs that stay in global memory: fetch them under the copies
  if (tid < D) {
    pr = tid < 4 ? cPrior[tid] : fprior[tid - 4];
    dpr = tid < 4 ? (double)a.cDeltaF[tid] : fdp[tid - 4];
    bt = a.btop[tid];
    if (a.HM) bm = a.bM[tid];
  }
  __syncthreads();   // mbarrier initialised and armed before anyone polls it
  {
    unsigned done = 0;
    while (!done) {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
  }
  // ---- diagonal, rhs --------------------------------------------------------------------------------
  double *hmd = key;   // HM * delta, bM_top = bM + HM * delta (:1070-1091)
  if (a.HM) {
    for (int r = warp; r < D; r += SOLVE_THREADS / 32) {
      double s = 0;
      for (int c = lane; c < D; c += 32) s += HMsrc[(size_t)r * D + c] * delta[c];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) hmd[r] = s;
    }
    __syncthreads();
  }
  if (tid < D) {
    const int r = tid;
    const double scb = Ssrc[(size_t)r * DP + D], ht = Hsrc[(size_t)r * D + r], scd = Ssrc[(size_t)r * DP + r];
    double v = bt + pr * dpr;                            // AccumulatedTopHessian.cpp:292-300 (the L pass carries the priors)
    double dg = ht + pr;
    if (a.HM) { v += bm + hmd[r]; dg += HMsrc[(size_t)r * D + r]; }
    v -= scb;
    bb[r] = v;
    if (a.bfinal) a.bfinal[r] = v;
    dg *= (1 + lambda);
    dg -= scd * sc;
    dtmp[r] = dg;
    const double sr = 1.0 / sqrt(dg + 10.0);             // :1143-1146
    S[r] = sr;
  }
  __syncthreads();
  if (tid < D) key[tid] = fabs(S[tid] * dtmp[tid] * S[tid]);
  __syncthreads();
  SOLVE_TS(1);
  // ---- pivot order = descending |diag| ---------------------------------------------------------------------
  if (tid < D) {
    const double mine = key[tid];
    int rank = 0;
    for (int j = 0; j < D; j++) { const double o = key[j]; rank += (o > mine || (o == mine && j < tid)) ? 1 : 0; }
    perm[rank] = tid;
  }
  __syncthreads();
  SOLVE_TS(2);
  // ---- assemble the permuted, scaled lower triangle (+ rhs row) in shared memory, then pull the register tiles ------
  // rows r and D-1-r together hold D+1 lower-triangle entries: a branch-free enumeration of exactly D(D+1)/2 entries
  {
    const int total = (D >> 1) * DP;
#pragma unroll 5
    for (int e = tid; e < total; e += SOLVE_THREADS) {
      const int rr = e / DP, cc = e - rr * DP;
      const bool lo = cc <= rr;
      const int i = lo ? rr : D - 1 - rr, j = lo ? cc : cc - rr - 1;
      const int pi = perm[i], pj = perm[j];
      const int r = max(pi, pj), c = min(pi, pj);
      const double h1 = Hsrc[(size_t)r * D + c];   // symmetrised by the stitch (AccumulatedTopHessian.h:107-126 epilogue)
      const double hs = Ssrc[(size_t)c * DP + r];
      const double hm = a.HM ? HMsrc[(size_t)r * D + c] : 0.0;
      double u = h1 + hm;
      u -= hs * sc;
      u = r == c ? dtmp[r] : u;
      if (a.Hfinal) { a.Hfinal[(size_t)r * D + c] = u; a.Hfinal[(size_t)c * D + r] = u; }
      As[i * D + j] = S[r] * u * S[c];
    }
    for (int j = tid; j < D; j += SOLVE_THREADS) { const int c = perm[j]; As[D * D + j] = S[c] * bb[c]; }
  }
  __syncthreads();
  SOLVE_TS(3);
  // ---- right-looking LDL^T in panels of 4 columns with look-ahead ---------------------------------------------------
  // update warps (tid < 256) own the trailing tiles in registers; panel warps (tid >= 256) own one row each of the
  // current panel.  BAR_PUB: panel p published (update -> panel), BAR_LY: L/Y of panel p ready (panel -> update).
  const int npanels = D >> 2;
  if (tid < SOLVE_UPD) {
    double reg[T][T];
#pragma unroll
    for (int ia = 0; ia < T; ia++)
#pragma unroll
      for (int jb = 0; jb < T; jb++) {
        const int i = ty + 16 * ia, j = tx + 16 * jb;
        reg[ia][jb] = (jb <= ia && i <= D && j < D && j <= i) ? As[i * D + j] : 0.0;
      }
    if (!a.pipe && tx < 4) {
#pragma unroll
      for (int ia = 0; ia < T; ia++) {
        const int i = ty + 16 * ia;
        if (i >= tx) pan[i * 4 + tx] = reg[ia][0];
      }
    }
    for (int e = tid; e < D * PST; e += SOLVE_UPD) M[e] = 0.0;   // the staged H is consumed: the region becomes the factor storage
    nb_arrive(BAR_PUB);
#pragma unroll 1
    for (int kb = 0; 4 * kb < npanels; kb++) {
      const int nact = (DP - 16 * kb + 15) >> 4;   // live tile rows (rows <= D)
#pragma unroll 1
      for (int q = 0; q < 4; q++) {
        const int p = 4 * kb + q;
        if (p >= npanels) break;
        if (p + 1 == npanels) { nb_sync(BAR_LY); break; }   // nothing left to update: the rhs row was finished by the panel warps
        long long *dbg = (a.dbg && (tid == 0 || tid == 224) && p >= 4 && p < 6) ? a.dbg + 16 + 8 * (p - 4) + 3 : nullptr;
        if (a.pipe) {
          if (dbg && tid == 0) dbg[0] = clock64();
          UpdatePipeDispatch<T, T>::run(nact, q >= 1, reg, M + (size_t)p * PST * 4, pan + (size_t)(p & 3) * (16 * T) * 4, hand + (size_t)(p & 1) * (16 * T) * 4, kb, q, ty, tx);
          if (dbg && tid == 0) dbg[2] = clock64() + (long long)(reg[T - 1][T - 1] == 123.0);
          continue;
        }
        // Y operand of the rank-4 update: block pivots -> the published panel itself (A21), else the Y rows of the panel warps
        UpdateDispatch<T, T>::run(nact, q == 3, reg, M + (size_t)p * PST * 4, (a.block_pivots ? pan : Yb) + (size_t)(p & 1) * (16 * T) * 4,
                                  pan + (size_t)((p + 1) & 1) * (16 * T) * 4, kb, q, ty, tx, dbg);
        if (dbg && tid == 0) dbg[2] = clock64() + (long long)(reg[T - 1][T - 1] == 123.0);
      }
      // next 16 columns: tile (a,b) takes over from tile (a+1,b+1)
#pragma unroll
      for (int ia = 0; ia < T - 1; ia++)
#pragma unroll
        for (int jb = 0; jb <= ia; jb++) reg[ia][jb] = reg[ia + 1][jb + 1];
#pragma unroll
      for (int jb = 0; jb < T; jb++) reg[T - 1][jb] = 0.0;
    }
  } else {
    const int pt = tid - SOLVE_UPD;
    if (a.pipe) panel_pipeline<T>(As, M, pan, hand, D, PST, pt, npanels, a.dbg);
    else
#pragma unroll 1
    for (int p = 0; p < npanels; p++) {
      long long *dbg = (a.dbg && pt == 8 && p >= 4 && p < 6) ? a.dbg + 16 + 8 * (p - 4) : nullptr;
      if (a.block_pivots) panel_rows<true>(pan + (size_t)(p & 1) * (16 * T) * 4, M + (size_t)p * PST * 4, Yb + (size_t)(p & 1) * (16 * T) * 4, D, 4 * p, pt, dbg);
      else panel_rows<false>(pan + (size_t)(p & 1) * (16 * T) * 4, M + (size_t)p * PST * 4, Yb + (size_t)(p & 1) * (16 * T) * 4, D, 4 * p, pt, dbg);
      if (dbg) dbg[2] = clock64();
      nb_arrive(BAR_LY);
    }
  }
  SOLVE_TS(4);
  // ---- back substitution L^T w = z in one warp (lane l keeps rows l, l+32, ...), the next row of L always in flight.
  // L[j][i] sits at M4[i >> 2][j][i & 3]; the rhs row D holds z.
  __syncthreads();
  if (warp == 0 && a.backsub_rowwise == 0) {
    backsub_blocked<(16 * T + 31) / 32>(M, z, D, PST, lane);
  } else if (warp == 0) {
    constexpr int W = (16 * T + 31) / 32;
    double w[W];
    int off[W];   // offset of column (lane + 32 m) inside a factor row
#pragma unroll
    for (int m = 0; m < W; m++) {
      const int i = lane + 32 * m;
      off[m] = (i >> 2) * PST * 4 + (i & 3);
      w[m] = i < D ? M[off[m] + D * 4] : 0.0;
    }
    if (W > 3) backsub_segment<W, (W > 3 ? 3 : 0)>(w, M, off, min(D - 1, 127), 96, lane);
    if (W > 2) backsub_segment<W, (W > 2 ? 2 : 0)>(w, M, off, min(D - 1, 95), 64, lane);
    if (W > 1) backsub_segment<W, (W > 1 ? 1 : 0)>(w, M, off, min(D - 1, 63), 32, lane);
    backsub_segment<W, 0>(w, M, off, min(D - 1, 31), 1, lane);
#pragma unroll
    for (int m = 0; m < W; m++) { const int i = lane + 32 * m; if (i < D) z[i] = w[m]; }
  }
  __syncthreads();
  SOLVE_TS(5);
  double *y = key;   // x in original order
  bool bad = false;
  for (int j = tid; j < D; j += SOLVE_THREADS) { const int r = perm[j]; const double xi = S[r] * z[j]; y[r] = xi; a.x[r] = xi; if (!isfinite(xi)) bad = true; }
  if (bad && a.status) atomicOr(a.status, 1);
  __syncthreads();
  SOLVE_TS(6);
  // ---- xAd (EnergyFunctional.cpp:509-513) and xc ------------------------------------------------------
  if (a.xAd) {
    for (int e = tid; e < nf * nf * 8; e += SOLVE_THREADS) {
      const int c = e & 7, ht = e >> 3, h = ht / nf, t = ht % nf;   // xAd index = h*nf + t
      const float *AhF = a.adHostF + 64 * (size_t)(h + nf * t), *AtF = a.adTargetF + 64 * (size_t)(h + nf * t);
      float ah[8], at[8];
#pragma unroll
      for (int k = 0; k < 8; k++) { ah[k] = __ldg(AhF + k * 8 + c); at[k] = __ldg(AtF + k * 8 + c); }
      float sh = 0.f, st = 0.f;
#pragma unroll
      for (int k = 0; k < 8; k++) { sh = fmaf((float)y[4 + 8 * h + k], ah[k], sh); st = fmaf((float)y[4 + 8 * t + k], at[k], st); }
      a.xAd[e] = sh + st;
    }
    if (tid < 4) a.xAd[(size_t)nf * nf * 8 + tid] = (float)y[tid];
  }
  SOLVE_TS(7);
  // ---- frames / calibration part of doStepFromBackup, new precalc and deltas (same CTA, no extra launch) ---------
  if (a.do_step) {
    __syncthreads();
    frame_step_body<false>(a.step);
  }
  SOLVE_TS(8);
  TRACE_EXIT(a.trace);
}

// ------------------------------------------------------------------------------------------------
// Frames / calibration part of one Gauss-Newton step on the device (single CTA, fp64):
//   EnergyFunctional::resubstituteF_MT frame+calib steps (:500-507), FullSystem::backupState (:260-271),
//   doStepFromBackup frames+calib (:185-257), FrameHessian::setState (HessianBlocks.h:217-230),
//   FrameFramePrecalc::set (HessianBlocks.cpp:431-461), setDeltaF (EnergyFunctional.cpp:163-194).
// Mirrors host_ba.cpp (same float operation order); keeps the whole iteration on the stream with no host round trip.
__device__ __forceinline__ void d_mul33f(const float *A, const float *B, float *C) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[3 * i + j] = (A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j]) + A[3 * i + 2] * B[6 + j];
}

// RETARGET (after the loop, FullSystemOptimize.cpp:415-423): no step; the newest frame gets its current pose as the new
// evaluation point (setEvalPT(PRE_camToWorld, [0, aff])), and with it new adjoints (EnergyFunctional::setAdjointsF,
// EnergyFunctional.cpp:42-103; mirrors host_ba.cpp make_adjoints), precalc and deltas for every pair.
template <bool RETARGET>
__device__ void frame_step_body(const StepArgs &a) {
  using namespace sosba_math;
  __shared__ double s_fs[16 * SOSBA_FS];
  __shared__ Rigid s_c2w[16], s_w2c[16];
  __shared__ double s_scaled[16][2];
  __shared__ float s_K[4];
  const int nf = a.nf, tid = threadIdx.x;
  const double SC_T = 0.5, SC_R = 1.0, SC_A = 10.0, SC_B = 1000.0, SC_F = 50.0, SC_C = 50.0;
#define FS_TS(n) do { if (a.trace && tid == 0) a.trace[n] = clock64(); } while (0)
  FS_TS(0);
  for (int e = tid; e < nf * SOSBA_FS; e += blockDim.x) s_fs[e] = a.fs[e];
  __syncthreads();
  FS_TS(1);
  if (tid < nf) {
    double *F = s_fs + SOSBA_FS * tid;
    double *G = a.fs + SOSBA_FS * tid;
    double *state = F + 12, *backup = F + 32, *step = F + 42;
    double scaled[10];
    if (!RETARGET) {
      for (int i = 0; i < 8; i++) step[i] = -a.x[4 + 8 * tid + i];
      step[8] = step[9] = 0.0;
      for (int i = 0; i < 10; i++) {
        backup[i] = state[i];
        state[i] = backup[i] + (double)a.stepfac * step[i];
        G[12 + i] = state[i]; G[32 + i] = backup[i]; G[42 + i] = step[i];
      }
    }
    for (int i = 0; i < 3; i++) scaled[i] = SC_T * state[i];
    for (int i = 3; i < 6; i++) scaled[i] = SC_R * state[i];
    scaled[6] = SC_A * state[6]; scaled[7] = SC_B * state[7]; scaled[8] = SC_A * state[8]; scaled[9] = SC_B * state[9];
    Rigid ev = rigid_from34(F);
    Rigid c2w = rigid_mul(rigid_exp(scaled), ev);
    if (RETARGET && tid == nf - 1) {   // FrameHessian::setEvalPT(PRE_camToWorld, [0 0 0 0 0 0 a b 0 0])
      ev = c2w;
      rigid_to34(ev, F); rigid_to34(ev, G);
      for (int i = 0; i < 10; i++) {
        const double v = (i == 6 || i == 7) ? state[i] : 0.0;
        state[i] = v; F[22 + i] = v; G[12 + i] = v; G[22 + i] = v;
        scaled[i] = i < 3 ? SC_T * v : i < 6 ? SC_R * v : (i & 1) ? SC_B * v : SC_A * v;
      }
      c2w = rigid_mul(rigid_exp(scaled), ev);
    }
    s_c2w[tid] = c2w;
    s_w2c[tid] = rigid_inverse(c2w);
    s_scaled[tid][0] = scaled[6]; s_scaled[tid][1] = scaled[7];
    for (int i = 0; i < 8; i++) {
      a.wprior[4 + 8 * nf + 8 * tid + i] = state[i];                 // delta_prior = state - getPriorZero() (== 0)
      a.wprior[4 + 16 * nf + 8 * tid + i] = state[i] - F[22 + i];    // delta = state - state_zero
    }
  }
  if (tid == 32) {  // calibration: CalibHessian::setValue (HessianBlocks.h:487-501)
    double *C = a.cs;   // value[4] | value_zero[4] | value_backup[4] | step[4]
    float sf[4];
    for (int i = 0; i < 4; i++) {
      if (!RETARGET) {
        C[12 + i] = -a.x[i];
        C[8 + i] = C[i];
        C[i] = C[8 + i] + (double)a.stepfac * C[12 + i];
      }
      sf[i] = (float)((i < 2 ? SC_F : SC_C) * C[i]);
      s_K[i] = sf[i];
      a.calib[i] = sf[i];
      a.calib[6 + i] = (float)(C[i] - C[4 + i]);
    }
    a.calib[4] = 1.0f / sf[0];
    a.calib[5] = 1.0f / sf[1];
  }
  __syncthreads();
  FS_TS(2);
  if (!RETARGET && tid == 64) {  // step norms of doStepFromBackup, float accumulation in frame order
    float sumA = 0, sumB = 0, sumT = 0, sumR = 0;
    for (int f = 0; f < nf; f++) {
      const double *st = s_fs + SOSBA_FS * f + 42;
      sumA += st[6] * st[6];
      sumB += st[7] * st[7];
      sumT += st[0] * st[0] + st[1] * st[1] + st[2] * st[2];
      sumR += st[3] * st[3] + st[4] * st[4] + st[5] * st[5];
    }
    a.iter[0] = sumA / nf; a.iter[1] = sumB / nf; a.iter[2] = sumT / nf; a.iter[3] = sumR / nf;
  }
  const float fx = s_K[0], fy = s_K[1], cx = s_K[2], cy = s_K[3];
  const float K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
  const float Ki[9] = {1.0f / fx, 0, -cx / fx, 0, 1.0f / fy, -cy / fy, 0, 0, 1};
  for (int e = tid; e < nf * nf; e += blockDim.x) {
    const int h = e / nf, t = e % nf;
    const double *Fh = s_fs + SOSBA_FS * h, *Ft = s_fs + SOSBA_FS * t;
    // setDeltaF first (its 32 float4 adjoint loads are in flight while the fp64 rigid algebra runs)
    const int idx = h + t * nf;
    const float4 *AhF = (const float4 *)(a.adHostF + 64 * (size_t)idx), *AtF = (const float4 *)(a.adTargetF + 64 * (size_t)idx);
    float4 rh[16], rt[16];
    if (!RETARGET) {
#pragma unroll
      for (int q = 0; q < 16; q++) { rh[q] = __ldg(AhF + q); rt[q] = __ldg(AtF + q); }
    } else {   // setAdjointsF for this pair, from the (new) evaluation points
      double Adj[36];
      rigid_adj(rigid_inverse(rigid_from34(Ft)), Adj);   // worldToTarget at the evaluation point
      double AH[64], AT[64];
      for (int i = 0; i < 64; i++) AH[i] = AT[i] = 0.0;
      for (int i = 0; i < 8; i++) AH[9 * i] = AT[9 * i] = 1.0;
      for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) { AH[8 * i + j] = Adj[6 * j + i]; AT[8 * i + j] = -Adj[6 * j + i]; }
      float eF = (float)Fh[52], eT = (float)Ft[52];
      if (eF == 0 || eT == 0) eT = eF = 1;
      const float a0 = (float)(exp(Ft[22 + 6] * SC_A - Fh[22 + 6] * SC_A) * eT / eF);   // aff_g2l_0 of both frames
      AT[8 * 6 + 6] = -a0; AH[8 * 6 + 6] = a0; AT[8 * 7 + 7] = -1; AH[8 * 7 + 7] = a0;
      const double rs[8] = {SC_T, SC_T, SC_T, SC_R, SC_R, SC_R, SC_A, SC_B};
      double *gH = a.adHost + 64 * (size_t)idx, *gT = a.adTarget + 64 * (size_t)idx;
      float *gHf = const_cast<float *>(a.adHostF) + 64 * (size_t)idx, *gTf = const_cast<float *>(a.adTargetF) + 64 * (size_t)idx;
      float fh[64], ft[64];
      for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) {
          const double vh = AH[8 * i + j] * rs[i], vt = AT[8 * i + j] * rs[i];
          gH[8 * i + j] = vh; gT[8 * i + j] = vt;
          fh[8 * i + j] = (float)vh; ft[8 * i + j] = (float)vt;
          gHf[8 * i + j] = fh[8 * i + j]; gTf[8 * i + j] = ft[8 * i + j];
        }
#pragma unroll
      for (int q = 0; q < 16; q++) {
        rh[q] = make_float4(fh[4 * q], fh[4 * q + 1], fh[4 * q + 2], fh[4 * q + 3]);
        rt[q] = make_float4(ft[4 * q], ft[4 * q + 1], ft[4 * q + 2], ft[4 * q + 3]);
      }
    }
    float pre[SOSBA_PRECALC_FLOATS];
    if (RETARGET) {   // PRE_RTll_0 / PRE_tTll_0 depend on the evaluation points only: constant while the loop runs
      const Rigid l0 = rigid_mul(rigid_inverse(rigid_from34(Ft)), rigid_from34(Fh));
      for (int i = 0; i < 9; i++) pre[SOSBA_PC_RTLL0 + i] = (float)l0.R[i];
      for (int i = 0; i < 3; i++) pre[SOSBA_PC_TTLL0 + i] = (float)l0.t[i];
    }
    const Rigid l = rigid_mul(s_w2c[t], s_c2w[h]);
    float R[9], tt[3], KR[9];
    for (int i = 0; i < 9; i++) R[i] = (float)l.R[i];
    for (int i = 0; i < 3; i++) tt[i] = (float)l.t[i];
    d_mul33f(K, R, KR);
    d_mul33f(KR, Ki, pre + SOSBA_PC_KRKI);
    for (int i = 0; i < 3; i++) pre[SOSBA_PC_KT + i] = (K[3 * i] * tt[0] + K[3 * i + 1] * tt[1]) + K[3 * i + 2] * tt[2];
    float expF = (float)Fh[52], expT = (float)Ft[52];     // AffLight::fromToVecExposure (NumType.h:157-168)
    if (expF == 0 || expT == 0) expT = expF = 1;
    const double aa = exp(s_scaled[t][0] - s_scaled[h][0]) * expT / expF;
    const double bbv = s_scaled[t][1] - aa * s_scaled[h][1];
    pre[SOSBA_PC_AFF] = (float)aa;
    pre[SOSBA_PC_AFF + 1] = (float)bbv;
    pre[SOSBA_PC_B0] = (float)(Fh[22 + 7] * SC_B);
    pre[SOSBA_PC_DIST] = (float)sqrt(l.t[0] * l.t[0] + l.t[1] * l.t[1] + l.t[2] * l.t[2]);
    pre[28] = pre[29] = pre[30] = pre[31] = 0.f;
    float4 *p4 = (float4 *)(a.precalc + (size_t)e * SOSBA_PRECALC_FLOATS);
#pragma unroll
    for (int q = RETARGET ? 0 : 3; q < 8; q++) p4[q] = make_float4(pre[4 * q], pre[4 * q + 1], pre[4 * q + 2], pre[4 * q + 3]);
    // adHTdeltaF[h + t*nf] = delta_h^T adHostF + delta_t^T adTargetF, summed over k in order
    float sh[8] = {0, 0, 0, 0, 0, 0, 0, 0}, st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const float dh = (float)(Fh[12 + k] - Fh[22 + k]), dt = (float)(Ft[12 + k] - Ft[22 + k]);
      const float4 h0 = rh[2 * k], h1 = rh[2 * k + 1], t0 = rt[2 * k], t1 = rt[2 * k + 1];
      sh[0] += dh * h0.x; sh[1] += dh * h0.y; sh[2] += dh * h0.z; sh[3] += dh * h0.w;
      sh[4] += dh * h1.x; sh[5] += dh * h1.y; sh[6] += dh * h1.z; sh[7] += dh * h1.w;
      st[0] += dt * t0.x; st[1] += dt * t0.y; st[2] += dt * t0.z; st[3] += dt * t0.w;
      st[4] += dt * t1.x; st[5] += dt * t1.y; st[6] += dt * t1.z; st[7] += dt * t1.w;
    }
    float4 *o4 = (float4 *)(a.adHTdeltaF + 8 * (size_t)idx);
    o4[0] = make_float4(sh[0] + st[0], sh[1] + st[1], sh[2] + st[2], sh[3] + st[3]);
    o4[1] = make_float4(sh[4] + st[4], sh[5] + st[5], sh[6] + st[6], sh[7] + st[7]);
  }
  FS_TS(3);
}


// the new evaluation point of the newest keyframe at the end of FullSystem::optimize, with all dependent window tables
__global__ void __launch_bounds__(256) k_frame_retarget(StepArgs a, ThArgs th, int do_th, int *zero_words, int n_zero) {
  PDL_ENTER();
  if (blockIdx.x == 1) {   // spare CTA: the pending threshold selection of the last linearisation, then the sums of the next one
    energy_th_body(th);
    __syncthreads();
    if (zero_words && (int)threadIdx.x < n_zero) zero_words[threadIdx.x] = 0;
    return;
  }
  if (!do_th && zero_words && (int)threadIdx.x < n_zero) zero_words[threadIdx.x] = 0;
  frame_step_body<true>(a);
}

// The step of one loop body in ONE launch: the points (back-substitution + doStepFromBackup, CTAs 0..n-2) and, concurrently
// in the spare last CTA, the frames / calibration / precalc / deltas.  Both only need x from k_solve.
__global__ void __launch_bounds__(256) k_step(ResubArgs ra, StepArgs sa) {
  __shared__ __align__(16) float s_xad[16 * 16 * 8 + 4];
  PDL_ENTER_T(ra.trace);
  if (ra.gate && *ra.gate) return;
  if (blockIdx.x == gridDim.x - 1) { frame_step_body<false>(sa); if (ra.trace) TRACE_EXIT(ra.trace + 4); return; }   // record + 1: the frame CTA on its own
  // xAd[h*nf+t] = x_h^T adHostF[h+nf*t] + x_t^T adTargetF[h+nf*t] and xc (EnergyFunctional.cpp:509-513): every CTA builds
  // the small table in shared memory (adjoints are L2-resident) instead of waiting for k_solve to do it serially
  {
    const int nf = ra.nf, tid = threadIdx.x;
    const double *x = sa.x;
    for (int e = tid; e < nf * nf * 8; e += blockDim.x) {
      const int c = e & 7, ht = e >> 3, h = ht / nf, t = ht % nf;
      const float *AhF = sa.adHostF + 64 * (size_t)(h + nf * t), *AtF = sa.adTargetF + 64 * (size_t)(h + nf * t);
      float ah[8], at[8];
#pragma unroll
      for (int k = 0; k < 8; k++) { ah[k] = __ldg(AhF + k * 8 + c); at[k] = __ldg(AtF + k * 8 + c); }
      float sh = 0.f, st = 0.f;
#pragma unroll
      for (int k = 0; k < 8; k++) { sh = fmaf((float)x[4 + 8 * h + k], ah[k], sh); st = fmaf((float)x[4 + 8 * t + k], at[k], st); }
      s_xad[e] = sh + st;
    }
    if (tid < 4) s_xad[nf * nf * 8 + tid] = (float)x[tid];
    __syncthreads();
    ra.xAd = s_xad;
  }
  resubstitute_body(ra, blockIdx.x, gridDim.x - 1);
  TRACE_EXIT(ra.trace);
}

// resubstitute with a caller-provided x: only the xAd part of the kernel above
__global__ void __launch_bounds__(256) k_make_xad(const double *__restrict__ x, int nf, const float *__restrict__ adHostF,
                                                  const float *__restrict__ adTargetF, float *__restrict__ xAd) {
  const int tid = threadIdx.x;
  for (int e = tid; e < nf * nf * 8; e += blockDim.x) {
    const int c = e & 7, ht = e >> 3, h = ht / nf, t = ht % nf;
    const float *AhF = adHostF + 64 * (size_t)(h + nf * t), *AtF = adTargetF + 64 * (size_t)(h + nf * t);
    float sh = 0.f, st = 0.f;
    for (int k = 0; k < 8; k++) { sh += (float)x[4 + 8 * h + k] * AhF[k * 8 + c]; st += (float)x[4 + 8 * t + k] * AtF[k * 8 + c]; }
    xAd[e] = sh + st;
  }
  if (tid < 4) xAd[(size_t)nf * nf * 8 + tid] = (float)x[tid];
}

}  // namespace

static size_t solve_smem_base(int D, int T) {   // factor storage, As, six D-vectors, panel + Y buffers (all even counts: 16-byte alignment holds)
  return ((size_t)D * (16 * T + 1) + (size_t)(D + 1) * D + 6 * (size_t)D + 6 * (size_t)(16 * T) * 4) * sizeof(double);
}

template <int T>
static int launch_solve_t(sosba *h, SolveArgs &a) {
  // per instantiation AND per device (the attribute belongs to the device's copy of the kernel): opt-in limit minus the
  // kernel's static shared memory
  static size_t dyn_max_dev[64] = {};
  const int di = (h->device >= 0 && h->device < 64) ? h->device : 0;
  if (!dyn_max_dev[di] || di != h->device) {
    int optin = 0;
    cudaFuncAttributes fa;
    SOSBA_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    SOSBA_CUDA(cudaFuncGetAttributes(&fa, k_solve<T>));
    const size_t lim = (size_t)optin - fa.sharedSizeBytes - 1024;
    SOSBA_CUDA(cudaFuncSetAttribute(k_solve<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lim));
    dyn_max_dev[di] = lim;
  }
  const size_t dyn_max = dyn_max_dev[di];
  const int D = a.D;
  // stage accSC (and HM) in shared memory when they fit beside the factor
  size_t smem = solve_smem_base(D, T);
  const size_t bytesS = ((((size_t)D + 1) * (D + 1) + 1) & ~(size_t)1) * 8;
  a.stage_sc = a.stage_hm = 0;
  if (smem + bytesS <= dyn_max) { a.stage_sc = 1; smem += bytesS; }
  if (a.stage_sc && a.HM && smem + (size_t)D * D * 8 <= dyn_max) { a.stage_hm = 1; smem += (size_t)D * D * 8; }
  if (smem > dyn_max) { sosba_set_error("window too large for the single-CTA solve (D=%d)", D); return SOSBA_E_ARG; }
  a.smem_words = (int)(smem / 4);
  SOSBA_CUDA(launch_pdl(k_solve<T>, a.do_th ? 2 : 1, SOLVE_THREADS, smem, h->stream, a));
  SOSBA_CUDA(cudaGetLastError());
  h->launches++;
  return SOSBA_OK;
}

int launch_solve(sosba *h, const SolveArgs &a0) {
  SolveArgs a = a0;
  if ((a.D & 3) || a.D + 1 > 16 * 7) { sosba_set_error("single-CTA solve supports 4 + 8 nf <= 108 (nf <= 13), got D=%d", a.D); return SOSBA_E_ARG; }
  static const bool rowwise = [] { const char *e = getenv("SOSBA_SOLVE_BACKSUB"); return e && e[0] == 'r'; }();
  a.backsub_rowwise = rowwise ? 1 : 0;
  static const bool scalar_pivots = [] { const char *e = getenv("SOSBA_SOLVE_PIVOTS"); return e && e[0] == 's'; }();   // "scalar": column-by-column LDL^T
  a.block_pivots = scalar_pivots ? 0 : 1;
  // measured at D = 68: factorisation 25.2 k cycles pipelined against 24.2 k with the hand-off scheme (the panel warps' own
  // window updates, one warp per scheduler beside two update warps, take the time the hand-off took) -> opt-in
  static const bool want_pipe = [] { const char *e = getenv("SOSBA_SOLVE_PIPE"); return e && e[0] == '1'; }();
  a.pipe = (a.block_pivots && want_pipe) ? 1 : 0;
  int T = (a.D + 1 + 15) / 16;
  if (const char *f = getenv("SOSBA_SOLVE_FORCE_T")) T = std::max(T, atoi(f));   // debug: run a wider instantiation
  if (T <= 3) return launch_solve_t<3>(h, a);
  if (T <= 5) return launch_solve_t<5>(h, a);
  return launch_solve_t<7>(h, a);
}

void launch_make_xad(sosba *h, const double *d_x, int nf, const float *adHostF, const float *adTargetF, float *xAd) {
  k_make_xad<<<1, 256, 0, h->stream>>>(d_x, nf, adHostF, adTargetF, xAd);
  h->launches++;
}

void launch_frame_retarget(sosba *h, const StepArgs &a, const ThArgs *th, int *zero_words, int n_zero) {
  ThArgs t = {};
  if (th) t = *th;
  launch_pdl(k_frame_retarget, th ? 2 : 1, 256, 0, h->stream, a, t, th ? 1 : 0, zero_words, n_zero);
  h->launches++;
}

void launch_step(sosba *h, const ResubArgs &ra, const StepArgs &sa) {
  const int blocks = (ra.P * 8 + 255) / 256;
  launch_pdl(k_step, blocks + 1, 256, 0, h->stream, ra, sa);
  h->launches++;
}
