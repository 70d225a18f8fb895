// pm_kernels_tiled.cuh — the bandwidth path of the pressure solve for large unmasked grids.
//
// One launch ("pass") advances the pressure field by up to T sweeps (temporal blocking) and
// produces the infinity-norm residual of every iterate it passes through, moving p once in and
// once out of HBM and f once in:
//   * the pressure buffers of the solve live in HBM in the split-row layout (pm_common.cuh: even storage
//     columns of a row first, then the odd ones); the (TY+2H) x 128 tile of p around a (TY x TX) output
//     block is staged into shared memory by ONE TMA tensor load (cp.async.bulk.tensor.3d over
//     {pair, parity, row}, zero-filled outside the allocation) signalled on an mbarrier and arrives in
//     the order the sweeps exchange neighbours in -- no rewrite, no barrier before the first half-sweep;
//     f goes straight from HBM into registers with 128-bit row loads meanwhile;
//   * every thread keeps its RPT x 2 cells of p and f in registers for the whole pass; shared
//     memory only carries the values neighbouring threads exchange (one 64-bit load and one
//     64-bit store per cell update), so the sweeps are bounded by the FP64 pipe, not by smem;
//   * halo H = 2T (red-black: one ring per colour half-sweep) or T (Jacobi): cells closer than s
//     rings to the tile edge are stale after s half-sweeps and are never written back;
//   * residuals cost no extra loads: a Jacobi or red update reads exactly the operands of the
//     residual of the iterate it replaces; a black update's operands are the residual operands of
//     the iterate it creates (reference residual trees: cavity-01.cpp:664-673, channel-01.cpp:676-678);
//   * the output block is written in the split-row layout (consecutive lanes to consecutive doubles of either
//     half of the row); wall ghosts (channel form) are refreshed by the thread that owns the wall-adjacent
//     cell, in the tile and in HBM (channel-01.cpp:531-541);
//   * per-iterate residual maxima: per-warp shared slots -> one global atomicMax per iterate and tile into one of
//     32 slot lines -> k_tiled_fold behind the pass.
// The reference's loop test (cavity-01.cpp:635) is evaluated on the device at the start of every
// pass from the residuals of the previous pass; passes after convergence exit at once and leave
// both buffers untouched, and the host re-runs at most one partial pass to land on the exact iterate.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "pm_common.cuh"
#include "pm_tile_cfg.cuh"

// ---- mbarrier / TMA primitives (sm_90+ PTX; SASS: SYNCS.*, UTMALDG) -----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int cx, int cy, int cz) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(cx), "r"(cy), "r"(cz)
      : "memory");
}

// ---- thread-block cluster primitives: distributed shared memory (SASS: mapa -> UMOV/PRMT on the CTA id, st.async -> STAS)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}
// One double into another CTA's shared memory; its arrival is counted (8 bytes) on that CTA's mbarrier.
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
               "l"(__double_as_longlong(v)), "r"(remote_bar)
               : "memory");
}
// relaxed: the arrive orders nothing itself (a release would drain every outstanding load behind a MEMBAR.GPU); the
// barrier initialisation it announces was published by fence.mbarrier_init.release.cluster
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t phase) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(phase)
        : "memory");
  } while (!ok);
}

#ifdef PM_TILE_PROFILE
// Debug build only (make variant EXTRA=-DPM_TILE_PROFILE): cycles per phase summed over the CTAs, thread 0's clock.
__device__ unsigned long long g_tile_prof[8];
#define PM_PROF(slot) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_tile_prof[slot], (unsigned long long)(t_ - prof_t)); prof_t = t_; } } while (0)
#else
#define PM_PROF(slot) do { } while (0)
#endif

// Loop test of the reference for the pass that starts at iterate m0: has any iterate of the
// previous pass (m0-T .. m0-1) met the tolerance?  The words it needs are fetched with T + 2
// independent loads issued at kernel entry (one L2 round trip, overlapped with the tile loads) and
// evaluated once the tile's own loads are in flight.
template <int T>
struct StopWords {
  int done;
  double tol;
  unsigned long long res[T];
};
template <int T>
__device__ __forceinline__ StopWords<T> stop_words_load(const PpeState* st, const unsigned long long* res_bits, int m0) {
  StopWords<T> w;
  w.done = st->done;
  w.tol = st->tol;
#pragma unroll
  for (int q = 0; q < T; ++q) w.res[q] = res_bits[max(m0 - T + q, 0)];
  return w;
}
// Returns true if the pass must not run; *first = the first iterate that met the tolerance (-1: decided by an earlier pass).
template <int T>
__device__ __forceinline__ bool stop_words_eval(const StopWords<T>& w, int m0, int* first) {
  *first = -1;
  if (w.done) return true;
#pragma unroll
  for (int q = 0; q < T; ++q) {
    const int m = m0 - T + q;
    if (m >= 1 && !(__longlong_as_double((long long)w.res[q]) > w.tol)) { *first = m; return true; }
  }
  return false;
}

// INT = the whole tile lies strictly inside the domain: every indicator of the cavity form is 1
// (neighbor_count == 4) and no cell of the channel form touches a wall ghost.
template <class A, int FORM, bool INT>
__device__ __forceinline__ double cell_update(const KP& k, int j, int i, double pc, double pe, double pw, double pn, double ps, double f) {
  if (FORM == 0) {
    if (INT) {  // cavity-01.cpp:651-654 with eps_* == 1 (1*x == x exactly)
      const double s = A::sub(A::add(A::add(pe, pw), A::add(pn, ps)), A::mul(f, k.hh));
      return A::add(A::mul(pc, k.om1), A::mul(k.wnc[4], s));
    }
    return upd_cavity<A>(k, j, i, pc, pe, pw, pn, ps, f);
  }
  return upd_channel<A>(k, pc, pe, pw, pn, ps, f);
}
template <class A, int FORM, bool INT>
__device__ __forceinline__ double cell_residual(const KP& k, int j, int i, double pc, double pe, double pw, double pn, double ps, double f) {
  if (FORM == 0) {
    if (INT) {  // cavity-01.cpp:670-673 with eps_* == 1
      double s = A::add(A::sub(pe, pc), A::sub(pw, pc));
      s = A::add(s, A::sub(pn, pc));
      s = A::add(s, A::sub(ps, pc));
      return A::sub(A::mul(k.idx2, s), f);
    }
    return res_cavity<A>(k, j, i, pc, pe, pw, pn, ps, f, k.idx2);
  }
  return res_channel<A>(k, pc, pe, pw, pn, ps, f);
}

// Per-thread view of the tile: RPT rows x 2 columns of p and f in registers for the whole pass.
template <int RPT>
struct Cells {
  double p0[RPT], p1[RPT];
  double f0[RPT], f1[RPT];
  __device__ __forceinline__ double f0v(int r) const { return f0[r]; }
  __device__ __forceinline__ double f1v(int r) const { return f1[r]; }
};

// m = max(m, |x|) where `on`, with the semantics of std::max(m, std::abs(x)) (a NaN never replaces m).
// fabs() folds into the compare and the select as an operand modifier: DSETP + FSEL + SEL per call.
__device__ __forceinline__ void acc_max(double& m, double x, bool on) {
  if (on && fabs(x) > m) m = fabs(x);
}

// Warp maximum of non-negative, non-NaN doubles (their bit patterns order like unsigned integers): two
// REDUX.MAX over the high and the low word instead of five shuffle rounds of 64-bit compares.
__device__ __forceinline__ double warp_max_nonneg(double v) {
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  return __hiloint2double((int)mh, (int)ml);
}

// Production arithmetic, interior tiles: the residual in sum form.  FORM 0 takes the four-neighbour sum s,
//   r = h^-2 * (s - 4 p) - f;  FORM 1 the two pair sums,  r = idx2*(pE+pW) + idy2*(pN+pS) - 2(idx2+idy2) p - f.
// The same value as the reference's residual trees (cavity-01.cpp:670-673, channel-01.cpp:676-678) up to
// rounding; the reference's own update tree sums the neighbours the same way (cavity-01.cpp:651-654).
template <int FORM>
__device__ __forceinline__ double res_sum4(const KP& k, double pc, double s, double f) {
  static_assert(FORM == 0, "four-neighbour sum: isotropic form only");
  return fma(k.idx2, fma(-4.0, pc, s), -f);
}
__device__ __forceinline__ double res_sum22(const KP& k, double pc, double sew, double sns, double f) {
  return fma(-k.denom, pc, fma(k.idx2, sew, fma(k.idy2, sns, -f)));
}

// Edge-row exchange between the CTAs of a cluster (stacked in y).  After colour half-sweep h every CTA pushes the
// cells of that colour of its first row into the row above the tile of the CTA below (tile row SH there) and of its
// last row into the row below the tile of the CTA above (tile row -1 there), one st.async per thread and direction;
// the 64 x 8 bytes of a half-sweep complete one phase of an mbarrier in the receiving CTA, which the threads that
// read that row wait on before half-sweep h + 1.  Two barriers per direction, alternating with h, so a phase can never
// receive bytes of the next one: the data of half-sweep h + 2 depends on what this CTA sends after it has waited
// for half-sweep h.  For the same reason a row is never overwritten before its last reader is done.
// The last half-sweep of a pass sends nothing (nobody reads it), so every byte sent is waited for before the
// receiving CTA can finish: no cluster barrier at the end.
// The remote addresses are re-derived from the thread's tile pointer at every use, so the exchange keeps only two flags
// per thread across the sweeps.
struct Xchg {
  uint64_t* bars;  // this CTA's four mbarriers: [0..1] count bytes arriving from below, [2..3] from above
  int nhs;         // half-sweeps of this pass
  bool dn, up;     // this thread holds cells of the tile's first row and the CTA below exchanges / last row and the CTA above
};
#define PM_XCHG_BYTES 512u  // 64 cells of one colour per row
// After half-sweep h in which the .x cell of the thread's first row was the target iff x_first: push the edge cells.
// tpx = the thread's .x cell of its first row inside the tile; segment 0 holds the tile's first rows, segment NSEG-1 its last.
template <class C>
__device__ __forceinline__ void xchg_send(const Xchg& x, const double* tpx, const Cells<C::RPT>& c, bool x_first, int h) {
  constexpr int SW = C::SW, SH = C::SH, RPT = C::RPT;
  if (h >= x.nhs - 1) return;  // nothing follows that would read it
  const uint32_t b = uint32_t(h & 1) * 8u;
  if (x.dn) {  // my first row -> tile row SH of the CTA below; counted on its "from above" pair
    const uint32_t rank = cluster_ctarank() - 1u;
    const uint32_t row = mapa_u32(smem_u32(tpx + SH * SW), rank), rbar = mapa_u32(smem_u32(x.bars + 2), rank);
    st_async_f64(row + (x_first ? 0u : uint32_t(SW / 2) * 8u), x_first ? c.p0[0] : c.p1[0], rbar + b);
  }
  if (x.up) {  // my last row -> tile row -1 of the CTA above; counted on its "from below" pair
    const uint32_t rank = cluster_ctarank() + 1u;
    const uint32_t row = mapa_u32(smem_u32(tpx - ((C::NSEG - 1) * RPT + 1) * SW), rank), rbar = mapa_u32(smem_u32(x.bars), rank);
    // RPT is even: in the last row the roles of .x and .y are swapped
    st_async_f64(row + (x_first ? uint32_t(SW / 2) * 8u : 0u), x_first ? c.p1[RPT - 1] : c.p0[RPT - 1], rbar + b);
  }
}
// Once per pass, by every thread, before the first push: all CTAs of the cluster are ready to receive.
__device__ __forceinline__ void xchg_gate(int h) {
  if (h == 0) cluster_wait_acquire();
}
// Before the edge rows of half-sweep h + 1 read them: the edge cells the neighbours pushed after their half-sweep h.
// (Tried: waits without a "memory" clobber, the guarded loads tied to a token the wait returns -- no faster, and ptxas
// 12.9 lost the token's value on the path where the first try succeeds.)
template <class C>
__device__ __forceinline__ void xchg_wait(const Xchg& x, int h) {
  if (h < 0 || !(x.dn || x.up)) return;
  const bool lead = (threadIdx.x & 63) == 0;
#ifdef PM_TILE_PROFILE
  const long long w0 = clock64();
#endif
  const bool again = h + 2 < x.nhs - 1;  // the same barrier serves half-sweep h + 2 if that one sends
  if (x.dn) {
    const uint32_t bar = smem_u32(x.bars + (h & 1));
    mbar_wait_u32(bar, uint32_t(h >> 1) & 1u);
    if (lead && again) mbar_expect_tx_u32(bar, PM_XCHG_BYTES);
  }
  if (x.up) {
    const uint32_t bar = smem_u32(x.bars + 2 + (h & 1));
    mbar_wait_u32(bar, uint32_t(h >> 1) & 1u);
    if (lead && again) mbar_expect_tx_u32(bar, PM_XCHG_BYTES);
  }
#ifdef PM_TILE_PROFILE
  if (lead) atomicAdd(&g_tile_prof[6], (unsigned long long)(clock64() - w0));  // summed over both edge segments
#endif
}

// Shared-memory layout of the exchange tile ("split rows"): within each 128-double row the 64 even columns
// come first, then the 64 odd ones: column c lives at (c & 1) * 64 + (c >> 1).  Lane q of a warp owns columns
// 2q (.x) and 2q+1 (.y); its west neighbour 2q-1 and east neighbour 2q+2 then sit at consecutive doubles
// across the lanes, so every 64-bit access of a sweep is bank-conflict free (the natural layout, stride
// 16 B across lanes, costs two wavefronts where one suffices).  `tpx` / `tpy` point at the thread's .x / .y cell
// of its first row.  The tile has one spare row above and below, so the never-used neighbour reads of ring
// cells stay in bounds and every shared access is tpx/tpy + constant.
//
// One colour half-sweep (red-black) over the thread's rows.
// PX: the target of row r is the .x cell iff (r & 1) == PX.  PRE: accumulate the residual of the iterate
// being replaced (operands before the update); POST: of the iterate being created (operands after).
// mOut: bit r / bit 16+r = the .x / .y cell of row r belongs to the output block (residual is taken there).
//
// INT tiles update EVERY cell they hold, ring included: a ring cell only ever feeds cells that are already
// stale for the same half-sweep count (see the header), so the predicate would buy nothing; boundary
// tiles keep it because ghost cells and cells beyond the domain must keep their values.
template <class A, int FORM, bool INT, class C, int PX, bool PRE, bool POST>
__device__ __forceinline__ void rb_half(const KP& k, double* tpx, double* tpy, Cells<C::RPT>& c, int i0, int jg0, unsigned mW, unsigned mOut,
                                        bool colW0, bool colW1, bool commit, double& rmax_pre, double& rmax_post) {
  constexpr int SW = C::SW, RPT = C::RPT;
  const unsigned mOx = mOut & 0xffffu, mOy = mOut >> 16;  // output-block rows of the .x / .y column
  double post_raw = 0.0;  // max |r| before the update over this half-sweep's cells (production interior tiles)
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const bool tx = (r & 1) == PX;  // compile-time after unrolling
    const int j = jg0 + r;
    const int i = tx ? i0 : i0 + 1;
    double* cell = (tx ? tpx : tpy) + r * SW;     // the target
    double* side = (tx ? tpy - 1 : tpx + 1) + r * SW;  // its outer horizontal neighbour: column 2q-1 resp. 2q+2
    const double pc = tx ? c.p0[r] : c.p1[r];
    const double fc = tx ? c.f0v(r) : c.f1v(r);
    double pw, pe, pn, ps;
    if (tx) { pw = side[0]; pe = c.p1[r]; }
    else { pw = c.p0[r]; pe = side[0]; }
    if (r + 1 < RPT) pn = tx ? c.p0[r + 1] : c.p1[r + 1];
    else pn = cell[SW];
    if (r >= 1) ps = tx ? c.p0[r - 1] : c.p1[r - 1];
    else ps = cell[-SW];
    const bool out = ((tx ? mOx : mOy) >> r) & 1u;
    if (INT && !A::exact) {
      // Production arithmetic, interior tile: the relaxation in residual form.  With all indicators 1,
      //   p*(1-w) + (w/4)*(S - f*h*h) == p + (w*h*h/4) * r,   r = h^-2 * sum(p_nb - p) - f   (the reference's residual),
      // so one FMA on the residual replaces the update tree; and because a colour-1 cell's neighbours do not
      // change during its half-sweep, its residual after the update is exactly (1-w) * r.  9 FP64 ops per cell
      // instead of 14; the iterates differ from the reference tree by rounding only (exact_arith=1 keeps the tree).
      const double rs = cell_residual<A, FORM, true>(k, j, i, pc, pe, pw, pn, ps, fc);
      const double nvf = fma(k.cw, rs, pc);
      if (tx) c.p0[r] = nvf; else c.p1[r] = nvf;
      cell[0] = nvf;
      acc_max(PRE ? rmax_pre : post_raw, rs, out);
      continue;
    }
    if (FORM == 0 && !INT && !A::exact) {
      // Production arithmetic, cavity tiles at a wall: the same residual form.  A neighbour behind a wall (eps == 0,
      // cavity-01.cpp:644-648) is replaced by the cell itself, so that  sum - 4 p  ==  sum over the real neighbours - nc p,
      // and the relaxation factor is omega h^2 / nc; the residual after a colour-1 update is (1 - omega) r for every nc.
      // One tree per cell instead of three (residual before, update, residual after): wall tiles were 3x slower than
      // interior ones, and since the streaming pass they sit on the critical path of every pass.
      const int ew = i > 1, ee = i < k.nx, en = j < k.ny;
      const double pw_ = ew ? pw : pc, pe_ = ee ? pe : pc, pn_ = en ? pn : pc;
      const double rs = fma(k.idx2, fma(-4.0, pc, (pe_ + pn_) + (pw_ + ps)), -fc);
      const int nc = ew + ee + en;
      const double cwc = nc == 3 ? k.cw : (nc == 2 ? k.cw3 : k.cw2);
      const bool upd = commit && ((mW >> r) & 1u) && (tx ? colW0 : colW1);
      if (upd) {
        const double nvf = fma(cwc, rs, pc);
        if (tx) c.p0[r] = nvf; else c.p1[r] = nvf;
        cell[0] = nvf;
      }
      if (PRE) acc_max(rmax_pre, rs, out);
      else acc_max(post_raw, rs, out && upd);
      continue;
    }
    if (PRE) acc_max(rmax_pre, cell_residual<A, FORM, INT>(k, j, i, pc, pe, pw, pn, ps, fc), out);
    const double nv = cell_update<A, FORM, INT>(k, j, i, pc, pe, pw, pn, ps, fc);
    if (INT) {
      if (tx) c.p0[r] = nv; else c.p1[r] = nv;
      cell[0] = nv;
      if (POST) acc_max(rmax_post, cell_residual<A, FORM, INT>(k, j, i, nv, pe, pw, pn, ps, fc), out);
    } else {
      const bool upd = commit && ((mW >> r) & 1u) && (tx ? colW0 : colW1);
      if (upd) {
        if (tx) c.p0[r] = nv; else c.p1[r] = nv;
        cell[0] = nv;
        if (FORM == 1) {  // refresh the wall ghosts this cell owns (channel-01.cpp:531-541)
          if (i == 1) { pw = nv; side[0] = nv; }  // i == 1 is always a .x cell: its west ghost is the outer neighbour
          if (i == k.nx) {
            pe = 0.0;
            if (tx) { c.p1[r] = 0.0; tpy[r * SW] = 0.0; }  // the east ghost is the .y cell of this pair
            else side[0] = 0.0;
          }
          if (j == 1) {
            ps = nv;
            cell[-SW] = nv;
            if (r >= 1) { if (tx) c.p0[r - 1] = nv; else c.p1[r - 1] = nv; }
          }
          if (j == k.ny) {
            pn = nv;
            cell[SW] = nv;
            if (r + 1 < RPT) { if (tx) c.p0[r + 1] = nv; else c.p1[r + 1] = nv; }
          }
        }
        if (POST) acc_max(rmax_post, cell_residual<A, FORM, INT>(k, j, i, nv, pe, pw, pn, ps, fc), out);
      }
    }
  }
  if ((INT || FORM == 0) && !A::exact && POST) {
    const double m = fabs(k.om1) * post_raw;  // residual of the iterate just created: (1 - omega) * r
    if (m > rmax_post) rmax_post = m;
  }
}

// The same half-sweep for interior tiles with production arithmetic, written for a short instruction
// stream: the relaxation in residual form, p += cw * r with r the residual (so the norm costs no extra
// flops), neighbours summed instead of differenced, and -- in the isotropic (cavity) form -- the diagonal
// pair sum p(r, .y) + p(r+1, .x) resp. p(r, .x) + p(r+1, .y) shared by the two same-colour cells of a row
// pair that both have it as two of their four neighbours: 5.5 FP64 operations per cell and iterate.
// A colour-1 cell's neighbours do not change during its half-sweep, so its residual after the update is
// exactly (1 - omega) * r: POST takes max |r| before the update and scales it once.
// The iterates differ from the reference trees by rounding only (exact_arith = 1 keeps the trees).
template <int FORM, class C, int PX, bool PRE>
__device__ __forceinline__ void rb_half_lean(const KP& k, double* tpx, double* tpy, Cells<C::RPT>& c, unsigned mOut,
                                             double& rmax_pre, double& rmax_post, const Xchg& x, int h) {
  constexpr int SW = C::SW, RPT = C::RPT;
  static_assert(RPT % 2 == 0, "rows are processed in pairs");
  double mx = PRE ? rmax_pre : 0.0;
  auto row_pair = [&](const int a) {
    const int b = a + 1;
    // row a: target tA (.x iff PX == 0); row b: the other column
    double rA, rB;
    if (PX == 0) {
      const double wA = tpy[a * SW - 1];                                   // column 2q-1
      const double sA = a >= 1 ? c.p0[a - 1] : tpx[(a - 1) * SW];
      const double eB = tpx[b * SW + 1];                                   // column 2q+2
      const double nB = b + 1 < RPT ? c.p1[b + 1] : tpy[(b + 1) * SW];
      if (FORM == 0) {
        const double d = c.p1[a] + c.p0[b];  // east + north of A == south + west of B
        rA = res_sum4<0>(k, c.p0[a], d + (wA + sA), c.f0v(a));
        rB = res_sum4<0>(k, c.p1[b], d + (eB + nB), c.f1v(b));
      } else {
        rA = res_sum22(k, c.p0[a], wA + c.p1[a], c.p0[b] + sA, c.f0v(a));
        rB = res_sum22(k, c.p1[b], c.p0[b] + eB, nB + c.p1[a], c.f1v(b));
      }
      c.p0[a] = fma(k.cw, rA, c.p0[a]);
      c.p1[b] = fma(k.cw, rB, c.p1[b]);
      tpx[a * SW] = c.p0[a];
      tpy[b * SW] = c.p1[b];
      acc_max(mx, rA, (mOut >> a) & 1u);
      acc_max(mx, rB, (mOut >> (16 + b)) & 1u);
    } else {
      const double eA = tpx[a * SW + 1];
      const double sA = a >= 1 ? c.p1[a - 1] : tpy[(a - 1) * SW];
      const double wB = tpy[b * SW - 1];
      const double nB = b + 1 < RPT ? c.p0[b + 1] : tpx[(b + 1) * SW];
      if (FORM == 0) {
        const double d = c.p0[a] + c.p1[b];  // west + north of A == south + east of B
        rA = res_sum4<0>(k, c.p1[a], d + (eA + sA), c.f1v(a));
        rB = res_sum4<0>(k, c.p0[b], d + (wB + nB), c.f0v(b));
      } else {
        rA = res_sum22(k, c.p1[a], c.p0[a] + eA, c.p1[b] + sA, c.f1v(a));
        rB = res_sum22(k, c.p0[b], wB + c.p1[b], nB + c.p0[a], c.f0v(b));
      }
      c.p1[a] = fma(k.cw, rA, c.p1[a]);
      c.p0[b] = fma(k.cw, rB, c.p0[b]);
      tpy[a * SW] = c.p1[a];
      tpx[b * SW] = c.p0[b];
      acc_max(mx, rA, (mOut >> (16 + a)) & 1u);
      acc_max(mx, rB, (mOut >> b) & 1u);
    }
  };
  if (C::CS > 1) {
    // The thread's first and last row LAST: what they read from the neighbour CTAs was pushed at the end of the previous
    // half-sweep and has had this half-sweep's inner rows to arrive in; their new values leave right away and have the
    // neighbours' inner rows of the next half-sweep to travel.
#pragma unroll
    for (int a = 2; a < RPT - 2; a += 2) row_pair(a);
    xchg_wait<C>(x, h - 1);
    row_pair(0);
    if (RPT > 2) row_pair(RPT - 2);
    xchg_gate(h);
    xchg_send<C>(x, tpx, c, PX == 0, h);
  } else {
#pragma unroll
    for (int a = 0; a < RPT; a += 2) row_pair(a);
  }
  if (PRE) rmax_pre = mx;
  else {
    const double m = fabs(k.om1) * mx;  // residual of the iterate just created
    if (m > rmax_post) rmax_post = m;
  }
}

// One Jacobi sweep from the shared-memory tile at (sx, sy) into the one at (dx, dy): neighbours in other threads are
// read from the source tile, every thread writes all of its cells (new, or unchanged) to the destination tile, so one
// barrier per sweep separates the iterates; the thread's own column is walked upwards with the old value of the row
// below carried in two registers (no staging of a whole sweep in registers: that spilled).
template <class A, int FORM, bool INT, class C>
__device__ __forceinline__ void jacobi_sweep(const KP& k, const double* sx, const double* sy, double* dx, double* dy, Cells<C::RPT>& c, int i0, int jg0,
                                             unsigned mW, unsigned mOut, bool colW0, bool colW1, bool commit, double& rmax_pre) {
  constexpr int SW = C::SW, RPT = C::RPT;
  double o0 = 0.0, o1 = 0.0;  // previous iterate of the row below
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int j = jg0 + r;
    const double c0 = c.p0[r], c1 = c.p1[r];
    const double wl = sy[r * SW - 1];
    const double er = sx[r * SW + 1];
    double pn0, pn1, ps0, ps1;
    if (r + 1 < RPT) { pn0 = c.p0[r + 1]; pn1 = c.p1[r + 1]; }
    else { pn0 = sx[(r + 1) * SW]; pn1 = sy[(r + 1) * SW]; }
    if (r >= 1) { ps0 = o0; ps1 = o1; }
    else { ps0 = sx[(r - 1) * SW]; ps1 = sy[(r - 1) * SW]; }
    const double r0v = cell_residual<A, FORM, INT>(k, j, i0, c0, c1, wl, pn0, ps0, c.f0v(r));
    const double r1v = cell_residual<A, FORM, INT>(k, j, i0 + 1, c1, er, c0, pn1, ps1, c.f1v(r));
    acc_max(rmax_pre, r0v, (mOut >> r) & 1u);
    acc_max(rmax_pre, r1v, (mOut >> (16 + r)) & 1u);
    o0 = c0; o1 = c1;
    if (!commit) continue;  // the residual-only pass (uniform across the block)
    double n0, n1;
    if (INT && !A::exact) {  // residual form of the relaxation, see rb_half
      n0 = fma(k.cw, r0v, c0);
      n1 = fma(k.cw, r1v, c1);
    } else {
      n0 = cell_update<A, FORM, INT>(k, j, i0, c0, c1, wl, pn0, ps0, c.f0v(r));
      n1 = cell_update<A, FORM, INT>(k, j, i0 + 1, c1, er, c0, pn1, ps1, c.f1v(r));
      if (!INT) {
        const bool rowW = (mW >> r) & 1u;
        if (!(rowW && colW0)) n0 = c0;
        if (!(rowW && colW1)) n1 = c1;
      }
    }
    c.p0[r] = n0; c.p1[r] = n1;
    dx[r * SW] = n0;
    dy[r * SW] = n1;
  }
  if (!commit) return;
  __syncthreads();  // the destination tile holds the new iterate (wall ghosts still the old ones)
  if (FORM == 1 && !INT) {
    // wall ghosts from the new values (channel-01.cpp:531-541), written by the thread that owns the wall-adjacent cell;
    // then every thread takes its cells back from the tile (a ghost may be one of them)
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int j = jg0 + r;
      const bool rowW = (mW >> r) & 1u;
      if (!rowW) continue;
      if (colW0 && i0 == 1) dy[r * SW - 1] = c.p0[r];
      if (colW0 && i0 == k.nx) dy[r * SW] = 0.0;
      if (colW1 && i0 + 1 == k.nx) dx[r * SW + 1] = 0.0;
      if (j == 1) { if (colW0) dx[(r - 1) * SW] = c.p0[r]; if (colW1) dy[(r - 1) * SW] = c.p1[r]; }
      if (j == k.ny) { if (colW0) dx[(r + 1) * SW] = c.p0[r]; if (colW1) dy[(r + 1) * SW] = c.p1[r]; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      c.p0[r] = dx[r * SW];
      c.p1[r] = dy[r * SW];
    }
  }
}

// All sweeps of one pass for one thread.  The sweep loop is NOT unrolled (the instruction footprint of
// the hot path stays within the instruction cache); per-iterate residual maxima therefore go to shared
// memory at the end of every sweep: two REDUX per warp, then a plain store into the warp's own slot
// red[warp * (T + 1) + t] (a 64-bit shared atomicMax is a compare-and-swap loop).
// PAR0 = colour of the .x cell of the thread's first row = (j0 & 1): i0 is always odd and TY, RPT are even,
// so it is the same for every thread of every tile of a launch.
template <class A, int FORM, int METHOD, int T, bool INT, int PAR0>
__device__ __forceinline__ void run_sweeps(const KP& k, double* tpx, double* tpy, Cells<TileCfg<METHOD, T>::RPT>& c, int i0, int jg0,
                                           unsigned mW, unsigned mOut, bool colW0, bool colW1, int nsw,
                                           unsigned long long* __restrict__ red, const Xchg& x) {
  using C = TileCfg<METHOD, T>;
  constexpr bool XC = C::CS > 1;
  const int lane = threadIdx.x & 31;
  unsigned long long* redw = red + (threadIdx.x >> 5) * (T + 1);
  double r_cur = 0.0, r_next = 0.0;  // residual maxima of iterate m0+t and m0+t+1
  const int nloop = nsw > 0 ? nsw : 1;
  if (XC && nsw == 0) cluster_wait_acquire();  // the residual-only pass pushes nothing; complete the barrier all the same
#pragma unroll 1
  for (int t = 0; t < nloop; ++t) {
    const bool commit = t < nsw;
    if (METHOD == PM_PPE_SOR_RB) {
      // colour 0 first ((i + j) even), as the oracle's red-black restatement
      if (INT && !A::exact) {  // interior tiles are only run with nsw > 0: commit is always true here
        rb_half_lean<FORM, C, PAR0, true>(k, tpx, tpy, c, mOut, r_cur, r_next, x, 2 * t);
        __syncthreads();
        rb_half_lean<FORM, C, 1 - PAR0, false>(k, tpx, tpy, c, mOut, r_cur, r_next, x, 2 * t + 1);
        __syncthreads();
      } else {
        if (XC) xchg_wait<C>(x, 2 * t - 1);
        rb_half<A, FORM, INT, C, PAR0, true, false>(k, tpx, tpy, c, i0, jg0, mW, mOut, colW0, colW1, commit, r_cur, r_next);
        if (commit) {
          if (XC) { xchg_gate(2 * t); xchg_send<C>(x, tpx, c, PAR0 == 0, 2 * t); }
          __syncthreads();
          if (XC) xchg_wait<C>(x, 2 * t);
          rb_half<A, FORM, INT, C, 1 - PAR0, false, true>(k, tpx, tpy, c, i0, jg0, mW, mOut, colW0, colW1, true, r_cur, r_next);
          if (XC) xchg_send<C>(x, tpx, c, PAR0 == 1, 2 * t + 1);
          __syncthreads();
        }
      }
    } else {
      const int so = (t & 1) * C::TILE_DOUBLES, dof = C::TILE_DOUBLES - so;  // the two tiles take turns
      jacobi_sweep<A, FORM, INT, C>(k, tpx + so, tpy + so, tpx + dof, tpy + dof, c, i0, jg0, mW, mOut, colW0, colW1, commit, r_cur);
    }
    const double v = warp_max_nonneg(r_cur);
    if (lane == 0) redw[t] = (unsigned long long)__double_as_longlong(v);
    r_cur = r_next;
    r_next = 0.0;
  }
  // colour-1 part of the last iterate created (red-black); zero for Jacobi
  const double v = warp_max_nonneg(r_cur);
  if (lane == 0) redw[nloop] = (unsigned long long)__double_as_longlong(v);
}

// Where the residual maxima of a pass are collected (see the end of tile_process): PM_FOLD_SLOTS lines of 16
// words, a ring of per-iterate maxima per slot; slot = linear CTA index % slots.
#define PM_FOLD_SLOTS 32
// One warp, launched behind the launches of a pass: res_bits[m] = max(res_bits[m], max over slots) for the
// iterates lo..hi the pass has completed; their ring entries are cleared for reuse.
__global__ void k_tiled_fold(unsigned long long* __restrict__ part, unsigned long long* __restrict__ res_bits, int lo, int hi) {
  static_assert(PM_FOLD_SLOTS == 32, "one lane per slot");
  const int lane = threadIdx.x;
  for (int m = lo; m <= hi; ++m) {
    unsigned long long* e = part + size_t(lane) * 16 + (m & 7);
    const unsigned long long v = *e;
    *e = 0ull;
    const unsigned hi32 = unsigned(v >> 32), mh = __reduce_max_sync(0xffffffffu, hi32);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi32 == mh ? unsigned(v) : 0u);
    const unsigned long long vm = (unsigned long long)mh << 32 | ml;
    if (lane == 0 && vm > res_bits[m]) res_bits[m] = vm;
  }
}

// ---------------------------------------------------------------------------
// Tiles that contain solid cells (backwards-step obstacle mask, backwards_step-01.cpp:893-939).  They are few -- the rim
// of the obstacle -- so this path is written for clarity, not speed: p stays in the shared-memory tile, every thread
// walks its RPT x 2 cells, the mask is read from global memory.  Tiles without a single fluid cell are copied through.
//
// Order of one reference iteration: sweep over the fluid cells, then applyPressureGhosts -- wall ghosts, then every
// solid cell <- mean of its fluid neighbours (:709-739) -- then the residual.  Here the solid-cell pass of iterate m is
// done at the START of the sweep that turns iterate m into m + 1 (nothing reads a solid cell in between), so the
// iterates stored between passes carry solid cells that lag by one such pass; pm_ppe_solve applies the last one (and
// the wall ghosts behind solid cells, which only the stored field cares about) when the solve ends.  With that
// placement the solid-cell pass depends on the same ring of neighbours as a colour half-sweep: solid cells of colour 0
// read colour-1 fluid cells (final since the last half-sweep), solid cells of colour 1 read colour-0 fluid cells.
// The lag costs one ring of the halo, though: a solid cell on the very edge of a tile arrives one ghost pass behind and
// cannot be brought up to date (a neighbour is missing), so the stale front starts one ring further in.  A pass over a
// masked problem therefore runs T - 1 sweeps on the geometry built for T (TiledPlan::run).  The residual of iterate m0 + t is taken in full, with the reference's tree, right after that
// pass; its colour-1 part at the end of a pass would need solid ghosts one ring further out than are valid.
// ---------------------------------------------------------------------------
struct MaskedGeom {
  int ib, jb;         // tile origin (i, jl)
  int jI_lo, jI_hi;   // jl range this rank has data for
  int jO_lo, jO_hi;   // jl range of the output block (clipped to this rank's rows)
  int iO_lo, iO_hi;   // i range of the output block
};
template <class A, int METHOD, int T>
__device__ __noinline__ void masked_tile(const KP& k, double* __restrict__ tile, const double* __restrict__ f, const uint8_t* __restrict__ M,
                                         double* __restrict__ pout, unsigned long long* __restrict__ red, const MaskedGeom g, int m0, int nsw, int cls) {
  using C = TileCfg<METHOD, T>;
  constexpr int SW = C::SW, SH = C::SH, RPT = C::RPT;
  static_assert(C::CS == 1 && SH * 4 * 4 <= SW * 8, "the fluid bits of the tile live in the spare row below it");
  const int tid = threadIdx.x, lane = tid & 31;
  const int q = tid & 63, rr0 = (tid >> 6) * RPT;
  auto sm = [&](int rt, int ct) -> double& { return tile[rt * SW + (ct & 1) * (SW / 2) + (ct >> 1)]; };  // split rows, see above
  unsigned long long* redw = red + (tid >> 5) * (T + 1);
  const int nloop = nsw > 0 ? nsw : 1;
  if (cls == 2) {
    // Fluid flags of the whole tile as bits, in the order of the split rows: word rt*4 + (ct&1)*2 + (ct>>6), bit (ct>>1)&31
    // -- one ballot per row and column parity.  A flag is set for interior cells this rank has data for that are fluid.
    uint32_t* fbits = reinterpret_cast<uint32_t*>(tile - SW);
    auto fluid = [&](int rt, int ct) -> bool { return (fbits[rt * 4 + (ct & 1) * 2 + (ct >> 6)] >> ((ct >> 1) & 31)) & 1u; };
    double fr[RPT][2];       // f of the thread's cells
    unsigned own_in = 0u, own_fl = 0u;  // bit 2r+e: the cell is an interior cell with data / and fluid
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int rt = rr0 + r, jl = g.jb + rt;
      const bool rowI = jl >= g.jI_lo && jl <= g.jI_hi;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = g.ib + 2 * q + e;
        const bool in = rowI && i >= 1 && i <= k.nx;
        const size_t cm = pm_idx(k, in ? jl : g.jI_lo, in ? i : 1);
        const bool fl = in && M[cm];
        fr[r][e] = fl ? __ldg(f + cm) : 0.0;
        const unsigned w = __ballot_sync(0xffffffffu, fl);
        if (lane == 0) fbits[rt * 4 + e * 2 + (q >> 5)] = w;
        own_in |= unsigned(in) << (2 * r + e);
        own_fl |= unsigned(fl) << (2 * r + e);
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int t = 0; t < nloop; ++t) {
      // ---- solid cells <- mean of their fluid neighbours (not before the first sweep of a solve: the reference's
      // first sweep reads whatever the solid cells hold) ----
      if (m0 + t > 0 && own_in != own_fl) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          const int rt = rr0 + r;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (!((own_in & ~own_fl) >> (2 * r + e) & 1u)) continue;
            const int ct = 2 * q + e;
            double sum = 0.0;
            int n = 0;  // same order as backwards_step-01.cpp:716-735: west, east, south, north (is_fluid is false outside the interior)
            if (ct >= 1 && fluid(rt, ct - 1)) { sum = __dadd_rn(sum, sm(rt, ct - 1)); ++n; }
            if (ct <= SW - 2 && fluid(rt, ct + 1)) { sum = __dadd_rn(sum, sm(rt, ct + 1)); ++n; }
            if (rt >= 1 && fluid(rt - 1, ct)) { sum = __dadd_rn(sum, sm(rt - 1, ct)); ++n; }
            if (rt <= SH - 2 && fluid(rt + 1, ct)) { sum = __dadd_rn(sum, sm(rt + 1, ct)); ++n; }
            if (n > 0) sm(rt, ct) = __ddiv_rn(sum, double(n));
          }
        }
      }
      __syncthreads();
      // ---- residual of iterate m0 + t over the fluid cells of the output block (channel-01.cpp:676-678) ----
      double rmax = 0.0;
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const int rt = rr0 + r, jl = g.jb + rt;
        if (jl < g.jO_lo || jl > g.jO_hi) continue;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ct = 2 * q + e, i = g.ib + ct;
          if (i < g.iO_lo || i > g.iO_hi || !((own_fl >> (2 * r + e)) & 1u)) continue;
          const double rs = res_channel<A>(k, sm(rt, ct), sm(rt, ct + 1), sm(rt, ct - 1), sm(rt + 1, ct), sm(rt - 1, ct), fr[r][e]);
          rmax = fmax(rmax, fabs(rs));
        }
      }
      {
        const double v = warp_max_nonneg(rmax);
        if (lane == 0) redw[t] = (unsigned long long)__double_as_longlong(v);
      }
      if (t >= nsw) break;  // the residual-only pass
      __syncthreads();      // every residual has read its neighbours before the first half-sweep overwrites them
      // ---- the two colour half-sweeps over the fluid cells; wall ghosts refreshed by the cell that owns them ----
#pragma unroll 1
      for (int colour = 0; colour < 2; ++colour) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          const int rt = rr0 + r, jl = g.jb + rt, j = k.j0 + jl;
          if (rt < 1 || rt > SH - 2) continue;
          const int e = (colour + j + g.ib) & 1;  // the cell of the pair (2q, 2q+1) with (i + j) % 2 == colour
          const int ct = 2 * q + e, i = g.ib + ct;
          if (ct < 1 || ct > SW - 2 || !((own_fl >> (2 * r + e)) & 1u)) continue;
          const double nv = upd_channel<A>(k, sm(rt, ct), sm(rt, ct + 1), sm(rt, ct - 1), sm(rt + 1, ct), sm(rt - 1, ct), e ? fr[r][1] : fr[r][0]);
          sm(rt, ct) = nv;
          if (i == 1) sm(rt, ct - 1) = nv;      // channel-01.cpp:531-541
          if (i == k.nx) sm(rt, ct + 1) = 0.0;
          if (j == 1) sm(rt - 1, ct) = nv;
          if (j == k.ny) sm(rt + 1, ct) = nv;
        }
        __syncthreads();
      }
    }
  }
  // ---- write the output block and the wall ghosts its cells own (a tile without fluid cells is copied through) ----
  if (nsw > 0) {
#pragma unroll 1
    for (int r = 0; r < RPT; ++r) {
      const int rt = rr0 + r, jl = g.jb + rt, j = k.j0 + jl;
      if (jl < g.jO_lo || jl > g.jO_hi) continue;
#pragma unroll 1
      for (int e = 0; e < 2; ++e) {
        const int ct = 2 * q + e, i = g.ib + ct;
        if (i < g.iO_lo || i > g.iO_hi) continue;
        const double v = sm(rt, ct);
        pout[pm_sidx(k, jl, i)] = v;
        if (i == 1) pout[pm_sidx(k, jl, 0)] = sm(rt, ct - 1);
        if (i == k.nx) pout[pm_sidx(k, jl, k.nx + 1)] = sm(rt, ct + 1);
        if (j == 1) pout[pm_sidx(k, jl - 1, i)] = sm(rt - 1, ct);
        if (j == k.ny) pout[pm_sidx(k, jl + 1, i)] = sm(rt + 1, ct);
      }
    }
  }
}

// Everything one CTA does for one tile once its TMA load has been issued on `bar`: masks, f loads, wait,
// the sweeps, the write-out and the per-iterate residual atomics.
template <class A, int FORM, int METHOD, int T, int PAR0>
__device__ __forceinline__ void tile_process(const KP& k, double* tile, uint64_t* bar, uint32_t phase, unsigned long long* red,
                                             double* __restrict__ pout, const double* __restrict__ f, PpeState* __restrict__ st,
                                             unsigned long long* __restrict__ res_bits, unsigned long long* __restrict__ fold_part, int m0, int nsw, int bx, int by,
                                             int crank, uint64_t* xbar, const int* xact, const uint8_t* __restrict__ mask, int tclass,
                                             const StopWords<T>& stopw, bool check_stop) {
  using C = TileCfg<METHOD, T>;
  constexpr int H = C::H, SW = C::SW, SH = C::SH, RPT = C::RPT, TX = C::TX, TY = C::TY, CS = C::CS;
  const int tid = threadIdx.x;
#ifdef PM_TILE_PROFILE
  long long prof_t = clock64();
#endif
  // (bx, by) = the cluster's output block; CTA `crank` of the cluster holds tile rows crank*SH .. crank*SH + SH-1 of
  // the cluster's (CS*SH) x SW tile.  Only the two ends of the stack have ring rows in y.
  const bool has_dn = crank > 0, has_up = crank < CS - 1;
  const int x0 = 1 + bx * TX, y0 = 1 + by * TY;  // first output cell (i, jl)
  const int ib = x0 - H, jb = y0 - H + crank * SH;  // this CTA's tile origin (i, jl)
  const int out_lo = has_dn ? 0 : H, out_hi = has_up ? SH - 1 : SH - H - 1;  // tile rows that belong to the output block
  const int q = tid & 63, sg = tid >> 6;
  const int c0 = 2 * q, i0 = ib + c0;
  const int rr0 = sg * RPT;
  const int jl0 = jb + rr0, jg0 = k.j0 + jl0;
  const bool colI0 = i0 >= 1 && i0 <= k.nx, colI1 = i0 + 1 >= 1 && i0 + 1 <= k.nx;
  const bool colW0 = colI0 && c0 >= 1, colW1 = colI1 && c0 + 1 <= SW - 2;
  const bool colO0 = colI0 && c0 >= H && c0 < H + TX, colO1 = colI1 && c0 + 1 >= H && c0 + 1 < H + TX;
  // Row masks as bit ranges over the thread's RPT rows (bit r = row rr0 + r):
  //   mI: the row holds domain cells this rank has data for;  mW: ... and is not a ring row of the cluster's tile;
  //   mO: the row belongs to the output block and to this rank.
  auto bit_range = [](int lo, int hi) -> unsigned {  // bits lo..hi of an RPT-bit mask, empty if hi < lo
    lo = max(lo, 0);
    hi = min(hi, RPT - 1);
    return hi >= lo ? ((2u << hi) - 1u) & ~((1u << lo) - 1u) : 0u;
  };
  const int jI_lo = max(1 - k.j0, 1 - H), jI_hi = min(k.ny - k.j0, k.nyl + H);  // jl range with data
  const unsigned mI = bit_range(jI_lo - jl0, jI_hi - jl0);
  const int w_lo = has_dn ? 0 : 1, w_hi = has_up ? SH - 1 : SH - 2;  // tile rows that may be updated
  const unsigned mW = mI & bit_range(w_lo - rr0, w_hi - rr0);
  const unsigned mO = bit_range(max(1 - jl0, out_lo - rr0), min(k.nyl - jl0, out_hi - rr0));
  const unsigned mOut = (colO0 ? mO : 0u) | ((colO1 ? mO : 0u) << 16);
  // Every updatable cell of the tile strictly inside the domain, with data for all of its neighbours (uniform over the block)?
  const bool interior = ib + 1 >= 2 && ib + SW - 2 <= k.nx - 1 && k.j0 + jb + w_lo >= 2 && k.j0 + jb + w_hi <= k.ny - 1 &&
                        jb + w_hi + 1 <= k.nyl + H;
  if (tid < C::NWARPS * (T + 1)) red[tid] = 0ull;  // ordered before the warps' stores by the barriers below

  // f: HBM -> registers, 128-bit row loads, overlapping the TMA transfer of p
  Cells<RPT> c;
  {
    const double* fp = f + pm_idx(k, jl0, i0);
    if (interior) {
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(fp + size_t(r) * k.pitch));
        c.f0[r] = v.x;
        c.f1[r] = v.y;
      }
    } else {
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const bool rowI = (mI >> r) & 1u;
        double2 v = make_double2(0.0, 0.0);
        if (rowI && colI0 && colI1) v = __ldg(reinterpret_cast<const double2*>(fp + size_t(r) * k.pitch));
        else if (rowI && colI0) v.x = __ldg(fp + size_t(r) * k.pitch);
        else if (rowI && colI1) v.y = __ldg(fp + size_t(r) * k.pitch + 1);
        c.f0[r] = v.x;
        c.f1[r] = v.y;
      }
    }
  }
  PM_PROF(0);  // masks, f loads issued
  if (check_stop) {  // the reference's loop test (uniform over the grid); the tile loads above are already in flight
    int first;
    if (stop_words_eval<T>(stopw, m0, &first)) {
      if (first >= 0 && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {
        st->iters = first;
        st->done = 1;
      }
      mbar_wait(bar, phase);  // never leave with a bulk copy still landing in this CTA's shared memory
      if (CS > 1) { cluster_arrive_relaxed(); cluster_wait_acquire(); }  // no CTA of the cluster is left waiting for this one
      return;
    }
  }
  mbar_wait(bar, phase);
  // Cluster barrier, first half: this CTA's barriers are initialised and its tile (with the rows the neighbours will
  // overwrite) has landed.  Second half: before the first push into a neighbour (xchg_gate).
  if (CS > 1) cluster_arrive_relaxed();
  PM_PROF(1);  // wait for the TMA tile
  // Obstacle mask (step case): tiles with solid cells take their own path; the classes were computed with the mask.
  bool masked = false;
  if constexpr (FORM == 1 && METHOD == PM_PPE_SOR_RB && CS == 1) {
    const int cls = tclass;
    if (cls != 0) {
      const MaskedGeom g{ib, jb, jI_lo, jI_hi, max(1, y0), min(k.nyl, y0 + TY - 1), max(1, x0), min(k.nx, x0 + TX - 1)};
      masked_tile<A, METHOD, T>(k, tile, f, mask, pout, red, g, m0, nsw, cls);
      masked = true;
    }
  }
  if (!masked) {
    // The tile arrived in the split-row layout: every thread takes its own cells; neighbours are read in place.
    // No barrier: before the first half-sweep's barrier a thread only ever writes cells it owns.
    double* tpx = tile + rr0 * SW + q;
    double* tpy = tpx + SW / 2;
  #pragma unroll
    for (int r = 0; r < RPT; ++r) {
      c.p0[r] = tpx[r * SW];
      c.p1[r] = tpy[r * SW];
    }

    PM_PROF(2);  // own cells
    Xchg x;
    x.bars = xbar;
    x.nhs = 2 * nsw;
    x.dn = CS > 1 && sg == 0 && xact[0];
    x.up = CS > 1 && sg == C::NSEG - 1 && xact[1];
    // the residual-only pass (nsw == 0) commits nothing and takes the general code path
    if (interior && nsw > 0) run_sweeps<A, FORM, METHOD, T, true, PAR0>(k, tpx, tpy, c, i0, jg0, mW, mOut, colW0, colW1, nsw, red, x);
    else run_sweeps<A, FORM, METHOD, T, false, PAR0>(k, tpx, tpy, c, i0, jg0, mW, mOut, colW0, colW1, nsw, red, x);

    PM_PROF(3);  // the sweeps
    // ---- write the output block (split-row layout: .x cells to the even half of the row, .y cells to the odd
    // half, consecutive lanes to consecutive doubles), plus the wall ghosts its cells own ----
    if (nsw > 0) {
      const int P = k.pitch;
      double* ox = pout + size_t(k.padr + jl0) * size_t(P) + size_t((PM_OFFC + i0 + k.psh) >> 1);  // PM_OFFC + i0 is even
      double* oy = ox + (P >> 1);
  #pragma unroll
      for (int r = 0; r < RPT; ++r) {
        if (!((mO >> r) & 1u)) continue;
        if (colO0) ox[size_t(r) * P] = c.p0[r];
        if (colO1) oy[size_t(r) * P] = c.p1[r];
        if (FORM == 1 && !interior) {
          const int j = jg0 + r, jl = jl0 + r;
          if (colO0 && i0 == 1) pout[pm_sidx(k, jl, 0)] = c.p0[r];
          if (colO0 && i0 == k.nx) pout[pm_sidx(k, jl, k.nx + 1)] = 0.0;
          if (colO1 && i0 + 1 == k.nx) pout[pm_sidx(k, jl, k.nx + 1)] = 0.0;
          if (j == 1) { if (colO0) pout[pm_sidx(k, jl - 1, i0)] = c.p0[r]; if (colO1) pout[pm_sidx(k, jl - 1, i0 + 1)] = c.p1[r]; }
          if (j == k.ny) { if (colO0) pout[pm_sidx(k, jl + 1, i0)] = c.p0[r]; if (colO1) pout[pm_sidx(k, jl + 1, i0 + 1)] = c.p1[r]; }
        }
      }
    }
  }

  // ---- residual norms ----
  // Per iterate and tile one global atomicMax, spread over PM_FOLD_SLOTS cache lines (a single line would take
  // (T+1) * tiles atomics per pass and stall every load that touches it); k_tiled_fold, launched behind the pass,
  // folds the slots into res_bits[m] for the iterates the pass completes.  Ring index m & 7: at most T + 1 <= 5
  // iterates are open at a time.
  __syncthreads();  // also: nobody touches the tile in shared memory after this point
  if (tid <= T) {
    const int m = m0 + tid;
    // Jacobi: entry t is the full residual of iterate m0+t.  Red-black: entry t collects the colour-0 part of
    // iterate m0+t (before sweep t+1 replaces it) and the colour-1 part taken right after sweep t created it.
    unsigned long long v = 0ull;
#pragma unroll
    for (int w = 0; w < C::NWARPS; ++w) v = max(v, red[w * (T + 1) + tid]);
    unsigned long long* mine = fold_part + size_t((blockIdx.y * gridDim.x + blockIdx.x) & (PM_FOLD_SLOTS - 1)) * 16;
    if (v != 0ull && m >= 1 && m <= k.max_iters) atomicMax(&mine[m & 7], v);
  }
  PM_PROF(4);  // write-out, residual atomics
#ifdef PM_TILE_PROFILE
  if (tid == 0) atomicAdd(&g_tile_prof[7], 1ull);
#endif
}

// One tile per CTA, two CTAs per SM: one CTA's loads overlap the other's sweeps.  Red-black: CS CTAs stacked in y
// form a cluster (launch attribute), blockIdx.y = cluster row * CS + rank in the cluster.
template <class A, int FORM, int METHOD, int T, int PAR0>
__global__ void __launch_bounds__((TileCfg<METHOD, T>::THREADS), PM_TILE_MINBLOCKS)
    k_ppe_tiled(const __grid_constant__ KP k, const __grid_constant__ CUtensorMap tmap_in, double* __restrict__ pout,
                const double* __restrict__ f, PpeState* __restrict__ st, unsigned long long* __restrict__ res_bits,
                unsigned long long* __restrict__ fold_part, const uint8_t* __restrict__ mask, const uint8_t* __restrict__ tcls,
                const int* __restrict__ torder, int m0, int nsw, int force, int tile_row0, int order_ntx) {
  using C = TileCfg<METHOD, T>;
  constexpr int CS = C::CS;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double* tile = reinterpret_cast<double*>(smem_raw) + C::SW;  // tile rows -1 .. SH: the row below and above come along
  __shared__ __align__(8) uint64_t mbar;
  __shared__ __align__(16) uint64_t xbar[4];  // edge rows from the CTA below [0..1] / above [2..3], alternating with the half-sweep
  __shared__ int xact[2];                     // does this CTA exchange edge rows with the CTA below [0] / above [1]
  __shared__ unsigned long long red[C::NWARPS * (T + 1)];  // per warp: bit patterns of the residual maxima of iterates m0 .. m0+T

  const int tid = threadIdx.x;
  // cluster rows rotated by one: the last row of the launch (all boundary tiles at the top wall, several times
  // slower than interior tiles) starts in the first wave instead of forming the tail
  const int ncy = int(gridDim.y) / CS, cy = int(blockIdx.y) / CS;
  const int crank = CS > 1 ? int(cluster_ctarank()) : 0;
  int bx = blockIdx.x, by = tile_row0 + (cy == 0 ? ncy - 1 : cy - 1);
  int tclass = 0;  // obstacle mask: 0 fluid cells only, 1 no fluid cell, 2 both
  if (CS == 1 && torder != nullptr) {  // whole-grid launch: the slow tiles (fluid and solid cells) first, the copied ones last
    const int t = torder[blockIdx.y * gridDim.x + blockIdx.x];  // tile index | class << 28: one dependent load before the TMA can go out
    const int ntx = order_ntx > 0 ? order_ntx : int(gridDim.x);  // a list launch (tiled_launch_list) has its own grid shape
    tclass = t >> 28;
    by = (t & 0x0fffffff) / ntx;
    bx = (t & 0x0fffffff) - by * ntx;
  } else if (CS == 1 && tcls != nullptr) {
    tclass = tcls[by * int(gridDim.x) + bx];  // in flight during the tile load
  }
#ifdef PM_TILE_PROFILE
  const long long prof_k = clock64();
#endif
  // The tile load goes out before anything else; the words of the reference's loop test follow as
  // independent loads and are looked at only after the f loads of the tile have been issued.
  if (tid == 0) {
    mbar_init(&mbar, 1);
    if (CS > 1) {
#pragma unroll
      for (int b = 0; b < 4; ++b) mbar_init(&xbar[b], 1);
    }
    fence_mbar_init();
    mbar_expect_tx(&mbar, (C::SH + 2 * C::XR) * C::SW * 8);
    tma_load_3d(tile - C::XR * C::SW, &tmap_in, &mbar, (PM_OFFC + 1 + bx * C::TX - C::H + k.psh) >> 1, 0,
                k.padr + 1 + by * C::TY - C::H + crank * C::SH - C::XR);
    if (CS > 1) {
      // A pair of stacked CTAs exchanges only where both edge rows are interior rows of the domain: a wall ghost row is
      // kept up to date inside the tile by the thread that owns the wall-adjacent cell (FORM 1) or never changes (FORM 0).
      const int j_first = k.j0 + 1 + by * C::TY - C::H + crank * C::SH, j_last = j_first + C::SH - 1;
      xact[0] = crank > 0 && j_first - 1 >= 1 && j_first <= k.ny;
      xact[1] = crank < CS - 1 && j_last >= 1 && j_last + 1 <= k.ny;
      // arm the phases of half-sweeps 0 and 1 where a neighbour will send (the last half-sweep of a pass sends nothing)
      const int nhs = 2 * nsw;
      if (crank > 0) {
        if (0 < nhs - 1) mbar_expect_tx(&xbar[0], PM_XCHG_BYTES);
        if (1 < nhs - 1) mbar_expect_tx(&xbar[1], PM_XCHG_BYTES);
      }
      if (crank < CS - 1) {
        if (0 < nhs - 1) mbar_expect_tx(&xbar[2], PM_XCHG_BYTES);
        if (1 < nhs - 1) mbar_expect_tx(&xbar[3], PM_XCHG_BYTES);
      }
    }
  }
  StopWords<T> stopw;
  if (!force) stopw = stop_words_load<T>(st, res_bits, m0);
  __syncthreads();
#ifdef PM_TILE_PROFILE
  if (tid == 0) atomicAdd(&g_tile_prof[5], (unsigned long long)(clock64() - prof_k));
#endif
  tile_process<A, FORM, METHOD, T, PAR0>(k, tile, &mbar, 0u, red, pout, f, st, res_bits, fold_part, m0, nsw, bx, by, crank, xbar, xact, mask, tclass, stopw, !force);
}

// Whole-plane conversion between the natural and the split-row layout (a permutation inside every row).
// to_split: dst[row][split(c)] = src[row][c]; otherwise dst[row][c] = src[row][split(c)].  One thread per column pair.
__global__ void k_split_rows(const __grid_constant__ KP k, const double* __restrict__ src, double* __restrict__ dst, int rows, int to_split) {
  const int pr = blockIdx.x * blockDim.x + threadIdx.x;  // pair index
  const int row = blockIdx.y * blockDim.y + threadIdx.y;
  const int half = k.pitch >> 1;
  if (pr >= half || row >= rows) return;
  const size_t base = size_t(row) * size_t(k.pitch);
  int h = pr + (k.psh >> 1);  // where the pair (2 pr, 2 pr + 1) sits in either half, see pm_split_col
  if (h >= half) h -= half;
  if (to_split) {
    const double2 v = *reinterpret_cast<const double2*>(src + base + 2 * pr);
    dst[base + h] = v.x;
    dst[base + half + h] = v.y;
  } else {
    *reinterpret_cast<double2*>(dst + base + 2 * pr) = make_double2(src[base + h], src[base + half + h]);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
#ifndef PM_TILED_DEVICE_ONLY  // (kernel experiments compile the device code alone)
struct TiledPlan {
  int sweeps = 1;       // T: temporal-blocking depth the tile geometry (halo) was built for
  int run = 1;          // sweeps actually done per pass: T, or T - 1 with an obstacle mask (see masked_tile: the solid cells of the
                        // stored iterates lag one ghost pass behind, which costs one ring of the halo)
  int halo = 2;         // H
  int tx = 0, ty = 0;   // output block
  int sh = 0, threads = 0;  // tile rows, threads per CTA
  int tiles_x = 0, tiles_y = 0;  // output blocks; in y one block per CLUSTER of cs stacked CTAs
  int cs = 1;                    // CTAs per cluster
  int box_rows = 0;              // rows of the TMA box: the tile, plus the row below and above it in a cluster
  int smem_bytes = 0;
  int psh = 0;          // column shift of the split-row layout (KP::psh)
  CUtensorMap map[2];   // p ping / p pong
  double* p[2] = {nullptr, nullptr};
  const void* kernel = nullptr;
  unsigned long long* fold_part = nullptr;  // PM_FOLD_SLOTS * 16 words
  size_t fold_bytes = 0;
  const uint8_t* mask = nullptr;        // obstacle mask plane (step case), else null
  uint8_t* tile_class = nullptr;        // per output block: 0 fluid cells only, 1 no fluid cell, 2 both (step case), else null
  int* tile_order = nullptr;            // launch order of the output blocks of a whole-grid launch: class 2, then 0, then 1
};

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline bool tiled_supported(const pm_config& c, const KP& k) {
  // the obstacle mask: red-black only, independent tiles only (masked_tile does not take part in a cluster's exchange)
  if (c.case_id == PM_CASE_STEP && (c.ppe_method != PM_PPE_SOR_RB || PM_TILE_CS != 1)) return false;
  if (c.ppe_method != PM_PPE_JACOBI && c.ppe_method != PM_PPE_SOR_RB) return false;
  if (k.pitch < 128) return false;
  return true;
}

template <class A, int FORM, int METHOD, int T>
static const void* tiled_kernel_ptr(int par0) {
  return par0 ? reinterpret_cast<const void*>(&k_ppe_tiled<A, FORM, METHOD, T, 1>) : reinterpret_cast<const void*>(&k_ppe_tiled<A, FORM, METHOD, T, 0>);
}

template <int METHOD, int T>
static void tiled_geometry(TiledPlan* pl) {
  using C = TileCfg<METHOD, T>;
  pl->sweeps = T; pl->halo = C::H; pl->tx = C::TX; pl->ty = C::TY; pl->sh = C::SH; pl->threads = C::THREADS;
  pl->smem_bytes = C::SMEM_BYTES; pl->psh = C::PSH; pl->cs = C::CS; pl->box_rows = C::SH + 2 * C::XR;
}

template <class A, int FORM>
static const void* tiled_pick(int method, int T, int par0, TiledPlan* pl) {
  if (method == PM_PPE_SOR_RB) {
    switch (T) {
      case 1: tiled_geometry<PM_PPE_SOR_RB, 1>(pl); return tiled_kernel_ptr<A, FORM, PM_PPE_SOR_RB, 1>(par0);
      case 2: tiled_geometry<PM_PPE_SOR_RB, 2>(pl); return tiled_kernel_ptr<A, FORM, PM_PPE_SOR_RB, 2>(par0);
      case 3: tiled_geometry<PM_PPE_SOR_RB, 3>(pl); return tiled_kernel_ptr<A, FORM, PM_PPE_SOR_RB, 3>(par0);
      case 4: tiled_geometry<PM_PPE_SOR_RB, 4>(pl); return tiled_kernel_ptr<A, FORM, PM_PPE_SOR_RB, 4>(par0);
    }
  } else {
    switch (T) {
      case 1: tiled_geometry<PM_PPE_JACOBI, 1>(pl); return tiled_kernel_ptr<A, FORM, PM_PPE_JACOBI, 1>(par0);
      case 2: tiled_geometry<PM_PPE_JACOBI, 2>(pl); return tiled_kernel_ptr<A, FORM, PM_PPE_JACOBI, 2>(par0);
      case 4: tiled_geometry<PM_PPE_JACOBI, 4>(pl); return tiled_kernel_ptr<A, FORM, PM_PPE_JACOBI, 4>(par0);
    }
  }
  return nullptr;
}

static inline bool tiled_create(TiledPlan* pl, const pm_config& c, const KP& k, double* p0, double* p1, int rows_alloc, std::string* err) {
  // measured at 8192^2: production red-black 13.8 ms/step at T = 4 (14.6 at 3); exact arithmetic 22.3 at T = 3 (23.6 at 4)
  // measured at 8192^2, Jacobi on 32 x 128 tiles, two shared-memory tiles: 13.6 ms/step at T = 4 (17.6 at T = 2)
  int T = c.sweeps_per_pass > 0 ? c.sweeps_per_pass : (c.ppe_method == PM_PPE_SOR_RB ? (c.exact_arith ? 3 : 4) : 4);
  const bool masked = c.case_id == PM_CASE_STEP;
  if (masked) T = c.sweeps_per_pass > 0 ? std::min(4, c.sweeps_per_pass + 1) : 4;  // geometry one sweep deeper than what a pass runs
  const bool cav = c.case_id == PM_CASE_CAVITY;
  const void* kern = nullptr;
  const int par0 = k.j0 & 1;
  if (c.exact_arith) kern = cav ? tiled_pick<Exact, 0>(c.ppe_method, T, par0, pl) : tiled_pick<Exact, 1>(c.ppe_method, T, par0, pl);
  else kern = cav ? tiled_pick<Fast, 0>(c.ppe_method, T, par0, pl) : tiled_pick<Fast, 1>(c.ppe_method, T, par0, pl);
  if (!kern) { *err = "sweeps_per_pass " + std::to_string(T) + " not built for this method (red-black: 1,2,3,4; jacobi: 1,2,4)"; return false; }
  pl->kernel = kern;
  pl->run = masked ? pl->sweeps - 1 : pl->sweeps;
  pl->tiles_x = (k.nx + pl->tx - 1) / pl->tx;
  pl->tiles_y = (k.nyl + pl->ty - 1) / pl->ty;
  pl->p[0] = p0; pl->p[1] = p1;

  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) { *err = "cuTensorMapEncodeTiled not available from the driver"; return false; }
  PFN_tmapEncodeTiled encode = reinterpret_cast<PFN_tmapEncodeTiled>(fn);
  for (int b = 0; b < 2; ++b) {  // {pair, parity, row} view of a split-row plane
    const cuuint64_t gdim[3] = {cuuint64_t(k.pitch / 2), 2u, cuuint64_t(rows_alloc)};
    const cuuint64_t gstr[2] = {cuuint64_t(k.pitch / 2) * 8, cuuint64_t(k.pitch) * 8};
    const cuuint32_t box[3] = {64u, 2u, cuuint32_t(pl->box_rows)};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = encode(&pl->map[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, pl->p[b], gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)); return false; }
  }
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pl->smem_bytes);
  if (e != cudaSuccess) { *err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return false; }
  pl->fold_bytes = size_t(PM_FOLD_SLOTS) * 16 * sizeof(unsigned long long);
  e = cudaMalloc(&pl->fold_part, pl->fold_bytes);
  if (e != cudaSuccess) { *err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return false; }
  return true;
}
static inline void tiled_destroy(TiledPlan* pl) {
  if (pl->fold_part) cudaFree(pl->fold_part);
  if (pl->tile_class) cudaFree(pl->tile_class);
  if (pl->tile_order) cudaFree(pl->tile_order);
  pl->fold_part = nullptr;
  pl->tile_class = nullptr;
  pl->tile_order = nullptr;
}
// Classes of the output blocks for an obstacle mask (global array, (ny+2) x (nx+2) bytes): looked at over the whole
// tile around each block, ring included, clipped to the interior cells this rank has data for.
static inline bool tiled_classify(TiledPlan* pl, const KP& k, const uint8_t* global_mask, cudaStream_t stream, std::string* err) {
  const int ntiles = pl->tiles_x * pl->tiles_y;
  std::vector<uint8_t> cls(size_t(ntiles), 0);
  const int cols = k.nx + 2;
  for (int by = 0; by < pl->tiles_y; ++by)
    for (int bx = 0; bx < pl->tiles_x; ++bx) {
      const int ib = 1 + bx * pl->tx - pl->halo, jb = 1 + by * pl->ty - pl->halo;
      const int jl_lo = std::max(std::max(1 - k.j0, 1 - pl->halo), jb), jl_hi = std::min(std::min(k.ny - k.j0, k.nyl + pl->halo), jb + pl->sh - 1);
      const int i_lo = std::max(1, ib), i_hi = std::min(k.nx, ib + 128 - 1);
      bool any_fluid = false, any_solid = false;
      for (int jl = jl_lo; jl <= jl_hi && !(any_fluid && any_solid); ++jl) {
        const uint8_t* row = global_mask + size_t(k.j0 + jl) * cols;
        for (int i = i_lo; i <= i_hi; ++i) {
          if (row[i]) any_fluid = true; else any_solid = true;
        }
      }
      cls[size_t(by) * pl->tiles_x + bx] = !any_solid ? 0 : (!any_fluid ? 1 : 2);
    }
  std::vector<int> order;
  order.reserve(size_t(ntiles));
  for (int want : {2, 0, 1})
    for (int t = 0; t < ntiles; ++t)
      if (cls[size_t(t)] == want) order.push_back(t | (want << 28));
  cudaError_t e = cudaSuccess;
  if (!pl->tile_class) e = cudaMalloc(&pl->tile_class, size_t(ntiles));
  if (e == cudaSuccess && !pl->tile_order) e = cudaMalloc(&pl->tile_order, size_t(ntiles) * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpyAsync(pl->tile_class, cls.data(), size_t(ntiles), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(pl->tile_order, order.data(), size_t(ntiles) * sizeof(int), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) { *err = std::string("tile classes: ") + cudaGetErrorString(e); return false; }
  return true;
}
// Before the first pass of a solve (stream order): empty slots.
static inline cudaError_t tiled_begin_solve(const TiledPlan* pl, cudaStream_t stream) {
  return cudaMemsetAsync(pl->fold_part, 0, pl->fold_bytes, stream);
}

// Launch one pass: reads iterate m0 from buffer `in`, writes iterate m0+nsw to the other buffer.
static inline cudaError_t tiled_launch(const TiledPlan* pl, const KP& k, int in, const double* f, PpeState* st, unsigned long long* res,
                                       int m0, int nsw, int force, int tile_row0, int tile_rows, cudaStream_t stream) {
  double* pout = pl->p[in ^ 1];
  const int* order = (tile_row0 == 0 && tile_rows == pl->tiles_y) ? pl->tile_order : nullptr;
  int order_ntx = 0;
  void* args[] = {(void*)&k, (void*)&pl->map[in], (void*)&pout, (void*)&f, (void*)&st, (void*)&res, (void*)&pl->fold_part,
                  (void*)&pl->mask, (void*)&pl->tile_class, (void*)&order, (void*)&m0, (void*)&nsw, (void*)&force, (void*)&tile_row0, (void*)&order_ntx};
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(pl->tiles_x, tile_rows * pl->cs);
  lc.blockDim = dim3(pl->threads);
  lc.dynamicSmemBytes = size_t(pl->smem_bytes);
  lc.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = unsigned(pl->cs); at[0].val.clusterDim.z = 1;
  lc.attrs = at;
  lc.numAttrs = pl->cs > 1 ? 1 : 0;
  return cudaLaunchKernelExC(&lc, pl->kernel, args);
}
// The same pass over an explicit list of n tiles (device array: tile index | class << 28), one CTA each: the frame of tiles
// around the rectangle the streaming kernel takes (pm_kernels_stream.cuh).  Independent tiles only.
static inline cudaError_t tiled_launch_list(const TiledPlan* pl, const KP& k, int in, const double* f, PpeState* st, unsigned long long* res,
                                            int m0, int nsw, int force, const int* list, int n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (pl->cs != 1) return cudaErrorInvalidValue;
  double* pout = pl->p[in ^ 1];
  int tile_row0 = 0, order_ntx = pl->tiles_x;
  void* args[] = {(void*)&k, (void*)&pl->map[in], (void*)&pout, (void*)&f, (void*)&st, (void*)&res, (void*)&pl->fold_part,
                  (void*)&pl->mask, (void*)&pl->tile_class, (void*)&list, (void*)&m0, (void*)&nsw, (void*)&force, (void*)&tile_row0, (void*)&order_ntx};
  return cudaLaunchKernel(pl->kernel, dim3(n, 1), dim3(pl->threads), args, size_t(pl->smem_bytes), stream);
}
// Behind all launches of the pass that started at iterate m0 with nsw sweeps.
static inline cudaError_t tiled_fold_launch(const TiledPlan* pl, const KP& k, unsigned long long* res, int m0, int nsw, cudaStream_t stream) {
  const int lo = std::max(m0, 1), hi = std::min(m0 + std::max(nsw, 1) - 1, k.max_iters);
  if (hi < lo) return cudaSuccess;
  k_tiled_fold<<<1, 32, 0, stream>>>(pl->fold_part, res, lo, hi);
  return cudaGetLastError();
}
#endif  // PM_TILED_DEVICE_ONLY
