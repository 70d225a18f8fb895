// pm_kernels_lex.cuh — PM_PPE_SOR_LEX: the reference's own lexicographic Gauss-Seidel/SOR ordering
// (cavity-01.cpp:640-656; channel-01.cpp:657-668; backwards_step-01.cpp:898-911), made parallel without
// changing a single operand: cell (j, i) needs the NEW west and south values and the OLD east and north
// values, so all cells of one anti-diagonal d = i + j are independent and the diagonals are processed in
// order.  One persistent CTA runs the whole solve — sweep, ghost refresh, residual, loop test — with the
// pressure field resident in shared memory when it fits (all BASELINE configs[0..2] do), so no host round
// trip happens per iteration.  Arithmetic is always the Exact policy: this mode exists to reproduce the
// reference bit for bit (it is ~min(nx,ny)-way parallel only; production uses red-black).
#pragma once
#include "pm_common.cuh"

template <int FORM, bool MASK>
__global__ void __launch_bounds__(1024, 1)
    k_ppe_lex(const __grid_constant__ KP k, double* pg, const double* __restrict__ f, const uint8_t* __restrict__ M,
              PpeState* __restrict__ st, unsigned long long* __restrict__ res_bits, int use_smem) {
  extern __shared__ double sp[];
  __shared__ double red[32];
  __shared__ double s_res;
  const int nx = k.nx, ny = k.ny;
  const int tid = threadIdx.x, nth = blockDim.x;
  double* P;  // deliberately not restrict: other threads write what this thread reads after each barrier
  int PP;
  if (use_smem) {
    PP = nx + 2;
    for (int idx = tid; idx < (ny + 2) * (nx + 2); idx += nth) {
      const int j = idx / (nx + 2), i = idx - j * (nx + 2);
      sp[idx] = pg[pm_idx(k, j, i)];
    }
    P = sp;
  } else {
    PP = k.pitch;
    P = pg + pm_idx(k, 0, 0);
  }
  __syncthreads();

  const double tol = st->tol;
  double res = st->res_init;
  int it = 0;
  while (res > tol && it < k.max_iters) {  // cavity-01.cpp:635
    ++it;
    for (int d = 2; d <= nx + ny; ++d) {
      for (int i = 1 + tid; i <= nx; i += nth) {
        const int j = d - i;
        if (j < 1 || j > ny) continue;
        const size_t g = pm_idx(k, j, i);
        if (MASK && !M[g]) continue;
        double* c = P + size_t(j) * PP + i;
        const double pc = c[0], pe = c[1], pw = c[-1], pn = c[PP], ps = c[-PP], fc = f[g];
        c[0] = FORM == 0 ? upd_cavity<Exact>(k, j, i, pc, pe, pw, pn, ps, fc) : upd_channel<Exact>(k, pc, pe, pw, pn, ps, fc);
      }
      __syncthreads();
    }
    if (FORM == 1) {  // applyPressureGhosts, channel-01.cpp:531-541 / backwards_step-01.cpp:685-740
      for (int t = 1 + tid; t <= max(nx, ny); t += nth) {
        if (t <= ny) {
          P[size_t(t) * PP] = P[size_t(t) * PP + 1];
          P[size_t(t) * PP + nx + 1] = 0.0;
        }
        if (t <= nx) {
          P[t] = P[size_t(PP) + t];
          P[size_t(ny + 1) * PP + t] = P[size_t(ny) * PP + t];
        }
      }
      __syncthreads();
      if (MASK) {  // solid cells read fluid neighbours only, so one parallel phase reproduces the serial loop
        for (int idx = tid; idx < nx * ny; idx += nth) {
          const int j = 1 + idx / nx, i = 1 + idx - (j - 1) * nx;
          const size_t g = pm_idx(k, j, i);
          if (M[g]) continue;
          double* c = P + size_t(j) * PP + i;
          double s = 0.0;
          int n = 0;
          if (i > 1 && M[g - 1]) { s = __dadd_rn(s, c[-1]); ++n; }
          if (i < nx && M[g + 1]) { s = __dadd_rn(s, c[1]); ++n; }
          if (j > 1 && M[g - k.pitch]) { s = __dadd_rn(s, c[-PP]); ++n; }
          if (j < ny && M[g + k.pitch]) { s = __dadd_rn(s, c[PP]); ++n; }
          if (n > 0) c[0] = __ddiv_rn(s, double(n));
        }
        __syncthreads();
      }
    }
    double a = 0.0;
    for (int idx = tid; idx < nx * ny; idx += nth) {
      const int j = 1 + idx / nx, i = 1 + idx - (j - 1) * nx;
      const size_t g = pm_idx(k, j, i);
      if (MASK && !M[g]) continue;
      const double* c = P + size_t(j) * PP + i;
      const double r = FORM == 0 ? res_cavity<Exact>(k, j, i, c[0], c[1], c[-1], c[PP], c[-PP], f[g], k.idx2)
                                 : res_channel<Exact>(k, c[0], c[1], c[-1], c[PP], c[-PP], f[g]);
      a = fmax(a, fabs(r));
    }
    const double m = block_max(a, red);
    if (tid == 0) s_res = m;
    __syncthreads();
    res = s_res;
  }

  if (use_smem) {
    for (int idx = tid; idx < (ny + 2) * (nx + 2); idx += nth) {
      const int j = idx / (nx + 2), i = idx - j * (nx + 2);
      pg[pm_idx(k, j, i)] = sp[idx];
    }
  }
  if (tid == 0) {
    st->iters = it;
    st->done = 1;
    if (it >= 1) res_bits[it] = (unsigned long long)__double_as_longlong(res);
  }
}


// Column-per-thread wavefront (nx <= 1024): the same order, with the dependent chain of one anti-diagonal
// step cut down to what really depends on the previous step.  Thread t owns column i = t + 1 and walks it
// upwards one cell per step (cell (i, d - i) at step d).  Of the five operands of a cell, only the NEW west
// value (just produced by lane t-1) and the NEW south value (this thread's own previous result) are on the
// critical path: west arrives by a warp shuffle (lane 0 of each warp and column 1 read shared memory), south
// stays in a register, and the OLD centre / east / north values and f of the next step are prefetched before
// the barrier.  Values and evaluation order are exactly those of the reference sweep.
template <int FORM, bool MASK>
__global__ void __launch_bounds__(1024, 1)
    k_ppe_lex_cols(const __grid_constant__ KP k, double* pg, const double* __restrict__ f, const uint8_t* __restrict__ M,
                   PpeState* __restrict__ st, unsigned long long* __restrict__ res_bits) {
  extern __shared__ double sp[];
  __shared__ double red[32];
  __shared__ double s_res;
  const int nx = k.nx, ny = k.ny, PP = nx + 2;
  const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31;
  for (int idx = tid; idx < (ny + 2) * PP; idx += nth) {
    const int j = idx / PP, i = idx - j * PP;
    sp[idx] = pg[pm_idx(k, j, i)];
  }
  double* P = sp;  // not restrict: neighbours write what this thread reads after each barrier
  __syncthreads();

  const int i = tid + 1;            // this thread's column (threads beyond nx only help outside the sweep)
  const bool col = i <= nx;
  const double tol = st->tol;
  double res = st->res_init;
  int it = 0;
  while (res > tol && it < k.max_iters) {
    ++it;
    // operands of the first cell of the column (row 1), fetched before the wave reaches it
    double pc = 0.0, pe = 0.0, pn = 0.0, mine = 0.0;  // mine: this thread's previous result (cell (i, j-1))
    // f (and the mask) come from global memory: a ring of the next four rows keeps their latency off the wave
    double fr[4] = {0.0, 0.0, 0.0, 0.0};
    bool mr[4] = {true, true, true, true};
    if (col) {
      pc = P[PP + i]; pe = P[PP + i + 1]; pn = P[2 * PP + i];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (1 + q <= ny) {
          fr[q] = f[pm_idx(k, 1 + q, i)];
          if (MASK) mr[q] = M[pm_idx(k, 1 + q, i)] != 0;
        }
      mine = P[i];  // south ghost of row 1
    }
    for (int d = 2; d <= nx + ny; ++d) {
      const int j = d - i;
      const bool act = col && j >= 1 && j <= ny;
      // west: lane-1 produced (i-1, j) in the previous step; warp-boundary lanes and column 1 read shared memory
      double pw = __shfl_up_sync(0xffffffffu, mine, 1);
      if (act && (lane == 0 || i == 1)) pw = P[j * PP + i - 1];
      if (act) {
        const double ps = mine;  // (i, j-1): own previous result, or the south ghost for j == 1
        double nv = pc;
        if (!MASK || mr[0]) nv = FORM == 0 ? upd_cavity<Exact>(k, j, i, pc, pe, pw, pn, ps, fr[0]) : upd_channel<Exact>(k, pc, pe, pw, pn, ps, fr[0]);
        P[j * PP + i] = nv;
        mine = nv;  // a solid cell passes its unchanged value on
        if (j + 1 <= ny) {  // operands of the next cell of the column: all still of the previous iterate
          pc = pn;
          pe = P[(j + 1) * PP + i + 1];
          pn = P[(j + 2) * PP + i];
        }
        fr[0] = fr[1]; fr[1] = fr[2]; fr[2] = fr[3];
        mr[0] = mr[1]; mr[1] = mr[2]; mr[2] = mr[3];
        if (j + 4 <= ny) {
          fr[3] = f[pm_idx(k, j + 4, i)];
          if (MASK) mr[3] = M[pm_idx(k, j + 4, i)] != 0;
        }
      }
      __syncthreads();
    }
    if (FORM == 1) {  // applyPressureGhosts, channel-01.cpp:531-541 / backwards_step-01.cpp:685-740
      for (int t = 1 + tid; t <= max(nx, ny); t += nth) {
        if (t <= ny) {
          P[size_t(t) * PP] = P[size_t(t) * PP + 1];
          P[size_t(t) * PP + nx + 1] = 0.0;
        }
        if (t <= nx) {
          P[t] = P[size_t(PP) + t];
          P[size_t(ny + 1) * PP + t] = P[size_t(ny) * PP + t];
        }
      }
      __syncthreads();
      if (MASK) {
        for (int idx = tid; idx < nx * ny; idx += nth) {
          const int j = 1 + idx / nx, ii = 1 + idx - (j - 1) * nx;
          const size_t g = pm_idx(k, j, ii);
          if (M[g]) continue;
          double* c = P + size_t(j) * PP + ii;
          double s = 0.0;
          int n = 0;
          if (ii > 1 && M[g - 1]) { s = __dadd_rn(s, c[-1]); ++n; }
          if (ii < nx && M[g + 1]) { s = __dadd_rn(s, c[1]); ++n; }
          if (j > 1 && M[g - k.pitch]) { s = __dadd_rn(s, c[-PP]); ++n; }
          if (j < ny && M[g + k.pitch]) { s = __dadd_rn(s, c[PP]); ++n; }
          if (n > 0) c[0] = __ddiv_rn(s, double(n));
        }
        __syncthreads();
      }
    }
    double a = 0.0;
    for (int idx = tid; idx < nx * ny; idx += nth) {
      const int j = 1 + idx / nx, ii = 1 + idx - (j - 1) * nx;
      const size_t g = pm_idx(k, j, ii);
      if (MASK && !M[g]) continue;
      const double* c = P + size_t(j) * PP + ii;
      const double r = FORM == 0 ? res_cavity<Exact>(k, j, ii, c[0], c[1], c[-1], c[PP], c[-PP], f[g], k.idx2)
                                 : res_channel<Exact>(k, c[0], c[1], c[-1], c[PP], c[-PP], f[g]);
      a = fmax(a, fabs(r));
    }
    const double m = block_max(a, red);
    if (tid == 0) s_res = m;
    __syncthreads();
    res = s_res;
  }
  for (int idx = tid; idx < (ny + 2) * PP; idx += nth) {
    const int j = idx / PP, ii = idx - j * PP;
    pg[pm_idx(k, j, ii)] = sp[idx];
  }
  if (tid == 0) {
    st->iters = it;
    st->done = 1;
    if (it >= 1) res_bits[it] = (unsigned long long)__double_as_longlong(res);
  }
}

// ---------------------------------------------------------------------------
// Small grids (BASELINE configs[0..2]: <= 16 K cells, thousands of iterations per step): the same persistent
// single-CTA solve for the Jacobi and red-black orderings.  The general path needs 3-5 launches per
// iteration (~8 us); here an iteration is two half-sweeps over shared memory, the ghost refresh and the
// residual tree, separated by __syncthreads (SURVEY 7.3 "small grids": latency-bound regime), and the
// reference's loop test runs on the device every iteration exactly as written (cavity-01.cpp:635).
// ---------------------------------------------------------------------------
template <class A, int FORM, bool MASK, int METHOD>
__global__ void __launch_bounds__(1024, 1)
    k_ppe_small(const __grid_constant__ KP k, double* pg, const double* __restrict__ f, const uint8_t* __restrict__ M,
                PpeState* __restrict__ st, unsigned long long* __restrict__ res_bits) {
  extern __shared__ double sp[];
  __shared__ double red[32];
  __shared__ double s_res;
  const int nx = k.nx, ny = k.ny, PP = nx + 2;
  const int tid = threadIdx.x, nth = blockDim.x;
  for (int idx = tid; idx < (ny + 2) * PP; idx += nth) {
    const int j = idx / PP, i = idx - j * PP;
    sp[idx] = pg[pm_idx(k, j, i)];
  }
  double* P = sp;
  __syncthreads();

  auto relax = [&](int j, int i, const double* c, double fc) -> double {
    return FORM == 0 ? upd_cavity<A>(k, j, i, c[0], c[1], c[-1], c[PP], c[-PP], fc) : upd_channel<A>(k, c[0], c[1], c[-1], c[PP], c[-PP], fc);
  };
  const double tol = st->tol;
  double res = st->res_init;
  int it = 0;
  const int hw = (nx + 1) / 2;
  constexpr int JMAX = 20;  // Jacobi: new values wait in registers until every thread has read (nx*ny <= JMAX*1024 enforced by the host)
  while (res > tol && it < k.max_iters) {
    ++it;
    if (METHOD == PM_PPE_SOR_RB) {
      for (int colour = 0; colour < 2; ++colour) {
        for (int idx = tid; idx < ny * hw; idx += nth) {
          const int j = 1 + idx / hw;
          const int i = 1 + ((colour + j + 1) & 1) + 2 * (idx - (j - 1) * hw);
          if (i > nx) continue;
          const size_t g = pm_idx(k, j, i);
          if (MASK && !M[g]) continue;
          double* c = P + size_t(j) * PP + i;
          c[0] = relax(j, i, c, f[g]);
        }
        __syncthreads();
      }
    } else {
      double nv[JMAX];
#pragma unroll
      for (int q = 0; q < JMAX; ++q) {
        const int idx = tid + q * nth;
        nv[q] = 0.0;
        if (idx < nx * ny) {
          const int j = 1 + idx / nx, i = 1 + idx - (j - 1) * nx;
          const size_t g = pm_idx(k, j, i);
          const double* c = P + size_t(j) * PP + i;
          nv[q] = (MASK && !M[g]) ? c[0] : relax(j, i, c, f[g]);
        }
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < JMAX; ++q) {
        const int idx = tid + q * nth;
        if (idx < nx * ny) {
          const int j = 1 + idx / nx, i = 1 + idx - (j - 1) * nx;
          P[size_t(j) * PP + i] = nv[q];
        }
      }
      __syncthreads();
    }
    if (FORM == 1) {  // applyPressureGhosts, channel-01.cpp:531-541 / backwards_step-01.cpp:685-740
      for (int t = 1 + tid; t <= max(nx, ny); t += nth) {
        if (t <= ny) {
          P[size_t(t) * PP] = P[size_t(t) * PP + 1];
          P[size_t(t) * PP + nx + 1] = 0.0;
        }
        if (t <= nx) {
          P[t] = P[size_t(PP) + t];
          P[size_t(ny + 1) * PP + t] = P[size_t(ny) * PP + t];
        }
      }
      __syncthreads();
      if (MASK) {
        for (int idx = tid; idx < nx * ny; idx += nth) {
          const int j = 1 + idx / nx, i = 1 + idx - (j - 1) * nx;
          const size_t g = pm_idx(k, j, i);
          if (M[g]) continue;
          double* c = P + size_t(j) * PP + i;
          double s = 0.0;
          int n = 0;
          if (i > 1 && M[g - 1]) { s = __dadd_rn(s, c[-1]); ++n; }
          if (i < nx && M[g + 1]) { s = __dadd_rn(s, c[1]); ++n; }
          if (j > 1 && M[g - k.pitch]) { s = __dadd_rn(s, c[-PP]); ++n; }
          if (j < ny && M[g + k.pitch]) { s = __dadd_rn(s, c[PP]); ++n; }
          if (n > 0) c[0] = __ddiv_rn(s, double(n));
        }
        __syncthreads();
      }
    }
    double a = 0.0;
    for (int idx = tid; idx < nx * ny; idx += nth) {
      const int j = 1 + idx / nx, i = 1 + idx - (j - 1) * nx;
      const size_t g = pm_idx(k, j, i);
      if (MASK && !M[g]) continue;
      const double* c = P + size_t(j) * PP + i;
      const double r = FORM == 0 ? res_cavity<A>(k, j, i, c[0], c[1], c[-1], c[PP], c[-PP], f[g], k.idx2)
                                 : res_channel<A>(k, c[0], c[1], c[-1], c[PP], c[-PP], f[g]);
      a = fmax(a, fabs(r));
    }
    const double m = block_max(a, red);
    if (tid == 0) s_res = m;
    __syncthreads();
    res = s_res;
  }
  for (int idx = tid; idx < (ny + 2) * PP; idx += nth) {
    const int j = idx / PP, i = idx - j * PP;
    pg[pm_idx(k, j, i)] = sp[idx];
  }
  if (tid == 0) {
    st->iters = it;
    st->done = 1;
    if (it >= 1) res_bits[it] = (unsigned long long)__double_as_longlong(res);
  }
}

// ---------------------------------------------------------------------------
// Small grids, red-black, on a thread-block CLUSTER: the same persistent solve spread over 8 SMs.
// CTA c of the cluster owns a band of rows of p and f in its own shared memory plus one halo row above and
// below; after each colour half-sweep every CTA pushes the freshly updated colour of its first and last row straight
// into its neighbours' halo rows (distributed shared memory: st.async, counted on an mbarrier of the receiving CTA)
// and waits for its own halo rows to arrive -- no cluster barrier inside the loop (barrier.cluster's release costs a
// MEMBAR.GPU: the pull-based version spent 3.7 us per iteration in three of them).  The per-CTA residual maxima travel
// the same way to all CTAs of the cluster, and every CTA evaluates the reference's loop test (cavity-01.cpp:635) on the same
// number, so the cluster leaves the loop together.
// ---------------------------------------------------------------------------
#include <cooperative_groups.h>
namespace pm_cg = cooperative_groups;

// CTAs per cluster.  Measured on configs[0..2] (us per iteration): one CTA 14.6 / 14.2 / 11.7, 4 CTAs 4.2 / 4.5 / 7.6,
// 8 CTAs 2.9 / 3.2 / 6.5, 16 CTAs (a non-portable size: cudaFuncAttributeNonPortableClusterSizeAllowed) 2.4 / 2.4 / 4.6.
#ifndef PM_CLUSTER
#define PM_CLUSTER 16
#endif

__device__ __forceinline__ void pm_band(int ny, int c, int* ja, int* nr) {  // rows ja+1 .. ja+nr of the domain
  const int base = ny / PM_CLUSTER, rem = ny % PM_CLUSTER;
  *nr = base + (c < rem ? 1 : 0);
  *ja = c * base + min(c, rem);
}

#ifndef PM_CLUSTER_CPT
#define PM_CLUSTER_CPT 2  // cells per thread and colour (512 threads x 2 x 2 colours x 16 CTAs = 32 K cells)
#endif

template <class A, int FORM, bool MASK>
__global__ void __launch_bounds__(512, 1)
    k_ppe_cluster(const __grid_constant__ KP k, double* pg, const double* __restrict__ f, const uint8_t* __restrict__ M,
                  PpeState* __restrict__ st, unsigned long long* __restrict__ res_bits) {
  extern __shared__ double smem[];
  __shared__ double red[32];
  __shared__ __align__(16) double rslot[2][PM_CLUSTER];  // the CTAs' residual maxima, double-buffered by iteration parity
  __shared__ __align__(8) uint64_t hbar[3][2];           // halo rows arrived: [exchange: colour 0, colour 1, every cell][from below, from above]
  __shared__ __align__(8) uint64_t rbar[2];              // residual maxima arrived, by iteration parity
  pm_cg::cluster_group cluster = pm_cg::this_cluster();
  const int c = int(cluster.block_rank());
  const int nx = k.nx, ny = k.ny, PP = nx + 2;
  const int tid = threadIdx.x, nth = blockDim.x;
  int j0, nr;
  pm_band(ny, c, &j0, &nr);
  double* P = smem;  // local rows 0 .. nr+1  <->  domain rows j0 .. j0+nr+1
  for (int idx = tid; idx < (nr + 2) * PP; idx += nth) {
    const int l = idx / PP, i = idx - l * PP;
    P[idx] = pg[pm_idx(k, j0 + l, i)];
  }
  const bool has_dn = c > 0, has_up = c + 1 < PM_CLUSTER;
  int nr_dn = 0, jdummy;
  if (has_dn) pm_band(ny, c - 1, &jdummy, &nr_dn);
  // Where this CTA's edge rows land: the lower neighbour's local row nr_dn + 1, the upper neighbour's local row 0.
  const uint32_t p_dn = has_dn ? mapa_u32(smem_u32(P + size_t(nr_dn + 1) * PP), uint32_t(c - 1)) : 0u;
  const uint32_t p_up = has_up ? mapa_u32(smem_u32(P), uint32_t(c + 1)) : 0u;
  // cells of one colour in domain row j (kind 2: every cell), times 8 bytes: what one exchange delivers per direction
  auto row_bytes = [&](int kind, int j) -> uint32_t {
    if (kind == 2) return uint32_t(nx) * 8u;
    const int first = 1 + ((kind + j + 1) & 1);  // first i >= 1 with (i + j) % 2 == kind
    return uint32_t((nx - first) / 2 + 1) * 8u;
  };
  constexpr int NKIND = MASK ? 3 : 2;
  if (tid == 0) {
    for (int kd = 0; kd < 3; ++kd) { mbar_init(&hbar[kd][0], 1); mbar_init(&hbar[kd][1], 1); }
    mbar_init(&rbar[0], 1);
    mbar_init(&rbar[1], 1);
    fence_mbar_init();
    for (int kd = 0; kd < NKIND; ++kd) {
      if (has_dn) mbar_expect_tx(&hbar[kd][0], row_bytes(kd, j0));            // my local row 0 = domain row j0
      if (has_up) mbar_expect_tx(&hbar[kd][1], row_bytes(kd, j0 + nr + 1));   // my local row nr+1
    }
    mbar_expect_tx(&rbar[0], PM_CLUSTER * 8);
    mbar_expect_tx(&rbar[1], PM_CLUSTER * 8);
  }

  // The cells this thread relaxes, fixed for the whole solve: per colour up to PM_CLUSTER_CPT cells with their
  // shared-memory offset, their (j, i) and their source value in registers, so an iteration is nothing but
  // loads, the reference's expression tree and one store per cell (no index arithmetic in the loop).
  const int hw = (nx + 1) / 2;
  int off[2][PM_CLUSTER_CPT], ji[2][PM_CLUSTER_CPT];
  double fv[2][PM_CLUSTER_CPT];
#pragma unroll
  for (int colour = 0; colour < 2; ++colour)
#pragma unroll
    for (int q = 0; q < PM_CLUSTER_CPT; ++q) {
      const int idx = tid + q * nth;
      off[colour][q] = -1; ji[colour][q] = 0; fv[colour][q] = 0.0;
      if (idx < nr * hw) {
        const int l = 1 + idx / hw, j = j0 + l;
        const int i = 1 + ((colour + j + 1) & 1) + 2 * (idx - (l - 1) * hw);
        if (i <= nx && (!MASK || M[pm_idx(k, j, i)])) {
          off[colour][q] = l * PP + i;
          ji[colour][q] = (j << 16) | i;
          fv[colour][q] = f[pm_idx(k, j, i)];
        }
      }
    }
  cluster.sync();  // every CTA's band is loaded and its barriers are armed: from here on the CTAs only push

  // Halo rows by push (no cluster barrier in the loop): after a half-sweep every CTA writes the cells of that colour of
  // its first and last row straight into the neighbours' halo rows (st.async, counted on the neighbour's mbarrier) and
  // waits for its own halo rows to arrive.  One barrier per kind of exchange and direction: the bytes of iteration n + 1
  // cannot arrive before phase n has completed, because the neighbour sends them only after it has received what this CTA
  // sends later in iteration n (and the residual exchange closes every iteration).
  auto exchange = [&](int kind, int it) {
    __syncthreads();  // the half-sweep's values are in shared memory
    for (int t = tid; t < 2 * nx; t += nth) {
      const int up = t >= nx, i = 1 + (up ? t - nx : t);
      const int l = up ? nr : 1;
      if (kind < 2 && ((i + j0 + l) & 1) != kind) continue;
      const double v = P[size_t(l) * PP + i];
      if (!up && has_dn) st_async_f64(p_dn + uint32_t(i) * 8u, v, mapa_u32(smem_u32(&hbar[kind][1]), uint32_t(c - 1)));
      if (up && has_up) st_async_f64(p_up + uint32_t(i) * 8u, v, mapa_u32(smem_u32(&hbar[kind][0]), uint32_t(c + 1)));
    }
    const uint32_t ph = uint32_t(it - 1) & 1u;
    if (tid < 32) {  // one warp watches the barriers; the block barrier below passes what it has seen on to the others
      if (has_dn) mbar_wait(&hbar[kind][0], ph);
      if (has_up) mbar_wait(&hbar[kind][1], ph);
    }
    __syncthreads();  // both rows have arrived: safe to arm the next phase and to read them
    if (tid == 0) {
      if (has_dn) mbar_expect_tx(&hbar[kind][0], row_bytes(kind, j0));
      if (has_up) mbar_expect_tx(&hbar[kind][1], row_bytes(kind, j0 + nr + 1));
    }
  };
  auto half_sweep = [&](int colour) {
#pragma unroll
    for (int q = 0; q < PM_CLUSTER_CPT; ++q) {
      const int o = colour ? off[1][q] : off[0][q];
      if (o < 0) continue;
      const int jj = colour ? ji[1][q] : ji[0][q];
      const int j = jj >> 16, i = jj & 0xffff;
      const double fc = colour ? fv[1][q] : fv[0][q];
      double* qd = P + o;
      const double r = FORM == 0 ? upd_cavity<A>(k, j, i, qd[0], qd[1], qd[-1], qd[PP], qd[-PP], fc) : upd_channel<A>(k, qd[0], qd[1], qd[-1], qd[PP], qd[-PP], fc);
      qd[0] = r;
      if (FORM == 1 && !MASK) {  // wall ghosts owned by this cell (channel-01.cpp:531-541)
        if (i == 1) qd[-1] = r;
        if (i == nx) qd[1] = 0.0;
        if (j == 1) qd[-PP] = r;
        if (j == ny) qd[PP] = r;
      }
    }
  };

  const double tol = st->tol;
  double res = st->res_init;
  int it = 0;
  while (res > tol && it < k.max_iters) {
    ++it;
    half_sweep(0);
    exchange(0, it);
    half_sweep(1);
    exchange(1, it);
    if (FORM == 1 && MASK) {  // backwards_step-01.cpp:685-740: wall ghosts first, then the solid cells
      for (int t = 1 + tid; t <= max(nx, nr); t += nth) {
        if (t <= nr) {
          P[size_t(t) * PP] = P[size_t(t) * PP + 1];
          P[size_t(t) * PP + nx + 1] = 0.0;
        }
        if (t <= nx) {
          if (c == 0) P[t] = P[size_t(PP) + t];
          if (c == PM_CLUSTER - 1) P[size_t(nr + 1) * PP + t] = P[size_t(nr) * PP + t];
        }
      }
      __syncthreads();
      for (int idx = tid; idx < nr * nx; idx += nth) {
        const int l = 1 + idx / nx, i = 1 + idx - (l - 1) * nx, j = j0 + l;
        const size_t g = pm_idx(k, j, i);
        if (M[g]) continue;
        double* q = P + size_t(l) * PP + i;
        double s = 0.0;
        int n = 0;
        if (i > 1 && M[g - 1]) { s = __dadd_rn(s, q[-1]); ++n; }
        if (i < nx && M[g + 1]) { s = __dadd_rn(s, q[1]); ++n; }
        if (j > 1 && M[g - k.pitch]) { s = __dadd_rn(s, q[-PP]); ++n; }
        if (j < ny && M[g + k.pitch]) { s = __dadd_rn(s, q[PP]); ++n; }
        if (n > 0) q[0] = __ddiv_rn(s, double(n));
      }
      exchange(2, it);  // solid cells of the neighbours' boundary rows changed too
    }
    double a = 0.0;
#pragma unroll
    for (int colour = 0; colour < 2; ++colour)
#pragma unroll
      for (int q = 0; q < PM_CLUSTER_CPT; ++q) {
        const int o = off[colour][q];
        if (o < 0) continue;
        const int j = ji[colour][q] >> 16, i = ji[colour][q] & 0xffff;
        const double* qd = P + o;
        const double r = FORM == 0 ? res_cavity<A>(k, j, i, qd[0], qd[1], qd[-1], qd[PP], qd[-PP], fv[colour][q], k.idx2)
                                   : res_channel<A>(k, qd[0], qd[1], qd[-1], qd[PP], qd[-PP], fv[colour][q]);
        a = fmax(a, fabs(r));
      }
    // the maximum over the cluster: every CTA pushes its own into the slot array of all eight (itself included)
    // two REDUX per warp (pm_kernels_tiled.cuh: warp_max_nonneg) instead of five rounds of 64-bit shuffles
    const int b = it & 1;
    {
      const double wm = warp_max_nonneg(a);
      if ((tid & 31) == 0) red[tid >> 5] = wm;
      __syncthreads();
      if (tid < 32) {
        const double v = warp_max_nonneg(tid < (nth >> 5) ? red[tid] : 0.0);
        if (tid == 0) red[0] = v;
      }
      __syncthreads();
    }
    if (tid < PM_CLUSTER)
      st_async_f64(mapa_u32(smem_u32(&rslot[b][c]), uint32_t(tid)), red[0], mapa_u32(smem_u32(&rbar[b]), uint32_t(tid)));
    if (tid < 32) mbar_wait(&rbar[b], uint32_t((it - 1) >> 1) & 1u);
    __syncthreads();
    const double g = (tid & 31) < PM_CLUSTER ? rslot[b][tid & 31] : 0.0;
    res = warp_max_nonneg(g);  // every warp reads the 8 maxima itself, so all threads of the cluster see the same number
    __syncthreads();    // every thread has read the slots: arm their next use (two iterations on)
    if (tid == 0) mbar_expect_tx(&rbar[b], PM_CLUSTER * 8);
  }
  cluster.sync();  // nobody leaves while a neighbour may still be pushing into this CTA's shared memory

  // write back the band, the wall-ghost columns, and the ghost rows the first / last CTA own
  const int la = c == 0 ? 0 : 1, lb = c == PM_CLUSTER - 1 ? nr + 1 : nr;
  for (int idx = tid; idx < (lb - la + 1) * PP; idx += nth) {
    const int l = la + idx / PP, i = idx - (l - la) * PP;
    pg[pm_idx(k, j0 + l, i)] = P[size_t(l) * PP + i];
  }
  if (c == 0 && tid == 0) {
    st->iters = it;
    st->done = 1;
    if (it >= 1) res_bits[it] = (unsigned long long)__double_as_longlong(res);
  }
}
