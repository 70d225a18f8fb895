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
