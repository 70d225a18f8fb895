// pm_kernels_residual.cuh — the residual-only pass of a capped red-black solve on the split-row buffers.
//
// The colour-0 half of an iterate's residual is only known when the next sweep reads its operands, so after the last pass of a
// solve that hits its iteration cap one more look at p and f is needed (cavity-01.cpp:659-677 evaluates the residual after
// every sweep).  k_ppe_tiled does it with nsw = 0 -- a whole tile load per 32 x 112 output block, 0.30 ms at 8192^2; this
// kernel reads p and f once, row-wise: one thread per column pair and row computes the residual of the pair's colour-0 cell
// with the very expressions the tiled kernel's wall path uses (rb_half, production arithmetic), so the value is bit-identical.
// Loop test, residual slots and fold: as in the passes (stop_words_*, fold_part, k_tiled_fold).
#pragma once
#include "pm_kernels_tiled.cuh"

template <int FORM>
__global__ void __launch_bounds__(512)
    k_residual_split(const __grid_constant__ KP k, const double* __restrict__ p, const double* __restrict__ f, PpeState* __restrict__ st,
                     unsigned long long* __restrict__ res_bits, unsigned long long* __restrict__ fold_part, int m0, int force) {
  constexpr int T = 4;
  __shared__ unsigned long long wmax[16];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (!force) {  // the reference's loop test for the iterates of the last pass (uniform over the grid)
    const StopWords<T> w = stop_words_load<T>(st, res_bits, m0);
    int first;
    if (stop_words_eval<T>(w, m0, &first)) {
      if (first >= 0 && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {
        st->iters = first;
        st->done = 1;
      }
      return;
    }
  }
  const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x) + 1;  // the pair (i, i + 1), i odd
  const int jl = blockIdx.y * blockDim.y + threadIdx.y + 1;
  double r = 0.0;
  if (i <= k.nx && jl <= k.nyl) {
    const int j = k.j0 + jl;
    const int t = ((i + j) & 1) == 0 ? i : i + 1;  // the colour-0 cell of the pair ((i + j) even, the oracle's first colour)
    if (t <= k.nx) {
      const double pc = p[pm_sidx(k, jl, t)], pw = p[pm_sidx(k, jl, t - 1)], pe = p[pm_sidx(k, jl, t + 1)];
      const double pn = p[pm_sidx(k, jl + 1, t)], ps = p[pm_sidx(k, jl - 1, t)];
      const double fc = f[pm_idx(k, jl, t)];
      if (FORM == 0) {  // rb_half, cavity tiles at a wall: a neighbour behind a wall is replaced by the cell itself
        const int ew = t > 1, ee = t < k.nx, en = j < k.ny;
        const double pw_ = ew ? pw : pc, pe_ = ee ? pe : pc, pn_ = en ? pn : pc;
        r = fma(k.idx2, fma(-4.0, pc, (pe_ + pn_) + (pw_ + ps)), -fc);
      } else {
        r = res_channel<Fast>(k, pc, pe, pw, pn, ps, fc);
      }
      r = fabs(r);
      if (!(r == r)) r = 0.0;  // a NaN never replaces the maximum (std::max(m, std::abs(x)))
    }
  }
  const double v = warp_max_nonneg(r);
  if ((tid & 31) == 0) wmax[tid >> 5] = (unsigned long long)__double_as_longlong(v);
  __syncthreads();
  if (tid == 0) {
    unsigned long long m = 0ull;
    for (int w = 0; w < int(blockDim.x * blockDim.y) / 32; ++w) m = max(m, wmax[w]);
    if (m != 0ull && m0 >= 1 && m0 <= k.max_iters)
      atomicMax(fold_part + size_t((blockIdx.y * gridDim.x + blockIdx.x) & (PM_FOLD_SLOTS - 1)) * 16 + (m0 & 7), m);
  }
}
