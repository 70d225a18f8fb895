// pm_common.cuh — device-side vocabulary shared by all kernels of libpm.so (sm_100a).
//
// Storage: every field (u, v, u*, v*, f, p ping, p pong) lives in one pitched FP64 plane
// with the SAME geometry, so neighbours of any field are +-1 and +-pitch in one index space
// (SURVEY H4).  Element (jl, i) of the local slab is at  (padr + jl) * pitch + offc + i  with
// offc = 15, so the first interior column i = 1 starts a 128-byte line; rows jl = 0 and
// jl = nyl+1 are the ghost rows (physical wall ghosts on the first/last rank, halo rows
// received from the neighbour slab otherwise); padr extra rows above and below hold the
// deeper halos of the temporally blocked pressure sweeps.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/pm.h"

#define PM_OFFC 15
#define PM_PADR 8

struct KP {
  // geometry
  int nx, ny;      // global interior
  int nyl, j0;     // local interior rows; global j = j0 + jl
  int pitch, padr; // doubles per row; pad rows
  int case_id, has_mask, first_rank, last_rank;
  int inlet_j_max;
  int psh;         // split-row layout of the tiled solve: column shift (0 or 2), see pm_split_col
  // constants, each computed on the host with the reference's own expression (see pm_capi.cu)
  double idx, idy, idx2, idy2;  // 1/dx, 1/dy, 1/(dx*dx), 1/(dy*dy)
  double hh;                    // cavity: grid_spacing*grid_spacing   (cavity-01.cpp:653)
  double nu, dt, uref, two_uref;
  double omega, om1;            // omega, 1.0 - omega
  double wnc[5];                // cavity: omega / neighbor_count      (cavity-01.cpp:651)
  double denom, rdenom;         // channel: 2*(idx2+idy2), and its reciprocal for the fast policy
  double cw;                    // fast policy, residual form of the relaxation: p += cw * r  (cavity: omega*h*h/4; channel: omega/denom)
  double cw3, cw2;              // cavity wall cells: omega*h*h / neighbor_count for 3 and 2 neighbours (cavity-01.cpp:644-651)
  double src_coef;              // cavity: (1/dt)*rho ; channel: rho/dt
  double cu, cv;                // correction coefficients
  double tol_factor, abs_tol;
  int max_iters;
  int fluid_count_global;       // cells entering the source mean (channel: nx*ny)
};

__host__ __device__ __forceinline__ size_t pm_idx(const KP& k, int jl, int i) {
  return size_t(k.padr + jl) * size_t(k.pitch) + size_t(PM_OFFC + i);
}

// Split-row layout of the pressure buffers of the tiled solve: within each row of `pitch` doubles the even
// storage columns come first, then the odd ones, so that one TMA box {64 pairs, 2 parities, rows} lands in
// shared memory in the order the sweeps exchange neighbours in.  Column c sits at
//   (c & 1) * pitch/2 + ((c + psh) >> 1)  (mod pitch/2),
// where the shift psh (0 or 2, chosen from the halo depth) makes the first pair of every tile even: TMA needs
// 16-byte aligned box rows.  The columns that wrap around are pad columns.
__host__ __device__ __forceinline__ int pm_split_col(const KP& k, int c) {
  const int half = k.pitch >> 1;
  int h = (c + k.psh) >> 1;
  if (h >= half) h -= half;
  return (c & 1) * half + h;
}
__host__ __device__ __forceinline__ size_t pm_sidx(const KP& k, int jl, int i) {
  return size_t(k.padr + jl) * size_t(k.pitch) + size_t(pm_split_col(k, PM_OFFC + i));
}

// ---- arithmetic policies -------------------------------------------------
// Exact: every operation individually IEEE-rounded, never contracted into an FMA
// (the __d*_rn intrinsics are documented as never merged) -> bit-identical to the
// oracle built with -ffp-contract=off.  Fast: plain operators, ptxas may fuse.
struct Exact {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static constexpr bool exact = true;
};
struct Fast {
  static __device__ __forceinline__ double add(double a, double b) { return a + b; }
  static __device__ __forceinline__ double sub(double a, double b) { return a - b; }
  static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
  static __device__ __forceinline__ double div(double a, double b) { return a / b; }
  static constexpr bool exact = false;
};

// ---- per-cell expression trees of the pressure solve ----------------------
// Cavity form, cavity-01.cpp:644-654.  (j, i) are GLOBAL indices.
template <class A>
__device__ __forceinline__ double upd_cavity(const KP& k, int j, int i, double pc, double pe, double pw,
                                             double pn, double ps, double f) {
  const int ew = i > 1, ee = i < k.nx, en = j < k.ny;
  const int nc = ew + ee + en + 1;
  if (A::exact) {
    const double a = A::add(A::mul(double(ee), pe), A::mul(double(ew), pw));
    const double b = A::add(A::mul(double(en), pn), ps);  // eps_s == 1: 1*x == x exactly
    const double s = A::sub(A::add(a, b), A::mul(f, k.hh));
    return A::add(A::mul(pc, k.om1), A::mul(k.wnc[nc], s));
  } else {
    const double a = (ee ? pe : 0.0) + (ew ? pw : 0.0);
    const double b = (en ? pn : 0.0) + ps;
    return pc * k.om1 + k.wnc[nc] * ((a + b) - f * k.hh);
  }
}
// Cavity residual, cavity-01.cpp:664-673.
template <class A>
__device__ __forceinline__ double res_cavity(const KP& k, int j, int i, double pc, double pe, double pw,
                                             double pn, double ps, double f, double h2i) {
  const int ew = i > 1, ee = i < k.nx, en = j < k.ny;
  if (A::exact) {
    double s = A::add(A::mul(double(ee), A::sub(pe, pc)), A::mul(double(ew), A::sub(pw, pc)));
    s = A::add(s, A::mul(double(en), A::sub(pn, pc)));
    s = A::add(s, A::sub(ps, pc));
    return A::sub(A::mul(h2i, s), f);
  } else {
    double s = (ee ? pe - pc : 0.0) + (ew ? pw - pc : 0.0);
    s += (en ? pn - pc : 0.0);
    s += ps - pc;
    return h2i * s - f;
  }
}
// Channel/step form, channel-01.cpp:659-666.
template <class A>
__device__ __forceinline__ double upd_channel(const KP& k, double pc, double pe, double pw, double pn, double ps,
                                              double f) {
  const double sum = A::add(A::mul(k.idx2, A::add(pe, pw)), A::mul(k.idy2, A::add(pn, ps)));
  const double pgs = A::exact ? A::div(A::sub(sum, f), k.denom) : (sum - f) * k.rdenom;
  return A::add(A::mul(k.om1, pc), A::mul(k.omega, pgs));
}
// Channel residual, channel-01.cpp:676-678.
template <class A>
__device__ __forceinline__ double res_channel(const KP& k, double pc, double pe, double pw, double pn, double ps,
                                              double f) {
  const double tc = A::mul(2.0, pc);
  const double lap = A::add(A::mul(A::add(A::sub(pe, tc), pw), k.idx2), A::mul(A::add(A::sub(pn, tc), ps), k.idy2));
  return A::sub(lap, f);
}

// ---- reductions: warp shuffles, then a block-level tree, then one atomic ----
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// All threads of the block must call.  Result valid in thread 0.
__device__ __forceinline__ double block_max(double v, double* sh /* >= 32 doubles */) {
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nth = blockDim.x * blockDim.y * blockDim.z;
  const int lane = tid & 31, w = tid >> 5, nw = (nth + 31) >> 5;
  v = warp_max(v);
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < nw ? sh[lane] : 0.0;
    v = warp_max(v);
  }
  __syncthreads();
  return v;
}
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nth = blockDim.x * blockDim.y * blockDim.z;
  const int lane = tid & 31, w = tid >> 5, nw = (nth + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < nw ? sh[lane] : 0.0;
    v = warp_sum(v);
  }
  __syncthreads();
  return v;
}
// max over non-negative doubles through their bit patterns (monotone for x >= 0; NaNs were
// dropped by fmax before, matching std::max(m, std::abs(x)) in the reference).
__device__ __forceinline__ void atomic_max_nonneg(unsigned long long* addr, double v) {
  if (v > 0.0) atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}

// ---- device-resident state of one pressure solve --------------------------
// res_bits[m] = bit pattern of max|r| of iterate m (m = 1..max_iters), accumulated by atomics.
struct PpeState {
  unsigned long long maxf_bits;  // max|f| after the source pass (pre-mean)
  unsigned long long maxf2_bits; // max|f| after mean removal (what the tolerance rule sees)
  double mean;                   // mean of f removed by the channel/step source pass
  double tol;
  double res_init;
  int done;        // sticky: the reference's while-condition became false
  int iters;       // iteration_count at that moment
  int kbase;       // iteration index base for graph-replayed launches
  int pad_;
  double ke_sum;   // diagnostics scratch
  unsigned long long div_bits;
  unsigned long long chain[2];  // exact source mean over slabs: running serial sum (bits) and cell count handed from rank to rank
};

// The reference's loop test `while (res > tol && it < max)` evaluated for entering iteration k
// (k >= 1) given the residual of iterate k-1.  Deterministic for every thread that evaluates it.
__device__ __forceinline__ bool ppe_stop_before(const PpeState* st, const unsigned long long* res_bits, int k, int max_iters) {
  if (st->done) return true;
  if (k > max_iters) return true;
  if (k >= 2) {
    const double r = __longlong_as_double((long long)res_bits[k - 1]);
    if (!(r > st->tol)) return true;
  }
  return false;
}

// splitmix64-keyed synthetic field, identical to oracle/ref_cpu.cpp synth().
__host__ __device__ __forceinline__ uint64_t pm_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ double pm_synth(uint64_t seed, int field, uint64_t flat) {
  const uint64_t z = pm_splitmix64(seed ^ pm_splitmix64((uint64_t(field) << 56) ^ flat));
  return double(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}
