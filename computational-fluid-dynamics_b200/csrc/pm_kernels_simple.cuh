// pm_kernels_simple.cuh — one kernel per reference loop, global-memory
// stencils with coalesced row access.  This is the general path: every case, the obstacle mask,
// any grid size, slabs.  The bandwidth path for large unmasked grids is pm_kernels_tiled.cuh.
#pragma once
#include "pm_common.cuh"

#define PM_BX 128
#define PM_BY 4

// ---------------------------------------------------------------------------
// k1  velocity BC ghost fill
// ---------------------------------------------------------------------------
// cavity-01.cpp:523-543.  One thread per ghost cell; reads only interior faces.
__global__ void k_bc_cavity(const __grid_constant__ KP k, double* __restrict__ U, double* __restrict__ V) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t <= k.nx) {
    if (k.last_rank) U[pm_idx(k, k.nyl + 1, t)] = k.two_uref - U[pm_idx(k, k.nyl, t)];
    if (k.first_rank) U[pm_idx(k, 0, t)] = -U[pm_idx(k, 1, t)];
  }
  if (t <= k.nyl) {  // v rows jl = 0..nyl (row 0 is a halo row on rank > 0; its ghosts follow the same rule)
    V[pm_idx(k, t, k.nx + 1)] = -V[pm_idx(k, t, k.nx)];
    V[pm_idx(k, t, 0)] = -V[pm_idx(k, t, 1)];
  }
}

// channel-01.cpp:513-529 and backwards_step-01.cpp:616-652 (walls; solid faces are k_bc_solid).
// The reference applies inlet/outlet columns first, then the wall rows for i = 0..nx, which read
// the just-set column values at the corners.  Here a row thread derives the
// corner value it would have read, so no thread depends on another.
__global__ void k_bc_channel(const __grid_constant__ KP k, double* __restrict__ U, double* __restrict__ V) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int nx = k.nx, nyl = k.nyl;
  // column phase: local rows
  if (t <= nyl + 1) {
    const int jl = t, j = k.j0 + jl;
    if (j >= 1 && j <= k.ny && jl >= 0) {  // u rows 1..ny (local rows 0..nyl+1 may be halo rows of a neighbour: same rule)
      const double uin = (k.case_id == PM_CASE_STEP && j > k.inlet_j_max) ? 0.0 : k.uref;
      U[pm_idx(k, jl, 0)] = uin;
      U[pm_idx(k, jl, nx)] = U[pm_idx(k, jl, nx - 1)];
    }
    if (j >= 0 && j <= k.ny && jl <= nyl) {  // v rows 0..ny
      V[pm_idx(k, jl, 0)] = 0.0;
      const double vo = V[pm_idx(k, jl, nx)];  // read BEFORE the wall rows zero v[0][nx] / v[ny][nx]
      V[pm_idx(k, jl, nx + 1)] = vo;
      if (j == 0 || j == k.ny) V[pm_idx(k, jl, nx)] = 0.0;  // the wall-row write for i = nx, issued by this thread to keep the order
    }
  }
  // row phase
  if (t <= nx) {
    const int i = t;
    if (k.first_rank) {
      if (i >= 1 && i < nx) V[pm_idx(k, 0, i)] = 0.0;
      double s;  // value of U[1][i] after the column phase
      if (i == 0) s = (k.case_id == PM_CASE_STEP && 1 > k.inlet_j_max) ? 0.0 : k.uref;
      else if (i == nx) s = U[pm_idx(k, 1, nx - 1)];
      else s = U[pm_idx(k, 1, i)];
      U[pm_idx(k, 0, i)] = -s;
    }
    if (k.last_rank) {
      if (i >= 1 && i < nx) V[pm_idx(k, nyl, i)] = 0.0;
      double s;
      if (i == 0) s = (k.case_id == PM_CASE_STEP && k.ny > k.inlet_j_max) ? 0.0 : k.uref;
      else if (i == nx) s = U[pm_idx(k, nyl, nx - 1)];
      else s = U[pm_idx(k, nyl, i)];
      U[pm_idx(k, nyl + 1, i)] = -s;
    }
  }
}

// backwards_step-01.cpp:654-682: zero the faces between a solid cell and a fluid neighbour.
// Runs after k_bc_channel (the wall ghosts read the pre-zero values, as in the reference).
// Slabs: local rows 0 and nyl+1 are the neighbour slabs' edge rows (their mask rows were uploaded with the
// halo): a solid cell there owns the v face below / above it that lies in THIS slab, so the halo rows are
// visited too; faces that would need mask rows beyond the halo are the neighbour's own business.
__global__ void k_bc_solid(const __grid_constant__ KP k, const uint8_t* __restrict__ M, double* __restrict__ U,
                           double* __restrict__ V) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int jl = blockIdx.y * blockDim.y + threadIdx.y;  // 0 .. nyl+1
  if (i > k.nx || jl > k.nyl + 1) return;
  const int j = k.j0 + jl;
  if (j < 1 || j > k.ny) return;  // the reference visits interior cells only
  if (M[pm_idx(k, jl, i)]) return;
  if (i < k.nx && M[pm_idx(k, jl, i + 1)]) U[pm_idx(k, jl, i)] = 0.0;
  if (i > 1 && M[pm_idx(k, jl, i - 1)]) U[pm_idx(k, jl, i - 1)] = 0.0;
  if (jl <= k.nyl && j < k.ny && M[pm_idx(k, jl + 1, i)]) V[pm_idx(k, jl, i)] = 0.0;
  if (jl >= 1 && j > 1 && M[pm_idx(k, jl - 1, i)]) V[pm_idx(k, jl - 1, i)] = 0.0;
}

// ---------------------------------------------------------------------------
// k2+k3  predictor: u* and v* in one pass (cavity-01.cpp:553-602; channel-01.cpp:553-602;
//        backwards_step-01.cpp:752-819)
// ---------------------------------------------------------------------------
// u* at the east face of cell (j, i).  uc = u[j][i], uw/ue_ = u[j][i-1] / u[j][i+1], uN/uS = u[j+1][i] / u[j-1][i];
// vc = v[j][i], ve_ = v[j][i+1], vS = v[j-1][i], vSE = v[j-1][i+1].
template <class A>
__device__ __forceinline__ double pred_u(const KP& k, double uc, double uw, double ue_, double uN, double uS, double vc, double ve_,
                                         double vS, double vSE) {
  const double t2 = A::mul(2.0, uc);
  const double diff = A::mul(k.nu, A::add(A::mul(A::add(A::sub(ue_, t2), uw), k.idx2), A::mul(A::add(A::sub(uN, t2), uS), k.idy2)));
  const double ue = A::mul(0.5, A::add(uc, ue_));
  const double uwf = A::mul(0.5, A::add(uw, uc));
  const double cx = A::mul(A::sub(A::mul(ue, ue), A::mul(uwf, uwf)), k.idx);
  const double vn = A::mul(0.5, A::add(vc, ve_));
  const double vsf = A::mul(0.5, A::add(vS, vSE));
  const double un = A::mul(0.5, A::add(uN, uc));
  const double usf = A::mul(0.5, A::add(uS, uc));
  const double cy = A::mul(A::sub(A::mul(vn, un), A::mul(vsf, usf)), k.idy);
  return A::add(uc, A::mul(k.dt, A::sub(A::sub(diff, cx), cy)));
}
// v* at the north face of cell (j, i).  vc = v[j][i], ve_/vw_ = v[j][i+1] / v[j][i-1], vN/vS = v[j+1][i] / v[j-1][i];
// uc = u[j][i], uN = u[j+1][i], uw = u[j][i-1], uNW = u[j+1][i-1].
template <class A>
__device__ __forceinline__ double pred_v(const KP& k, double vc, double ve_, double vw_, double vN, double vS, double uc, double uN,
                                         double uw, double uNW) {
  const double t2 = A::mul(2.0, vc);
  const double diff = A::mul(k.nu, A::add(A::mul(A::add(A::sub(ve_, t2), vw_), k.idx2), A::mul(A::add(A::sub(vN, t2), vS), k.idy2)));
  const double vn = A::mul(0.5, A::add(vc, vN));
  const double vsf = A::mul(0.5, A::add(vS, vc));
  const double cy = A::mul(A::sub(A::mul(vn, vn), A::mul(vsf, vsf)), k.idy);
  const double ue = A::mul(0.5, A::add(uc, uN));
  const double uwf = A::mul(0.5, A::add(uw, uNW));
  const double ve = A::mul(0.5, A::add(vc, ve_));
  const double vw = A::mul(0.5, A::add(vw_, vc));
  const double cx = A::mul(A::sub(A::mul(ue, ve), A::mul(uwf, vw)), k.idx);
  return A::add(vc, A::mul(k.dt, A::sub(A::sub(diff, cy), cx)));
}

// ---------------------------------------------------------------------------
// Row kernels (every case): two cells per thread -- columns i, i+1 with i odd, so the
// pair starts at an even storage column -- every field moved with 128-bit loads and stores, the same per-cell
// expression trees as the general kernels.  Block 128 x 2 threads = 256 columns x 2 rows.
// ---------------------------------------------------------------------------
#define PM_RX 128
#define PM_RY 2
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

// MASK: the obstacle mask of the step case (backwards_step-01.cpp:755-761,790-796: a face between two solid cells is 0).
template <class A, bool MASK>
__global__ void __launch_bounds__(PM_RX* PM_RY)
    k_predict_rows(const __grid_constant__ KP k, const double* __restrict__ u, const double* __restrict__ v, const uint8_t* __restrict__ M,
                   double* __restrict__ us, double* __restrict__ vs) {
  const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x) + 1;
  const int jl = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i > k.nx || jl > k.nyl) return;
  const int j = k.j0 + jl, P = k.pitch;
  const size_t c = pm_idx(k, jl, i);
  // the reads past column nx (u) / nx+1 (v) land in the zeroed pad columns of the plane and are never used
  const double2 U = ld2(u + c), UN = ld2(u + c + P), US = ld2(u + c - P);
  const double2 V = ld2(v + c), VN = ld2(v + c + P), VS = ld2(v + c - P);
  const double uW = u[c - 1], uE2 = u[c + 2], uNW = u[c + P - 1];
  const double vW = v[c - 1], vE2 = v[c + 2], vSE2 = v[c - P + 2];
  const bool a_u = i <= k.nx - 1, b_u = i + 1 <= k.nx - 1;  // u* columns 1..nx-1
  const bool a_v = j <= k.ny - 1, b_v = a_v && i + 1 <= k.nx;  // v* rows 1..ny-1, columns 1..nx
  bool ua_on = true, ub_on = true, va_on = true, vb_on = true;
  if (MASK) {  // PM_OFFC + i is even: the pair's mask bytes sit on a 2-byte boundary
    const uchar2 mc = *reinterpret_cast<const uchar2*>(M + c), mn = *reinterpret_cast<const uchar2*>(M + c + P);
    const uint8_t me2 = M[c + 2];
    ua_on = mc.x || mc.y; ub_on = mc.y || me2;
    va_on = mc.x || mn.x; vb_on = mc.y || mn.y;
  }
  const double ua = ua_on ? pred_u<A>(k, U.x, uW, U.y, UN.x, US.x, V.x, V.y, VS.x, VS.y) : 0.0;
  const double ub = ub_on ? pred_u<A>(k, U.y, U.x, uE2, UN.y, US.y, V.y, vE2, VS.y, vSE2) : 0.0;
  if (a_u && b_u) st2(us + c, ua, ub);
  else if (a_u) us[c] = ua;
  if (a_v) {
    const double va = va_on ? pred_v<A>(k, V.x, V.y, vW, VN.x, VS.x, U.x, UN.x, uW, uNW) : 0.0;
    const double vb = vb_on ? pred_v<A>(k, V.y, vE2, V.x, VN.y, VS.y, U.y, UN.y, U.x, UN.x) : 0.0;
    if (b_v) st2(vs + c, va, vb);
    else vs[c] = va;
  }
}

// ---------------------------------------------------------------------------
// k4  divergence source + max|f|  (cavity-01.cpp:622-630; channel-01.cpp:613-619;
//     backwards_step-01.cpp:830-841).  Also per-block partial sums for the mean.
// ---------------------------------------------------------------------------
// Row version (see k_predict_rows).  MASK: solid cells get f = 0 and stay out of the maximum and the sum.
template <class A, bool MASK>
__global__ void __launch_bounds__(PM_RX* PM_RY)
    k_source_rows(const __grid_constant__ KP k, const double* __restrict__ us, const double* __restrict__ vs, const uint8_t* __restrict__ M,
                  double* __restrict__ f, PpeState* __restrict__ st, double* __restrict__ partial /* one per block, or null */) {
  __shared__ double sh[32];
  const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x) + 1;
  const int jl = blockIdx.y * blockDim.y + threadIdx.y + 1;
  double a = 0.0, sum = 0.0;
  if (i <= k.nx && jl <= k.nyl) {
    const size_t c = pm_idx(k, jl, i);
    const double2 U = ld2(us + c), V = ld2(vs + c), VS = ld2(vs + c - k.pitch);
    const double uW = us[c - 1];
    bool fa_on = true, fb_on = true;
    if (MASK) {
      const uchar2 mc = *reinterpret_cast<const uchar2*>(M + c);
      fa_on = mc.x; fb_on = mc.y;
    }
    const double fa = fa_on ? A::mul(k.src_coef, A::add(A::mul(A::sub(U.x, uW), k.idx), A::mul(A::sub(V.x, VS.x), k.idy))) : 0.0;
    if (i + 1 <= k.nx) {
      const double fb = fb_on ? A::mul(k.src_coef, A::add(A::mul(A::sub(U.y, U.x), k.idx), A::mul(A::sub(V.y, VS.y), k.idy))) : 0.0;
      st2(f + c, fa, fb);
      a = fmax(fabs(fa), fabs(fb));
      sum = fa + fb;
    } else {
      f[c] = fa;
      a = fabs(fa);
      sum = fa;
    }
  }
  const double m = block_max(a, sh);
  if (threadIdx.x == 0 && threadIdx.y == 0) atomic_max_nonneg(&st->maxf_bits, m);
  if (partial) {
    const double s = block_sum(sum, sh);
    if (threadIdx.x == 0 && threadIdx.y == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = s;
  }
}

// f -= mean on (fluid) cells when max|f| > 0, and max|f| of the result (channel-01.cpp:621-628, :643-646): row version.
template <class A, bool MASK>
__global__ void __launch_bounds__(PM_RX* PM_RY)
    k_sub_mean_rows(const __grid_constant__ KP k, double* __restrict__ f, const uint8_t* __restrict__ M, PpeState* __restrict__ st) {
  __shared__ double sh[32];
  const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x) + 1;
  const int jl = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const bool apply = st->maxf_bits != 0ull;
  const double mean = st->mean;
  double a = 0.0;
  if (i <= k.nx && jl <= k.nyl) {
    const size_t c = pm_idx(k, jl, i);
    const bool two = i + 1 <= k.nx;
    bool xa = true, xb = two;
    if (MASK) {
      const uchar2 mc = *reinterpret_cast<const uchar2*>(M + c);
      xa = mc.x; xb = two && mc.y;
    }
    double2 F = two ? ld2(f + c) : make_double2(f[c], 0.0);
    if (apply) {
      if (xa) F.x = A::sub(F.x, mean);
      if (xb) F.y = A::sub(F.y, mean);
      if (two) st2(f + c, F.x, F.y);
      else f[c] = F.x;
    }
    a = fmax(xa ? fabs(F.x) : 0.0, xb ? fabs(F.y) : 0.0);
  }
  const double m = block_max(a, sh);
  if (threadIdx.x == 0 && threadIdx.y == 0) atomic_max_nonneg(&st->maxf2_bits, m);
}

// Predictor + source in one pass for the cavity, where nothing happens between the two (cavity-01.cpp:388-389 and
// :622-630; channel/step apply boundary conditions to u*, v* in between): u*, v* are written as k_predict_rows writes
// them and f is formed from the values still in registers, so u*, v* are not read back (16 B per cell less).
// A thread owns the columns i, i+1 (i odd) of PM_FUSE_ROWS consecutive rows: v* of the row below stays in registers,
// u* of the column to the west comes through shared memory (the first thread of a block recomputes it), and the
// row below the strip is recomputed once per strip.  Boundary faces the predictor never writes (u*[j][0], u*[j][nx],
// v*[0][i], v*[ny][i]) are read from memory, exactly as the separate source pass would see them.
#define PM_FUSE_ROWS 16
// FORM 1 (channel, no obstacle mask): the reference applies the boundary conditions to u*, v* between the two passes
// (channel-01.cpp:368-371); the source only sees four of the values they set -- the inlet face u*[j][0] = U, the outlet
// face u*[j][nx] = u*[j][nx-1], the wall faces v*[0][i] = v*[ny][i] = 0 (channel-01.cpp:513-529) -- and takes them inline;
// k_bc_channel still runs behind this kernel for the stored u*, v*.  `partial`: one sum of f per block for the source mean.
template <class A, int FORM>
__global__ void __launch_bounds__(PM_RX)
    k_predict_source(const __grid_constant__ KP k, const double* __restrict__ u, const double* __restrict__ v,
                            double* __restrict__ us, double* __restrict__ vs, double* __restrict__ f, PpeState* __restrict__ st,
                            double* __restrict__ partial, double* __restrict__ fsplit /* f again in the split-row layout, or null */) {
  __shared__ double s_ub[2][PM_RX];
  __shared__ double sh[32];
  const int tid = threadIdx.x;
  const int i = 2 * (blockIdx.x * blockDim.x + tid) + 1;
  const int jl0 = blockIdx.y * PM_FUSE_ROWS + 1;
  const int P = k.pitch;
  const bool has = i <= k.nx, has_b = i + 1 <= k.nx;
  const bool a_u = i <= k.nx - 1, b_u = i + 1 <= k.nx - 1;  // columns whose u* the predictor writes
  double a_max = 0.0, f_sum = 0.0;
  double va_prev = 0.0, vb_prev = 0.0;  // v* of the row below the current one
  if (has) {
    const int jl = jl0 - 1, j = k.j0 + jl;
    const size_t c = pm_idx(k, jl, i);
    if (j == 0) {  // bottom wall faces: never written by the predictor
      if (FORM == 0) {
        const double2 Vm = ld2(vs + c);
        va_prev = Vm.x; vb_prev = Vm.y;
      }  // FORM 1: v*[0][i] = 0
    } else {       // last row of the strip below (another block writes it): same trees, same operands
      const double2 U = ld2(u + c), UN = ld2(u + c + P);
      const double2 V = ld2(v + c), VN = ld2(v + c + P), VS = ld2(v + c - P);
      const double uW = u[c - 1], uNW = u[c + P - 1], vW = v[c - 1], vE2 = v[c + 2];
      va_prev = pred_v<A>(k, V.x, V.y, vW, VN.x, VS.x, U.x, UN.x, uW, uNW);
      vb_prev = pred_v<A>(k, V.y, vE2, V.x, VN.y, VS.y, U.y, UN.y, U.x, UN.x);
    }
  }
  for (int r = 0; r < PM_FUSE_ROWS; ++r) {
    const int jl = jl0 + r;
    if (jl > k.nyl) break;  // uniform over the block
    const int j = k.j0 + jl;
    const size_t c = pm_idx(k, jl, i);
    double ua = 0.0, ub = 0.0, va = 0.0, vb = 0.0, uw = 0.0;
    if (has) {
      const double2 U = ld2(u + c), UN = ld2(u + c + P), US = ld2(u + c - P);
      const double2 V = ld2(v + c), VN = ld2(v + c + P), VS = ld2(v + c - P);
      const double uW = u[c - 1], uE2 = u[c + 2], uNW = u[c + P - 1];
      const double vW = v[c - 1], vE2 = v[c + 2], vSE2 = v[c - P + 2];
      const bool a_v = j <= k.ny - 1;
      // u*: computed where the predictor writes, memory elsewhere (east wall face of the last column)
      if (a_u) ua = pred_u<A>(k, U.x, uW, U.y, UN.x, US.x, V.x, V.y, VS.x, VS.y);
      else if (FORM == 0) ua = us[c];  // FORM 1: the outlet face takes the value of the face to its west, below
      if (b_u) ub = pred_u<A>(k, U.y, U.x, uE2, UN.y, US.y, V.y, vE2, VS.y, vSE2);
      else if (has_b) ub = FORM == 0 ? us[c + 1] : ua;
      if (a_u && b_u) st2(us + c, ua, ub);
      else if (a_u) us[c] = ua;
      if (a_v) {
        va = pred_v<A>(k, V.x, V.y, vW, VN.x, VS.x, U.x, UN.x, uW, uNW);
        vb = pred_v<A>(k, V.y, vE2, V.x, VN.y, VS.y, U.y, UN.y, U.x, UN.x);
        if (has_b) st2(vs + c, va, vb);
        else vs[c] = va;
      } else if (FORM == 0) {  // top wall faces
        const double2 Vm = ld2(vs + c);
        va = Vm.x; vb = Vm.y;
      }  // FORM 1: v*[ny][i] = 0
      if (tid == 0) {  // u* of the column to the west of the block
        if (i == 1) uw = FORM == 0 ? us[c - 1] : k.uref;  // west wall face / inlet face
        else uw = pred_u<A>(k, uW, u[c - 2], U.x, uNW, u[c - P - 1], vW, V.x, v[c - P - 1], VS.x);
      }
    }
    s_ub[r & 1][tid] = ub;
    __syncthreads();
    if (has) {
      if (tid > 0) uw = s_ub[r & 1][tid - 1];
      if (FORM == 1 && !a_u) ua = uw;  // outlet face
      const double fa = A::mul(k.src_coef, A::add(A::mul(A::sub(ua, uw), k.idx), A::mul(A::sub(va, va_prev), k.idy)));
      if (has_b) {
        const double fb = A::mul(k.src_coef, A::add(A::mul(A::sub(ub, ua), k.idx), A::mul(A::sub(vb, vb_prev), k.idy)));
        st2(f + c, fa, fb);
        a_max = fmax(a_max, fmax(fabs(fa), fabs(fb)));
        f_sum += fa + fb;
        if (fsplit) fsplit[pm_sidx(k, jl, i + 1)] = fb;
      } else {
        f[c] = fa;
        a_max = fmax(a_max, fabs(fa));
        f_sum += fa;
      }
      // the streaming pressure pass reads f in the layout of its p buffers (pm_common.cuh): saves a permutation pass
      if (fsplit) fsplit[pm_sidx(k, jl, i)] = fa;

      va_prev = va;
      vb_prev = vb;
    }
  }
  const double m = block_max(a_max, sh);
  if (tid == 0) atomic_max_nonneg(&st->maxf_bits, m);
  if (partial) {
    const double sum = block_sum(f_sum, sh);
    if (tid == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = sum;
  }
}

// k5 (fast policy): fixed-shape tree over the per-block partial sums, one block.
__global__ void k_mean_from_partials(const double* __restrict__ partial, int n, int count, PpeState* __restrict__ st) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int t = threadIdx.x; t < n; t += blockDim.x) s += partial[t];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) st->mean = count > 0 ? s / double(count) : 0.0;
}
__global__ void k_mean_from_sum(int count, PpeState* __restrict__ st) { st->mean = count > 0 ? st->ke_sum / double(count) : 0.0; }
// k5 (exact policy): the reference's serial row-major sum (channel-01.cpp:622-625), one warp,
// every lane adding the same 32 shuffled values in index order.  Slabs: the running sum and count enter
// through st->chain (zero on the first rank) and leave through it, so the chain of ranks performs the
// reference's additions in the reference's order; the last rank's mean is the global one.
__global__ void k_mean_serial(const __grid_constant__ KP k, const double* __restrict__ f, const uint8_t* __restrict__ M,
                              PpeState* __restrict__ st, int chained) {
  const int lane = threadIdx.x;
  double s = chained ? __longlong_as_double((long long)st->chain[0]) : 0.0;
  long long cnt = chained ? (long long)st->chain[1] : 0;
  for (int jl = 1; jl <= k.nyl; ++jl)
    for (int ib = 1; ib <= k.nx; ib += 32) {
      const int i = ib + lane;
      double x = 0.0;
      int use = 0;
      if (i <= k.nx) {
        const size_t c = pm_idx(k, jl, i);
        use = !k.has_mask || M[c];
        x = f[c];
      }
      const int n = min(32, k.nx - ib + 1);
      for (int q = 0; q < n; ++q) {
        const double xq = __shfl_sync(0xffffffffu, x, q);
        const int uq = __shfl_sync(0xffffffffu, use, q);
        if (uq) { s = __dadd_rn(s, xq); ++cnt; }
      }
    }
  if (lane == 0) {
    st->mean = cnt > 0 ? __ddiv_rn(s, double(cnt)) : 0.0;
    st->chain[0] = (unsigned long long)__double_as_longlong(s);
    st->chain[1] = (unsigned long long)cnt;
  }
}
// max|f| only (cavity, or when pm_ppe_solve is called on an uploaded f).
__global__ void __launch_bounds__(PM_BX* PM_BY)
    k_max_f(const __grid_constant__ KP k, const double* __restrict__ f, const uint8_t* __restrict__ M,
            PpeState* __restrict__ st) {
  __shared__ double sh[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int jl = blockIdx.y * blockDim.y + threadIdx.y + 1;
  double a = 0.0;
  if (i <= k.nx && jl <= k.nyl) {
    const size_t c = pm_idx(k, jl, i);
    if (!k.has_mask || M[c]) a = fabs(f[c]);
  }
  const double m = block_max(a, sh);
  if (threadIdx.x == 0 && threadIdx.y == 0) atomic_max_nonneg(&st->maxf2_bits, m);
}

// k6: tolerance and loop entry (cavity-01.cpp:618,632; channel-01.cpp:647-649). One thread.
__global__ void k_ppe_begin(const __grid_constant__ KP k, PpeState* __restrict__ st) {
  const double mx = __longlong_as_double((long long)st->maxf2_bits);
  double tol, r0;
  if (k.case_id == PM_CASE_CAVITY) {
    tol = __dmul_rn(k.tol_factor, mx);
    r0 = 1.0;
  } else {
    tol = fmax(__dmul_rn(k.tol_factor, (mx > 0 ? mx : 1.0)), k.abs_tol);
    r0 = __dadd_rn(tol, 1.0);
  }
  st->tol = tol;
  st->res_init = r0;
  st->iters = 0;
  st->done = !(r0 > tol) || !(0 < k.max_iters);
  st->kbase = 0;
}

// ---------------------------------------------------------------------------
// k7  one colour of a red-black SOR sweep, in place.  colour = (i + j) & 1 with GLOBAL j.
//     The thread that updates a wall-adjacent cell also refreshes the ghost behind it
//     (channel-01.cpp:531-541): that ghost is read by no other thread.
// ---------------------------------------------------------------------------
template <class A, int FORM /*0 cavity, 1 channel*/>
__global__ void __launch_bounds__(PM_BX* PM_BY)
    k_rb_colour(const __grid_constant__ KP k, double* __restrict__ p, const double* __restrict__ f,
                const uint8_t* __restrict__ M, PpeState* __restrict__ st, const unsigned long long* __restrict__ res_bits,
                int kiter_rel, int colour, int first_of_iter, int fuse_ghosts) {
  const int kiter = st->kbase + kiter_rel;
  if (ppe_stop_before(st, res_bits, kiter, k.max_iters)) {
    if (first_of_iter && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0 && !st->done &&
        kiter <= k.max_iters) {
      st->iters = kiter - 1;
      st->done = 1;
    }
    return;
  }
  const int jl = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (jl > k.nyl) return;
  const int j = k.j0 + jl;
  const int i = 1 + ((colour + j + 1) & 1) + 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (i > k.nx) return;
  const size_t c = pm_idx(k, jl, i);
  if (k.has_mask && !M[c]) return;
  const int P = k.pitch;
  const double pc = p[c], pe = p[c + 1], pw = p[c - 1], pn = p[c + P], ps = p[c - P], fc = f[c];
  double r;
  if (FORM == 0) r = upd_cavity<A>(k, j, i, pc, pe, pw, pn, ps, fc);
  else r = upd_channel<A>(k, pc, pe, pw, pn, ps, fc);
  p[c] = r;
  if (FORM == 1 && fuse_ghosts) {
    if (i == 1) p[c - 1] = r;
    if (i == k.nx) p[c + 1] = 0.0;
    if (j == 1) p[c - P] = r;
    if (j == k.ny) p[c + P] = r;
  }
}

// Jacobi sweep src -> dst (every neighbour from the previous iterate).
template <class A, int FORM>
__global__ void __launch_bounds__(PM_BX* PM_BY)
    k_jacobi(const __grid_constant__ KP k, const double* __restrict__ src, double* __restrict__ dst,
             const double* __restrict__ f, const uint8_t* __restrict__ M, PpeState* __restrict__ st,
             const unsigned long long* __restrict__ res_bits, int kiter_rel, int fuse_ghosts) {
  const int kiter = st->kbase + kiter_rel;
  if (ppe_stop_before(st, res_bits, kiter, k.max_iters)) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0 && !st->done && kiter <= k.max_iters) {
      st->iters = kiter - 1;
      st->done = 1;
    }
    return;
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int jl = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i > k.nx || jl > k.nyl) return;
  const int j = k.j0 + jl;
  const size_t c = pm_idx(k, jl, i);
  const int P = k.pitch;
  const double pc = src[c];
  double r = pc;
  if (!k.has_mask || M[c]) {
    const double pe = src[c + 1], pw = src[c - 1], pn = src[c + P], ps = src[c - P], fc = f[c];
    if (FORM == 0) r = upd_cavity<A>(k, j, i, pc, pe, pw, pn, ps, fc);
    else r = upd_channel<A>(k, pc, pe, pw, pn, ps, fc);
  }
  dst[c] = r;
  if (FORM == 1 && fuse_ghosts) {
    if (i == 1) dst[c - 1] = r;
    if (i == k.nx) dst[c + 1] = 0.0;
    if (j == 1) dst[c - P] = r;
    if (j == k.ny) dst[c + P] = r;
  }
}

// k8 (mask case): wall ghosts as their own pass, then solid-cell extrapolation
// (backwards_step-01.cpp:685-740).  The wall ghosts read the pre-extrapolation values.
// split: p is a buffer of the tiled solve (split-row layout); gated: part of the iteration loop (skipped once the
// reference's loop test has ended it) -- the tiled solve calls both kernels once, ungated, behind its last pass.
__global__ void k_pghost_walls(const __grid_constant__ KP k, double* __restrict__ p, PpeState* __restrict__ st,
                               const unsigned long long* __restrict__ res_bits, int kiter_rel, int split, int gated) {
  if (gated && ppe_stop_before(st, res_bits, st->kbase + kiter_rel, k.max_iters)) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x + 1;
  auto at = [&](int jl, int i) -> double& { return p[split ? pm_sidx(k, jl, i) : pm_idx(k, jl, i)]; };
  if (t <= k.nyl) {
    at(t, 0) = at(t, 1);
    at(t, k.nx + 1) = 0.0;
  }
  if (t <= k.nx) {
    if (k.first_rank) at(0, t) = at(1, t);
    if (k.last_rank) at(k.nyl + 1, t) = at(k.nyl, t);
  }
}
__global__ void __launch_bounds__(PM_BX* PM_BY)
    k_pghost_solid(const __grid_constant__ KP k, double* __restrict__ p, const uint8_t* __restrict__ M,
                   PpeState* __restrict__ st, const unsigned long long* __restrict__ res_bits, int kiter_rel, int split, int gated) {
  if (gated && ppe_stop_before(st, res_bits, st->kbase + kiter_rel, k.max_iters)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int jl = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i > k.nx || jl > k.nyl) return;
  const int j = k.j0 + jl;
  const size_t c = pm_idx(k, jl, i);
  if (M[c]) return;
  const int P = k.pitch;
  auto at = [&](int jj, int ii) -> double& { return p[split ? pm_sidx(k, jj, ii) : pm_idx(k, jj, ii)]; };
  double s = 0.0;
  int n = 0;
  if (i > 1 && M[c - 1]) { s = __dadd_rn(s, at(jl, i - 1)); ++n; }
  if (i < k.nx && M[c + 1]) { s = __dadd_rn(s, at(jl, i + 1)); ++n; }
  if (j > 1 && M[c - P]) { s = __dadd_rn(s, at(jl - 1, i)); ++n; }
  if (j < k.ny && M[c + P]) { s = __dadd_rn(s, at(jl + 1, i)); ++n; }
  if (n > 0) at(jl, i) = __ddiv_rn(s, double(n));
}

// k9  residual max-norm of the current iterate -> res_bits[kiter]
//     (cavity-01.cpp:659-677; channel-01.cpp:673-681; backwards_step-01.cpp:917-930)
template <class A, int FORM>
__global__ void __launch_bounds__(PM_BX* PM_BY)
    k_residual(const __grid_constant__ KP k, const double* __restrict__ p, const double* __restrict__ f,
               const uint8_t* __restrict__ M, PpeState* __restrict__ st, unsigned long long* __restrict__ res_bits,
               int kiter_rel, int force) {
  __shared__ double sh[32];
  const int kiter = st->kbase + kiter_rel;
  if (!force && ppe_stop_before(st, res_bits, kiter, k.max_iters)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int jl = blockIdx.y * blockDim.y + threadIdx.y + 1;
  double a = 0.0;
  if (i <= k.nx && jl <= k.nyl) {
    const size_t c = pm_idx(k, jl, i);
    if (!k.has_mask || M[c]) {
      const int P = k.pitch;
      const double pc = p[c], pe = p[c + 1], pw = p[c - 1], pn = p[c + P], ps = p[c - P], fc = f[c];
      double r;
      if (FORM == 0) r = res_cavity<A>(k, k.j0 + jl, i, pc, pe, pw, pn, ps, fc, k.idx2);
      else r = res_channel<A>(k, pc, pe, pw, pn, ps, fc);
      a = fabs(r);
    }
  }
  const double m = block_max(a, sh);
  if (threadIdx.x == 0 && threadIdx.y == 0) atomic_max_nonneg(&res_bits[kiter], m);
}

// ---------------------------------------------------------------------------
// k10  velocity correction (cavity-01.cpp:695-711; channel-01.cpp:693-702;
//      backwards_step-01.cpp:944-976)
// ---------------------------------------------------------------------------
// Row version (see k_predict_rows).  psplit: p is the final buffer of the tiled solve in its split-row layout, where the
// pair (i, i+1) is one double in each half of the row.  MASK: faces between two solid cells are 0, except the last u
// column and the last v row, which the reference always corrects (backwards_step-01.cpp:950-975).
template <class A, bool MASK>
__global__ void __launch_bounds__(PM_RX* PM_RY)
    k_correct_rows(const __grid_constant__ KP k, const double* __restrict__ us, const double* __restrict__ vs,
                   const double* __restrict__ p, const uint8_t* __restrict__ M, double* __restrict__ u, double* __restrict__ v, int psplit) {
  const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x) + 1;
  const int jl = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i > k.nx || jl > k.nyl) return;
  const int j = k.j0 + jl, P = k.pitch;
  const size_t c = pm_idx(k, jl, i);
  double pa, pb, pe2, pna, pnb;
  if (psplit) {
    const size_t sa = pm_sidx(k, jl, i), sb = pm_sidx(k, jl, i + 1);
    pa = p[sa]; pb = p[sb]; pe2 = p[pm_sidx(k, jl, i + 2)];
    pna = p[sa + P]; pnb = p[sb + P];
  } else {
    const double2 Pc = ld2(p + c), Pn = ld2(p + c + P);
    pa = Pc.x; pb = Pc.y; pe2 = p[c + 2];
    pna = Pn.x; pnb = Pn.y;
  }
  const bool a_u = i <= k.nx - 1, b_u = i + 1 <= k.nx - 1;
  const bool a_v = j <= k.ny - 1, b_v = a_v && i + 1 <= k.nx;
  bool ua_on = true, ub_on = true, va_on = true, vb_on = true;
  if (MASK) {
    const uchar2 mc = *reinterpret_cast<const uchar2*>(M + c), mn = *reinterpret_cast<const uchar2*>(M + c + P);
    const uint8_t me2 = M[c + 2];
    ua_on = (i == k.nx - 1) || mc.x || mc.y; ub_on = (i + 1 == k.nx - 1) || mc.y || me2;
    va_on = (j == k.ny - 1) || mc.x || mn.x; vb_on = (j == k.ny - 1) || mc.y || mn.y;
  }
  if (a_u) {
    const double2 U = ld2(us + c);
    const double ua = ua_on ? A::sub(U.x, A::mul(k.cu, A::sub(pb, pa))) : 0.0;
    if (b_u) st2(u + c, ua, ub_on ? A::sub(U.y, A::mul(k.cu, A::sub(pe2, pb))) : 0.0);
    else u[c] = ua;
  }
  if (a_v) {
    const double2 V = ld2(vs + c);
    const double va = va_on ? A::sub(V.x, A::mul(k.cv, A::sub(pna, pa))) : 0.0;
    if (b_v) st2(v + c, va, vb_on ? A::sub(V.y, A::mul(k.cv, A::sub(pnb, pb))) : 0.0);
    else v[c] = va;
  }
}

// ---------------------------------------------------------------------------
// k11  diagnostics (cavity-01.cpp:741-766; channel-01.cpp:733-759)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(PM_BX* PM_BY)
    k_diag(const __grid_constant__ KP k, const double* __restrict__ u, const double* __restrict__ v,
           const uint8_t* __restrict__ M, PpeState* __restrict__ st, double* __restrict__ partial) {
  __shared__ double sh[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int jl = blockIdx.y * blockDim.y + threadIdx.y + 1;
  double ke = 0.0, d = 0.0;
  if (i <= k.nx && jl <= k.nyl) {
    const size_t c = pm_idx(k, jl, i);
    if (!k.has_mask || M[c]) {
      const double uc = __dmul_rn(0.5, __dadd_rn(u[c - 1], u[c]));
      const double vc = __dmul_rn(0.5, __dadd_rn(v[c - k.pitch], v[c]));
      ke = __dmul_rn(0.5, __dadd_rn(__dmul_rn(uc, uc), __dmul_rn(vc, vc)));
      if (k.case_id == PM_CASE_CAVITY)
        d = __dmul_rn(__dsub_rn(__dadd_rn(__dsub_rn(u[c], u[c - 1]), v[c]), v[c - k.pitch]), k.idx);
      else
        d = __dadd_rn(__dmul_rn(__dsub_rn(u[c], u[c - 1]), k.idx), __dmul_rn(__dsub_rn(v[c], v[c - k.pitch]), k.idy));
      d = fabs(d);
    }
  }
  const double m = block_max(d, sh);
  if (threadIdx.x == 0 && threadIdx.y == 0) atomic_max_nonneg(&st->div_bits, m);
  const double s = block_sum(ke, sh);
  if (threadIdx.x == 0 && threadIdx.y == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = s;
}
// Export-ready fields of the VTK writers, one dense row-major array of nyl x nx doubles each (this rank's rows):
// interpolateToCellCenters (cavity-01.cpp:717-733; channel-01.cpp:708-731; backwards_step-01.cpp:981-1009: solid cells stay 0),
// the magnitude sqrt(uc*uc + vc*vc) and the vorticity loops of the writers with their own expression trees
// (cavity-01.cpp:187-223: (diff * dx_inv) * 0.5 inside, one-sided at the edges; channel-01.cpp:171-182: (0.5 * diff) * idx;
// backwards_step-01.cpp:203-236: only where the cell and its four neighbours are fluid and off the domain edge, else 0).
// Every operation individually rounded, like the host code it replaces (-ffp-contract=off).  p: natural plane or split-row buffer.
__global__ void __launch_bounds__(PM_BX* PM_BY)
    k_export(const __grid_constant__ KP k, const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ p, int psplit,
             const uint8_t* __restrict__ M, double* __restrict__ out, size_t field_stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int jl = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i > k.nx || jl > k.nyl) return;
  const int P = k.pitch, j = k.j0 + jl;
  const size_t c = pm_idx(k, jl, i);
  const bool mask = k.has_mask != 0;
  auto fluid = [&](size_t q) { return !mask || M[q] != 0; };
  auto ucen = [&](size_t q) { return fluid(q) ? __dmul_rn(0.5, __dadd_rn(u[q - 1], u[q])) : 0.0; };
  auto vcen = [&](size_t q) { return fluid(q) ? __dmul_rn(0.5, __dadd_rn(v[q - P], v[q])) : 0.0; };
  const bool fl = fluid(c);
  const double uc = ucen(c), vc = vcen(c);
  double vort = 0.0;
  if (k.case_id == PM_CASE_CAVITY) {
    double dvdx, dudy;
    if (i == 1) dvdx = __dmul_rn(__dsub_rn(vcen(c + 1), vc), k.idx);
    else if (i == k.nx) dvdx = __dmul_rn(__dsub_rn(vc, vcen(c - 1)), k.idx);
    else dvdx = __dmul_rn(__dmul_rn(__dsub_rn(vcen(c + 1), vcen(c - 1)), k.idx), 0.5);
    if (j == 1) dudy = __dmul_rn(__dsub_rn(ucen(c + P), uc), k.idx);
    else if (j == k.ny) dudy = __dmul_rn(__dsub_rn(uc, ucen(c - P)), k.idx);
    else dudy = __dmul_rn(__dmul_rn(__dsub_rn(ucen(c + P), ucen(c - P)), k.idx), 0.5);
    vort = __dsub_rn(dvdx, dudy);
  } else if (k.case_id == PM_CASE_CHANNEL) {
    double dvdx, dudy;
    if (i == 1) dvdx = __dmul_rn(__dsub_rn(vcen(c + 1), vc), k.idx);
    else if (i == k.nx) dvdx = __dmul_rn(__dsub_rn(vc, vcen(c - 1)), k.idx);
    else dvdx = __dmul_rn(__dmul_rn(0.5, __dsub_rn(vcen(c + 1), vcen(c - 1))), k.idx);
    if (j == 1) dudy = __dmul_rn(__dsub_rn(ucen(c + P), uc), k.idy);
    else if (j == k.ny) dudy = __dmul_rn(__dsub_rn(uc, ucen(c - P)), k.idy);
    else dudy = __dmul_rn(__dmul_rn(0.5, __dsub_rn(ucen(c + P), ucen(c - P))), k.idy);
    vort = __dsub_rn(dvdx, dudy);
  } else {
    const bool ok = fl && !(i == 1 || i == k.nx || j == 1 || j == k.ny) && M[c - 1] && M[c + 1] && M[c - P] && M[c + P];
    if (ok) {
      const double dvdx = __dmul_rn(__dmul_rn(0.5, __dsub_rn(vcen(c + 1), vcen(c - 1))), k.idx);
      const double dudy = __dmul_rn(__dmul_rn(0.5, __dsub_rn(ucen(c + P), ucen(c - P))), k.idy);
      vort = __dsub_rn(dvdx, dudy);
    }
  }
  const size_t o = size_t(jl - 1) * size_t(k.nx) + size_t(i - 1);
  out[o] = uc;
  out[o + field_stride] = vc;
  out[o + 2 * field_stride] = fl ? __dsqrt_rn(__dadd_rn(__dmul_rn(uc, uc), __dmul_rn(vc, vc))) : 0.0;
  out[o + 3 * field_stride] = fl ? p[psplit ? pm_sidx(k, jl, i) : c] : 0.0;
  out[o + 4 * field_stride] = vort;
}

__global__ void k_sum_partials(const double* __restrict__ partial, int n, PpeState* __restrict__ st) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int t = threadIdx.x; t < n; t += blockDim.x) s += partial[t];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) st->ke_sum = s;
}

// The four corner ghosts of p are never written by any sweep; a ping-pong (Jacobi) solve carries
// them into the second buffer so both hold the same field outside the stencil's reach.
__global__ void k_copy_corners(const __grid_constant__ KP k, const double* __restrict__ src, double* __restrict__ dst, int split) {
  const int t = threadIdx.x;
  if (t >= 4) return;
  const int jl = (t & 2) ? k.nyl + 1 : 0, i = (t & 1) ? k.nx + 1 : 0;
  if ((jl == 0 && !k.first_rank) || (jl != 0 && !k.last_rank)) return;
  const size_t q = split ? pm_sidx(k, jl, i) : pm_idx(k, jl, i);  // split: buffers of the tiled solve (split-row layout)
  dst[q] = src[q];
}

// Small words device -> anywhere the device can address, pinned host memory included (zero-copy).  The solver's
// state read-backs go this way instead of cudaMemcpyAsync: a copy-engine transfer of a few bytes would queue
// behind the bulk host copies of pm_host_step_* on the same engine and stall the pressure solve for milliseconds.
__global__ void k_publish_words(const unsigned long long* __restrict__ src, unsigned long long* __restrict__ dst, int nwords) {
  for (int q = threadIdx.x; q < nwords; q += blockDim.x) dst[q] = src[q];
  __threadfence_system();
}
// Plane-sized device copy on the SMs (128-bit), for the same reason.
__global__ void k_copy_plane(const double2* __restrict__ src, double2* __restrict__ dst, size_t n2) {
  for (size_t q = size_t(blockIdx.x) * blockDim.x + threadIdx.x; q < n2; q += size_t(gridDim.x) * blockDim.x) dst[q] = src[q];
}

// ---------------------------------------------------------------------------
// synthetic state and layout conversion
// ---------------------------------------------------------------------------
// value(field, j, i) keyed by the reference flat index j*cols + i with GLOBAL j.
__global__ void k_fill_random(const __grid_constant__ KP k, double* __restrict__ a, int field, int rows_global,
                              int cols, uint64_t seed, double amplitude) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int jl = blockIdx.y * blockDim.y + threadIdx.y;  // local rows 0..nyl+1
  if (i >= cols || jl > k.nyl + 1) return;
  const int j = k.j0 + jl;
  if (j >= rows_global) return;
  a[pm_idx(k, jl, i)] = __dmul_rn(amplitude, pm_synth(seed, field, uint64_t(j) * uint64_t(cols) + uint64_t(i)));
}
