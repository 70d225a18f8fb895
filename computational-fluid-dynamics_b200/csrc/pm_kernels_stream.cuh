// pm_kernels_stream.cuh — the streaming form of the temporally blocked red-black pass (production arithmetic,
// interior of large unmasked grids): the same T = 4 sweeps and per-iterate residual norms as k_ppe_tiled, bit for bit,
// with 15-22 % redundant cell updates instead of 42 %, no block barrier and no separate load / sweep / store phases.
//
// Reference loop: cavity-01.cpp:635-677, channel-01.cpp:652-681 (sweep + residual of every iterate).
//
// One WARP owns a strip of 128 columns (112 of them output, a halo of H = 2T = 8 on either side, the k_ppe_tiled tile
// columns) and walks it upwards through a chunk of R output rows.  The 2T colour half-sweeps of the pass form a
// pipeline over the rows: at tick tau the row tau enters from HBM, half-sweep h is applied to row tau - h (h = 0..7,
// ascending, each reading the row above as half-sweep h - 1 left it in this tick and the row below as half-sweep
// h + 1 left it in the previous one) and row tau - 7 leaves for HBM.  So a row is read once and written once per
// four sweeps, like a tile, but only the columns of the strip edge and the 16 rows at the two ends of a chunk are swept in vain.
//   * every lane keeps 4 adjacent columns {E0, O0, E1, O1} of the 8 rows in flight in registers (rows rotate through
//     8 register slots; the tick loop is unrolled 8 times so every slot index is a compile-time constant);
//   * the only horizontal neighbour a lane does not hold comes by one 64-bit warp shuffle per row and half-sweep;
//   * p and f (both in the split-row layout: a lane's two even columns are 16 contiguous bytes, and so are the two odd ones)
//     arrive through cp.async rings private to the warp, PM_STREAM_P rows ahead: p is taken into registers once, f is
//     read where a half-sweep needs it (one conflict-free LDS.128 for the two cells of that colour);
//   * per half-sweep one running max |r| (residual form of the relaxation: the norm costs a compare), folded into the
//     per-iterate slots exactly as k_ppe_tiled does (colour-0 part before the update, colour-1 part as |1 - omega| * max |r|);
//   * arithmetic: the association of the four-neighbour sum follows rb_half_lean's shared diagonal sums
//     ((i + jl) even: (E + N) + (W + S), odd: (W + N) + (E + S)), so both kernels produce identical bits and can
//     share one pass (k_ppe_tiled takes the frame of tiles that touch a wall or a slab edge).
#pragma once
#include <cstdio>
#include "pm_kernels_tiled.cuh"

#ifndef PM_STREAM_P
#define PM_STREAM_P 3  // rows in flight beyond the row whose north neighbours the first half-sweep reads; the p ring holds P + 1 rows
#endif
#ifndef PM_STREAM_MINB
#define PM_STREAM_MINB 12
#endif
static_assert(PM_STREAM_P == 3, "the rings of rows on their way (P + 1 rows) must divide the 8-tick unroll");

struct StreamGeom {
  int bx0, nbx;  // first strip (tile column of the tiled plan) and number of strips
  int ya, ye;    // first output row (jl, odd: 1 + tile row * TY) and one past the last
  int rows;      // output rows per chunk (even)
  int nchunks;
};

#define PM_STREAM_PRING_BYTES ((PM_STREAM_P + 1) * 1024)
// p rows on their way | f rows on their way | f of the 8 rows in flight (slot = tick & 7, like the register slots): every
// shared-memory address of the loop is the lane's base plus a compile-time constant
#define PM_STREAM_RING_BYTES (2 * PM_STREAM_PRING_BYTES + 8 * 1024)
// + the loop counters: the loop body fills the register file of a 12-warp SM, and what ptxas spills to local memory
// around it comes back from DRAM here (the L1 left beside 12 x 16 KB of shared memory does not hold 12 warps' frames)
#define PM_STREAM_SMEM_BYTES (PM_STREAM_RING_BYTES + 16)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Kernel parameters the ticks read straight from the constant bank (nothing of this is carried in registers).
struct StreamIO {
  const double* pin;  // + element offset = the lane's even-column pair of a row; odd pair at + pitch / 2
  const double* fin;
  double* pout;
  int pitch;
  int rows;           // R: output rows of this chunk
};

struct StreamCtx {
  double r[8][4];   // rows in flight: slot = tick & 7, cells {E0, O0, E1, O1}
  double acc[8];    // max |r| seen by half-sweep h over the output cells of the chunk
  const double2* pr;  // this lane's 16 bytes in the p ring: row slot s at pr[s * 64], odd columns at + 32
  double2* fr;        // likewise in the f ring of rows on their way (P + 1 slots), followed by the f rows in flight (8 slots)
  uint32_t pdst, fdst;  // shared-space addresses of pr / fr for cp.async
  uint32_t gsrc;  // element offset of the next row to fetch, row tau + P + 1 (planes stay below 2^32 elements: stream_supported);
                  // the row that leaves in tick tau, row tau - 7, is P + 8 rows below it
  unsigned act;  // bit h: the row half-sweep h works on lies in the output rows
  // columns 8 .. 119 of the strip: lanes 2 .. 29, read off the lane's ring address (the rings start 1 KiB-aligned) rather than
  // kept in a predicate across the loop
  __device__ __forceinline__ bool lane_ok() const { return ((pdst >> 4) & 31u) - 2u < 28u; }
};

// r = residual with the iterate's own operands; the relaxation is p += cw * r  (rb_half_lean)
template <int FORM, bool PAIR_A>
__device__ __forceinline__ double stream_res(const KP& k, double pc, double pw, double pe, double pn, double ps, double f) {
  if (FORM == 0) {
    const double s = PAIR_A ? (pe + pn) + (pw + ps) : (pw + pn) + (pe + ps);
    return res_sum4<0>(k, pc, s, f);
  }
  return res_sum22(k, pc, pw + pe, pn + ps, f);
}

// m = std::max(m, std::abs(x)) (a NaN never replaces m), branch-free; |x| by clearing the sign bit
// in the integer pipe (abs.f64 became an FP64-pipe DADD -RZ, |x| per call): LOP3 + DSETP.GT + 2 FSEL.
// (fmax() drags NaN quieting and register copies along; an `if` around two of these turns into branches.)
__device__ __forceinline__ void stream_max(double& m, double x) {
  asm("{\n .reg .pred q;\n .reg .b32 lo, hi;\n .reg .f64 a;\n mov.b64 {lo, hi}, %1;\n and.b32 hi, hi, 0x7fffffff;\n mov.b64 a, {lo, hi};\n"
      " setp.gt.f64 q, a, %0;\n selp.f64 %0, a, %0, q;\n}"
      : "+d"(m)
      : "d"(x));
}
__device__ __forceinline__ void stream_max_if(double& m, double x, unsigned on) {
  asm("{\n .reg .pred q, o;\n .reg .b32 lo, hi;\n .reg .f64 a;\n setp.ne.u32 o, %2, 0;\n mov.b64 {lo, hi}, %1;\n and.b32 hi, hi, 0x7fffffff;\n mov.b64 a, {lo, hi};\n"
      " setp.gt.and.f64 q, a, %0, o;\n selp.f64 %0, a, %0, q;\n}"
      : "+d"(m)
      : "d"(x), "r"(on));
}

// DYN: the ticks at the two ends of a chunk, where some half-sweeps work on rows outside the output rows (bit h of c.act);
// in between every half-sweep's row is an output row, and the lanes that hold halo columns are dropped at the very end.
// tau0: the tick of U == 0 in this pass of the loop (uniform; only the DYN ticks look at it).
template <int FORM, int PAR0, int U, bool DYN>
__device__ __forceinline__ void stream_tick(const KP& k, const StreamIO& io, StreamCtx& c, int tau0) {
  const int half = io.pitch >> 1;
  constexpr int P = PM_STREAM_P, PR = P + 1;
  constexpr bool tgtE = ((U + PAR0) & 1) == 0;  // this tick's half-sweeps all update the even storage columns (i odd) of their rows
  constexpr int G0 = tgtE ? 0 : 1, G1 = tgtE ? 2 : 3;
  cp_async_wait<P - 1>();  // rows tau and tau + 1 have landed
  // the cells below half-sweep 7's targets: row tau - 8, about to be replaced in its slot by row tau
  const double cy0 = c.r[U][G0], cy1 = c.r[U][G1];
  {
    const double2 e = c.pr[(U % PR) * 64], o = c.pr[(U % PR) * 64 + 32];
    c.r[U][0] = e.x; c.r[U][2] = e.y;
    c.r[U][1] = o.x; c.r[U][3] = o.y;
  }
  const double2 nx = c.pr[((U + 1) % PR) * 64 + (tgtE ? 0 : 32)];  // row tau + 1 as it came from HBM: north of half-sweep 0
  // f of row tau moves to its slot among the rows in flight (the lane's own 32 bytes: no synchronisation)
  const double2 fe = c.fr[(U % PR) * 64], fo = c.fr[(U % PR) * 64 + 32];
  c.fr[(PR + U) * 64] = fe;
  c.fr[(PR + U) * 64 + 32] = fo;
  // Fetch row tau + P + 1 into the slots row tau has just been read from -- right away, not at the end of the tick: the
  // load/store unit takes the reads above and these copies in order, and the row gains a tick to arrive in.
  {
    const uint32_t pd = c.pdst + uint32_t((U % PR) * 1024);
    const uint32_t fd = c.fdst + uint32_t((U % PR) * 1024);
    if (!DYN || tau0 + U + PR < io.rows + 16) {  // nothing beyond the chunk's last row is needed (or may exist)
      cp_async16(pd, io.pin + c.gsrc);
      cp_async16(pd + 512u, io.pin + c.gsrc + half);
      cp_async16(fd, io.fin + c.gsrc);
      cp_async16(fd + 512u, io.fin + c.gsrc + half);
    }
    cp_async_commit();
    c.gsrc += uint32_t(io.pitch);
  }
  if (DYN) c.act = ((c.act << 1) & 0xffu) | unsigned(unsigned(tau0 + U - 8) < unsigned(io.rows));
  // the one neighbour per row that lives in another lane: none of these cells changes during this tick
  auto outer = [&](int h) {
    const int sl = (U - h) & 7;
    return tgtE ? __shfl_up_sync(0xffffffffu, c.r[sl][3], 1) : __shfl_down_sync(0xffffffffu, c.r[sl][0], 1);
  };
  double xq = outer(0);
#pragma unroll
  for (int h = 0; h < 8; ++h) {
    const int sl = (U - h) & 7, sn = (U - h + 1) & 7, ss = (U - h - 1) & 7;
    const bool jl_odd = ((U - h) & 1) == 0;               // the chunk's first row is odd
    const bool pair_a = tgtE ? jl_odd : !jl_odd;          // (i + jl) even
    const double2 fv = h == 0 ? (tgtE ? fe : fo) : c.fr[(PR + sl) * 64 + (tgtE ? 0 : 32)];
    const double p0 = c.r[sl][G0], p1 = c.r[sl][G1];
    const double xh = xq;
    if (h < 7) xq = outer(h + 1);  // one half-sweep ahead
    double w0, e0, w1, e1;
    if (tgtE) { w0 = xh; e0 = c.r[sl][1]; w1 = c.r[sl][1]; e1 = c.r[sl][3]; }
    else { w0 = c.r[sl][0]; e0 = c.r[sl][2]; w1 = c.r[sl][2]; e1 = xh; }
    const double n0 = h == 0 ? nx.x : c.r[sn][G0], n1 = h == 0 ? nx.y : c.r[sn][G1];
    const double s0 = h == 7 ? cy0 : c.r[ss][G0], s1 = h == 7 ? cy1 : c.r[ss][G1];
    double r0, r1;
    if (pair_a) {
      r0 = stream_res<FORM, true>(k, p0, w0, e0, n0, s0, fv.x);
      r1 = stream_res<FORM, true>(k, p1, w1, e1, n1, s1, fv.y);
    } else {
      r0 = stream_res<FORM, false>(k, p0, w0, e0, n0, s0, fv.x);
      r1 = stream_res<FORM, false>(k, p1, w1, e1, n1, s1, fv.y);
    }
    c.r[sl][G0] = fma(k.cw, r0, p0);
    c.r[sl][G1] = fma(k.cw, r1, p1);
    if (DYN) {
      stream_max_if(c.acc[h], r0, c.act & (1u << h));
      stream_max_if(c.acc[h], r1, c.act & (1u << h));
    } else {
      stream_max(c.acc[h], r0);
      stream_max(c.acc[h], r1);
    }
  }
  // row tau - 7 has passed all half-sweeps
  if (c.lane_ok() && (!DYN || ((c.act >> 7) & 1u))) {
    const int so = (U + 1) & 7;
    const uint32_t gdst = c.gsrc - uint32_t(PR + 8) * uint32_t(io.pitch);  // gsrc already points at row tau + P + 2
    *reinterpret_cast<double2*>(io.pout + gdst) = make_double2(c.r[so][0], c.r[so][2]);
    *reinterpret_cast<double2*>(io.pout + gdst + half) = make_double2(c.r[so][1], c.r[so][3]);
  }
}

template <int FORM, int PAR0>
__global__ void __launch_bounds__(32, PM_STREAM_MINB)
    k_ppe_stream(const __grid_constant__ KP k, const double* __restrict__ pin, double* __restrict__ pout, const double* __restrict__ fsplit,
                 PpeState* __restrict__ st, unsigned long long* __restrict__ res_bits, unsigned long long* __restrict__ fold_part,
                 const __grid_constant__ StreamGeom g, int m0, int force) {
  constexpr int T = 4, H = 2 * T, P = PM_STREAM_P, PR = P + 1;
  using C = TileCfg<PM_PPE_SOR_RB, T>;
  static_assert(C::H == H && C::SW == 128, "strips are the tile columns of the tiled plan");
  extern __shared__ __align__(1024) unsigned char stream_smem[];
  const int lane = threadIdx.x;
  StopWords<T> stopw;
  if (!force) stopw = stop_words_load<T>(st, res_bits, m0);

  const int item = blockIdx.x;
  const int chunk = item / g.nbx, strip = item - chunk * g.nbx;
  const int ya = g.ya + chunk * g.rows, yb = min(ya + g.rows, g.ye);
  const int ib = 1 + (g.bx0 + strip) * C::TX - H;           // first column (i) of the strip, odd
  const int pe = ((PM_OFFC + ib + k.psh) >> 1) + 2 * lane;   // the lane's pair of even storage columns inside the even half of a row
  const int jstart = ya - H;                                 // first row (jl) that enters

  StreamCtx c;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    c.r[s][0] = c.r[s][1] = c.r[s][2] = c.r[s][3] = 0.0;
    c.acc[s] = 0.0;
  }
  c.pr = reinterpret_cast<const double2*>(stream_smem) + lane;
  c.fr = reinterpret_cast<double2*>(stream_smem + PM_STREAM_PRING_BYTES) + lane;
  c.pdst = smem_u32(c.pr);
  c.fdst = smem_u32(c.fr);
  const StreamIO io{pin, fsplit, pout, k.pitch, yb - ya};
  const int half = k.pitch >> 1;
  c.gsrc = uint32_t(k.padr + jstart) * uint32_t(k.pitch) + uint32_t(pe);
  c.act = 0u;
  static_assert(H == 8, "rows + 16 rows enter: the output rows and H below / above");
  // rows 0 .. P
#pragma unroll
  for (int q = 0; q < PR; ++q) {
    const uint32_t pd = c.pdst + uint32_t(q * 1024), fd = c.fdst + uint32_t(q * 1024);
    cp_async16(pd, io.pin + c.gsrc);
    cp_async16(pd + 512u, io.pin + c.gsrc + half);
    cp_async16(fd, io.fin + c.gsrc);
    cp_async16(fd + 512u, io.fin + c.gsrc + half);
    cp_async_commit();
    c.gsrc += uint32_t(io.pitch);
  }
  if (!force) {  // the reference's loop test (uniform over the grid), as in k_ppe_tiled
    int first;
    if (stop_words_eval<T>(stopw, m0, &first)) {
      if (first >= 0 && item == 0 && lane == 0) {
        st->iters = first;
        st->done = 1;
      }
      cp_async_wait<0>();
      return;
    }
  }
  // Tick tau needs rows tau and tau + 1; the last output row leaves at tick rows + 14.  Eight ticks per pass of the loop;
  // passes 2 .. rows / 8 (ticks 16 .. rows + 7) have every half-sweep on an output row and every fetch inside the chunk.
  volatile int* loopc = reinterpret_cast<volatile int*>(stream_smem + PM_STREAM_RING_BYTES);
  {
    const int nticks = io.rows + 2 * H - 1;
    const int nit0 = (nticks + 7) >> 3;
    loopc[0] = 0;                             // pass of the loop (every lane writes the same words)
    loopc[1] = nit0;                          // passes
    loopc[2] = min(io.rows >> 3, nit0 - 1);   // last steady one
  }
#pragma unroll 1
  for (;;) {
    __syncwarp();  // every lane writes the same words; all lanes have read a word before any lane overwrites it
    const int it = loopc[0], nit = loopc[1], it_steady_last = loopc[2];
    __syncwarp();
    if (it >= nit) break;
    loopc[0] = it + 1;
    if (it >= 2 && it <= it_steady_last) {
      stream_tick<FORM, PAR0, 0, false>(k, io, c, 0);
      stream_tick<FORM, PAR0, 1, false>(k, io, c, 0);
      stream_tick<FORM, PAR0, 2, false>(k, io, c, 0);
      stream_tick<FORM, PAR0, 3, false>(k, io, c, 0);
      stream_tick<FORM, PAR0, 4, false>(k, io, c, 0);
      stream_tick<FORM, PAR0, 5, false>(k, io, c, 0);
      stream_tick<FORM, PAR0, 6, false>(k, io, c, 0);
      stream_tick<FORM, PAR0, 7, false>(k, io, c, 0);
    } else {
      if (it_steady_last >= 2 && it == it_steady_last + 1) {  // the steady passes do not keep the bits: all were set
        c.act = 0xffu;
      }
      stream_tick<FORM, PAR0, 0, true>(k, io, c, 8 * it);
      stream_tick<FORM, PAR0, 1, true>(k, io, c, 8 * it);
      stream_tick<FORM, PAR0, 2, true>(k, io, c, 8 * it);
      stream_tick<FORM, PAR0, 3, true>(k, io, c, 8 * it);
      stream_tick<FORM, PAR0, 4, true>(k, io, c, 8 * it);
      stream_tick<FORM, PAR0, 5, true>(k, io, c, 8 * it);
      stream_tick<FORM, PAR0, 6, true>(k, io, c, 8 * it);
      stream_tick<FORM, PAR0, 7, true>(k, io, c, 8 * it);
    }
  }
  cp_async_wait<0>();
  // per-iterate maxima: entry t = colour-0 part of iterate m0 + t (half-sweep 2t, operands before the update) and the
  // colour-1 part taken when half-sweep 2t - 1 created it ((1 - omega) * r, see rb_half_lean)
  const double sc = fabs(k.om1);
#pragma unroll
  for (int t = 0; t <= T; ++t) {
    double v = t < T ? c.acc[2 * t] : 0.0;
    if (t > 0) {
      const double m = sc * c.acc[2 * t - 1];
      if (m > v) v = m;
    }
    v = warp_max_nonneg(c.lane_ok() ? v : 0.0);  // the lanes of the strip's halo columns hold no output cell
    const int m = m0 + t;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    if (lane == 0 && bits != 0ull && m >= 1 && m <= k.max_iters)
      atomicMax(fold_part + size_t(item & (PM_FOLD_SLOTS - 1)) * 16 + (m & 7), bits);
  }
}

#ifdef PM_TILED_DEVICE_ONLY
template <int FORM>
static const void* stream_kernel_ptr(int) { return reinterpret_cast<const void*>(&k_ppe_stream<FORM, 0>); }
#else
// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
template <int FORM>
static const void* stream_kernel_ptr(int par0) {
  return par0 ? reinterpret_cast<const void*>(&k_ppe_stream<FORM, 1>) : reinterpret_cast<const void*>(&k_ppe_stream<FORM, 0>);
}

// Which tiles of the tiled plan's rows [row_lo, row_hi) the streaming kernel takes (the largest rectangle of tiles that lie
// strictly inside the domain with data for every neighbour: k_ppe_tiled's `interior`, which is separable in x and y),
// and the list of the others for tiled_launch_list.
struct StreamPlan {
  bool on = false;
  StreamGeom g{};
  const void* kernel = nullptr;
  int* frame = nullptr;
  int nframe = 0;
  int items = 0;
};

static inline void stream_destroy(StreamPlan* sp) {
  if (sp->frame) cudaFree(sp->frame);
  sp->frame = nullptr;
  sp->on = false;
}

static inline bool stream_supported(const pm_config& c, const KP& k, const TiledPlan& pl) {
  const int nyl_rows = k.nyl + 2 + 2 * k.padr;
  return !c.exact_arith && c.ppe_method == PM_PPE_SOR_RB && !k.has_mask && pl.sweeps == 4 && pl.run == 4 && pl.cs == 1 &&
         double(nyl_rows) * double(k.pitch) < 4.0e9 &&  // 32-bit element offsets inside a plane
         std::getenv("PM_NO_STREAM") == nullptr;
}

// The plan's arithmetic, free of CUDA calls (pm_stream_plan exposes it to the CPU tests): the rectangle of tiles
// [bx0, bx0 + nbx) x [by0, by0 + nby) inside the tile rows [row_lo, row_hi) whose tiles all satisfy k_ppe_tiled's `interior`
// (separable in x and y), and the chunk height.
struct StreamShape { int bx0, nbx, by0, nby, rows, nchunks, items, nframe; };
static inline bool stream_shape(int nx, int ny, int nyl, int j0, int tiles_x, int row_lo, int row_hi, int tx, int ty, int sh, int H, int slots,
                                int force_rows, StreamShape* o) {
  const int SW = 128;
  auto x_ok = [&](int bx) { const int ib = 1 + bx * tx - H; return ib + 1 >= 2 && ib + SW - 2 <= nx - 1; };
  auto y_ok = [&](int by) {
    const int jb = 1 + by * ty - H;
    return j0 + jb + 1 >= 2 && j0 + jb + sh - 2 <= ny - 1 && jb + sh - 1 <= nyl + H && jb >= 1 - H;
  };
  int bx0 = -1, bx1 = -2, by0 = -1, by1 = -2;
  for (int bx = 0; bx < tiles_x; ++bx)
    if (x_ok(bx)) { if (bx0 < 0) bx0 = bx; bx1 = bx; }
  for (int by = row_lo; by < row_hi; ++by)
    if (y_ok(by)) { if (by0 < 0) by0 = by; by1 = by; }
  // (both predicates hold on one contiguous range)
  for (int bx = bx0; bx0 >= 0 && bx <= bx1; ++bx) if (!x_ok(bx)) { bx0 = -1; break; }
  for (int by = by0; by0 >= 0 && by <= by1; ++by) if (!y_ok(by)) { by0 = -1; break; }
  if (bx0 < 0 || by0 < 0) return false;  // nothing to stream: the tiled kernel keeps every tile
  const int nbx = bx1 - bx0 + 1, nby = by1 - by0 + 1;
  if (nbx * nby < 64) return false;
  // Chunk height: whole tile rows, and as few whole waves of warps as possible.  All warps of a wave run at the same
  // pace, so a launch that is a few warps over a wave takes a wave longer (measured at 8192^2: 1728 warps on 1776
  // slots 10.2 ms/step, 1872 warps 12.8); within that, the higher the chunk the less of it is spent on its 16 halo rows.
  const int rows = nby * ty;
  int best = ty;
  double best_cost = 1e300;
  for (int w = 1; w <= 64; ++w) {
    const int nch_max = std::max(1, int((long long)w * slots / nbx));
    int R = ((rows + nch_max - 1) / nch_max + ty - 1) / ty * ty;
    R = std::min(R, rows);
    const int nch = (rows + R - 1) / R;
    if ((long long)nbx * nch > (long long)w * slots) continue;
    const double cost = double(w) * (R + 2 * H + 8);
    if (cost < best_cost) { best_cost = cost; best = R; }
    if (R <= 2 * ty) break;
  }
  if (force_rows >= ty && force_rows % ty == 0) best = force_rows;
  o->bx0 = bx0; o->nbx = nbx; o->by0 = by0; o->nby = nby;
  o->rows = best;
  o->nchunks = (rows + best - 1) / best;
  o->items = nbx * o->nchunks;
  o->nframe = (row_hi - row_lo) * tiles_x - nbx * nby;
  return true;
}

static inline bool stream_create(StreamPlan* sp, const TiledPlan& pl, const pm_config& c, const KP& k, int row_lo, int row_hi, cudaStream_t stream,
                                 std::string* err) {
  sp->on = false;
  const int H = pl.halo;
  int dev = 0, sms = 148, per_sm = PM_STREAM_MINB;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool cav = c.case_id == PM_CASE_CAVITY;
  sp->kernel = cav ? stream_kernel_ptr<0>(k.j0 & 1) : stream_kernel_ptr<1>(k.j0 & 1);
  cudaError_t e = cudaFuncSetAttribute(sp->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PM_STREAM_SMEM_BYTES);
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sp->kernel, 32, PM_STREAM_SMEM_BYTES);
  if (e != cudaSuccess || per_sm < 1) { *err = std::string("streaming kernel attributes: ") + cudaGetErrorString(e); return false; }
  int force_rows = 0;
  if (const char* ev = std::getenv("PM_STREAM_ROWS")) force_rows = std::atoi(ev);
  int slots = sms * per_sm;
  if (const char* ev = std::getenv("PM_STREAM_SLOTS")) { const int v = std::atoi(ev); if (v > 0) slots = v; }
  StreamShape sh{};
  if (!stream_shape(k.nx, k.ny, k.nyl, k.j0, pl.tiles_x, row_lo, row_hi, pl.tx, pl.ty, pl.sh, H, slots, force_rows, &sh)) return true;
  const int bx0 = sh.bx0, bx1 = sh.bx0 + sh.nbx - 1, by0 = sh.by0, by1 = sh.by0 + sh.nby - 1, nbx = sh.nbx, best = sh.rows;
  const int rows = sh.nby * pl.ty;
  sp->g.bx0 = bx0; sp->g.nbx = nbx;
  sp->g.ya = 1 + by0 * pl.ty; sp->g.ye = 1 + (by1 + 1) * pl.ty;
  sp->g.rows = best;
  sp->g.nchunks = (rows + best - 1) / best;
  sp->items = nbx * sp->g.nchunks;
  std::vector<int> frame;
  for (int by = row_lo; by < row_hi; ++by)
    for (int bx = 0; bx < pl.tiles_x; ++bx)
      if (!(bx >= bx0 && bx <= bx1 && by >= by0 && by <= by1)) frame.push_back(by * pl.tiles_x + bx);
  sp->nframe = int(frame.size());
  if (sp->frame) { cudaFree(sp->frame); sp->frame = nullptr; }
  if (sp->nframe > 0) {
    e = cudaMalloc(&sp->frame, frame.size() * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpyAsync(sp->frame, frame.data(), frame.size() * sizeof(int), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { *err = std::string("stream plan: ") + cudaGetErrorString(e); return false; }
  }
  sp->on = true;
  if (std::getenv("PM_DEBUG_STREAM"))
    fprintf(stderr, "[pm] streaming pass: strips %d..%d, rows %d..%d in chunks of %d -> %d warps; %d frame tiles\n", bx0, bx1, sp->g.ya, sp->g.ye - 1,
            sp->g.rows, sp->items, sp->nframe);
  return true;
}

// One pass (T = 4 sweeps) over the plan's rectangle: reads iterate m0 from buffer `in`, writes iterate m0 + 4 to the other.
static inline cudaError_t stream_launch(const StreamPlan* sp, const TiledPlan* pl, const KP& k, int in, const double* fsplit, PpeState* st,
                                        unsigned long long* res, int m0, int force, cudaStream_t stream) {
  const double* pin = pl->p[in];
  double* pout = pl->p[in ^ 1];
  void* args[] = {(void*)&k, (void*)&pin, (void*)&pout, (void*)&fsplit, (void*)&st, (void*)&res, (void*)&pl->fold_part, (void*)&sp->g, (void*)&m0, (void*)&force};
  static_assert(12 * (PM_STREAM_SMEM_BYTES + 1024) <= 227 * 1024, "12 warps per SM");
  return cudaLaunchKernel(sp->kernel, dim3(sp->items), dim3(32), args, PM_STREAM_SMEM_BYTES, stream);
}
#endif  // PM_TILED_DEVICE_ONLY
