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
#include "pm_kernels_tiled.cuh"

#ifndef PM_STREAM_P
#define PM_STREAM_P 3  // rows in flight beyond the row whose north neighbours the first half-sweep reads; the p ring holds P + 1 rows
#endif
#ifndef PM_STREAM_MINB
#define PM_STREAM_MINB 10
#endif
static_assert(PM_STREAM_P == 3 || PM_STREAM_P == 7, "the p ring (P + 1 rows) must divide the 8-tick unroll; the f ring holds 16 rows");

struct StreamGeom {
  int bx0, nbx;  // first strip (tile column of the tiled plan) and number of strips
  int ya, ye;    // first output row (jl, odd: 1 + tile row * TY) and one past the last
  int rows;      // output rows per chunk (even)
  int nchunks;
};

#define PM_STREAM_PRING_BYTES ((PM_STREAM_P + 1) * 1024)
#define PM_STREAM_SMEM_BYTES (PM_STREAM_PRING_BYTES + 16 * 1024)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct StreamCtx {
  double r[8][4];   // rows in flight: slot = tick & 7, cells {E0, O0, E1, O1}
  double acc[8];    // max |r| seen by half-sweep h over the output cells of the chunk
  const double2* pr;  // this lane's 16 bytes in the p ring: row slot s at pr[s * 64], odd columns at + 32
  const double2* fr;  // likewise in the f ring (16 rows)
  int flo, fhi;       // f ring: index offset for the static slots below 8 / from 8 up (the ring is twice the unroll)
  uint32_t pdst, fdst;  // shared-space addresses of pr / fr for cp.async
  const double* pin;    // + element offset = the lane's even-column pair of a row; odd pair at + half
  const double* fin;
  double* pout;
  size_t gsrc;   // element offset of the next row to fetch
  size_t gdst;   // element offset of the row that leaves in this tick (row tau - 7)
  int half;      // pitch / 2
  int pitch;
  int tau;       // tick
  int rows;      // R: output rows of this chunk
  int nfetch;    // rows still to fetch after the next one (the fetch pointer stops at the chunk's last row)
  unsigned act;  // bit h: the row half-sweep h works on lies in the output rows and this lane holds output columns
  bool lane_ok;
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

template <int FORM, int PAR0, int U>
__device__ __forceinline__ void stream_tick(const KP& k, StreamCtx& c) {
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
  c.act = ((c.act << 1) & 0xffu) | unsigned(c.lane_ok && unsigned(c.tau - 8) < unsigned(c.rows));
  // the one neighbour per row that lives in another lane: none of these cells changes during this tick
  double xn[8];
#pragma unroll
  for (int h = 0; h < 8; ++h) {
    constexpr int dummy = 0; (void)dummy;
    const int sl = (U - h) & 7;
    xn[h] = tgtE ? __shfl_up_sync(0xffffffffu, c.r[sl][3], 1) : __shfl_down_sync(0xffffffffu, c.r[sl][0], 1);
  }
#pragma unroll
  for (int h = 0; h < 8; ++h) {
    const int sl = (U - h) & 7, sn = (U - h + 1) & 7, ss = (U - h - 1) & 7;
    const bool jl_odd = ((U - h) & 1) == 0;               // the chunk's first row is odd
    const bool pair_a = tgtE ? jl_odd : !jl_odd;          // (i + jl) even
    const int fs = (U - h) & 15;                          // static f slot; the ring's other half every second pass of the loop
    const double2 fv = c.fr[(fs < 8 ? c.flo : c.fhi) + fs * 64 + (tgtE ? 0 : 32)];
    const double p0 = c.r[sl][G0], p1 = c.r[sl][G1];
    double w0, e0, w1, e1;
    if (tgtE) { w0 = xn[h]; e0 = c.r[sl][1]; w1 = c.r[sl][1]; e1 = c.r[sl][3]; }
    else { w0 = c.r[sl][0]; e0 = c.r[sl][2]; w1 = c.r[sl][2]; e1 = xn[h]; }
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
    const bool on = (c.act >> h) & 1u;
    acc_max(c.acc[h], r0, on);
    acc_max(c.acc[h], r1, on);
  }
  // row tau - 7 has passed all half-sweeps
  if ((c.act >> 7) & 1u) {
    const int so = (U + 1) & 7;
    *reinterpret_cast<double2*>(c.pout + c.gdst) = make_double2(c.r[so][0], c.r[so][2]);
    *reinterpret_cast<double2*>(c.pout + c.gdst + c.half) = make_double2(c.r[so][1], c.r[so][3]);
  }
  c.gdst += size_t(c.pitch);
  // fetch row tau + P + 1 into the slots row tau (p) and row tau - 8 - (7 - P) (f) have left
  {
    const int fs = (U + PR) & 15;
    const uint32_t pd = c.pdst + uint32_t((U % PR) * 1024);
    const uint32_t fd = c.fdst + uint32_t(((fs < 8 ? c.flo : c.fhi) + fs * 64) * 16);
    cp_async16(pd, c.pin + c.gsrc);
    cp_async16(pd + 512u, c.pin + c.gsrc + c.half);
    cp_async16(fd, c.fin + c.gsrc);
    cp_async16(fd + 512u, c.fin + c.gsrc + c.half);
    cp_async_commit();
    if (c.nfetch > 0) c.gsrc += size_t(c.pitch);
    --c.nfetch;
  }
  ++c.tau;
}

template <int FORM, int PAR0>
__global__ void __launch_bounds__(32, PM_STREAM_MINB)
    k_ppe_stream(const __grid_constant__ KP k, const double* __restrict__ pin, double* __restrict__ pout, const double* __restrict__ fsplit,
                 PpeState* __restrict__ st, unsigned long long* __restrict__ res_bits, unsigned long long* __restrict__ fold_part,
                 const __grid_constant__ StreamGeom g, int m0, int force) {
  constexpr int T = 4, H = 2 * T, P = PM_STREAM_P, PR = P + 1;
  using C = TileCfg<PM_PPE_SOR_RB, T>;
  static_assert(C::H == H && C::SW == 128, "strips are the tile columns of the tiled plan");
  extern __shared__ __align__(16) unsigned char stream_smem[];
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
  c.fr = reinterpret_cast<const double2*>(stream_smem + PM_STREAM_PRING_BYTES) + lane;
  c.flo = c.fhi = 0;
  c.pdst = smem_u32(c.pr);
  c.fdst = smem_u32(c.fr);
  c.pin = pin;
  c.fin = fsplit;
  c.pout = pout;
  c.pitch = k.pitch;
  c.half = k.pitch >> 1;
  c.gsrc = size_t(k.padr + jstart) * size_t(k.pitch) + size_t(pe);
  c.gdst = size_t(k.padr + jstart - 7) * size_t(k.pitch) + size_t(pe);
  c.tau = 0;
  c.rows = yb - ya;
  c.act = 0u;
  c.lane_ok = lane >= 2 && lane <= 29;  // columns 8 .. 119 of the strip
  const int nrows = c.rows + 2 * H;      // rows that enter: the output rows and H below / above
  c.nfetch = nrows - 1;
  // rows 0 .. P
#pragma unroll
  for (int q = 0; q < PR; ++q) {
    const uint32_t pd = c.pdst + uint32_t(q * 1024), fd = c.fdst + uint32_t(q * 1024);
    cp_async16(pd, c.pin + c.gsrc);
    cp_async16(pd + 512u, c.pin + c.gsrc + c.half);
    cp_async16(fd, c.fin + c.gsrc);
    cp_async16(fd + 512u, c.fin + c.gsrc + c.half);
    cp_async_commit();
    if (c.nfetch > 0) c.gsrc += size_t(c.pitch);
    --c.nfetch;
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
  // tick tau: rows tau (+1 as north) must be there; the last output row leaves at tick rows + 14
  const int nticks = c.rows + 2 * H - 1;
#pragma unroll 1
  for (int it = 0; it * 8 < nticks; ++it) {
    stream_tick<FORM, PAR0, 0>(k, c);
    stream_tick<FORM, PAR0, 1>(k, c);
    stream_tick<FORM, PAR0, 2>(k, c);
    stream_tick<FORM, PAR0, 3>(k, c);
    stream_tick<FORM, PAR0, 4>(k, c);
    stream_tick<FORM, PAR0, 5>(k, c);
    stream_tick<FORM, PAR0, 6>(k, c);
    stream_tick<FORM, PAR0, 7>(k, c);
    // the f ring holds 16 rows: every second pass of the loop works on its other half
    c.flo = 512 - c.flo;
    c.fhi = -c.flo;
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
    v = warp_max_nonneg(v);
    const int m = m0 + t;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    if (lane == 0 && bits != 0ull && m >= 1 && m <= k.max_iters)
      atomicMax(fold_part + size_t(item & (PM_FOLD_SLOTS - 1)) * 16 + (m & 7), bits);
  }
}

template <int FORM>
static const void* stream_kernel_ptr(int par0) {
  return par0 ? reinterpret_cast<const void*>(&k_ppe_stream<FORM, 1>) : reinterpret_cast<const void*>(&k_ppe_stream<FORM, 0>);
}
