// pm_tile_cfg.cuh — compile-time geometry of the tiles of the tiled pressure solve (pm_kernels_tiled.cuh).
// Separate from the kernels so that host-only checks (tests/helpers/split_layout_check.cu) can include it.
#pragma once
#include "pm_common.cuh"

// Tile shape.  Red-black: 4 segments x 12 rows = 48 x 128 cells on 256 threads -- eight warps per CTA spread evenly
// over the four SM sub-partitions (ten did not: 3/3/2/2), and the taller tile raises the share of output cells from
// 63 % to 68 % at T = 3; it costs the full 128 registers per thread (96 hold the thread's 24 cells of p and f).
// Jacobi sweeps from one shared-memory tile into a second one (no staging of new values in registers) on 4 segments x 8 rows
// = 32 x 128 cells, 256 threads: 5 x 8 on 320 threads sat at the 96-register cap and spilled.
#ifndef PM_TILE_NSEG
#define PM_TILE_NSEG 4   // row segments per tile; 64 threads (column pairs) each
#endif
#ifndef PM_TILE_RPT
#define PM_TILE_RPT 12   // rows per thread (even: keeps the colour of a thread's first row uniform over the launch)
#endif
#ifndef PM_TILE_NSEG_JACOBI
#define PM_TILE_NSEG_JACOBI 4
#endif
#ifndef PM_TILE_RPT_JACOBI
#define PM_TILE_RPT_JACOBI 8
#endif
#ifndef PM_TILE_MINBLOCKS
#define PM_TILE_MINBLOCKS 2
#endif
// Red-black tiles can be stacked into thread-block clusters of PM_TILE_CS CTAs in y: the CTAs of a stack exchange their
// edge rows through distributed shared memory after every colour half-sweep, so only the two ends of the stack carry
// a halo in y.  At T = 4 (halo 8) the share of output cells goes from 58 % (one 48 x 128 tile) to 80 % (four).
// Measured at 8192^2: the 27 % fewer cell updates are eaten by the lock step of the stack (a CTA's half-sweep
// grows from 1060 to 1500-1700 cycles): 13.8 ms/step either way.  Default 1 (independent tiles); lib/libpm_cs4.so is the
// cluster build the GPU tests also run.
#ifndef PM_TILE_CS
#define PM_TILE_CS 1
#endif

template <int METHOD, int T>
struct TileCfg {
  static constexpr int H = (METHOD == PM_PPE_SOR_RB) ? 2 * T : ((T + 1) / 2) * 2;  // even: keeps 16-byte alignment of row pairs
  static constexpr int SW = 128;                  // tile width in doubles == 64 column pairs == one TMA box row (1 KiB)
  static constexpr int NSEG = (METHOD == PM_PPE_SOR_RB) ? PM_TILE_NSEG : PM_TILE_NSEG_JACOBI;
  static constexpr int RPT = (METHOD == PM_PPE_SOR_RB) ? PM_TILE_RPT : PM_TILE_RPT_JACOBI;  // rows per thread
  static constexpr int THREADS = 64 * NSEG;
  static constexpr int NWARPS = THREADS / 32;
  static constexpr int SH = NSEG * RPT;           // tile height
  static constexpr int CS = (METHOD == PM_PPE_SOR_RB) ? PM_TILE_CS : 1;  // CTAs per cluster (stacked in y)
  static constexpr int TX = SW - 2 * H;           // output block of a cluster
  static constexpr int TY = CS * SH - 2 * H;
  static constexpr int XR = CS > 1 ? 1 : 0;       // rows below and above the tile that the TMA load brings along (the neighbour CTAs' edge rows)
  // tile + one row above and below (the neighbour CTA's edge row, or spare); Jacobi sweeps from one such tile into a second
  static constexpr int TILE_DOUBLES = (SH + 2) * SW;
  static constexpr int SMEM_BYTES = ((METHOD == PM_PPE_SOR_RB) ? 1 : 2) * TILE_DOUBLES * 8;
  static_assert(TX > 0 && TY > 0, "halo too deep for the tile");
  static_assert(CS >= 1 && CS <= 8, "portable cluster sizes only");
  static_assert(RPT % 2 == 0 && TY % 2 == 0, "PAR0 (colour of a thread's first row) must not depend on the segment or the tile row");
  static_assert(H <= PM_PADR, "halo deeper than the pad rows of the planes");
  static_assert(H % 2 == 0 && TX % 4 == 0 && (PM_OFFC + 1) % 2 == 0 && PM_OFFC + 1 >= H,
                "the first storage column of every tile must be even (it is the .x cell of lane 0), non-negative, and the same mod 4 for all tiles");
  // shift of the split-row layout that makes (first storage column + PSH) / 2 even: 16-byte aligned TMA box rows
  static constexpr int PSH = (4 - ((PM_OFFC + 1 - H) & 3)) & 3;
  static_assert(PSH == 0 || PSH == 2, "");
};

