// Host-only check of the split-row layout helpers of csrc/pm_common.cuh / pm_tile_cfg.cuh (built with nvcc, run on the CPU):
//   * pm_split_col is a bijection of [0, pitch) for every pitch / shift in use;
//   * for every tile shape the first pair of every tile is even (16-byte aligned TMA box rows) and in range;
//   * lane q's .x / .y cells of a tile land at consecutive doubles of the even / odd half.
#include <cstdio>
#include <vector>
#include <algorithm>
#include "../../computational-fluid-dynamics_b200/csrc/pm_tile_cfg.cuh"

template <int METHOD, int T>
static int check_shape(int nx) {
  using C = TileCfg<METHOD, T>;
  KP k{};
  k.nx = nx;
  k.pitch = std::max(144, ((PM_OFFC + nx + 2 + PM_PADR + 15) / 16) * 16);
  k.padr = PM_PADR;
  k.psh = C::PSH;
  int bad = 0;
  std::vector<int> seen(k.pitch, 0);
  for (int c = 0; c < k.pitch; ++c) {
    const int s = pm_split_col(k, c);
    if (s < 0 || s >= k.pitch || seen[s]++) ++bad;
    if ((s >= k.pitch / 2) != bool(c & 1)) ++bad;  // even columns in the first half, odd ones in the second
  }
  const int tiles_x = (nx + C::TX - 1) / C::TX;
  for (int bx = 0; bx < tiles_x; ++bx) {
    const int cs = PM_OFFC + 1 + bx * C::TX - C::H;  // first storage column of the tile
    if (cs < 0 || (cs & 1)) ++bad;
    const int pair0 = (cs + k.psh) >> 1;             // TMA x coordinate
    if (pair0 & 1) ++bad;                            // 16-byte alignment of the box rows
    for (int q = 0; q < 64; ++q) {
      const int cx = cs + 2 * q, cy = cx + 1;
      if (cy >= k.pitch) break;                      // zero-filled by TMA beyond the row
      if (pm_split_col(k, cx) != pair0 + q && pair0 + q < k.pitch / 2) ++bad;
      if (pm_split_col(k, cy) != k.pitch / 2 + pair0 + q && pair0 + q < k.pitch / 2) ++bad;
    }
  }
  if (bad) std::printf("METHOD %d T %d nx %d: %d violations\n", METHOD, T, nx, bad);
  return bad;
}

int main() {
  int bad = 0;
  for (int nx : {48, 93, 117, 250, 300, 1024, 8192, 16384}) {
    bad += check_shape<PM_PPE_SOR_RB, 1>(nx) + check_shape<PM_PPE_SOR_RB, 2>(nx) + check_shape<PM_PPE_SOR_RB, 3>(nx) + check_shape<PM_PPE_SOR_RB, 4>(nx);
    bad += check_shape<PM_PPE_JACOBI, 1>(nx) + check_shape<PM_PPE_JACOBI, 2>(nx) + check_shape<PM_PPE_JACOBI, 4>(nx);
  }
  std::printf(bad ? "FAIL\n" : "split layout ok\n");
  return bad ? 1 : 0;
}
