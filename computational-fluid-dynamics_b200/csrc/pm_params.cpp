// pm_params.cpp — host-only parameter derivation of the three reference solvers.
// Compiled with -ffp-contract=off: dt, omega and the step count feed the numerics and must carry the
// same bits as the reference's constructor initialisers.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "../../include/pm.h"

namespace {
constexpr double kPi = 3.14159265358979323846;

// compute_optimal_omega, cavity-01.cpp:74-78
double sor_omega_square(int n) {
  const double rho_j = std::cos(kPi / (n + 1));
  return 2.0 / (1.0 + std::sqrt(1.0 - rho_j * rho_j));
}
// computeOptimalOmega2D, channel-01.cpp:76-81 / backwards_step-01.cpp:77-82
double sor_omega_rect(int nx, int ny) {
  const double cx = std::cos(kPi / (nx + 1));
  const double cy = std::cos(kPi / (ny + 1));
  const double rho_j = 0.5 * (cx + cy);
  return 2.0 / (1.0 + std::sqrt(std::max(1e-14, 1.0 - rho_j * rho_j)));
}
}  // namespace

extern "C" int pm_config_init(pm_config* cfg, int case_id, int nx, int ny, double re, double dt) {
  if (cfg == nullptr) return PM_ERR_INVALID_ARGUMENT;
  if (case_id != PM_CASE_CAVITY && case_id != PM_CASE_CHANNEL && case_id != PM_CASE_STEP) return PM_ERR_INVALID_ARGUMENT;
  pm_config c;
  std::memset(&c, 0, sizeof c);
  c.struct_size = sizeof(pm_config);
  c.case_id = case_id;
  c.u_ref = 1.0;        // lid_velocity / INLET_VELOCITY
  c.rho = 1.0;          // density / DENSITY
  c.max_iters = 10000;  // max_sor_iterations / MAX_SOR_ITERS
  c.ppe_method = PM_PPE_SOR_RB;
  c.exact_arith = 0;
  c.kernel_path = PM_PATH_AUTO;
  c.device = -1;
  c.rank = 0;
  c.nranks = 1;

  if (case_id == PM_CASE_CAVITY) {
    // cavity-01.cpp:309-320 (constants), :356-363 (derived)
    const double length = 1.0, height = 1.0;
    const int n = nx > 0 ? nx : 63;
    c.re = re > 0 ? re : 1000.0;
    c.cfl = 0.5;
    c.final_time = 20.0;
    c.tol_factor = 1e-9;
    c.print_interval = 100;
    c.save_interval = 100;
    c.nu = c.rho * c.u_ref * length / c.re;
    const double h = length / n;
    c.dx = h;
    c.dy = h;
    c.omega = sor_omega_square(n);
    c.dt = dt > 0 ? dt : c.cfl * std::min(0.25 * h * h / c.nu, h / c.u_ref);
    c.nx = static_cast<int>(length * n);
    // The reference is square (j_max = int(cavity_height*n)); ny != nx keeps the same h (tall cavity).
    c.ny = ny > 0 ? ny : static_cast<int>(height * n);
    c.lx = length;
    c.ly = c.ny * h;
  } else {
    // channel-01.cpp:287-300,337-344; backwards_step-01.cpp:319-334,378-387
    const bool is_step = case_id == PM_CASE_STEP;
    const double length = is_step ? 8.0 : 3.0;
    const double height = is_step ? 2.0 : 1.0;  // HEIGHT / HEIGHT_TOTAL
    const double height_inlet = 1.0;            // HEIGHT / HEIGHT_INLET (enters nu)
    c.nx = nx > 0 ? nx : (is_step ? 8 * 32 : 93);
    c.ny = ny > 0 ? ny : (is_step ? 32 : 31);
    c.re = re > 0 ? re : 100.0;
    c.cfl = is_step ? 0.2 : 0.25;
    c.final_time = is_step ? 15.0 : 10.0;
    c.tol_factor = 1e-7;
    c.abs_tol = 1e-10;
    c.print_interval = is_step ? 10 : 100;
    c.save_interval = is_step ? 10 : 100;
    c.nu = c.u_ref * height_inlet / c.re;
    c.dx = length / c.nx;
    c.dy = height / c.ny;
    c.omega = sor_omega_rect(c.nx, c.ny);
    const double hmin = std::min(c.dx, c.dy);
    c.dt = dt > 0 ? dt : c.cfl * std::min(0.25 * hmin * hmin / c.nu, hmin / std::max(1e-12, c.u_ref));
    c.lx = length;
    c.ly = height;
    if (is_step) {
      c.step_i_location = static_cast<int>(2.0 / c.dx);        // STEP_LOCATION / dx, :386
      c.inlet_j_max = static_cast<int>(height_inlet / c.dy);   // :493
    }
  }
  c.total_steps = static_cast<int>(c.final_time / c.dt);
  *cfg = c;
  return PM_OK;
}

// Rows of the global grid owned by `rank`: interior rows j0+1 .. j0+ny_local (block distribution,
// the first ny % nranks ranks get one extra row).
extern "C" int pm_slab_range(int ny, int nranks, int rank, int* j0, int* ny_local) {
  if (ny <= 0 || nranks <= 0 || rank < 0 || rank >= nranks || nranks > ny) return PM_ERR_INVALID_ARGUMENT;
  const int base = ny / nranks, rem = ny % nranks;
  const int mine = base + (rank < rem ? 1 : 0);
  const int start = rank * base + std::min(rank, rem);
  if (j0) *j0 = start;
  if (ny_local) *ny_local = mine;
  return PM_OK;
}

extern "C" double pm_cheby_omega(double omega, int q) {
  const double rho2 = pmi_cheby_rho2(omega);
  double w = 1.0;
  for (int n = 1; n <= q; ++n) w = pmi_cheby_next_omega(rho2, n, w);
  return w;
}

extern "C" double pm_omega_mixed_bc(int case_id, int nx, int ny, double dx, double dy) {
  const double pi = 3.14159265358979323846;
  double rho;
  if (case_id == PM_CASE_CAVITY) {
    rho = 0.5 * (1.0 + std::cos(pi / (2.0 * ny + 1.0)));
  } else {
    const double ix2 = 1.0 / (dx * dx), iy2 = 1.0 / (dy * dy);
    rho = (ix2 * std::cos(pi / (2.0 * nx + 1.0)) + iy2) / (ix2 + iy2);
  }
  return 2.0 / (1.0 + std::sqrt(std::max(1e-14, 1.0 - rho * rho)));
}
