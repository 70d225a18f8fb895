// oracle/ref_cpu.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A runtime-parameter restatement of the projection-method time step of
// tjjones6/Computational-Fluid-Dynamics (cavity-01.cpp, channel-01.cpp,
// backwards_step-01.cpp), one function per reference member function, each
// citing the reference lines whose expression trees it follows.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this; the product (libpm.so) never does.
//
// Parity status: PINNED.  The reference ships no golden vectors, so the oracle
// is pinned by running the reference itself: oracle/build_ref.sh compiles the
// unmodified reference sources (where they lie under /root/reference) into
// oracle/_ref/, tests/golden/make_golden.py records their fields, and
// tests/test_oracle.py requires this file to reproduce them bit for bit
// (lexicographic SOR, -O2 -ffp-contract=off).
//
// The Jacobi and red-black orderings do not exist in the reference
// (SURVEY §0 fact 2); they reuse the reference's per-cell expression trees
// and change only which iterate each neighbour is read from.
//
// Build: g++ -std=c++17 -O2 -ffp-contract=off -shared -fPIC ref_cpu.cpp -o libref_cpu.so
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "../include/pm.h"

namespace {

// Dense reference-shaped field, element (j,i) at j*cols+i  (Field = vector<vector<double>>, cavity-01.cpp:45)
struct Arr {
  int rows = 0, cols = 0;
  std::vector<double> a;
  void init(int r, int c) { rows = r; cols = c; a.assign(size_t(r) * c, 0.0); }
  double& operator()(int j, int i) { return a[size_t(j) * cols + i]; }
  double operator()(int j, int i) const { return a[size_t(j) * cols + i]; }
};

struct Oracle {
  pm_config c;
  int nx, ny;
  Arr u, v, us, vs, f, p;
  std::vector<uint8_t> fluid;  // (ny+2)*(nx+2)
  bool is_fluid(int j, int i) const { return fluid[size_t(j) * (nx + 2) + i] != 0; }
};

uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// U(-1,1) keyed by (seed, field, flat reference index) — same on the device (pm_fill_random).
double synth(uint64_t seed, int field, uint64_t flat) {
  const uint64_t z = splitmix64(seed ^ splitmix64((uint64_t(field) << 56) ^ flat));
  return double(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}

// ---------------------------------------------------------------- a1: omega
// cavity-01.cpp:74-78
double omega_cavity(int n) {
  const double pi = 3.14159265358979323846;
  const double rj = std::cos(pi / (n + 1));
  return 2.0 / (1.0 + std::sqrt(1.0 - rj * rj));
}
// channel-01.cpp:76-81, backwards_step-01.cpp:77-82
double omega_2d(int nx, int ny) {
  const double pi = 3.14159265358979323846;
  const double r = 0.5 * (std::cos(pi / (nx + 1)) + std::cos(pi / (ny + 1)));
  return 2.0 / (1.0 + std::sqrt(std::max(1e-14, 1.0 - r * r)));
}

// ---------------------------------------------------------------- a16: mask
// backwards_step-01.cpp:500-520: fluid iff interior and (i > step_i || j <= inlet_j_max)
void default_mask(Oracle& o) {
  o.fluid.assign(size_t(o.ny + 2) * (o.nx + 2), 0);
  for (int j = 1; j <= o.ny; ++j)
    for (int i = 1; i <= o.nx; ++i) {
      bool fl = true;
      if (o.c.case_id == PM_CASE_STEP) fl = (i > o.c.step_i_location) || (j <= o.c.inlet_j_max);
      o.fluid[size_t(j) * (o.nx + 2) + i] = fl ? 1 : 0;
    }
}

// ---------------------------------------------------------------- a3/a4/a5: velocity BC
// cavity-01.cpp:523-543
void bc_cavity(Oracle& o, Arr& U, Arr& V) {
  const int nx = o.nx, ny = o.ny;
  for (int i = 0; i <= nx; ++i) U(ny + 1, i) = 2.0 * o.c.u_ref - U(ny, i);
  for (int i = 0; i <= nx; ++i) U(0, i) = -U(1, i);
  for (int j = 0; j <= ny; ++j) V(j, nx + 1) = -V(j, nx);
  for (int j = 0; j <= ny; ++j) V(j, 0) = -V(j, 1);
}
// channel-01.cpp:513-529; backwards_step-01.cpp:616-683 (inlet split + solid-face zeroing)
void bc_channel(Oracle& o, Arr& U, Arr& V) {
  const int nx = o.nx, ny = o.ny;
  const bool step = o.c.case_id == PM_CASE_STEP;
  for (int j = 1; j <= ny; ++j) U(j, 0) = (!step || j <= o.c.inlet_j_max) ? o.c.u_ref : 0.0;
  for (int j = 0; j <= ny; ++j) V(j, 0) = 0.0;
  for (int j = 1; j <= ny; ++j) U(j, nx) = U(j, nx - 1);
  for (int j = 0; j <= ny; ++j) V(j, nx + 1) = V(j, nx);
  for (int i = 1; i <= nx; ++i) V(0, i) = 0.0;
  for (int i = 0; i <= nx; ++i) U(0, i) = -U(1, i);
  for (int i = 1; i <= nx; ++i) V(ny, i) = 0.0;
  for (int i = 0; i <= nx; ++i) U(ny + 1, i) = -U(ny, i);
  if (!step) return;
  for (int j = 1; j <= ny; ++j)
    for (int i = 1; i <= nx; ++i) {
      if (o.is_fluid(j, i)) continue;
      if (i < nx && o.is_fluid(j, i + 1)) U(j, i) = 0.0;
      if (i > 1 && o.is_fluid(j, i - 1)) U(j, i - 1) = 0.0;
      if (j < ny && o.is_fluid(j + 1, i)) V(j, i) = 0.0;
      if (j > 1 && o.is_fluid(j - 1, i)) V(j - 1, i) = 0.0;
    }
}
void apply_bc(Oracle& o, int which) {
  Arr& U = which ? o.us : o.u;
  Arr& V = which ? o.vs : o.v;
  if (o.c.case_id == PM_CASE_CAVITY) bc_cavity(o, U, V); else bc_channel(o, U, V);
}

// ---------------------------------------------------------------- a6: predictor
// cavity-01.cpp:548-603 (hi, h2i) == channel-01.cpp:546-603 (idx, idy, idx2, idy2);
// backwards_step-01.cpp:745-820 adds the face-validity test.
void predict(Oracle& o) {
  const int nx = o.nx, ny = o.ny;
  const double idx = 1.0 / o.c.dx, idy = 1.0 / o.c.dy;
  const double idx2 = 1.0 / (o.c.dx * o.c.dx), idy2 = 1.0 / (o.c.dy * o.c.dy);
  const double nu = o.c.nu, dt = o.c.dt;
  const bool step = o.c.case_id == PM_CASE_STEP;
  const Arr& u = o.u; const Arr& v = o.v;
  for (int j = 1; j <= ny; ++j)
    for (int i = 1; i <= nx - 1; ++i) {
      if (step && !(o.is_fluid(j, i) || o.is_fluid(j, i + 1))) { o.us(j, i) = 0.0; continue; }
      const double diff = nu * ((u(j, i + 1) - 2.0 * u(j, i) + u(j, i - 1)) * idx2 +
                                (u(j + 1, i) - 2.0 * u(j, i) + u(j - 1, i)) * idy2);
      const double ue = 0.5 * (u(j, i) + u(j, i + 1));
      const double uw = 0.5 * (u(j, i - 1) + u(j, i));
      const double cx = (ue * ue - uw * uw) * idx;
      const double vn = 0.5 * (v(j, i) + v(j, i + 1));
      const double vs_ = 0.5 * (v(j - 1, i) + v(j - 1, i + 1));
      const double un = 0.5 * (u(j + 1, i) + u(j, i));
      const double us_ = 0.5 * (u(j - 1, i) + u(j, i));
      const double cy = (vn * un - vs_ * us_) * idy;
      o.us(j, i) = u(j, i) + dt * (diff - cx - cy);
    }
  for (int j = 1; j <= ny - 1; ++j)
    for (int i = 1; i <= nx; ++i) {
      if (step && !(o.is_fluid(j, i) || o.is_fluid(j + 1, i))) { o.vs(j, i) = 0.0; continue; }
      const double diff = nu * ((v(j, i + 1) - 2.0 * v(j, i) + v(j, i - 1)) * idx2 +
                                (v(j + 1, i) - 2.0 * v(j, i) + v(j - 1, i)) * idy2);
      const double vn = 0.5 * (v(j, i) + v(j + 1, i));
      const double vs_ = 0.5 * (v(j - 1, i) + v(j, i));
      const double cy = (vn * vn - vs_ * vs_) * idy;
      const double ue = 0.5 * (u(j, i) + u(j + 1, i));
      const double uw = 0.5 * (u(j, i - 1) + u(j + 1, i - 1));
      const double ve = 0.5 * (v(j, i) + v(j, i + 1));
      const double vw = 0.5 * (v(j, i - 1) + v(j, i));
      const double cx = (ue * ve - uw * vw) * idx;
      o.vs(j, i) = v(j, i) + dt * (diff - cy - cx);
    }
}

// ---------------------------------------------------------------- a7/a8: source (+mean)
// cavity-01.cpp:622-630; channel-01.cpp:608-629; backwards_step-01.cpp:825-866
double source(Oracle& o) {
  const int nx = o.nx, ny = o.ny;
  double mx = 0.0;
  if (o.c.case_id == PM_CASE_CAVITY) {
    const double hi = 1.0 / o.c.dx, dti = 1.0 / o.c.dt;
    for (int j = 1; j <= ny; ++j)
      for (int i = 1; i <= nx; ++i) {
        o.f(j, i) = dti * o.c.rho * ((o.us(j, i) - o.us(j, i - 1)) * hi + (o.vs(j, i) - o.vs(j - 1, i)) * hi);
        mx = std::max(mx, std::abs(o.f(j, i)));
      }
    return mx;
  }
  const bool step = o.c.case_id == PM_CASE_STEP;
  const double idx = 1.0 / o.c.dx, idy = 1.0 / o.c.dy;
  const double coeff = o.c.rho / o.c.dt;
  for (int j = 1; j <= ny; ++j)
    for (int i = 1; i <= nx; ++i) {
      if (step && !o.is_fluid(j, i)) { o.f(j, i) = 0.0; continue; }
      o.f(j, i) = coeff * ((o.us(j, i) - o.us(j, i - 1)) * idx + (o.vs(j, i) - o.vs(j - 1, i)) * idy);
      mx = std::max(mx, std::abs(o.f(j, i)));
    }
  if (mx > 0) {
    double mean = 0.0; int cnt = 0;
    for (int j = 1; j <= ny; ++j)
      for (int i = 1; i <= nx; ++i)
        if (!step || o.is_fluid(j, i)) { mean += o.f(j, i); ++cnt; }
    if (cnt > 0) {
      mean /= static_cast<double>(cnt);
      for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i)
          if (!step || o.is_fluid(j, i)) o.f(j, i) -= mean;
    }
  }
  return mx;
}

// ---------------------------------------------------------------- a12: pressure ghosts
// channel-01.cpp:531-541; backwards_step-01.cpp:685-740
void pressure_ghosts(Oracle& o, Arr& p) {
  const int nx = o.nx, ny = o.ny;
  for (int j = 1; j <= ny; ++j) p(j, 0) = p(j, 1);
  for (int j = 1; j <= ny; ++j) p(j, nx + 1) = 0.0;
  for (int i = 1; i <= nx; ++i) { p(0, i) = p(1, i); p(ny + 1, i) = p(ny, i); }
  if (o.c.case_id != PM_CASE_STEP) return;
  for (int j = 1; j <= ny; ++j)
    for (int i = 1; i <= nx; ++i) {
      if (o.is_fluid(j, i)) continue;
      double s = 0.0; int n = 0;
      if (i > 1 && o.is_fluid(j, i - 1)) { s += p(j, i - 1); ++n; }
      if (i < nx && o.is_fluid(j, i + 1)) { s += p(j, i + 1); ++n; }
      if (j > 1 && o.is_fluid(j - 1, i)) { s += p(j - 1, i); ++n; }
      if (j < ny && o.is_fluid(j + 1, i)) { s += p(j + 1, i); ++n; }
      if (n > 0) p(j, i) = s / n;
    }
}

// ---------------------------------------------------------------- a10/a11: one cell update
// cavity-01.cpp:644-654.  pE,pN / pW,pS are whatever iterate the ordering dictates.
inline double upd_cavity(const Oracle& o, int j, int i, double pc, double pe, double pw, double pn, double ps, double f) {
  const int ew = (i > 1) ? 1 : 0, ee = (i < o.nx) ? 1 : 0, en = (j < o.ny) ? 1 : 0, es = 1;
  const int nc = ew + ee + en + es;
  const double h = o.c.dx, w = o.c.omega;
  return pc * (1.0 - w) + (w / nc) * ((ee * pe + ew * pw) + (en * pn + es * ps) - f * (h * h));
}
// channel-01.cpp:659-666
inline double upd_channel(const Oracle& o, double idx2, double idy2, double denom, double pc, double pe, double pw, double pn, double ps, double f) {
  const double sum = idx2 * (pe + pw) + idy2 * (pn + ps);
  const double pgs = (sum - f) / denom;
  return (1.0 - o.c.omega) * pc + o.c.omega * pgs;
}
// a13: cavity-01.cpp:659-677; channel-01.cpp:673-681; backwards_step-01.cpp:917-930
double residual_rows(const Oracle& o, const Arr& p, int ja, int jb) {
  const int nx = o.nx;
  double mx = 0.0;
  if (o.c.case_id == PM_CASE_CAVITY) {
    const double h2i = 1.0 / (o.c.dx * o.c.dx);
    for (int j = ja; j <= jb; ++j)
      for (int i = 1; i <= nx; ++i) {
        const int ew = (i > 1) ? 1 : 0, ee = (i < nx) ? 1 : 0, en = (j < o.ny) ? 1 : 0, es = 1;
        const double r = h2i * (ee * (p(j, i + 1) - p(j, i)) + ew * (p(j, i - 1) - p(j, i)) +
                                en * (p(j + 1, i) - p(j, i)) + es * (p(j - 1, i) - p(j, i))) - o.f(j, i);
        mx = std::max(mx, std::abs(r));
      }
    return mx;
  }
  const bool step = o.c.case_id == PM_CASE_STEP;
  const double idx2 = 1.0 / (o.c.dx * o.c.dx), idy2 = 1.0 / (o.c.dy * o.c.dy);
  for (int j = ja; j <= jb; ++j)
    for (int i = 1; i <= nx; ++i) {
      if (step && !o.is_fluid(j, i)) continue;
      const double lap = (p(j, i + 1) - 2.0 * p(j, i) + p(j, i - 1)) * idx2 + (p(j + 1, i) - 2.0 * p(j, i) + p(j - 1, i)) * idy2;
      mx = std::max(mx, std::abs(lap - o.f(j, i)));
    }
  return mx;
}

// One sweep over rows ja..jb.  colour: -1 = all cells in lexicographic order (in place: the
// reference's p_prev copy / buffer swap is provably redundant);
// 0/1 = only cells with (i+j)%2 == colour (red-black half sweep, in place).
void sweep_inplace_rows(Oracle& o, Arr& p, int colour, int ja, int jb) {
  const int nx = o.nx;
  if (o.c.case_id == PM_CASE_CAVITY) {
    for (int j = ja; j <= jb; ++j)
      for (int i = 1; i <= nx; ++i) {
        if (colour >= 0 && ((i + j) & 1) != colour) continue;
        p(j, i) = upd_cavity(o, j, i, p(j, i), p(j, i + 1), p(j, i - 1), p(j + 1, i), p(j - 1, i), o.f(j, i));
      }
    return;
  }
  const bool step = o.c.case_id == PM_CASE_STEP;
  const double idx2 = 1.0 / (o.c.dx * o.c.dx), idy2 = 1.0 / (o.c.dy * o.c.dy);
  const double denom = 2.0 * (idx2 + idy2);
  for (int j = ja; j <= jb; ++j)
    for (int i = 1; i <= nx; ++i) {
      if (colour >= 0 && ((i + j) & 1) != colour) continue;
      if (step && !o.is_fluid(j, i)) continue;
      p(j, i) = upd_channel(o, idx2, idy2, denom, p(j, i), p(j, i + 1), p(j, i - 1), p(j + 1, i), p(j - 1, i), o.f(j, i));
    }
}
// Jacobi: every neighbour from `src` (previous iterate), result into `dst`; cells not updated are copied.
void sweep_jacobi_rows(Oracle& o, const Arr& src, Arr& dst, int ja, int jb) {
  const int nx = o.nx;
  const bool cav = o.c.case_id == PM_CASE_CAVITY, step = o.c.case_id == PM_CASE_STEP;
  const double idx2 = 1.0 / (o.c.dx * o.c.dx), idy2 = 1.0 / (o.c.dy * o.c.dy);
  const double denom = 2.0 * (idx2 + idy2);
  for (int j = ja; j <= jb; ++j)
    for (int i = 1; i <= nx; ++i) {
      if (step && !o.is_fluid(j, i)) { dst(j, i) = src(j, i); continue; }
      dst(j, i) = cav ? upd_cavity(o, j, i, src(j, i), src(j, i + 1), src(j, i - 1), src(j + 1, i), src(j - 1, i), o.f(j, i))
                      : upd_channel(o, idx2, idy2, denom, src(j, i), src(j, i + 1), src(j, i - 1), src(j + 1, i), src(j - 1, i), o.f(j, i));
    }
}

// ---------------------------------------------------------------- a9/a14: PPE driver loop
// cavity-01.cpp:609-690; channel-01.cpp:635-688; backwards_step-01.cpp:872-939.
// Expects o.f already built by source() (the cavity builds it inside the same function, :622-630).
void ppe(Oracle& o, pm_ppe_result* out) {
  const int nx = o.nx, ny = o.ny;
  const bool cav = o.c.case_id == PM_CASE_CAVITY, step = o.c.case_id == PM_CASE_STEP;
  double mx = 0.0;
  for (int j = 1; j <= ny; ++j)
    for (int i = 1; i <= nx; ++i)
      if (!step || o.is_fluid(j, i)) mx = std::max(mx, std::abs(o.f(j, i)));
  double tol, res;
  if (cav) {
    o.p.init(ny + 2, nx + 2);  // cold start, :610-611
    tol = o.c.tol_factor * mx; // :632
    res = 1.0;                 // :618
  } else {
    tol = std::max(o.c.tol_factor * (mx > 0 ? mx : 1.0), o.c.abs_tol);  // channel :647
    res = tol + 1.0;                                                  // :649
  }
  Arr tmp;
  if (o.c.ppe_method == PM_PPE_JACOBI) tmp = o.p;
  double cheby_w = 1.0;
  int it = 0;
  while (res > tol && it < o.c.max_iters) {
    ++it;
    if (o.c.ppe_method == PM_PPE_SOR_LEX) {
      sweep_inplace_rows(o, o.p, -1, 1, ny);
    } else if (o.c.ppe_method == PM_PPE_SOR_RB) {
      sweep_inplace_rows(o, o.p, 0, 1, ny);
      sweep_inplace_rows(o, o.p, 1, 1, ny);
    } else if (o.c.ppe_method == PM_PPE_SOR_CHEBY) {
      // Red-black SOR with Chebyshev acceleration (not in the reference; README.md:39 asks for a better Poisson solver):
      // the factor of colour half-sweep q is w_0 = 1, w_1 = 1/(1 - rho2/2), w_q = 1/(1 - rho2 w_{q-1}/4), with rho2 the
      // squared Jacobi spectral radius behind the reference's omega (cavity-01.cpp:74-78).  Formulas: include/pm.h.
      const double w_opt = o.c.omega, rho2 = pmi_cheby_rho2(w_opt);
      for (int colour = 0; colour < 2; ++colour) {
        const int q = 2 * (it - 1) + colour;
        cheby_w = q == 0 ? 1.0 : pmi_cheby_next_omega(rho2, q, cheby_w);
        o.c.omega = cheby_w;
        sweep_inplace_rows(o, o.p, colour, 1, ny);
      }
      o.c.omega = w_opt;
    } else {
      sweep_jacobi_rows(o, o.p, tmp, 1, ny);
      // ghosts/solid cells of tmp carry over from the previous iterate until refreshed below
      for (int j = 1; j <= ny; ++j) { tmp(j, 0) = o.p(j, 0); tmp(j, nx + 1) = o.p(j, nx + 1); }
      for (int i = 0; i <= nx + 1; ++i) { tmp(0, i) = o.p(0, i); tmp(ny + 1, i) = o.p(ny + 1, i); }
      std::swap(o.p.a, tmp.a);
    }
    if (!cav) pressure_ghosts(o, o.p);
    res = residual_rows(o, o.p, 1, ny);
  }
  if (out) {
    out->iterations = it; out->hit_cap = it >= o.c.max_iters; out->residual = res;
    out->tolerance = tol; out->max_source = mx;
  }
}

// ---------------------------------------------------------------- a15: correction
// cavity-01.cpp:695-711; channel-01.cpp:693-702; backwards_step-01.cpp:944-976
void correct(Oracle& o) {
  const int nx = o.nx, ny = o.ny;
  if (o.c.case_id == PM_CASE_CAVITY) {
    const double c = o.c.dt / o.c.dx;
    for (int j = 1; j <= ny; ++j)
      for (int i = 1; i <= nx - 1; ++i) o.u(j, i) = o.us(j, i) - c * o.c.rho * (o.p(j, i + 1) - o.p(j, i));
    for (int j = 1; j <= ny - 1; ++j)
      for (int i = 1; i <= nx; ++i) o.v(j, i) = o.vs(j, i) - c * o.c.rho * (o.p(j + 1, i) - o.p(j, i));
    return;
  }
  const bool step = o.c.case_id == PM_CASE_STEP;
  for (int j = 1; j <= ny; ++j)
    for (int i = 1; i <= nx - 1; ++i) {
      const bool valid = !step || (i == nx - 1) || o.is_fluid(j, i) || o.is_fluid(j, i + 1);
      o.u(j, i) = valid ? o.us(j, i) - (o.c.dt / (o.c.rho * o.c.dx)) * (o.p(j, i + 1) - o.p(j, i)) : 0.0;
    }
  for (int j = 1; j <= ny - 1; ++j)
    for (int i = 1; i <= nx; ++i) {
      const bool valid = !step || (j == ny - 1) || o.is_fluid(j, i) || o.is_fluid(j + 1, i);
      o.v(j, i) = valid ? o.vs(j, i) - (o.c.dt / (o.c.rho * o.c.dy)) * (o.p(j + 1, i) - o.p(j, i)) : 0.0;
    }
}

// ---------------------------------------------------------------- k11: diagnostics
// cavity-01.cpp:741-766; channel-01.cpp:733-759; backwards_step-01.cpp:1018-1051
void diagnostics(const Oracle& o, double* max_div, double* avg_ke) {
  const int nx = o.nx, ny = o.ny;
  const bool cav = o.c.case_id == PM_CASE_CAVITY, step = o.c.case_id == PM_CASE_STEP;
  double ke = 0.0, md = 0.0; int cnt = 0;
  for (int j = 1; j <= ny; ++j)
    for (int i = 1; i <= nx; ++i) {
      if (step && !o.is_fluid(j, i)) continue;
      const double uc = 0.5 * (o.u(j, i - 1) + o.u(j, i));
      const double vc = 0.5 * (o.v(j - 1, i) + o.v(j, i));
      ke += 0.5 * (uc * uc + vc * vc);
      ++cnt;
    }
  const double idx = 1.0 / o.c.dx, idy = 1.0 / o.c.dy;
  for (int j = 1; j <= ny; ++j)
    for (int i = 1; i <= nx; ++i) {
      if (step && !o.is_fluid(j, i)) continue;
      const double d = cav ? (o.u(j, i) - o.u(j, i - 1) + o.v(j, i) - o.v(j - 1, i)) * idx
                           : (o.u(j, i) - o.u(j, i - 1)) * idx + (o.v(j, i) - o.v(j - 1, i)) * idy;
      md = std::max(md, std::abs(d));
    }
  *max_div = md;
  *avg_ke = step ? (cnt > 0 ? ke / cnt : 0.0) : ke / (nx * ny);
}

// One projection step in the case's call order (cavity-01.cpp:387-390; channel-01.cpp:368-375).
void step_once(Oracle& o, pm_ppe_result* r) {
  if (o.c.case_id == PM_CASE_CAVITY) {
    apply_bc(o, 0); predict(o); source(o); ppe(o, r); correct(o);
  } else {
    predict(o); apply_bc(o, 1); source(o); ppe(o, r); correct(o); apply_bc(o, 0);
  }
}

Arr* field(Oracle& o, int id) {
  switch (id) {
    case PM_FIELD_U: return &o.u; case PM_FIELD_V: return &o.v; case PM_FIELD_P: return &o.p;
    case PM_FIELD_USTAR: return &o.us; case PM_FIELD_VSTAR: return &o.vs; case PM_FIELD_F: return &o.f;
  }
  return nullptr;
}

}  // namespace

extern "C" {

// a1/a2: parameter derivation with the reference's expression trees
// (cavity-01.cpp:356-363; channel-01.cpp:337-344; backwards_step-01.cpp:378-387).
int orc_config_init(pm_config* c, int case_id, int nx, int ny, double re, double dt) {
  if (!c) return PM_ERR_INVALID_ARGUMENT;
  std::memset(c, 0, sizeof(*c));
  c->struct_size = sizeof(pm_config);
  c->case_id = case_id;
  c->rho = 1.0; c->u_ref = 1.0; c->max_iters = 10000;
  c->ppe_method = PM_PPE_SOR_LEX; c->exact_arith = 1; c->nranks = 1; c->device = -1;
  if (case_id == PM_CASE_CAVITY) {
    const double L = 1.0, H = 1.0;
    const int n = nx > 0 ? nx : 63;
    c->re = re > 0 ? re : 1000.0; c->cfl = 0.5; c->final_time = 20.0;
    c->tol_factor = 1e-9; c->abs_tol = 0.0; c->print_interval = 100; c->save_interval = 100;
    c->nu = c->rho * c->u_ref * L / c->re;
    const double h = L / n;
    c->dx = c->dy = h;
    c->omega = omega_cavity(n);
    c->dt = dt > 0 ? dt : c->cfl * std::min(0.25 * h * h / c->nu, h / c->u_ref);
    c->nx = static_cast<int>(L * n);
    c->ny = ny > 0 ? ny : static_cast<int>(H * n);
    c->lx = L; c->ly = c->ny * h;
  } else if (case_id == PM_CASE_CHANNEL || case_id == PM_CASE_STEP) {
    const bool st = case_id == PM_CASE_STEP;
    const double L = st ? 8.0 : 3.0, H = st ? 2.0 : 1.0, Hin = 1.0;
    c->nx = nx > 0 ? nx : (st ? 8 * 32 : 93);
    c->ny = ny > 0 ? ny : (st ? 32 : 31);
    c->re = re > 0 ? re : 100.0; c->cfl = st ? 0.2 : 0.25; c->final_time = st ? 15.0 : 10.0;
    c->tol_factor = 1e-7; c->abs_tol = 1e-10;
    c->print_interval = st ? 10 : 100; c->save_interval = st ? 10 : 100;
    c->nu = c->u_ref * Hin / c->re;
    c->dx = L / c->nx; c->dy = H / c->ny;
    c->omega = omega_2d(c->nx, c->ny);
    const double m = std::min(c->dx, c->dy);
    c->dt = dt > 0 ? dt : c->cfl * std::min(0.25 * m * m / c->nu, m / std::max(1e-12, c->u_ref));
    c->lx = L; c->ly = H;
    if (st) {
      c->step_i_location = static_cast<int>(2.0 / c->dx);
      c->inlet_j_max = static_cast<int>(Hin / c->dy);
    }
  } else {
    return PM_ERR_INVALID_ARGUMENT;
  }
  c->total_steps = static_cast<int>(c->final_time / c->dt);
  return PM_OK;
}

void* orc_create(const pm_config* c) {
  if (!c || c->nx <= 0 || c->ny <= 0) return nullptr;
  Oracle* o = new (std::nothrow) Oracle();
  if (!o) return nullptr;
  o->c = *c; o->nx = c->nx; o->ny = c->ny;
  const int nx = c->nx, ny = c->ny;
  o->p.init(ny + 2, nx + 2); o->f.init(ny + 2, nx + 2);
  o->u.init(ny + 2, nx + 1); o->us.init(ny + 2, nx + 1);
  o->v.init(ny + 1, nx + 2); o->vs.init(ny + 1, nx + 2);
  default_mask(*o);
  return o;
}
void orc_destroy(void* h) { delete static_cast<Oracle*>(h); }
void orc_set_method(void* h, int method) { static_cast<Oracle*>(h)->c.ppe_method = method; }
void orc_set_max_iters(void* h, int k) { static_cast<Oracle*>(h)->c.max_iters = k; }
void orc_set_omega(void* h, double w) { static_cast<Oracle*>(h)->c.omega = w; }

double* orc_field(void* h, int id) { Arr* a = field(*static_cast<Oracle*>(h), id); return a ? a->a.data() : nullptr; }
size_t orc_field_count(void* h, int id) { Arr* a = field(*static_cast<Oracle*>(h), id); return a ? a->a.size() : 0; }
uint8_t* orc_mask(void* h) { return static_cast<Oracle*>(h)->fluid.data(); }

void orc_fill_random_scaled(void* h, uint64_t seed, double amplitude) {
  Oracle& o = *static_cast<Oracle*>(h);
  for (int id = 0; id < PM_FIELD_COUNT; ++id) {
    Arr* a = field(o, id);
    for (size_t k = 0; k < a->a.size(); ++k) a->a[k] = amplitude * synth(seed, id, k);
  }
}
void orc_fill_random(void* h, uint64_t seed) { orc_fill_random_scaled(h, seed, 1.0); }
double orc_synth(uint64_t seed, int fieldid, uint64_t flat) { return synth(seed, fieldid, flat); }

void orc_apply_bc(void* h, int which) { apply_bc(*static_cast<Oracle*>(h), which); }
void orc_predict(void* h) { predict(*static_cast<Oracle*>(h)); }
double orc_source(void* h) { return source(*static_cast<Oracle*>(h)); }
void orc_ppe_solve(void* h, pm_ppe_result* r) { ppe(*static_cast<Oracle*>(h), r); }
void orc_correct(void* h) { correct(*static_cast<Oracle*>(h)); }
void orc_step(void* h, int n, pm_ppe_result* r) {
  pm_ppe_result tmp{};
  for (int k = 0; k < n; ++k) step_once(*static_cast<Oracle*>(h), &tmp);
  if (r) *r = tmp;
}
void orc_diagnostics(void* h, double* md, double* ke) { diagnostics(*static_cast<Oracle*>(h), md, ke); }

// Row-range building blocks for the slab-decomposition tests (tests/test_slab_gloo.py):
// the caller owns rows ja..jb of p and exchanges halo rows between calls.
void orc_sweep_rows(void* h, int colour, int ja, int jb) { Oracle& o = *static_cast<Oracle*>(h); sweep_inplace_rows(o, o.p, colour, ja, jb); }
void orc_jacobi_rows(void* h, const double* src, int ja, int jb) {
  Oracle& o = *static_cast<Oracle*>(h);
  Arr s; s.rows = o.p.rows; s.cols = o.p.cols; s.a.assign(src, src + o.p.a.size());
  sweep_jacobi_rows(o, s, o.p, ja, jb);
}
void orc_pressure_ghosts(void* h) { Oracle& o = *static_cast<Oracle*>(h); pressure_ghosts(o, o.p); }
double orc_residual_rows(void* h, int ja, int jb) { Oracle& o = *static_cast<Oracle*>(h); return residual_rows(o, o.p, ja, jb); }

}  // extern "C"
