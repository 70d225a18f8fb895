// oracle/ref_wrap.cpp — phase-level C access to the UNMODIFIED reference solver.  TEST INFRASTRUCTURE ONLY.
//
// The reference translation unit is #included from where it lies (REF_SRC, e.g.
// /root/reference/cavity-01.cpp, or a /tmp copy whose compile-time constants
// were changed by oracle/build_ref.sh); nothing of it is copied into this repo.
// `private` is made `public` for that one include so the wrapper can call the
// solver's own member functions and read its own fields; main() is renamed.
// Output: oracle/_ref/libref_<name>.so (git-ignored, travels to the GPU box).
//
// Build (see oracle/build_ref.sh):
//   g++ -std=c++17 -O2 -ffp-contract=off -shared -fPIC -DREF_CASE=0
//       -DREF_SRC='"/root/reference/cavity-01.cpp"' ref_wrap.cpp -o _ref/libref_cavity_default.so
#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>
#include <sys/stat.h>
#include <unistd.h>

#define private public
#define main reference_main_unused
#include REF_SRC
#undef main
#undef private

#if REF_CASE == 0
using Solver = CavityFlow::CavitySolver;
using Field2 = CavityFlow::Field;
#elif REF_CASE == 1
using Solver = ChannelFlow::ChannelSolver;
using Field2 = ChannelFlow::Field;
#else
using Solver = BackwardsStepFlow::BackwardsStepSolver;
using Field2 = BackwardsStepFlow::Field;
#endif

namespace {
struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };
struct Quiet {  // silence the reference's banner / log lines while we drive it
  NullBuf nb; std::streambuf *o, *e;
  Quiet() : o(std::cout.rdbuf(&nb)), e(std::cerr.rdbuf(&nb)) {}
  ~Quiet() { std::cout.rdbuf(o); std::cerr.rdbuf(e); }
};
Field2* pick(Solver* s, int id) {
  switch (id) {
    case 0: return &s->u_corrected; case 1: return &s->v_corrected; case 2: return &s->pressure;
    case 3: return &s->u_tentative; case 4: return &s->v_tentative; case 5: return &s->source_term;
  }
  return nullptr;
}
}  // namespace

extern "C" {

int ref_case(void) { return REF_CASE; }

// Constructs the reference solver.  Its constructor creates ./vtk_output (and the channel/step
// ones write frame 0 there), so construction happens inside a scratch directory.
void* ref_create(void) {
  Quiet q;
  char tmpl[] = "/tmp/pm_ref_XXXXXX";
  char* d = mkdtemp(tmpl);
  char cwd[4096];
  if (!d || !getcwd(cwd, sizeof cwd) || chdir(d) != 0) return nullptr;
  Solver* s = nullptr;
  try { s = new Solver(); } catch (...) { s = nullptr; }
  if (chdir(cwd) != 0) { /* keep going */ }
  std::error_code ec; std::filesystem::remove_all(d, ec);
  return s;
}
void ref_destroy(void* h) { delete static_cast<Solver*>(h); }

int ref_nx(void* h) { return static_cast<Solver*>(h)->i_max; }
int ref_ny(void* h) { return static_cast<Solver*>(h)->j_max; }
double ref_dt(void* h) { return static_cast<Solver*>(h)->time_step; }
double ref_omega(void* h) { return static_cast<Solver*>(h)->optimal_omega; }
double ref_nu(void* h) { return static_cast<Solver*>(h)->kinematic_viscosity; }
int ref_total_steps(void* h) { return static_cast<Solver*>(h)->total_time_steps; }
#if REF_CASE == 0
double ref_dx(void* h) { return static_cast<Solver*>(h)->grid_spacing; }
double ref_dy(void* h) { return static_cast<Solver*>(h)->grid_spacing; }
int ref_max_iters(void*) { return Solver::max_sor_iterations; }
#else
double ref_dx(void* h) { return static_cast<Solver*>(h)->dx; }
double ref_dy(void* h) { return static_cast<Solver*>(h)->dy; }
int ref_max_iters(void*) { return Solver::MAX_SOR_ITERS; }
#endif

size_t ref_field_count(void* h, int id) {
  Field2* f = pick(static_cast<Solver*>(h), id);
  return f ? f->size() * (*f)[0].size() : 0;
}
void ref_get(void* h, int id, double* out) {
  Field2* f = pick(static_cast<Solver*>(h), id);
  for (auto& row : *f) { std::memcpy(out, row.data(), row.size() * sizeof(double)); out += row.size(); }
}
void ref_set(void* h, int id, const double* in) {
  Field2* f = pick(static_cast<Solver*>(h), id);
  for (auto& row : *f) { std::memcpy(row.data(), in, row.size() * sizeof(double)); in += row.size(); }
}
void ref_get_mask(void* h, uint8_t* out) {
  Solver* s = static_cast<Solver*>(h);
#if REF_CASE == 2
  for (auto& row : s->is_fluid) for (bool b : row) *out++ = b ? 1 : 0;
#else
  for (int j = 0; j <= s->j_max + 1; ++j)
    for (int i = 0; i <= s->i_max + 1; ++i) *out++ = (j >= 1 && j <= s->j_max && i >= 1 && i <= s->i_max) ? 1 : 0;
#endif
}

void ref_apply_bc(void* h, int which) {
  Solver* s = static_cast<Solver*>(h);
#if REF_CASE == 0
  (void)which; s->applyBoundaryConditions();
#else
  if (which) s->applyVelocityBC(s->u_tentative, s->v_tentative); else s->applyBoundaryConditions();
#endif
}
void ref_predict(void* h) { static_cast<Solver*>(h)->computeTentativeVelocities(); }
void ref_source(void* h) {
#if REF_CASE == 0
  (void)h;  // the cavity builds its source inside solverPressurePoisson (cavity-01.cpp:622-630)
#else
  static_cast<Solver*>(h)->buildSourceTerm();
#endif
}
void ref_ppe(void* h, int* iters, double* res) {
  Quiet q;
  auto r = static_cast<Solver*>(h)->solverPressurePoisson();
  if (iters) *iters = r.first;
  if (res) *res = r.second;
}
void ref_correct(void* h) { static_cast<Solver*>(h)->applyPressureCorrection(); }

// The body of run()'s time loop without logging/export (cavity-01.cpp:387-390; channel-01.cpp:368-375).
void ref_step(void* h, int n, int* iters, double* res) {
  Quiet q;
  Solver* s = static_cast<Solver*>(h);
  for (int k = 0; k < n; ++k) {
#if REF_CASE == 0
    s->applyBoundaryConditions();
    s->computeTentativeVelocities();
    auto r = s->solverPressurePoisson();
    s->applyPressureCorrection();
#else
    s->computeTentativeVelocities();
    s->applyVelocityBC(s->u_tentative, s->v_tentative);
    s->buildSourceTerm();
    auto r = s->solverPressurePoisson();
    s->applyPressureCorrection();
    s->applyBoundaryConditions();
#endif
    if (iters) *iters = r.first;
    if (res) *res = r.second;
  }
}
// Wall-clock seconds for n steps (bench.py --impl reference).
double ref_time_steps(void* h, int n) {
  auto t0 = std::chrono::steady_clock::now();
  ref_step(h, n, nullptr, nullptr);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
