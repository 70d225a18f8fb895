// driver_main.cpp — the three drop-in drivers (cavity, channel, backwards_step), host C++17 over the C-ABI.
//
// One source, compiled three times with -DPM_DRIVER_CASE=0|1|2.  Each binary keeps the surface of the
// reference program it replaces: the README's flags (--Re --Nx --Ny --dt; README.md:125-126), the banner and
// per-step log lines on stdout, the convergence warnings on stderr, the exit codes, and the files
// vtk_output/<case>_%06d.vtk + <case>_animation.pvd (cavity-01.cpp:464-518,741-774; channel-01.cpp:450-504,
// 733-769; backwards_step-01.cpp:551-607,1018-1061).  All numerics run on the GPU through include/pm.h; the
// host only formats text.  Flags beyond the README's are extras and default to the reference behaviour.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pm.h"

#ifndef PM_DRIVER_CASE
#define PM_DRIVER_CASE 0
#endif

namespace {
constexpr const char* RESET = "\033[0m";
constexpr const char* RED = "\033[31m";
constexpr const char* GREEN = "\033[32m";
constexpr const char* YELLOW = "\033[33m";
constexpr const char* BLUE = "\033[34m";
constexpr const char* CYAN = "\033[36m";

struct CaseText {
  const char* vtk_base;      // file stem
  const char* vtk_title;     // second header line of the VTK file
  const char* banner;        // first banner line
  const char* start_msg;
};
constexpr CaseText kText[3] = {
    {"cavity_flow", "Lid-Driven Cavity Flow Data - Time: ", "=== Lid-Driven Cavity Flow Simulation ===", "Starting simulation...\n"},
    {"channel_flow", "Channel Flow Data - Time: ", "=== Channel Flow Simulation ===", "Starting simulation...\n"},
    {"backwards_step", "Backwards Step Flow Data - Time: ", "=== Backwards Step Flow Simulation ===", "Starting backwards step simulation...\n"},
};

struct Options {
  int nx = 0, ny = 0;
  double re = 0.0, dt = 0.0;
  int steps = -1;
  double tfinal = -1.0;
  int ppe = PM_PPE_SOR_RB;
  std::string omega = "";  // "" = the reference's factor; "mixed" = pm_omega_mixed_bc; or a number
  int max_iters = -1;
  int exact = 0;
  int path = PM_PATH_AUTO;
  int sweeps = 0;
  int print_interval = -1, save_interval = -1;
  bool vtk = true;
  std::string outdir = "vtk_output";
  int device = -1;
  int stop_after = 0;  // leave the time loop after this many steps (banner and "Step n/N" still show the full run)
  int gpus = 1;  // j-slabs, one host thread + one handle per GPU of this box (NCCL between them)
};

[[noreturn]] void usage(const char* prog) {
  std::fprintf(stderr,
               "usage: %s [--Re R] [--Nx N] [--Ny N] [--dt T]\n"
               "          [--steps N | --tfinal T] [--ppe sor-rb|jacobi|sor-lex|cheby] [--omega reference|mixed|W] [--max-iters K] [--exact 0|1]\n"
               "          [--path auto|simple|tiled] [--sweeps T] [--print-interval N] [--save-interval N]\n"
               "          [--no-vtk] [--outdir DIR] [--device D] [--gpus N] [--stop-after N]\n"
               "Omitted flags keep the reference's compiled-in constants.\n",
               prog);
  std::exit(2);
}

Options parse(int argc, char** argv) {
  Options o;
  for (int a = 1; a < argc; ++a) {
    const std::string f = argv[a];
    auto val = [&]() -> const char* {
      if (a + 1 >= argc) usage(argv[0]);
      return argv[++a];
    };
    if (f == "--Re") o.re = std::atof(val());
    else if (f == "--Nx") o.nx = std::atoi(val());
    else if (f == "--Ny") o.ny = std::atoi(val());
    else if (f == "--dt") o.dt = std::atof(val());
    else if (f == "--steps") o.steps = std::atoi(val());
    else if (f == "--tfinal") o.tfinal = std::atof(val());
    else if (f == "--max-iters") o.max_iters = std::atoi(val());
    else if (f == "--exact") o.exact = std::atoi(val());
    else if (f == "--sweeps") o.sweeps = std::atoi(val());
    else if (f == "--print-interval") o.print_interval = std::atoi(val());
    else if (f == "--save-interval") o.save_interval = std::atoi(val());
    else if (f == "--device") o.device = std::atoi(val());
    else if (f == "--gpus") o.gpus = std::atoi(val());
    else if (f == "--stop-after") o.stop_after = std::atoi(val());
    else if (f == "--outdir") o.outdir = val();
    else if (f == "--no-vtk") o.vtk = false;
    else if (f == "--omega") o.omega = val();
    else if (f == "--ppe") {
      const std::string v = val();
      if (v == "sor-rb") o.ppe = PM_PPE_SOR_RB;
      else if (v == "jacobi") o.ppe = PM_PPE_JACOBI;
      else if (v == "sor-lex") o.ppe = PM_PPE_SOR_LEX;
      else if (v == "cheby" || v == "sor-cheby") o.ppe = PM_PPE_SOR_CHEBY;
      else usage(argv[0]);
    } else if (f == "--path") {
      const std::string v = val();
      if (v == "auto") o.path = PM_PATH_AUTO;
      else if (v == "simple") o.path = PM_PATH_SIMPLE;
      else if (v == "tiled") o.path = PM_PATH_TILED;
      else usage(argv[0]);
    } else usage(argv[0]);
  }
  return o;
}

void check(int status, pm_solver* s, const char* what) {
  if (status == PM_OK) return;
  const char* msg = pm_last_error(s);
  throw std::runtime_error(msg && *msg ? std::string(msg) : std::string(what) + ": " + pm_status_string(status));
}

// ---- host-side view of what a frame prints (only touched at save intervals): five dense ny x nx arrays formed on the
// device by pm_export_begin / pm_export_wait (cell-centre velocities, magnitude, pressure, vorticity) and the fluid mask ----
struct Snapshot {
  int nx = 0, ny = 0;
  std::vector<double> uc, vc, mag, p, vort;
  std::vector<uint8_t> fluid;
  size_t at(int j, int i) const { return size_t(j - 1) * nx + (i - 1); }
  bool F(int j, int i) const { return fluid[size_t(j) * (nx + 2) + i] != 0; }
};

// Legacy-VTK STRUCTURED_POINTS writer; text identical to the reference writers
// (cavity-01.cpp:95-231; channel-01.cpp:88-190; backwards_step-01.cpp:89-220): every double goes through the
// sticky fixed/precision(6) stream state, the step writer prints the literal "0.0" where the reference does.
void write_vtk(const std::string& path, const pm_config& c, const Snapshot& sn, double t) {
  std::FILE* f = std::fopen(path.c_str(), "w");
  if (!f) throw std::runtime_error("Cannot open file: " + path);
  const int nx = c.nx, ny = c.ny, cs = c.case_id;
  const bool step = cs == PM_CASE_STEP;
  std::fprintf(f, "# vtk DataFile Version 3.0\n%s%.6f\nASCII\nDATASET STRUCTURED_POINTS\n", kText[cs].vtk_title, t);
  std::fprintf(f, "DIMENSIONS %d %d 1\n", nx, ny);
  std::fprintf(f, "ORIGIN %.6f %.6f 0.0\n", c.dx * 0.5, c.dy * 0.5);
  std::fprintf(f, "SPACING %.6f %.6f 1.0\n", c.dx, c.dy);
  std::fprintf(f, "POINT_DATA %d\n", nx * ny);
  std::fprintf(f, "SCALARS TimeValue double 1\nLOOKUP_TABLE default\n");
  for (int n = 0; n < nx * ny; ++n) std::fprintf(f, "%.6f\n", t);
  if (step) {
    std::fprintf(f, "SCALARS FluidMask double 1\nLOOKUP_TABLE default\n");
    for (int j = 1; j <= ny; ++j)
      for (int i = 1; i <= nx; ++i) std::fprintf(f, "%.6f\n", sn.F(j, i) ? 1.0 : 0.0);
  }
  std::fprintf(f, "VECTORS velocity double\n");
  for (int j = 1; j <= ny; ++j)
    for (int i = 1; i <= nx; ++i) {
      if (step && !sn.F(j, i)) std::fprintf(f, "0.0 0.0 0.0\n");
      else std::fprintf(f, "%.6f %.6f 0.0\n", sn.uc[sn.at(j, i)], sn.vc[sn.at(j, i)]);
    }
  // solid cells hold 0 in every exported array (interpolateToCellCenters leaves them 0, the writers print 0.0 for them)
  std::fprintf(f, "SCALARS u_velocity double 1\nLOOKUP_TABLE default\n");
  for (int j = 1; j <= ny; ++j)
    for (int i = 1; i <= nx; ++i) std::fprintf(f, "%.6f\n", sn.uc[sn.at(j, i)]);
  std::fprintf(f, "SCALARS v_velocity double 1\nLOOKUP_TABLE default\n");
  for (int j = 1; j <= ny; ++j)
    for (int i = 1; i <= nx; ++i) std::fprintf(f, "%.6f\n", sn.vc[sn.at(j, i)]);
  std::fprintf(f, "SCALARS velocity_magnitude double 1\nLOOKUP_TABLE default\n");
  for (int j = 1; j <= ny; ++j)
    for (int i = 1; i <= nx; ++i) {
      if (step && !sn.F(j, i)) { std::fprintf(f, "0.0\n"); continue; }
      std::fprintf(f, "%.6f\n", sn.mag[sn.at(j, i)]);
    }
  std::fprintf(f, "SCALARS pressure double 1\nLOOKUP_TABLE default\n");
  for (int j = 1; j <= ny; ++j)
    for (int i = 1; i <= nx; ++i) std::fprintf(f, "%.6f\n", sn.p[sn.at(j, i)]);
  // vorticity: formed on the device with the writers' own expression trees (cavity-01.cpp:187-223: one-sided at the edges,
  // (diff * dx_inv) * 0.5 inside; channel-01.cpp:171-182: (0.5 * diff) * idx; backwards_step-01.cpp:203-236: only where the
  // cell and its four neighbours are fluid and off the domain edge -- elsewhere that writer prints the literal 0.0)
  std::fprintf(f, "SCALARS vorticity double 1\nLOOKUP_TABLE default\n");
  for (int j = 1; j <= ny; ++j)
    for (int i = 1; i <= nx; ++i) {
      if (step) {
        bool ok = sn.F(j, i) && !(i == 1 || i == nx || j == 1 || j == ny);
        if (ok && (!sn.F(j, i - 1) || !sn.F(j, i + 1) || !sn.F(j - 1, i) || !sn.F(j + 1, i))) ok = false;
        if (!ok) { std::fprintf(f, "0.0\n"); continue; }
      }
      std::fprintf(f, "%.6f\n", sn.vort[sn.at(j, i)]);
    }
  const bool bad = std::ferror(f) != 0;
  if (std::fclose(f) != 0 || bad) throw std::runtime_error("Error writing to file: " + path);
}

void write_pvd(const std::string& path, const std::vector<std::string>& files, const std::vector<double>& times) {
  std::FILE* f = std::fopen(path.c_str(), "w");
  if (!f) throw std::runtime_error("Cannot open collection file: " + path);
  std::fprintf(f, "<?xml version=\"1.0\"?>\n<VTKFile type=\"Collection\" version=\"0.1\" byte_order=\"LittleEndian\">\n  <Collection>\n");
  for (size_t n = 0; n < files.size(); ++n)
    std::fprintf(f, "    <DataSet timestep=\"%.6f\" group=\"\" part=\"0\" file=\"%s\"/>\n", times[n], files[n].c_str());
  std::fprintf(f, "  </Collection>\n</VTKFile>\n");
  if (std::fclose(f) != 0) throw std::runtime_error("Error writing collection file: " + path);
}

// Reusable barrier for the per-GPU host threads (C++17 has no std::barrier).
class Barrier {
 public:
  explicit Barrier(int n) : n_(n) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m_);
    const int gen = gen_;
    if (++count_ == n_) { count_ = 0; ++gen_; cv_.notify_all(); }
    else cv_.wait(lk, [&] { return gen != gen_; });
  }
 private:
  std::mutex m_;
  std::condition_variable cv_;
  int n_, count_ = 0, gen_ = 0;
};

struct Run {
  pm_config cfg{};
  pm_solver* s = nullptr;
  Options opt;
  int rank = 0, nranks = 1;
  Barrier* bar = nullptr;      // null on a single GPU
  Snapshot* shared = nullptr;  // rank 0's snapshot; every rank downloads its own rows into it
  Snapshot snap;
  std::vector<std::string> files;
  std::vector<double> times;
  int total_steps = 0, print_interval = 100, save_interval = 100;

  void sync_ranks() { if (bar) bar->wait(); }

  // A frame is taken in two halves: export_begin enqueues the device-side export (kernel + copy to pinned host memory on its
  // own stream) behind the current state; export_flush waits for it, and rank 0 writes the file.  Inside the time loop the
  // flush comes after the NEXT step has been issued, so the copy overlaps that step; the order of the lines on stdout
  // stays the reference's (the "Exported" line precedes everything the next step prints).
  struct Pending { bool active = false; int step = 0; double t = 0.0; } pend;

  void export_begin(int step, double t) {
    if (!opt.vtk) return;
    export_flush();
    try {
      check(pm_export_begin(s), s, "export");
      pend.active = true; pend.step = step; pend.t = t;
    } catch (const std::exception& e) {  // an export failure is logged, the run goes on (cavity-01.cpp:479-481)
      std::fprintf(stderr, "%sError exporting VTK data: %s%s\n", RED, e.what(), RESET);
    }
  }

  void export_flush() {
    if (!opt.vtk || !pend.active) return;
    pend.active = false;
    Snapshot& sn = shared ? *shared : snap;
    try {
      char name[256];
      std::snprintf(name, sizeof name, "%s_%06d.vtk", kText[cfg.case_id].vtk_base, pend.step);
      if (rank == 0) {  // size the host arrays once; pm_export_wait then fills only the caller's rows
        sn.nx = cfg.nx; sn.ny = cfg.ny;
        const size_t n = size_t(cfg.nx) * cfg.ny;
        for (std::vector<double>* a : {&sn.uc, &sn.vc, &sn.mag, &sn.p, &sn.vort}) a->resize(n);
      }
      sync_ranks();
      check(pm_export_wait(s, sn.uc.data(), sn.vc.data(), sn.mag.data(), sn.p.data(), sn.vort.data(), sn.uc.size()), s, "export");
      sync_ranks();
      if (rank == 0) {
        write_vtk(opt.outdir + "/" + name, cfg, sn, pend.t);
        files.emplace_back(name);
        times.push_back(pend.t);
        if (pend.step % print_interval == 0 || pend.step == 0) std::printf("%sExported VTK file: %s%s\n", BLUE, name, RESET);
      }
    } catch (const std::exception& e) {
      std::fprintf(stderr, "%sError exporting VTK data: %s%s\n", RED, e.what(), RESET);
    }
    sync_ranks();
  }

  void banner() const {
    const pm_config& c = cfg;
    std::printf("%s%s\n", CYAN, kText[c.case_id].banner);
    if (c.case_id == PM_CASE_CAVITY) {
      std::printf("Domain: %.6fx%.6f\n", c.lx, c.ly);
      std::printf("Grid: %dx%d (spacing=%.6f)\n", c.nx, c.nx, c.dx);
    } else {
      std::printf("Domain: %.6fx%.6f\n", c.lx, c.ly);
      if (c.case_id == PM_CASE_STEP) std::printf("Step: height=%.6f, location=%.6f\n", c.ly - 1.0, 2.0);
      std::printf("Grid: %dx%d (dx=%.6f, dy=%.6f)\n", c.nx, c.ny, c.dx, c.dy);
    }
    std::printf("Time: dt=%.6f, steps=%d, final_time=%.6f\n", c.dt, total_steps, c.final_time);
    std::printf("Reynolds=%.6f, kinematic viscosity=%.6f, CFL=%.6f\n", c.re, c.nu, c.cfl);
    std::printf("Relaxation factor=%.6f\n", c.omega);
    std::printf("VTK export interval=%d steps\n", save_interval);
    std::printf("==========================================\n%s\n", RESET);
  }

  void log_line(int step, double t, const pm_ppe_result& r) {
    double md = 0.0, ke = 0.0;
    check(pm_diagnostics(s, &md, &ke), s, "diagnostics");  // collective over the slabs
    if (rank != 0) return;
    if (cfg.case_id == PM_CASE_CAVITY)
      std::printf("Step %6d/%d | t=%6.2f | max(div)=%10.2e | avg_KE=%10.6f | SOR_iters=%4d\n", step, total_steps, t, md, ke, r.iterations);
    else
      std::printf("Step %6d/%d | t=%8.3f | max(div)=%10.2e | avg_KE=%10.6f | PPE iters=%4d | res=%10.2e\n", step, total_steps, t, md, ke,
                  r.iterations, r.residual);
  }

  int main_loop() {
    const int cs = cfg.case_id;
    const bool root = rank == 0;
    Snapshot& sn0 = shared ? *shared : snap;
    if (root) sn0.fluid.assign(size_t(cfg.ny + 2) * (cfg.nx + 2), 0);
    if (opt.vtk) check(pm_export_prepare(s), s, "export buffers");  // before any rank's first collective (see pm.h)
    sync_ranks();
    check(pm_download_mask(s, sn0.fluid.data(), sn0.fluid.size()), s, "mask");  // each rank fills its rows
    sync_ranks();
    if (root && cs == PM_CASE_STEP) {  // setupGeometry's report, printed before the stream turns fixed (backwards_step-01.cpp:495-531)
      const std::vector<uint8_t>& m = sn0.fluid;
      int fluid = 0;
      for (int j = 1; j <= cfg.ny; ++j)
        for (int i = 1; i <= cfg.nx; ++i) fluid += m[size_t(j) * (cfg.nx + 2) + i];
      std::printf("%sSetting up backwards step geometry:\n  Step location: x = %g (i = %d)\n  Inlet height: %g (j = 1 to %d)\n  Total height: %g (j = 1 to %d)\n%s",
                  CYAN, 2.0, cfg.step_i_location, 1.0, cfg.inlet_j_max, cfg.ly, cfg.ny, RESET);
      std::printf("%sGeometry setup complete. Fluid cells: %d/%d%s\n", BLUE, fluid, cfg.nx * cfg.ny, RESET);
    }
    if (root && opt.vtk) {
      try {
        std::filesystem::create_directories(opt.outdir);
      } catch (const std::filesystem::filesystem_error& e) {
        throw std::runtime_error("Failed to setup output directory: Failed to create directory: " + opt.outdir + " Error: " + e.what());
      }
      std::printf("%sCreated output directory: %s%s\n", BLUE, opt.outdir.c_str(), RESET);
    }
    if (root) banner();
    // frame 0: the channel/step constructors and the cavity's run() apply the BCs and export before stepping
    if (root && cs == PM_CASE_CAVITY) std::printf("%s%s%s", GREEN, kText[cs].start_msg, RESET);
    check(pm_apply_bc(s, 0), s, "apply_bc");
    export_begin(0, 0.0);
    export_flush();
    if (root && cs != PM_CASE_CAVITY) std::printf("%s%s%s", GREEN, kText[cs].start_msg, RESET);

    for (int step = 1; step <= total_steps; ++step) {
      const double t = step * cfg.dt;
      pm_ppe_result r{};
      check(pm_step(s, 1, &r), s, "step");
      export_flush();  // the frame taken after the previous step, if any: its copy ran beside this step
      if (root && r.hit_cap) {
        if (cs == PM_CASE_CAVITY)
          std::fprintf(stderr, "Warning: SOR solver did not converge in %d iterations. Final residual: %g\n", cfg.max_iters, r.residual);
        else
          std::fprintf(stderr, "%sWarning: PPE SOR hit max iterations, max_res=%g%s\n", YELLOW, r.residual, RESET);
      }
      if (step % print_interval == 0 || step == total_steps) log_line(step, t, r);
      if (step % save_interval == 0 || step == total_steps) export_begin(step, t);
      if (opt.stop_after > 0 && step >= opt.stop_after) break;
    }
    export_flush();
    if (root && opt.vtk) {
      try {
        const std::string pvd = std::string(kText[cs].vtk_base) + "_animation.pvd";
        write_pvd(opt.outdir + "/" + pvd, files, times);
        std::printf("%sCreated ParaView collection file: %s%s\n", CYAN, pvd.c_str(), RESET);
      } catch (const std::exception& e) {
        std::fprintf(stderr, "%sError creating ParaView collection: %s%s\n", RED, e.what(), RESET);
      }
    }
    if (root) std::printf("%sSimulation completed successfully!\nVTK files saved in directory: %s\nOpen '%s/%s_animation.pvd' in ParaView for animation\n%s", GREEN,
                opt.outdir.c_str(), opt.outdir.c_str(), kText[cs].vtk_base, RESET);
    return 0;
  }
};
}  // namespace

int main(int argc, char** argv) {
  Run run;
  const bool colour_err = PM_DRIVER_CASE != PM_CASE_CAVITY;  // cavity-01.cpp:787-789 prints the error uncoloured
  try {
    run.opt = parse(argc, argv);
    const Options& o = run.opt;
    check(pm_config_init(&run.cfg, PM_DRIVER_CASE, o.nx, o.ny, o.re, o.dt), nullptr, "config");
    pm_config& c = run.cfg;
    c.ppe_method = o.ppe;
    c.exact_arith = o.exact;
    c.kernel_path = o.path;
    c.sweeps_per_pass = o.sweeps;
    c.device = o.device;
    if (o.ppe == PM_PPE_JACOBI) c.omega = 1.0;  // plain Jacobi diverges for omega > 1; the verification mode runs unrelaxed
    // beyond the reference: a factor derived from the operator's own boundary conditions (include/pm.h)
    if (o.omega == "mixed" || (o.omega.empty() && o.ppe == PM_PPE_SOR_CHEBY)) c.omega = pm_omega_mixed_bc(c.case_id, c.nx, c.ny, c.dx, c.dy);
    else if (o.omega != "" && o.omega != "reference") c.omega = std::atof(o.omega.c_str());
    if (!(c.omega > 0.0 && c.omega < 2.0)) throw std::runtime_error("--omega must lie in (0, 2)");
    if (o.max_iters >= 0) c.max_iters = o.max_iters;
    if (o.tfinal > 0) { c.final_time = o.tfinal; c.total_steps = static_cast<int>(c.final_time / c.dt); }
    run.total_steps = o.steps >= 0 ? o.steps : c.total_steps;
    run.print_interval = o.print_interval > 0 ? o.print_interval : c.print_interval;
    run.save_interval = o.save_interval > 0 ? o.save_interval : c.save_interval;
    if (o.gpus <= 1) {
      const int st = pm_create(&c, &run.s);
      if (st != PM_OK) throw std::runtime_error(pm_last_error(nullptr));
      const int rc = run.main_loop();
      pm_destroy(run.s);
      return rc;
    }
    // One host thread and one handle per GPU; the handles talk NCCL among themselves inside pm_step.
    if (c.ppe_method == PM_PPE_SOR_LEX) throw std::runtime_error("sor-lex does not shard over GPUs; use sor-rb or jacobi with --gpus");
    if (pm_nccl_unique_id(c.nccl_id) != PM_OK) throw std::runtime_error(pm_last_error(nullptr));
    Barrier bar(o.gpus);
    std::vector<Run> runs(o.gpus, run);
    std::vector<std::string> errors(o.gpus);
    std::vector<std::thread> th;
    for (int r = 0; r < o.gpus; ++r) {
      Run& me = runs[r];
      me.rank = r; me.nranks = o.gpus; me.bar = &bar; me.shared = &runs[0].snap;
      me.cfg.rank = r; me.cfg.nranks = o.gpus; me.cfg.device = r;
      th.emplace_back([&me, &errors, r] {
        try {
          if (pm_create(&me.cfg, &me.s) != PM_OK) throw std::runtime_error(pm_last_error(nullptr));
          me.main_loop();
        } catch (const std::exception& e) {
          errors[r] = e.what();
          std::fprintf(stderr, "Error (rank %d): %s\n", r, e.what());
          std::_Exit(1);  // the other ranks are blocked in a collective; there is nothing to unwind to
        }
        if (me.s) pm_destroy(me.s);
      });
    }
    for (auto& t : th) t.join();
    return 0;
  } catch (const std::exception& e) {
    if (colour_err) std::fprintf(stderr, "%sError: %s%s\n", RED, e.what(), RESET);
    else std::fprintf(stderr, "Error: %s\n", e.what());
    if (run.s) pm_destroy(run.s);
    return 1;
  }
}
