/*
 * pm.h — C-ABI of the B200-native projection-method time step.
 *
 * Drop-in boundary for the hot path of tjjones6/Computational-Fluid-Dynamics.
 * The reference has no plugin/FFI seam: each solver's run() calls private
 * members that mutate the solver's own fields
 *   cavity-01.cpp:387-390   (BC, predictor, PPE, correction)
 *   channel-01.cpp:368-375  (predictor, BC(u*,v*), source, PPE, correction, BC)
 *   backwards_step-01.cpp:412-419 (same as channel, with the is_fluid mask)
 * This header is the seam a maintainer would bind instead of those calls
 * (see INTEGRATION.md).  Plain pointers and sizes only; no C++/torch types.
 *
 * Array shapes at the boundary are the reference's (dense, row-major,
 * field[j][i], j = row):
 *   p, f            : (ny+2) x (nx+2)   cavity-01.cpp:433-435
 *   u, u*           : (ny+2) x (nx+1)   cavity-01.cpp:436-437
 *   v, v*           : (ny+1) x (nx+2)   cavity-01.cpp:439-440
 *   mask (is_fluid) : (ny+2) x (nx+2) bytes, backwards_step-01.cpp:480-483
 *
 * Every entry point returns a pm_status; 0 is success.  No exception crosses
 * the boundary.  A handle is driven by one host thread; independent handles
 * may coexist.  There is no CPU fallback: without a usable CUDA device
 * pm_create fails with PM_ERR_CUDA.
 */
#ifndef PM_H_
#define PM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PM_ABI_VERSION 1

typedef enum pm_status {
  PM_OK = 0,
  PM_ERR_INVALID_ARGUMENT = 1, /* std::invalid_argument in the reference (create_field, cavity-01.cpp:57-59) */
  PM_ERR_RUNTIME = 2,          /* std::runtime_error (dt <= 0 cavity-01.cpp:423-425; step outside domain backwards_step-01.cpp:459-461) */
  PM_ERR_CUDA = 3,             /* CUDA runtime/driver failure, or no device */
  PM_ERR_NCCL = 4,             /* NCCL failure or libnccl not loadable */
  PM_ERR_UNSUPPORTED = 5       /* combination not implemented (e.g. sor-lex on >1 rank) */
} pm_status;

/* Which reference solver's semantics the handle reproduces. */
typedef enum pm_case {
  PM_CASE_CAVITY = 0,  /* cavity-01.cpp          */
  PM_CASE_CHANNEL = 1, /* channel-01.cpp         */
  PM_CASE_STEP = 2     /* backwards_step-01.cpp  */
} pm_case;

/* Ordering of the pressure-Poisson sweep.  The reference only has the
 * lexicographic one (cavity-01.cpp:640-656, channel-01.cpp:657-668). */
typedef enum pm_ppe_method {
  PM_PPE_JACOBI = 0,  /* deterministic verification mode */
  PM_PPE_SOR_RB = 1,  /* production: red-black SOR        */
  PM_PPE_SOR_LEX = 2, /* reference ordering, wavefront-parallel, single rank only */
  /* Red-black SOR with Chebyshev acceleration (the README's "better Poisson solver" item, README.md:39; not in the
   * reference code): the relaxation factor changes with every colour half-sweep q = 0, 1, 2, ...
   *   w_0 = 1,  w_1 = 1 / (1 - rho^2 / 2),  w_q = 1 / (1 - rho^2 w_{q-1} / 4),   rho^2 = 1 - (2 / omega - 1)^2
   * (rho = the Jacobi spectral radius pm_config.omega was derived from, cavity-01.cpp:74-78), and tends to omega.  The
   * error norm then falls from the first sweep on instead of growing first, as it does with omega from the start.
   * Same sweeps, ghosts, residual and loop test as PM_PPE_SOR_RB; general kernel path; shards over slabs. */
  PM_PPE_SOR_CHEBY = 3
} pm_ppe_method;

/* w_q of PM_PPE_SOR_CHEBY from w_{q-1} (q >= 1; w_0 = 1).  Inline helpers (not exports) so that the library and the CPU
 * oracle evaluate one expression tree. */
static inline double pmi_cheby_next_omega(double rho2, int q, double w_prev) {
  return q == 1 ? 1.0 / (1.0 - 0.5 * rho2) : 1.0 / (1.0 - 0.25 * rho2 * w_prev);
}
static inline double pmi_cheby_rho2(double omega) {
  const double t = 2.0 / omega - 1.0;
  return 1.0 - t * t;
}

typedef enum pm_field {
  PM_FIELD_U = 0,      /* u_corrected */
  PM_FIELD_V = 1,      /* v_corrected */
  PM_FIELD_P = 2,      /* pressure    */
  PM_FIELD_USTAR = 3,  /* u_tentative */
  PM_FIELD_VSTAR = 4,  /* v_tentative */
  PM_FIELD_F = 5,      /* source_term */
  PM_FIELD_COUNT = 6
} pm_field;

typedef enum pm_kernel_path {
  PM_PATH_AUTO = 0,
  PM_PATH_SIMPLE = 1, /* one global-memory pass per colour + separate residual pass */
  PM_PATH_TILED = 2,  /* TMA-staged shared-memory tiles, fused residual, temporal blocking */
  PM_PATH_PERSISTENT = 3 /* small grids: one persistent CTA runs the whole solve out of shared memory */
} pm_kernel_path;

typedef struct pm_config {
  uint32_t struct_size; /* sizeof(pm_config); checked by pm_create */
  int32_t case_id;      /* pm_case */
  int32_t nx, ny;       /* GLOBAL interior cell counts */
  double dx, dy;        /* cavity: dx == dy == grid_spacing */
  double nu;            /* kinematic viscosity */
  double dt;            /* time step */
  double u_ref;         /* lid velocity (cavity) / inlet velocity (channel, step) */
  double rho;           /* density */
  double omega;         /* relaxation factor */
  double tol_factor;    /* cavity 1e-9 (cavity-01.cpp:317); channel/step 1e-7 */
  double abs_tol;       /* channel/step 1e-10 (channel-01.cpp:297); unused by cavity */
  int32_t max_iters;    /* 10000 in the reference */
  int32_t ppe_method;   /* pm_ppe_method */
  int32_t exact_arith;  /* 1: no FMA contraction, IEEE divides, serial-order mean: bit-identical to the oracle */
  int32_t sweeps_per_pass; /* temporal-blocking depth of the tiled path; 0 = auto */
  int32_t kernel_path;  /* pm_kernel_path */
  int32_t step_i_location; /* step: last solid column  (backwards_step-01.cpp:386) */
  int32_t inlet_j_max;     /* step: last inlet row     (backwards_step-01.cpp:493) */
  int32_t device;       /* CUDA device ordinal; -1 = current device */
  int32_t rank, nranks; /* slab decomposition in j; single GPU: 0, 1 */
  int32_t poll_chunk;   /* PPE passes enqueued between host polls of the device-side stop flag; 0 = auto */
  uint8_t nccl_id[128]; /* ncclUniqueId shared by all ranks (pm_nccl_unique_id on rank 0); ignored when nranks == 1 */
  /* Informational: filled by pm_config_init for the drivers, not read by the kernels. */
  double lx, ly, re, cfl, final_time;
  int32_t total_steps, print_interval, save_interval;
  int32_t reserved_;
} pm_config;

typedef struct pm_ppe_result {
  int32_t iterations;  /* == reference's iteration_count */
  int32_t hit_cap;     /* iterations >= max_iters: the reference prints a warning (cavity-01.cpp:681-684) */
  double residual;     /* final max|r| (infinity norm) */
  double tolerance;    /* the tolerance the loop used */
  double max_source;   /* max|f| seen by the tolerance rule */
} pm_ppe_result;

typedef struct pm_solver pm_solver;

/* ---- configuration (host only; a1/a2 of SURVEY §8a) -------------------- */

/* Fill *cfg with the reference's compiled-in constants for `case_id`
 * (cavity-01.cpp:309-320,356-363; channel-01.cpp:287-300,337-344;
 * backwards_step-01.cpp:319-334,378-387), overriding Re, nx, ny, dt when the
 * argument is > 0 (the README's --Re --Nx --Ny --dt).  dt <= 0 keeps the
 * reference's CFL rule.  Derives nu, dx, dy, omega, dt, total_steps and the
 * step geometry with the reference's own expression trees. */
int pm_config_init(pm_config* cfg, int case_id, int nx, int ny, double re, double dt);

/* Rows of the global grid owned by `rank`: interior rows j0+1 .. j0+ny_local. */
int pm_slab_range(int ny, int nranks, int rank, int* j0, int* ny_local);

/* Relaxation factor of colour half-sweep q (0, 1, 2, ...) of PM_PPE_SOR_CHEBY for a problem whose fixed-omega solver would
 * use `omega` (host-only, no device needed). */
double pm_cheby_omega(double omega, int q);

/* A relaxation factor derived from the operator the solvers actually iterate on.  The reference takes omega from the
 * Jacobi spectral radius of the DIRICHLET problem (cavity-01.cpp:74-78, channel-01.cpp:76-81), but its pressure problem is
 * Neumann on all walls but one (cavity: the south ghost row is held at 0; channel / step: the outlet ghost column), whose
 * slowest Jacobi mode is constant along the all-Neumann direction and a quarter wave along the other:
 *   cavity:   rho = (1 + cos(pi / (2 ny + 1))) / 2
 *   channel:  rho = (cos(pi / (2 nx + 1)) / dx^2 + 1 / dy^2) / (1 / dx^2 + 1 / dy^2)        (step: the same)
 *   omega = 2 / (1 + sqrt(1 - rho^2)).
 * With it red-black SOR needs 3-4.4 x fewer iterations on the reference's own configurations (tests/test_oracle.py), and the
 * channel and step cases converge inside the reference's cap of 10000, which they never do with the reference's factor.
 * Not the default: the drivers take it with `--omega mixed` (and with `--ppe cheby`, whose schedule then tends to it). */
double pm_omega_mixed_bc(int case_id, int nx, int ny, double dx, double dy);

/* Host-only test hook (no device needed): how the streaming pressure pass would cut a slab of ny_local rows starting at global
 * row j0 of an nx x ny grid into strips and chunks for `slots` resident warps, over the tile rows [row_lo, row_hi) of the
 * tiled plan (row_hi < 0: all).  out[8] = {first strip, strips, first tile row, tile rows, rows per chunk, chunks, warps,
 * tiles left to the tiled kernel}.  Returns 1 if the pass streams, 0 if not, -1 on bad arguments. */
int pm_stream_plan(int nx, int ny, int ny_local, int j0, int row_lo, int row_hi, int slots, int* out);

/* ---- lifetime ---------------------------------------------------------- */
int pm_create(const pm_config* cfg, pm_solver** out);
int pm_destroy(pm_solver* s);
const char* pm_last_error(const pm_solver* s); /* s may be NULL: error of the last failed pm_create on this thread */
const char* pm_status_string(int status);
int pm_abi_version(void);

/* 128-byte ncclUniqueId for pm_config.nccl_id (call on one rank, broadcast to the others). */
int pm_nccl_unique_id(uint8_t out[128]);

/* ---- data movement (reference array shapes, GLOBAL arrays) ------------- */
/* `count` must equal the element count of the global field.  With nranks > 1
 * each rank copies only its own slab (plus halo rows) from/to the global array. */
int pm_upload(pm_solver* s, int field, const double* host, size_t count);
int pm_download(pm_solver* s, int field, double* host, size_t count);
/* Slab-local variants: `host` holds only this rank's rows, global rows j0 .. j0+ny_local+1 (clipped to
 * the field's row count), i.e. local rows 0..ny_local+1 including the two ghost/halo rows. */
int pm_slab_rows(pm_solver* s, int field, int* first_global_row, int* nrows, int* ncols);
int pm_upload_slab(pm_solver* s, int field, const double* host, size_t count);
int pm_download_slab(pm_solver* s, int field, double* host, size_t count);
int pm_upload_mask(pm_solver* s, const uint8_t* is_fluid, size_t count);   /* step only; default is the reference rectangle */
int pm_download_mask(pm_solver* s, uint8_t* is_fluid, size_t count);
/* Synthetic state generated on the device: value(field, j, i) = U(-1,1) from
 * splitmix64(seed, field, flat reference index).  Same generator in oracle/. */
int pm_fill_random(pm_solver* s, uint64_t seed);
/* Same, scaled: value = amplitude * U(-1,1).  A power-of-two amplitude keeps host and device bit-identical. */
int pm_fill_random_scaled(pm_solver* s, uint64_t seed, double amplitude);
int pm_fill_zero(pm_solver* s);

/* ---- phases (one per reference member function) ------------------------ */
/* which = 0: (u,v)  [applyBoundaryConditions]; 1: (u*,v*) [channel-01.cpp:369] */
int pm_apply_bc(pm_solver* s, int which);
int pm_predict(pm_solver* s);                      /* computeTentativeVelocities */
int pm_source(pm_solver* s);                       /* source term (+ mean removal for channel/step) */
int pm_ppe_solve(pm_solver* s, pm_ppe_result* r);  /* solverPressurePoisson (cavity: without the source loop, which pm_source does) */
int pm_correct(pm_solver* s);                      /* applyPressureCorrection */

/* nsteps whole projection steps in the case's own call order
 * (cavity-01.cpp:387-390 / channel-01.cpp:368-375).  *last may be NULL. */
int pm_step(pm_solver* s, int nsteps, pm_ppe_result* last);

/* ---- host-resident steps, streamed ---------------------------------------
 * One projection step per call on inputs and results that live in HOST memory (pinned for full PCIe speed):
 * this rank's slab rows of u and v in, of u, v and p out, in the layout of pm_upload_slab / pm_download_slab.
 *   pm_host_step_submit  enqueues the upload of (u, v) for one step and records where its results go;
 *   pm_host_step_run     runs the oldest submitted step -- pm_step(1) on the uploaded velocities; the pressure
 *                        state stays on the device from step to step, as in the reference's time loop -- and
 *                        enqueues the download of (u, v, p); returns when the pressure solve has finished;
 *   pm_host_step_drain   blocks until every enqueued download has landed in host memory.
 * Copies run on their own streams through three rotating (u, v) plane sets, so the upload of step n+1 and the
 * download of step n-1 overlap the kernels of step n.  At most two steps may be submitted and not yet run.
 * The host buffers of a step must stay valid and untouched until pm_host_step_drain returns.
 * While a step is submitted and not yet run, pm_step, pm_fill_zero / pm_fill_random* and uploads of u or v return
 * PM_ERR_INVALID_ARGUMENT; once every submitted step has run they wait for the pending downloads first. */
int pm_host_step_submit(pm_solver* s, const double* u_in, size_t u_count, const double* v_in, size_t v_count,
                        double* u_out, double* v_out, double* p_out, size_t p_count);
int pm_host_step_run(pm_solver* s, pm_ppe_result* r);
int pm_host_step_drain(pm_solver* s);

/* max|div u| over (fluid) cells and mean kinetic energy of the cell-centred
 * velocity (logStatistics, cavity-01.cpp:741-766). */
int pm_diagnostics(pm_solver* s, double* max_div, double* avg_ke);

/* What the VTK writers print, formed on the device (interpolateToCellCenters and the vorticity loops: cavity-01.cpp:717-733,
 * 187-223; channel-01.cpp:708-731,171-182; backwards_step-01.cpp:981-1009,203-236), bit for bit what the host loops give:
 * five dense row-major arrays of ny x nx doubles (j = 1..ny, i = 1..nx) -- u_center, v_center, velocity magnitude, pressure,
 * vorticity; solid cells (step) and vorticity cells the step writer skips hold 0.  pm_export_begin enqueues the kernel behind
 * the current state and starts the copy to pinned host memory on its own stream, so time steps issued afterwards overlap it;
 * pm_export_wait waits for the copy and fills this rank's rows of the caller's arrays (count = nx * ny each). */
int pm_export_prepare(pm_solver* s); /* allocate the staging buffers now (pm_export_begin does it lazily).  Needed only when one
                                       * process drives several handles from threads: allocation synchronises devices and must
                                       * not race another handle's NCCL call, so prepare all handles first, then a barrier */
int pm_export_begin(pm_solver* s);
int pm_export_wait(pm_solver* s, double* u_center, double* v_center, double* magnitude, double* pressure, double* vorticity, size_t count);

/* Block until all work queued on the handle has finished. */
int pm_sync(pm_solver* s);

/* ---- measurement helpers ------------------------------------------------ */
typedef struct pm_timing {
  double ppe_ms;        /* device time spent in pm_ppe_solve since creation (cumulative; take differences) */
  double other_ms;      /* device time of the non-pressure phases of the most recent projection step (pm_get_timing waits for it) */
  int64_t kernel_launches; /* kernels launched by this handle since creation */
  int64_t ppe_passes;   /* PPE kernel passes since creation */
} pm_timing;
int pm_get_timing(pm_solver* s, pm_timing* t);
/* CUDA-event stopwatch on the handle's own stream (torch.cuda.Event would only see torch's stream). */
int pm_timer_start(pm_solver* s);
int pm_timer_stop(pm_solver* s, double* elapsed_ms); /* synchronizes the stream */

#ifdef __cplusplus
}
#endif
#endif /* PM_H_ */
