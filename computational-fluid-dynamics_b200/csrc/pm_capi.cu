// pm_capi.cu — host side of libpm.so: the handle, the HBM layout, the per-phase launch sequences
// and the C-ABI of include/pm.h.  Host code is C++17; every device entry is a hand-written
// sm_100a kernel from pm_kernels_*.cuh.  There is no CPU fallback anywhere in this file.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/pm.h"
#include "pm_common.cuh"
#include "pm_kernels_simple.cuh"
#include "pm_kernels_tiled.cuh"
#include "pm_kernels_stream.cuh"
#include "pm_kernels_residual.cuh"
#include "pm_kernels_lex.cuh"
#include "pm_nccl.hpp"

// ---------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------
enum { PL_U = 0, PL_V, PL_US, PL_VS, PL_F, PL_P0, PL_P1, PL_COUNT };

// Streamed host-resident steps (pm_host_step_*): three rotating (u, v) plane sets, a pressure out-plane,
// one stream per copy direction.
struct HostPipe {
  bool ready = false;
  double* extra = nullptr;  // 5 planes: u[1], u[2], v[1], v[2], pout
  double* u[3] = {}, *v[3] = {};
  double* pout = nullptr;
  cudaStream_t h2d = nullptr, d2h = nullptr;
  cudaEvent_t ev_up[3] = {}, ev_down[3] = {}, ev_step = nullptr, ev_pdown = nullptr;
  bool down_pending[3] = {false, false, false}, pdown_pending = false;
  long long submitted = 0, run = 0;
  struct Job { double* u_out; double* v_out; double* p_out; } job[3] = {};
};

struct pm_solver {
  pm_config cfg{};
  KP kp{};
  int device = 0;
  cudaStream_t stream = nullptr, comm_stream = nullptr, edge_stream = nullptr;  // edge_stream: high priority, the slab's edge tile rows
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_halo = nullptr, ev_edge = nullptr, ev_pass = nullptr, ev_s0 = nullptr, ev_s1 = nullptr;
  bool step_timed = false;
  double last_step_ppe_ms = 0.0;
  size_t plane = 0;        // doubles per plane
  int rows_alloc = 0;
  double* base = nullptr;  // PL_COUNT planes
  double* pl[PL_COUNT] = {};
  int p_cur = PL_P0;       // plane holding the current pressure (natural layout)
  // Tiled path: the solve ping-pongs between two buffers of its own in the split-row layout (pm_common.cuh).  The
  // pressure is converted lazily: p_split / p_nat say which representation currently holds it.
  double* tp[2] = {nullptr, nullptr};
  int tp_cur = 0;
  bool p_split = false, p_nat = true;
  uint8_t* mask = nullptr; // same geometry, bytes
  PpeState* d_state = nullptr;
  unsigned long long* d_res = nullptr;  // max_iters + 2 entries
  double* d_partial = nullptr;
  int n_partial = 0;    // blocks of cell_grid (k_source, k_diag)
  int cap_partial = 0;  // allocated slots: max over the grids that write partial sums
  PpeState* h_state = nullptr;  // pinned
  unsigned long long* h_res = nullptr;  // pinned, max_iters + 2
  bool f_max_valid = false;
  int last_iters = 0;
  bool use_tiled = false;
  bool use_small = false;  // persistent single-CTA solve (small grids)
  bool no_cluster = false; // PM_NO_CLUSTER=1 in the environment (or no room for a cluster on this device): keep the persistent solve on one SM
  bool cluster_checked = false;
  int sweeps = 1;
  int cheby_q = 0;       // PM_PPE_SOR_CHEBY: colour half-sweeps launched in this solve, and the factor of the last one
  double cheby_w = 1.0;
  TiledPlan tiled{};
  // streaming pass (pm_kernels_stream.cuh) over the interior tiles of the plan; f in the split-row layout for it
  // PM_TRACE_PASS=1: device timestamps of the pieces of every pass of the slab path (tiled_pass), summed per solve and
  // printed by pm_destroy -- a text timeline of a pass (profiles/r02_pass_timeline_*.txt)
  struct PassTrace {
    bool on = false;
    static constexpr int NEV = 7, CAP = 64;  // events per pass: start, edge rows, halo rows, frame, interior, fold, allreduce
    cudaEvent_t ev[CAP][NEV] = {};
    int used = 0;
    double sum_us[NEV] = {};
    long long passes = 0;
  } trace;
  // device-side export (pm_export_begin / pm_export_wait): dense staging on the device and in pinned host memory
  struct Export {
    double* dev = nullptr;
    double* host = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ready = nullptr, done = nullptr;
    bool pending = false;
  } ex;
  StreamPlan splan{};
  double* fsplit = nullptr;
  bool fsplit_valid = false;  // the fused predictor + source pass has just written fsplit along with f
  PmNccl nccl{};
  pm_timing timing{};
  HostPipe hp{};
  std::string err;
};

static thread_local std::string g_create_error;

static int fail(pm_solver* s, int status, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (s) s->err = buf; else g_create_error = buf;
  return status;
}
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(s, PM_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define CKL(s_)                                                                                           \
  do {                                                                                                    \
    cudaError_t e_ = cudaGetLastError();                                                                  \
    if (e_ != cudaSuccess) return fail(s_, PM_ERR_CUDA, "kernel launch: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
    (s_)->timing.kernel_launches++;                                                                       \
  } while (0)
#define PMTRY(expr)            \
  do {                         \
    int st_ = (expr);          \
    if (st_ != PM_OK) return st_; \
  } while (0)

// `bytes` (a multiple of 8) from device memory to device-addressable memory, in stream order, without a copy engine.
static int publish_words(pm_solver* s, void* dst, const void* src, size_t bytes) {
  k_publish_words<<<1, 32, 0, s->stream>>>(static_cast<const unsigned long long*>(src), static_cast<unsigned long long*>(dst), int(bytes / 8));
  CKL(s);
  return PM_OK;
}

static inline dim3 cell_block() { return dim3(PM_BX, PM_BY); }
static inline dim3 cell_grid(const KP& k) { return dim3((k.nx + PM_BX - 1) / PM_BX, (k.nyl + PM_BY - 1) / PM_BY); }
static inline dim3 rows_block() { return dim3(PM_RX, PM_RY); }  // k_*_rows: two cells per thread
static inline dim3 rows_grid(const KP& k) { return dim3(((k.nx + 1) / 2 + PM_RX - 1) / PM_RX, (k.nyl + PM_RY - 1) / PM_RY); }
static inline dim3 half_grid(const KP& k) { return dim3(((k.nx + 1) / 2 + PM_BX - 1) / PM_BX, (k.nyl + PM_BY - 1) / PM_BY); }

static void field_dims(const pm_solver* s, int field, int* rows, int* cols) {
  const int nx = s->cfg.nx, ny = s->cfg.ny;
  switch (field) {
    case PM_FIELD_U: case PM_FIELD_USTAR: *rows = ny + 2; *cols = nx + 1; break;
    case PM_FIELD_V: case PM_FIELD_VSTAR: *rows = ny + 1; *cols = nx + 2; break;
    default: *rows = ny + 2; *cols = nx + 2; break;
  }
}
static double* field_plane(pm_solver* s, int field) {
  switch (field) {
    case PM_FIELD_U: return s->pl[PL_U];
    case PM_FIELD_V: return s->pl[PL_V];
    case PM_FIELD_P: return s->pl[s->p_cur];
    case PM_FIELD_USTAR: return s->pl[PL_US];
    case PM_FIELD_VSTAR: return s->pl[PL_VS];
    case PM_FIELD_F: return s->pl[PL_F];
  }
  return nullptr;
}

// ---------------------------------------------------------------------------
// misc exported helpers
// ---------------------------------------------------------------------------
extern "C" int pm_abi_version(void) { return PM_ABI_VERSION; }

extern "C" const char* pm_status_string(int status) {
  switch (status) {
    case PM_OK: return "ok";
    case PM_ERR_INVALID_ARGUMENT: return "invalid argument";
    case PM_ERR_RUNTIME: return "runtime error";
    case PM_ERR_CUDA: return "CUDA error";
    case PM_ERR_NCCL: return "NCCL error";
    case PM_ERR_UNSUPPORTED: return "unsupported";
  }
  return "unknown status";
}
extern "C" const char* pm_last_error(const pm_solver* s) { return s ? s->err.c_str() : g_create_error.c_str(); }

extern "C" int pm_nccl_unique_id(uint8_t out[128]) {
  std::string e;
  if (!pm_nccl_get_unique_id(out, &e)) return fail(nullptr, PM_ERR_NCCL, "%s", e.c_str());
  return PM_OK;
}

// ---------------------------------------------------------------------------
// create / destroy
// ---------------------------------------------------------------------------
// Everything in KP that depends on the relaxation factor (PM_PPE_SOR_CHEBY changes it with every colour half-sweep).
static void kp_set_omega(KP& k, double omega) {
  k.omega = omega; k.om1 = 1.0 - omega;
  for (int n = 1; n <= 4; ++n) k.wnc[n] = omega / n;
  k.wnc[0] = 0.0;
  k.cw = k.case_id == PM_CASE_CAVITY ? omega * k.hh / 4.0 : omega / k.denom;
  k.cw3 = omega * k.hh / 3.0;
  k.cw2 = omega * k.hh / 2.0;
}

static void fill_kp(pm_solver* s, int j0, int nyl) {
  const pm_config& c = s->cfg;
  KP& k = s->kp;
  std::memset(&k, 0, sizeof k);
  k.nx = c.nx; k.ny = c.ny; k.nyl = nyl; k.j0 = j0;
  // pitch: PM_OFFC left pad + nx+2 columns + right pad for tile halos, rounded to 16 doubles (128 B)
  k.pitch = std::max(144, ((PM_OFFC + c.nx + 2 + PM_PADR + 15) / 16) * 16);
  k.padr = PM_PADR;
  k.case_id = c.case_id;
  k.has_mask = c.case_id == PM_CASE_STEP;
  k.first_rank = c.rank == 0;
  k.last_rank = c.rank == c.nranks - 1;
  k.inlet_j_max = c.inlet_j_max;
  // the reference's per-call locals, same expression trees (cavity-01.cpp:549-550,613-615; channel-01.cpp:547-550,638-640)
  k.idx = 1.0 / c.dx; k.idy = 1.0 / c.dy;
  k.idx2 = 1.0 / (c.dx * c.dx); k.idy2 = 1.0 / (c.dy * c.dy);
  k.hh = c.dx * c.dx;
  k.nu = c.nu; k.dt = c.dt; k.uref = c.u_ref; k.two_uref = 2.0 * c.u_ref;
  k.denom = 2.0 * (k.idx2 + k.idy2);
  k.rdenom = 1.0 / k.denom;
  kp_set_omega(k, c.omega);
  if (c.case_id == PM_CASE_CAVITY) {
    const double dti = 1.0 / c.dt;
    k.src_coef = dti * c.rho;               // time_step_inv * density, cavity-01.cpp:624
    const double dt_over_h = c.dt / c.dx;   // :696
    k.cu = dt_over_h * c.rho; k.cv = k.cu;  // :701,:708
  } else {
    k.src_coef = c.rho / c.dt;              // channel-01.cpp:610
    k.cu = c.dt / (c.rho * c.dx);           // :697
    k.cv = c.dt / (c.rho * c.dy);           // :701
  }
  k.tol_factor = c.tol_factor; k.abs_tol = c.abs_tol;
  k.max_iters = c.max_iters;
  k.fluid_count_global = c.nx * c.ny;
}

static int upload_mask_rows(pm_solver* s, const uint8_t* global_mask) {
  const KP& k = s->kp;
  const int cols = k.nx + 2;
  int cnt = 0;
  for (int j = 1; j <= k.ny; ++j)
    for (int i = 1; i <= k.nx; ++i) cnt += global_mask[size_t(j) * cols + i] ? 1 : 0;
  s->kp.fluid_count_global = cnt;
  CK(cudaMemsetAsync(s->mask, 0, s->plane, s->stream));
  // local rows as deep as the pad rows reach (the tiled solve looks PM_PADR rows into the neighbour slabs) <- global rows j0 + jl
  const int jl_lo = std::max(-k.padr, -k.j0), jl_hi = std::min(k.nyl + 1 + k.padr, k.ny + 1 - k.j0);
  CK(cudaMemcpy2DAsync(s->mask + size_t(k.padr + jl_lo) * size_t(k.pitch) + PM_OFFC, size_t(k.pitch), global_mask + size_t(k.j0 + jl_lo) * cols, size_t(cols),
                       size_t(cols), size_t(jl_hi - jl_lo + 1), cudaMemcpyHostToDevice, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  if (s->use_tiled && k.has_mask) {
    std::string e;
    if (!tiled_classify(&s->tiled, k, global_mask, s->stream, &e)) return fail(s, PM_ERR_CUDA, "%s", e.c_str());
    s->tiled.mask = s->mask;
  }
  return PM_OK;
}

static int destroy_impl(pm_solver* s) {
  if (!s) return PM_OK;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  pm_nccl_destroy(&s->nccl);
  tiled_destroy(&s->tiled);
  if (s->hp.ready) {
    cudaStreamSynchronize(s->hp.h2d);
    cudaStreamSynchronize(s->hp.d2h);
    s->pl[PL_U] = s->hp.u[0];  // planes of s->base
    s->pl[PL_V] = s->hp.v[0];
    for (int q = 0; q < 3; ++q) { cudaEventDestroy(s->hp.ev_up[q]); cudaEventDestroy(s->hp.ev_down[q]); }
    cudaEventDestroy(s->hp.ev_step);
    cudaEventDestroy(s->hp.ev_pdown);
    cudaStreamDestroy(s->hp.h2d);
    cudaStreamDestroy(s->hp.d2h);
    cudaFree(s->hp.extra);
  }
  if (s->trace.on) {
    const auto& t = s->trace;
    if (t.passes > 0) {
      const double n = double(t.passes);
      fprintf(stderr,
              "[pm] rank %d of %d, %lld streamed passes, microseconds from the start of a pass (averages): edge tile rows done %.1f | "
              "halo rows exchanged %.1f | wall tiles of the middle rows done %.1f | streaming kernel done %.1f | residual slots folded %.1f | "
              "allreduce(max) of the residual words done = end of pass %.1f\n",
              s->cfg.rank, s->cfg.nranks, t.passes, t.sum_us[1] / n, t.sum_us[2] / n, t.sum_us[3] / n, t.sum_us[4] / n, t.sum_us[5] / n, t.sum_us[6] / n);
    }
    for (auto& row : s->trace.ev)
      for (cudaEvent_t ev : row)
        if (ev) cudaEventDestroy(ev);
  }
  if (s->ex.dev) cudaFree(s->ex.dev);
  if (s->ex.host) cudaFreeHost(s->ex.host);
  if (s->ex.ready) cudaEventDestroy(s->ex.ready);
  if (s->ex.done) cudaEventDestroy(s->ex.done);
  if (s->ex.stream) cudaStreamDestroy(s->ex.stream);
  if (s->base) cudaFree(s->base);
  if (s->tp[0]) cudaFree(s->tp[0]);
  if (s->fsplit) cudaFree(s->fsplit);
  stream_destroy(&s->splan);
  if (s->mask) cudaFree(s->mask);
  if (s->d_state) cudaFree(s->d_state);
  if (s->d_res) cudaFree(s->d_res);
  if (s->d_partial) cudaFree(s->d_partial);
  if (s->h_state) cudaFreeHost(s->h_state);
  if (s->h_res) cudaFreeHost(s->h_res);
  for (cudaEvent_t e : {s->ev_a, s->ev_b, s->ev_t0, s->ev_t1, s->ev_halo, s->ev_edge, s->ev_pass, s->ev_s0, s->ev_s1})
    if (e) cudaEventDestroy(e);
  if (s->stream) cudaStreamDestroy(s->stream);
  if (s->comm_stream) cudaStreamDestroy(s->comm_stream);
  if (s->edge_stream) cudaStreamDestroy(s->edge_stream);
  delete s;
  return PM_OK;
}

// Slabs: the tile rows whose output the neighbour slabs need within one pass (row 0, and as many rows from the top as cover
// the last H output rows) are launched ahead of the rest.  A single rank has none.
static void slab_edge_rows(const pm_solver* s, int* bot, int* top) {
  *bot = *top = 0;
  if (s->cfg.nranks == 1) return;
  const TiledPlan& pl = s->tiled;
  int t = 1;
  while (t < pl.tiles_y && s->kp.nyl - (pl.tiles_y - t) * pl.ty < pl.halo) ++t;
  *bot = 1;
  *top = t;
}

static int create_impl(pm_solver* s, const pm_config* cfg) {
  s->cfg = *cfg;
  const pm_config& c = s->cfg;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(s, PM_ERR_CUDA, "no CUDA device: libpm has no CPU path");
  if (c.device >= 0) {
    if (c.device >= ndev) return fail(s, PM_ERR_INVALID_ARGUMENT, "device %d out of range (%d devices)", c.device, ndev);
    s->device = c.device;
  } else {
    CK(cudaGetDevice(&s->device));
  }
  CK(cudaSetDevice(s->device));

  int j0 = 0, nyl = c.ny;
  if (pm_slab_range(c.ny, c.nranks, c.rank, &j0, &nyl) != PM_OK)
    return fail(s, PM_ERR_INVALID_ARGUMENT, "bad slab decomposition: ny=%d nranks=%d rank=%d", c.ny, c.nranks, c.rank);
  fill_kp(s, j0, nyl);
  const KP& k = s->kp;

  CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s->comm_stream, cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0;  // numerically lowest = greatest priority
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&s->edge_stream, cudaStreamNonBlocking, hi));
  }
  for (cudaEvent_t* e : {&s->ev_a, &s->ev_b, &s->ev_t0, &s->ev_t1, &s->ev_s0, &s->ev_s1}) CK(cudaEventCreate(e));
  for (cudaEvent_t* e : {&s->ev_halo, &s->ev_edge, &s->ev_pass}) CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));

  s->rows_alloc = nyl + 2 + 2 * k.padr;
  s->plane = size_t(k.pitch) * size_t(s->rows_alloc);
  CK(cudaMalloc(&s->base, s->plane * PL_COUNT * sizeof(double)));
  CK(cudaMemsetAsync(s->base, 0, s->plane * PL_COUNT * sizeof(double), s->stream));
  for (int q = 0; q < PL_COUNT; ++q) s->pl[q] = s->base + s->plane * q;
  CK(cudaMalloc(&s->mask, s->plane));
  CK(cudaMemsetAsync(s->mask, 0, s->plane, s->stream));
  CK(cudaMalloc(&s->d_state, sizeof(PpeState)));
  CK(cudaMemsetAsync(s->d_state, 0, sizeof(PpeState), s->stream));
  CK(cudaMalloc(&s->d_res, size_t(c.max_iters + 2) * sizeof(unsigned long long)));
  CK(cudaMemsetAsync(s->d_res, 0, size_t(c.max_iters + 2) * sizeof(unsigned long long), s->stream));
  {  // one partial sum per block of whichever grid writes them: the cell kernels (k_source, k_diag) or the row kernels (k_source_rows)
    const dim3 g = cell_grid(k), gr = rows_grid(k);
    s->n_partial = int(g.x * g.y);
    s->cap_partial = std::max(s->n_partial, int(gr.x * gr.y));
    CK(cudaMalloc(&s->d_partial, size_t(s->cap_partial) * sizeof(double)));
  }
  CK(cudaMallocHost(&s->h_state, sizeof(PpeState)));
  CK(cudaMallocHost(&s->h_res, size_t(c.max_iters + 2) * sizeof(unsigned long long)));

  // kernel path.  AUTO: small grids -> the persistent cluster / single-CTA solve; every other unmasked
  // Jacobi / red-black problem -> the TMA-tiled kernel; the rest (obstacle mask beyond the small-grid limit)
  // -> the general kernels.
  const bool tiled_ok = tiled_supported(c, k);
  if (c.kernel_path == PM_PATH_TILED && !tiled_ok)
    return fail(s, PM_ERR_UNSUPPORTED, "tiled path: jacobi / sor-rb (with the obstacle mask: sor-rb only)");
  const size_t pbytes = size_t(c.ny + 2) * size_t(c.nx + 2) * sizeof(double);
  const bool small_ok = c.nranks == 1 && c.ppe_method != PM_PPE_SOR_LEX && c.ppe_method != PM_PPE_SOR_CHEBY && pbytes <= size_t(200) * 1024 &&
                        size_t(c.nx) * size_t(c.ny) <= size_t(20) * 1024;
  if (c.kernel_path == PM_PATH_PERSISTENT && !small_ok)
    return fail(s, PM_ERR_UNSUPPORTED, "persistent path needs a single rank and a pressure field that fits shared memory");
  s->no_cluster = std::getenv("PM_NO_CLUSTER") != nullptr;
  s->use_small = small_ok && (c.kernel_path == PM_PATH_PERSISTENT || c.kernel_path == PM_PATH_AUTO);
  s->use_tiled = tiled_ok && !s->use_small && c.ppe_method != PM_PPE_SOR_LEX &&
                 (c.kernel_path == PM_PATH_TILED || (c.kernel_path == PM_PATH_AUTO && (c.nranks == 1 || c.ny / c.nranks >= 2 * PM_PADR)));
  // (the slab test looks at the smallest slab of the decomposition, not at this rank's: every rank must take the same path,
  // or their NCCL call sequences differ -- 31 rows on 2 ranks are 16 + 15)
  if (s->use_tiled) {
    std::string e;
    CK(cudaMalloc(&s->tp[0], s->plane * 2 * sizeof(double)));
    CK(cudaMemsetAsync(s->tp[0], 0, s->plane * 2 * sizeof(double), s->stream));
    s->tp[1] = s->tp[0] + s->plane;
    if (!tiled_create(&s->tiled, c, k, s->tp[0], s->tp[1], s->rows_alloc, &e))
      return fail(s, PM_ERR_CUDA, "tiled path setup: %s", e.c_str());
    s->sweeps = s->tiled.run;
    s->kp.psh = s->tiled.psh;
    if (std::getenv("PM_TRACE_PASS") != nullptr && c.nranks > 1) {
      for (auto& row : s->trace.ev)
        for (cudaEvent_t& ev : row) CK(cudaEventCreate(&ev));
      s->trace.on = true;
    }
    if (stream_supported(c, s->kp, s->tiled)) {
      int bot = 0, top = 0;
      slab_edge_rows(s, &bot, &top);
      if (bot + top < s->tiled.tiles_y && !stream_create(&s->splan, s->tiled, c, s->kp, bot, s->tiled.tiles_y - top, s->stream, &e))
        return fail(s, PM_ERR_CUDA, "streaming pass setup: %s", e.c_str());
      if (s->splan.on) CK(cudaMalloc(&s->fsplit, s->plane * sizeof(double)));
    }
  }

  // is_fluid: interior true; step: the reference rectangle (backwards_step-01.cpp:500-520).  After the kernel path is
  // known: the tiled solve also wants the class of every tile (fluid only / solid only / both).
  {
    std::vector<uint8_t> m(size_t(c.ny + 2) * (c.nx + 2), 0);
    for (int j = 1; j <= c.ny; ++j)
      for (int i = 1; i <= c.nx; ++i)
        m[size_t(j) * (c.nx + 2) + i] =
            (c.case_id != PM_CASE_STEP) || (i > c.step_i_location) || (j <= c.inlet_j_max) ? 1 : 0;
    PMTRY(upload_mask_rows(s, m.data()));
  }

  if (c.nranks > 1) {
    if (c.ppe_method == PM_PPE_SOR_LEX)
      return fail(s, PM_ERR_UNSUPPORTED, "sor-lex does not shard (SURVEY 8e); use jacobi or sor-rb with nranks > 1");
    std::string e;
    if (!pm_nccl_init(&s->nccl, c.nccl_id, c.nranks, c.rank, &e)) return fail(s, PM_ERR_NCCL, "%s", e.c_str());
  }
  CK(cudaStreamSynchronize(s->stream));
  return PM_OK;
}

extern "C" int pm_create(const pm_config* cfg, pm_solver** out) {
  if (!cfg || !out) return fail(nullptr, PM_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (cfg->struct_size != sizeof(pm_config))
    return fail(nullptr, PM_ERR_INVALID_ARGUMENT, "pm_config.struct_size %u != %zu (ABI mismatch)", cfg->struct_size, sizeof(pm_config));
  // create_field, cavity-01.cpp:57-59
  if (cfg->nx <= 0 || cfg->ny <= 0) return fail(nullptr, PM_ERR_INVALID_ARGUMENT, "Field dimensions must be positive");
  if (cfg->case_id < PM_CASE_CAVITY || cfg->case_id > PM_CASE_STEP) return fail(nullptr, PM_ERR_INVALID_ARGUMENT, "unknown case %d", cfg->case_id);
  if (cfg->ppe_method < PM_PPE_JACOBI || cfg->ppe_method > PM_PPE_SOR_CHEBY) return fail(nullptr, PM_ERR_INVALID_ARGUMENT, "unknown ppe method %d", cfg->ppe_method);
  if (cfg->nx < 2 || cfg->ny < 2) return fail(nullptr, PM_ERR_INVALID_ARGUMENT, "grid must be at least 2x2");
  if (cfg->kernel_path < PM_PATH_AUTO || cfg->kernel_path > PM_PATH_PERSISTENT) return fail(nullptr, PM_ERR_INVALID_ARGUMENT, "unknown kernel path %d", cfg->kernel_path);
  if (cfg->sweeps_per_pass < 0 || cfg->sweeps_per_pass > 4) return fail(nullptr, PM_ERR_INVALID_ARGUMENT, "sweeps_per_pass out of range");
  if (cfg->nranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->nranks) return fail(nullptr, PM_ERR_INVALID_ARGUMENT, "bad rank %d of %d", cfg->rank, cfg->nranks);
  if (cfg->max_iters < 0 || cfg->max_iters > (1 << 24)) return fail(nullptr, PM_ERR_INVALID_ARGUMENT, "max_iters out of range");
  // validateParameters, cavity-01.cpp:423-425
  if (!(cfg->dt > 0)) return fail(nullptr, PM_ERR_RUNTIME, "Computed time step is non-positive. Check physical parameters!");
  if (!(cfg->dx > 0) || !(cfg->dy > 0)) return fail(nullptr, PM_ERR_INVALID_ARGUMENT, "grid spacing must be positive");
  // backwards_step-01.cpp:459-461
  if (cfg->case_id == PM_CASE_STEP && (cfg->step_i_location <= 0 || cfg->step_i_location >= cfg->nx))
    return fail(nullptr, PM_ERR_RUNTIME, "Step location is outside computational domain!");
  pm_solver* s = new (std::nothrow) pm_solver();
  if (!s) return fail(nullptr, PM_ERR_RUNTIME, "out of host memory");
  const int st = create_impl(s, cfg);
  if (st != PM_OK) {
    g_create_error = s->err;
    destroy_impl(s);
    return st;
  }
  *out = s;
  return PM_OK;
}
extern "C" int pm_destroy(pm_solver* s) { return destroy_impl(s); }

// Host-only test hook: the streaming pass's plan for a slab of nyl rows starting at global row j0 of an nx x ny grid, tile rows
// [row_lo, row_hi) (row_hi < 0: all), `slots` resident warps.  out[8] = {bx0, nbx, by0, nby, rows per chunk, chunks, warps,
// frame tiles}.  Returns 1 if the pass streams, 0 if the tiled kernel keeps every tile.
extern "C" int pm_stream_plan(int nx, int ny, int nyl, int j0, int row_lo, int row_hi, int slots, int* out) {
  using C = TileCfg<PM_PPE_SOR_RB, 4>;
  if (!out || nx < 1 || nyl < 1 || slots < 1) return -1;
  const int tiles_x = (nx + C::TX - 1) / C::TX, tiles_y = (nyl + C::TY - 1) / C::TY;
  if (row_hi < 0) row_hi = tiles_y;
  StreamShape sh{};
  if (!stream_shape(nx, ny, nyl, j0, tiles_x, row_lo, row_hi, C::TX, C::TY, C::SH, C::H, slots, 0, &sh)) return 0;
  const int v[8] = {sh.bx0, sh.nbx, sh.by0, sh.nby, sh.rows, sh.nchunks, sh.items, sh.nframe};
  for (int q = 0; q < 8; ++q) out[q] = v[q];
  return 1;
}

extern "C" int pm_sync(pm_solver* s) {
  if (!s) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  CK(cudaStreamSynchronize(s->stream));
  return PM_OK;
}

// ---------------------------------------------------------------------------
// lazy conversion of the pressure between the natural planes and the split-row buffers of the tiled solve
// ---------------------------------------------------------------------------
static int convert_rows(pm_solver* s, const double* src, double* dst, int to_split) {
  const dim3 b(128, 2), g((s->kp.pitch / 2 + 127) / 128, (s->rows_alloc + 1) / 2);
  k_split_rows<<<g, b, 0, s->stream>>>(s->kp, src, dst, s->rows_alloc, to_split);
  CKL(s);
  return PM_OK;
}
static int ensure_p_natural(pm_solver* s) {  // before anything reads s->pl[s->p_cur]
  if (s->use_tiled && !s->p_nat) {
    PMTRY(convert_rows(s, s->tp[s->tp_cur], s->pl[s->p_cur], 0));
    s->p_nat = true;
  }
  return PM_OK;
}
static int ensure_p_split(pm_solver* s) {  // before the tiled solve reads s->tp[s->tp_cur]
  if (!s->p_split) {
    PMTRY(convert_rows(s, s->pl[s->p_cur], s->tp[s->tp_cur], 1));
    s->p_split = true;
  }
  return PM_OK;
}
static void p_natural_written(pm_solver* s) { s->p_nat = true; s->p_split = false; }

// The streamed host steps (pm_host_step_*) rotate u, v through three plane sets and keep downloads in flight.
// Entry points that write u or v outside that pipeline first refuse to run between a submit and its run, then
// wait for the pending downloads (which read the plane set the handle currently points at).
extern "C" int pm_host_step_drain(pm_solver* s);
static int pipe_quiesce(pm_solver* s, const char* who) {
  if (!s->hp.ready) return PM_OK;
  if (s->hp.submitted != s->hp.run)
    return fail(s, PM_ERR_INVALID_ARGUMENT, "%s: %lld step(s) submitted with pm_host_step_submit and not yet run", who, s->hp.submitted - s->hp.run);
  return pm_host_step_drain(s);
}

// ---------------------------------------------------------------------------
// data movement
// ---------------------------------------------------------------------------
// Local storage rows jl_a..jl_b of a field correspond to global rows j0+jl.
static void local_row_span(const pm_solver* s, int rows_global, bool with_halo, int* ja, int* jb) {
  const KP& k = s->kp;
  int a = with_halo ? 0 : (k.first_rank ? 0 : 1);
  int b = with_halo ? k.nyl + 1 : (k.last_rank ? k.nyl + 1 : k.nyl);
  b = std::min(b, rows_global - 1 - k.j0);
  *ja = a; *jb = b;
}

extern "C" int pm_upload(pm_solver* s, int field, const double* host, size_t count) {
  if (!s || !host) return PM_ERR_INVALID_ARGUMENT;
  int rows, cols;
  field_dims(s, field, &rows, &cols);
  double* dst = field_plane(s, field);
  if (!dst) return fail(s, PM_ERR_INVALID_ARGUMENT, "unknown field %d", field);
  if (count != size_t(rows) * cols) return fail(s, PM_ERR_INVALID_ARGUMENT, "field %d expects %zu elements, got %zu", field, size_t(rows) * cols, count);
  CK(cudaSetDevice(s->device));
  if (field == PM_FIELD_U || field == PM_FIELD_V) PMTRY(pipe_quiesce(s, "pm_upload"));
  int ja, jb;
  local_row_span(s, rows, true, &ja, &jb);
  const KP& k = s->kp;
  CK(cudaMemcpy2DAsync(dst + pm_idx(k, ja, 0), size_t(k.pitch) * 8, host + size_t(k.j0 + ja) * cols, size_t(cols) * 8,
                       size_t(cols) * 8, size_t(jb - ja + 1), cudaMemcpyHostToDevice, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  if (field == PM_FIELD_F) s->f_max_valid = false;
  if (field == PM_FIELD_P) p_natural_written(s);
  return PM_OK;
}
extern "C" int pm_download(pm_solver* s, int field, double* host, size_t count) {
  if (!s || !host) return PM_ERR_INVALID_ARGUMENT;
  int rows, cols;
  field_dims(s, field, &rows, &cols);
  double* src = field_plane(s, field);
  if (!src) return fail(s, PM_ERR_INVALID_ARGUMENT, "unknown field %d", field);
  if (count != size_t(rows) * cols) return fail(s, PM_ERR_INVALID_ARGUMENT, "field %d expects %zu elements, got %zu", field, size_t(rows) * cols, count);
  CK(cudaSetDevice(s->device));
  if (field == PM_FIELD_P) PMTRY(ensure_p_natural(s));
  int ja, jb;
  local_row_span(s, rows, false, &ja, &jb);
  const KP& k = s->kp;
  if (jb >= ja)
    CK(cudaMemcpy2DAsync(host + size_t(k.j0 + ja) * cols, size_t(cols) * 8, src + pm_idx(k, ja, 0), size_t(k.pitch) * 8,
                         size_t(cols) * 8, size_t(jb - ja + 1), cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return PM_OK;
}
extern "C" int pm_slab_rows(pm_solver* s, int field, int* first_global_row, int* nrows, int* ncols) {
  if (!s || field < 0 || field >= PM_FIELD_COUNT) return PM_ERR_INVALID_ARGUMENT;
  int rows, cols, ja, jb;
  field_dims(s, field, &rows, &cols);
  local_row_span(s, rows, true, &ja, &jb);
  if (first_global_row) *first_global_row = s->kp.j0 + ja;
  if (nrows) *nrows = jb - ja + 1;
  if (ncols) *ncols = cols;
  return PM_OK;
}
static int slab_copy(pm_solver* s, int field, double* host, size_t count, bool to_device) {
  if (!s || !host) return PM_ERR_INVALID_ARGUMENT;
  int rows, cols, ja, jb;
  field_dims(s, field, &rows, &cols);
  double* dev = field_plane(s, field);
  if (!dev) return fail(s, PM_ERR_INVALID_ARGUMENT, "unknown field %d", field);
  local_row_span(s, rows, true, &ja, &jb);
  const size_t n = size_t(jb - ja + 1);
  if (count != n * cols) return fail(s, PM_ERR_INVALID_ARGUMENT, "slab of field %d expects %zu elements, got %zu", field, n * cols, count);
  CK(cudaSetDevice(s->device));
  if (to_device && (field == PM_FIELD_U || field == PM_FIELD_V)) PMTRY(pipe_quiesce(s, "pm_upload_slab"));
  if (field == PM_FIELD_P && !to_device) PMTRY(ensure_p_natural(s));
  const KP& k = s->kp;
  if (to_device)
    CK(cudaMemcpy2DAsync(dev + pm_idx(k, ja, 0), size_t(k.pitch) * 8, host, size_t(cols) * 8, size_t(cols) * 8, n, cudaMemcpyHostToDevice, s->stream));
  else
    CK(cudaMemcpy2DAsync(host, size_t(cols) * 8, dev + pm_idx(k, ja, 0), size_t(k.pitch) * 8, size_t(cols) * 8, n, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  if (to_device && field == PM_FIELD_F) s->f_max_valid = false;
  if (to_device && field == PM_FIELD_P) p_natural_written(s);
  return PM_OK;
}
extern "C" int pm_upload_slab(pm_solver* s, int field, const double* host, size_t count) { return slab_copy(s, field, const_cast<double*>(host), count, true); }
extern "C" int pm_download_slab(pm_solver* s, int field, double* host, size_t count) { return slab_copy(s, field, host, count, false); }

extern "C" int pm_upload_mask(pm_solver* s, const uint8_t* is_fluid, size_t count) {
  if (!s || !is_fluid) return PM_ERR_INVALID_ARGUMENT;
  if (s->cfg.case_id != PM_CASE_STEP) return fail(s, PM_ERR_UNSUPPORTED, "only the step case carries an obstacle mask");
  if (count != size_t(s->cfg.ny + 2) * (s->cfg.nx + 2)) return fail(s, PM_ERR_INVALID_ARGUMENT, "mask expects (ny+2)*(nx+2) bytes");
  CK(cudaSetDevice(s->device));
  return upload_mask_rows(s, is_fluid);
}
extern "C" int pm_download_mask(pm_solver* s, uint8_t* is_fluid, size_t count) {
  if (!s || !is_fluid) return PM_ERR_INVALID_ARGUMENT;
  const int cols = s->cfg.nx + 2;
  if (count != size_t(s->cfg.ny + 2) * cols) return fail(s, PM_ERR_INVALID_ARGUMENT, "mask expects (ny+2)*(nx+2) bytes");
  CK(cudaSetDevice(s->device));
  int ja, jb;
  local_row_span(s, s->cfg.ny + 2, false, &ja, &jb);
  const KP& k = s->kp;
  CK(cudaMemcpy2DAsync(is_fluid + size_t(k.j0 + ja) * cols, size_t(cols), s->mask + pm_idx(k, ja, 0), size_t(k.pitch),
                       size_t(cols), size_t(jb - ja + 1), cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return PM_OK;
}

extern "C" int pm_fill_zero(pm_solver* s) {
  if (!s) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  PMTRY(pipe_quiesce(s, "pm_fill_zero"));  // u, v may currently live in one of the streaming plane sets
  CK(cudaMemsetAsync(s->base, 0, s->plane * PL_COUNT * sizeof(double), s->stream));
  if (s->hp.ready) CK(cudaMemsetAsync(s->hp.extra, 0, s->plane * 5 * sizeof(double), s->stream));
  if (s->tp[0]) CK(cudaMemsetAsync(s->tp[0], 0, s->plane * 2 * sizeof(double), s->stream));
  s->p_cur = PL_P0;
  s->tp_cur = 0;
  s->p_nat = true;
  s->p_split = s->use_tiled;  // zero in either layout
  s->f_max_valid = false;
  return PM_OK;
}
extern "C" int pm_fill_random_scaled(pm_solver* s, uint64_t seed, double amplitude);
extern "C" int pm_fill_random(pm_solver* s, uint64_t seed) { return pm_fill_random_scaled(s, seed, 1.0); }
extern "C" int pm_fill_random_scaled(pm_solver* s, uint64_t seed, double amplitude) {
  if (!s) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  PMTRY(pm_fill_zero(s));
  const KP& k = s->kp;
  for (int field = 0; field < PM_FIELD_COUNT; ++field) {
    int rows, cols;
    field_dims(s, field, &rows, &cols);
    const dim3 b(128, 4), g((cols + 127) / 128, (k.nyl + 2 + 3) / 4);
    k_fill_random<<<g, b, 0, s->stream>>>(k, field_plane(s, field), field, rows, cols, seed, amplitude);
    CKL(s);
  }
  p_natural_written(s);
  s->f_max_valid = false;
  return PM_OK;
}

// ---------------------------------------------------------------------------
// halo exchange between slabs (no-ops on a single rank)
// ---------------------------------------------------------------------------
// `depth` halo rows each way on `stream`: my top interior rows nyl-depth+1..nyl -> the upper rank's rows
// 1-depth..0; my bottom interior rows 1..depth -> the lower rank's rows nyl+1..nyl+depth.  Rows are
// contiguous in the pitched plane, so each direction is one message of depth*pitch doubles.
static int exchange_halo(pm_solver* s, double* plane, int depth, cudaStream_t stream) {
  if (s->cfg.nranks == 1) return PM_OK;
  const KP& k = s->kp;
  if (depth > k.padr || depth > k.nyl) return fail(s, PM_ERR_UNSUPPORTED, "halo depth %d exceeds pad rows %d or slab height %d", depth, k.padr, k.nyl);
  std::string e;
  const size_t n = size_t(k.pitch) * depth;
  double* send_up = plane + size_t(k.padr + k.nyl - depth + 1) * k.pitch;
  double* recv_up = plane + size_t(k.padr + k.nyl + 1) * k.pitch;
  double* send_dn = plane + size_t(k.padr + 1) * k.pitch;
  double* recv_dn = plane + size_t(k.padr + 1 - depth) * k.pitch;
  if (!pm_nccl_exchange(&s->nccl, stream, k.last_rank ? nullptr : send_up, k.last_rank ? nullptr : recv_up,
                        k.first_rank ? nullptr : send_dn, k.first_rank ? nullptr : recv_dn, n, &e))
    return fail(s, PM_ERR_NCCL, "%s", e.c_str());
  return PM_OK;
}
static int exchange_halo1(pm_solver* s, double* plane) { return exchange_halo(s, plane, 1, s->stream); }
static int allreduce_res(pm_solver* s, int m_first, int count) {
  if (s->cfg.nranks == 1 || count <= 0) return PM_OK;
  std::string e;
  if (!pm_nccl_allreduce_max_u64(&s->nccl, s->stream, s->d_res + m_first, size_t(count), &e)) return fail(s, PM_ERR_NCCL, "%s", e.c_str());
  return PM_OK;
}

// ---------------------------------------------------------------------------
// phases
// ---------------------------------------------------------------------------
extern "C" int pm_apply_bc(pm_solver* s, int which) {
  if (!s) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  const KP& k = s->kp;
  double* U = which ? s->pl[PL_US] : s->pl[PL_U];
  double* V = which ? s->pl[PL_VS] : s->pl[PL_V];
  const int n = std::max(k.nx, k.nyl) + 2;
  if (k.case_id == PM_CASE_CAVITY) {
    if (which) return PM_OK;  // the cavity never applies BCs to (u*,v*)
    k_bc_cavity<<<(n + 255) / 256, 256, 0, s->stream>>>(k, U, V);
    CKL(s);
  } else {
    k_bc_channel<<<(n + 255) / 256, 256, 0, s->stream>>>(k, U, V);
    CKL(s);
    if (k.has_mask) {
      const dim3 g((k.nx + PM_BX - 1) / PM_BX, (k.nyl + 2 + PM_BY - 1) / PM_BY);  // halo rows included
      k_bc_solid<<<g, cell_block(), 0, s->stream>>>(k, s->mask, U, V);
      CKL(s);
    }
  }
  return PM_OK;
}

extern "C" int pm_predict(pm_solver* s) {
  if (!s) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  const KP& k = s->kp;
  PMTRY(exchange_halo1(s, s->pl[PL_U]));
  PMTRY(exchange_halo1(s, s->pl[PL_V]));
  // two cells per thread, 128-bit rows; the obstacle mask rides along as byte pairs
  const dim3 g = rows_grid(k), b = rows_block();
  if (k.has_mask) {
    if (s->cfg.exact_arith) k_predict_rows<Exact, true><<<g, b, 0, s->stream>>>(k, s->pl[PL_U], s->pl[PL_V], s->mask, s->pl[PL_US], s->pl[PL_VS]);
    else k_predict_rows<Fast, true><<<g, b, 0, s->stream>>>(k, s->pl[PL_U], s->pl[PL_V], s->mask, s->pl[PL_US], s->pl[PL_VS]);
  } else {
    if (s->cfg.exact_arith) k_predict_rows<Exact, false><<<g, b, 0, s->stream>>>(k, s->pl[PL_U], s->pl[PL_V], s->mask, s->pl[PL_US], s->pl[PL_VS]);
    else k_predict_rows<Fast, false><<<g, b, 0, s->stream>>>(k, s->pl[PL_U], s->pl[PL_V], s->mask, s->pl[PL_US], s->pl[PL_VS]);
  }
  CKL(s);
  return PM_OK;
}

static int source_finish(pm_solver* s, int n_partial);

extern "C" int pm_source(pm_solver* s) {
  if (!s) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  const KP& k = s->kp;
  const bool cav = k.case_id == PM_CASE_CAVITY;
  const bool exact = s->cfg.exact_arith != 0;
  PMTRY(exchange_halo1(s, s->pl[PL_VS]));  // f[1][i] reads v*[0][i] of the slab below
  CK(cudaMemsetAsync(&s->d_state->maxf_bits, 0, 2 * sizeof(unsigned long long), s->stream));
  double* partial = (!cav && !exact) ? s->d_partial : nullptr;
  const dim3 g = rows_grid(k), b = rows_block();
  const int n_partial = int(g.x * g.y);  // blocks that write a partial sum
  if (n_partial > s->cap_partial) return fail(s, PM_ERR_RUNTIME, "partial-sum buffer too small: %d blocks, %d slots", n_partial, s->cap_partial);
  if (k.has_mask) {
    if (exact) k_source_rows<Exact, true><<<g, b, 0, s->stream>>>(k, s->pl[PL_US], s->pl[PL_VS], s->mask, s->pl[PL_F], s->d_state, partial);
    else k_source_rows<Fast, true><<<g, b, 0, s->stream>>>(k, s->pl[PL_US], s->pl[PL_VS], s->mask, s->pl[PL_F], s->d_state, partial);
  } else {
    if (exact) k_source_rows<Exact, false><<<g, b, 0, s->stream>>>(k, s->pl[PL_US], s->pl[PL_VS], s->mask, s->pl[PL_F], s->d_state, partial);
    else k_source_rows<Fast, false><<<g, b, 0, s->stream>>>(k, s->pl[PL_US], s->pl[PL_VS], s->mask, s->pl[PL_F], s->d_state, partial);
  }
  CKL(s);
  return source_finish(s, n_partial);
}

// What follows the pass that wrote f, max|f| and the per-block sums: the tolerance's max|f| (cavity), or the mean removal and
// the max|f| behind it (channel / step; channel-01.cpp:621-628, :643-646).
static int source_finish(pm_solver* s, int n_partial) {
  const KP& k = s->kp;
  const bool cav = k.case_id == PM_CASE_CAVITY;
  const bool exact = s->cfg.exact_arith != 0;
  if (cav) {
    // the tolerance rule reads max|f| of this very pass (cavity-01.cpp:628-632)
    PMTRY(publish_words(s, &s->d_state->maxf2_bits, &s->d_state->maxf_bits, sizeof(unsigned long long)));
    s->f_max_valid = true;
    return PM_OK;
  }
  if (s->cfg.nranks > 1) {
    std::string e;
    if (!pm_nccl_allreduce_max_u64(&s->nccl, s->stream, &s->d_state->maxf_bits, 1, &e)) return fail(s, PM_ERR_NCCL, "%s", e.c_str());
    if (exact) {
      // the reference's serial sum, continued from slab to slab in row order (channel-01.cpp:622-625)
      const int r = s->cfg.rank, last = s->cfg.nranks - 1;
      if (r > 0 && !pm_nccl_recv_words(&s->nccl, s->stream, s->d_state->chain, 2, r - 1, &e)) return fail(s, PM_ERR_NCCL, "%s", e.c_str());
      k_mean_serial<<<1, 32, 0, s->stream>>>(k, s->pl[PL_F], s->mask, s->d_state, r > 0);
      CKL(s);
      if (r < last && !pm_nccl_send_words(&s->nccl, s->stream, s->d_state->chain, 2, r + 1, &e)) return fail(s, PM_ERR_NCCL, "%s", e.c_str());
      static_assert(sizeof(double) == sizeof(unsigned long long), "");
      if (!pm_nccl_bcast_words(&s->nccl, s->stream, reinterpret_cast<unsigned long long*>(&s->d_state->mean), 1, last, &e))
        return fail(s, PM_ERR_NCCL, "%s", e.c_str());
      {
        const dim3 gm = rows_grid(k), bm = rows_block();
        if (k.has_mask) k_sub_mean_rows<Exact, true><<<gm, bm, 0, s->stream>>>(k, s->pl[PL_F], s->mask, s->d_state);
        else k_sub_mean_rows<Exact, false><<<gm, bm, 0, s->stream>>>(k, s->pl[PL_F], s->mask, s->d_state);
      }
      CKL(s);
    } else {
      k_sum_partials<<<1, 1024, 0, s->stream>>>(s->d_partial, n_partial, s->d_state);  // local sum -> ke_sum scratch
      CKL(s);
      if (!pm_nccl_allreduce_sum_f64(&s->nccl, s->stream, &s->d_state->ke_sum, 1, &e)) return fail(s, PM_ERR_NCCL, "%s", e.c_str());
      k_mean_from_sum<<<1, 1, 0, s->stream>>>(k.fluid_count_global, s->d_state);
      CKL(s);
      {
        const dim3 gm = rows_grid(k), bm = rows_block();
        if (k.has_mask) k_sub_mean_rows<Fast, true><<<gm, bm, 0, s->stream>>>(k, s->pl[PL_F], s->mask, s->d_state);
        else k_sub_mean_rows<Fast, false><<<gm, bm, 0, s->stream>>>(k, s->pl[PL_F], s->mask, s->d_state);
      }
      CKL(s);
    }
    if (!pm_nccl_allreduce_max_u64(&s->nccl, s->stream, &s->d_state->maxf2_bits, 1, &e)) return fail(s, PM_ERR_NCCL, "%s", e.c_str());
    s->f_max_valid = true;
    return PM_OK;
  }
  if (exact) {
    k_mean_serial<<<1, 32, 0, s->stream>>>(k, s->pl[PL_F], s->mask, s->d_state, 0);
    CKL(s);
    {
        const dim3 gm = rows_grid(k), bm = rows_block();
        if (k.has_mask) k_sub_mean_rows<Exact, true><<<gm, bm, 0, s->stream>>>(k, s->pl[PL_F], s->mask, s->d_state);
        else k_sub_mean_rows<Exact, false><<<gm, bm, 0, s->stream>>>(k, s->pl[PL_F], s->mask, s->d_state);
      }
  } else {
    k_mean_from_partials<<<1, 1024, 0, s->stream>>>(s->d_partial, n_partial, k.fluid_count_global, s->d_state);
    CKL(s);
    {
        const dim3 gm = rows_grid(k), bm = rows_block();
        if (k.has_mask) k_sub_mean_rows<Fast, true><<<gm, bm, 0, s->stream>>>(k, s->pl[PL_F], s->mask, s->d_state);
        else k_sub_mean_rows<Fast, false><<<gm, bm, 0, s->stream>>>(k, s->pl[PL_F], s->mask, s->d_state);
      }
  }
  CKL(s);
  s->f_max_valid = true;
  return PM_OK;
}

// Predictor and source in one pass: the cavity (nothing happens between them), and the channel without a mask (the boundary
// values the source needs are taken inline; the caller runs pm_apply_bc(1) behind this for the stored u*, v*).
static int predict_source_fused(pm_solver* s) {
  CK(cudaSetDevice(s->device));
  const KP& k = s->kp;
  const bool cav = k.case_id == PM_CASE_CAVITY;
  PMTRY(exchange_halo1(s, s->pl[PL_U]));
  PMTRY(exchange_halo(s, s->pl[PL_V], 2, s->stream));  // the row below the slab is recomputed: it reads v two rows down
  CK(cudaMemsetAsync(&s->d_state->maxf_bits, 0, 2 * sizeof(unsigned long long), s->stream));
  const dim3 g(((k.nx + 1) / 2 + PM_RX - 1) / PM_RX, (k.nyl + PM_FUSE_ROWS - 1) / PM_FUSE_ROWS);
  const int n_partial = int(g.x * g.y);
  double* partial = (!cav && !s->cfg.exact_arith) ? s->d_partial : nullptr;
  if (partial && n_partial > s->cap_partial) return fail(s, PM_ERR_RUNTIME, "partial-sum buffer too small: %d blocks, %d slots", n_partial, s->cap_partial);
  double *u = s->pl[PL_U], *v = s->pl[PL_V], *us = s->pl[PL_US], *vs = s->pl[PL_VS], *f = s->pl[PL_F];
  // single rank, cavity: f is final here, and the streaming pass never reads a ghost or pad row of it -> written split as well
  double* fsp = (cav && s->splan.on && s->cfg.nranks == 1) ? s->fsplit : nullptr;
  s->fsplit_valid = fsp != nullptr;
  if (cav) {
    if (s->cfg.exact_arith) k_predict_source<Exact, 0><<<g, PM_RX, 0, s->stream>>>(k, u, v, us, vs, f, s->d_state, partial, fsp);
    else k_predict_source<Fast, 0><<<g, PM_RX, 0, s->stream>>>(k, u, v, us, vs, f, s->d_state, partial, fsp);
  } else {
    if (s->cfg.exact_arith) k_predict_source<Exact, 1><<<g, PM_RX, 0, s->stream>>>(k, u, v, us, vs, f, s->d_state, partial, fsp);
    else k_predict_source<Fast, 1><<<g, PM_RX, 0, s->stream>>>(k, u, v, us, vs, f, s->d_state, partial, fsp);
  }
  CKL(s);
  if (!cav) PMTRY(pm_apply_bc(s, 1));
  return source_finish(s, n_partial);
}

extern "C" int pm_correct(pm_solver* s) {
  if (!s) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  const KP& k = s->kp;
  // after a tiled solve the pressure is read where the solve left it, in the split-row layout
  const int psplit = s->use_tiled && s->p_split && !s->p_nat;
  const double* p = psplit ? s->tp[s->tp_cur] : s->pl[s->p_cur];
  PMTRY(exchange_halo1(s, const_cast<double*>(p)));
  {
    const dim3 g = rows_grid(k), b = rows_block();
    if (k.has_mask) {
      if (s->cfg.exact_arith) k_correct_rows<Exact, true><<<g, b, 0, s->stream>>>(k, s->pl[PL_US], s->pl[PL_VS], p, s->mask, s->pl[PL_U], s->pl[PL_V], psplit);
      else k_correct_rows<Fast, true><<<g, b, 0, s->stream>>>(k, s->pl[PL_US], s->pl[PL_VS], p, s->mask, s->pl[PL_U], s->pl[PL_V], psplit);
    } else {
      if (s->cfg.exact_arith) k_correct_rows<Exact, false><<<g, b, 0, s->stream>>>(k, s->pl[PL_US], s->pl[PL_VS], p, s->mask, s->pl[PL_U], s->pl[PL_V], psplit);
      else k_correct_rows<Fast, false><<<g, b, 0, s->stream>>>(k, s->pl[PL_US], s->pl[PL_VS], p, s->mask, s->pl[PL_U], s->pl[PL_V], psplit);
    }
  }
  CKL(s);
  return PM_OK;
}

// ---------------------------------------------------------------------------
// pressure solve
// ---------------------------------------------------------------------------
template <class A, int FORM>
static int launch_iteration_simple(pm_solver* s, int krel, int kabs) {
  const KP& k = s->kp;
  const int method = s->cfg.ppe_method;
  const bool masked = k.has_mask != 0;
  const int fuse = masked ? 0 : 1;
  double* f = s->pl[PL_F];
  if (method == PM_PPE_JACOBI) {
    // iterate m lives in plane (m & 1) counted from the plane that held p at solve start
    const int buf[2] = {s->p_cur, s->p_cur == PL_P0 ? PL_P1 : PL_P0};
    double* src = s->pl[buf[(kabs - 1) & 1]];
    double* dst = s->pl[buf[kabs & 1]];
    k_jacobi<A, FORM><<<cell_grid(k), cell_block(), 0, s->stream>>>(k, src, dst, f, s->mask, s->d_state, s->d_res, krel, fuse);
    CKL(s);
    if (masked) {
      k_pghost_walls<<<(std::max(k.nx, k.nyl) + 255) / 256, 256, 0, s->stream>>>(k, dst, s->d_state, s->d_res, krel, 0, 1);
      CKL(s);
      PMTRY(exchange_halo1(s, dst));  // a solid cell in my edge row extrapolates from the neighbour slab's fresh fluid values
      k_pghost_solid<<<cell_grid(k), cell_block(), 0, s->stream>>>(k, dst, s->mask, s->d_state, s->d_res, krel, 0, 1);
      CKL(s);
    }
    PMTRY(exchange_halo1(s, dst));
    k_residual<A, FORM><<<cell_grid(k), cell_block(), 0, s->stream>>>(k, dst, f, s->mask, s->d_state, s->d_res, krel, 0);
    CKL(s);
    PMTRY(allreduce_res(s, kabs, 1));
  } else {
    double* p = s->pl[s->p_cur];
    KP k0 = k, k1 = k;  // per colour half-sweep: the relaxation factor of PM_PPE_SOR_CHEBY moves (include/pm.h)
    if (method == PM_PPE_SOR_CHEBY) {
      if (kabs == 1) { s->cheby_q = 0; s->cheby_w = 1.0; }
      if (s->cheby_q != 2 * (kabs - 1)) return fail(s, PM_ERR_RUNTIME, "sor-cheby: iterations must be launched in order");
      const double rho2 = pmi_cheby_rho2(s->cfg.omega);
      if (s->cheby_q > 0) s->cheby_w = pmi_cheby_next_omega(rho2, s->cheby_q, s->cheby_w);
      kp_set_omega(k0, s->cheby_w);
      s->cheby_w = pmi_cheby_next_omega(rho2, s->cheby_q + 1, s->cheby_w);
      kp_set_omega(k1, s->cheby_w);
      s->cheby_q += 2;
    }
    k_rb_colour<A, FORM><<<half_grid(k), cell_block(), 0, s->stream>>>(k0, p, f, s->mask, s->d_state, s->d_res, krel, 0, 1, fuse);
    CKL(s);
    PMTRY(exchange_halo1(s, p));
    k_rb_colour<A, FORM><<<half_grid(k), cell_block(), 0, s->stream>>>(k1, p, f, s->mask, s->d_state, s->d_res, krel, 1, 0, fuse);
    CKL(s);
    if (masked) {
      k_pghost_walls<<<(std::max(k.nx, k.nyl) + 255) / 256, 256, 0, s->stream>>>(k, p, s->d_state, s->d_res, krel, 0, 1);
      CKL(s);
      PMTRY(exchange_halo1(s, p));  // a solid cell in my edge row extrapolates from the neighbour slab's fresh fluid values
      k_pghost_solid<<<cell_grid(k), cell_block(), 0, s->stream>>>(k, p, s->mask, s->d_state, s->d_res, krel, 0, 1);
      CKL(s);
    }
    PMTRY(exchange_halo1(s, p));
    k_residual<A, FORM><<<cell_grid(k), cell_block(), 0, s->stream>>>(k, p, f, s->mask, s->d_state, s->d_res, krel, 0);
    CKL(s);
    PMTRY(allreduce_res(s, kabs, 1));
  }
  s->timing.ppe_passes++;
  return PM_OK;
}

static int launch_iteration_simple(pm_solver* s, int krel, int kabs) {
  const bool cav = s->kp.case_id == PM_CASE_CAVITY;
  if (s->cfg.exact_arith) return cav ? launch_iteration_simple<Exact, 0>(s, krel, kabs) : launch_iteration_simple<Exact, 1>(s, krel, kabs);
  return cav ? launch_iteration_simple<Fast, 0>(s, krel, kabs) : launch_iteration_simple<Fast, 1>(s, krel, kabs);
}

static int read_state(pm_solver* s) {
  static_assert(sizeof(PpeState) % 8 == 0, "PpeState is copied in 64-bit words");
  PMTRY(publish_words(s, s->h_state, s->d_state, sizeof(PpeState)));
  CK(cudaStreamSynchronize(s->stream));
  return PM_OK;
}


// Tiled path: pass n reads buffer in0 ^ (n & 1) holding iterate n*T and writes iterate n*T + nsw to the
// other buffer.  The loop test runs on the device (stop_words_eval); the host only polls the sticky flag.
//
// Slabs: the tile rows whose output the neighbours need (within H rows of the slab edge) are launched
// first; their H halo rows travel by ncclSend/ncclRecv on the comm stream while the interior tile rows
// run; the residual maxima of the pass are max-allreduced in stream order before the next pass tests them.
static void trace_mark(pm_solver* s, int slot, int which, cudaStream_t st) {
  if (slot >= 0) cudaEventRecord(s->trace.ev[slot][which], st);
}
// After a synchronisation of every stream: fold the recorded passes into the sums.
static void trace_collect(pm_solver* s) {
  auto& t = s->trace;
  for (int q = 0; q < t.used; ++q) {
    for (int w = 1; w < t.NEV; ++w) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, t.ev[q][0], t.ev[q][w]) == cudaSuccess) t.sum_us[w] += 1e3 * ms;
      else (void)cudaGetLastError();
    }
    ++t.passes;
  }
  t.used = 0;
}

static int tiled_pass(pm_solver* s, int in, int m0, int nsw, int force) {
  const KP& k = s->kp;
  const TiledPlan& pl = s->tiled;
  const StreamPlan& sp = s->splan;
  const double* f = s->pl[PL_F];
  // A full pass of production red-black: the interior tiles go to the streaming kernel, the frame around them (tiles at a
  // wall) to k_ppe_tiled on the high-priority stream, both at once.
  const bool streamed = sp.on && nsw == pl.sweeps;
  if (nsw == 0 && sp.on && std::getenv("PM_NO_RESIDUAL_ROWS") == nullptr) {
    // the residual-only pass behind a capped solve: one row-wise look at p and f (pm_kernels_residual.cuh), bit-identical to
    // k_ppe_tiled with nsw = 0; nothing is written, so no halo travels
    const dim3 b(128, 4), g((((k.nx + 1) / 2) + 127) / 128, (k.nyl + 3) / 4);
    if (k.case_id == PM_CASE_CAVITY) k_residual_split<0><<<g, b, 0, s->stream>>>(k, pl.p[in], f, s->d_state, s->d_res, pl.fold_part, m0, force);
    else k_residual_split<1><<<g, b, 0, s->stream>>>(k, pl.p[in], f, s->d_state, s->d_res, pl.fold_part, m0, force);
    CKL(s);
    CK(tiled_fold_launch(&pl, k, s->d_res, m0, nsw, s->stream));
    s->timing.kernel_launches += 2;
    if (s->cfg.nranks > 1 && !force && m0 >= 1) PMTRY(allreduce_res(s, m0, 1));
    s->timing.ppe_passes++;
    return PM_OK;
  }
  int bot = 0, top = 0;
  slab_edge_rows(s, &bot, &top);
  if (s->cfg.nranks == 1 || bot + top >= pl.tiles_y) {
    if (streamed && s->cfg.nranks == 1) {
      CK(cudaEventRecord(s->ev_pass, s->stream));
      CK(cudaStreamWaitEvent(s->edge_stream, s->ev_pass, 0));
      CK(tiled_launch_list(&pl, k, in, f, s->d_state, s->d_res, m0, nsw, force, sp.frame, sp.nframe, s->edge_stream));
      CK(cudaEventRecord(s->ev_edge, s->edge_stream));
      CK(stream_launch(&sp, &pl, k, in, s->fsplit, s->d_state, s->d_res, m0, force, s->stream));
      CK(cudaStreamWaitEvent(s->stream, s->ev_edge, 0));
      s->timing.kernel_launches += 2;
    } else {
      CK(tiled_launch(&pl, k, in, f, s->d_state, s->d_res, m0, nsw, force, 0, pl.tiles_y, s->stream));
      s->timing.kernel_launches++;
    }
    if (s->cfg.nranks > 1 && nsw > 0) PMTRY(exchange_halo(s, pl.p[in ^ 1], pl.halo, s->stream));
  } else {
    // The edge tile rows go out on a high-priority stream and the interior rows on the main stream at the same time:
    // the edge CTAs are scheduled first, the interior ones fill the rest of the machine (two edge launches alone would
    // leave half of it idle), and the halo rows travel while the interior is still being swept.
    const int tr = (s->trace.on && streamed && s->trace.used < s->trace.CAP) ? s->trace.used++ : -1;
    trace_mark(s, tr, 0, s->stream);
    CK(cudaEventRecord(s->ev_pass, s->stream));
    CK(cudaStreamWaitEvent(s->edge_stream, s->ev_pass, 0));
    CK(tiled_launch(&pl, k, in, f, s->d_state, s->d_res, m0, nsw, force, 0, bot, s->edge_stream));
    CK(tiled_launch(&pl, k, in, f, s->d_state, s->d_res, m0, nsw, force, pl.tiles_y - top, top, s->edge_stream));
    s->timing.kernel_launches += 2;
    CK(cudaEventRecord(s->ev_edge, s->edge_stream));
    trace_mark(s, tr, 1, s->edge_stream);
    if (nsw > 0) {
      CK(cudaStreamWaitEvent(s->comm_stream, s->ev_edge, 0));
      PMTRY(exchange_halo(s, pl.p[in ^ 1], pl.halo, s->comm_stream));
      CK(cudaEventRecord(s->ev_halo, s->comm_stream));
      trace_mark(s, tr, 2, s->comm_stream);
    }
    if (streamed) {  // the frame of the middle rows behind the edge rows on their stream, the rectangle on the main stream
      CK(tiled_launch_list(&pl, k, in, f, s->d_state, s->d_res, m0, nsw, force, sp.frame, sp.nframe, s->edge_stream));
      CK(cudaEventRecord(s->ev_edge, s->edge_stream));
      trace_mark(s, tr, 3, s->edge_stream);
      CK(stream_launch(&sp, &pl, k, in, s->fsplit, s->d_state, s->d_res, m0, force, s->stream));
      trace_mark(s, tr, 4, s->stream);
      s->timing.kernel_launches += 2;
    } else {
      CK(tiled_launch(&pl, k, in, f, s->d_state, s->d_res, m0, nsw, force, bot, pl.tiles_y - bot - top, s->stream));
      s->timing.kernel_launches++;
    }
    CK(cudaStreamWaitEvent(s->stream, s->ev_edge, 0));
    if (nsw > 0) CK(cudaStreamWaitEvent(s->stream, s->ev_halo, 0));
    CK(tiled_fold_launch(&pl, k, s->d_res, m0, nsw, s->stream));
    s->timing.kernel_launches++;
    trace_mark(s, tr, 5, s->stream);
    if (!force) {
      const int lo = std::max(m0, 1), hi = m0 + std::max(nsw, 1) - 1;
      PMTRY(allreduce_res(s, lo, hi - lo + 1));
    }
    trace_mark(s, tr, 6, s->stream);
    s->timing.ppe_passes++;
    return PM_OK;
  }
  // entries m0 .. m0+max(nsw,1)-1 are now complete on this rank (both colour parts): slots -> res_bits
  CK(tiled_fold_launch(&pl, k, s->d_res, m0, nsw, s->stream));
  s->timing.kernel_launches++;
  if (s->cfg.nranks > 1 && !force) {
    const int lo = std::max(m0, 1), hi = m0 + std::max(nsw, 1) - 1;
    PMTRY(allreduce_res(s, lo, hi - lo + 1));
  }
  s->timing.ppe_passes++;
  return PM_OK;
}

static int tiled_solve(pm_solver* s, int* iters_out, double* res_out) {
  const TiledPlan& pl = s->tiled;
  const int K = s->cfg.max_iters, T = pl.run;  // sweeps per pass
  const int in0 = s->tp_cur;
  CK(tiled_begin_solve(&pl, s->stream));
  if (s->cfg.nranks > 1) {  // halos of the inputs: f once per solve, p as deep as one pass reaches
    PMTRY(exchange_halo(s, s->pl[PL_F], pl.halo, s->stream));
    PMTRY(exchange_halo(s, pl.p[in0], pl.halo, s->stream));
  }
  if (s->splan.on && !s->fsplit_valid) PMTRY(convert_rows(s, s->pl[PL_F], s->fsplit, 1));  // f is constant over the solve
  s->fsplit_valid = false;  // whoever writes f next says so again
  int m = 0, n = 0;
  bool done = false;
  int chunk = s->cfg.poll_chunk > 0 ? s->cfg.poll_chunk : std::max(4, std::min(128, s->last_iters / (2 * T)));
  while (m < K && !done) {
    int launched = 0;
    while (m < K && launched < chunk) {
      const int nsw = std::min(T, K - m);
      PMTRY(tiled_pass(s, in0 ^ (n & 1), m, nsw, 0));
      m += nsw; ++n; ++launched;
    }
    PMTRY(read_state(s));
    done = s->h_state->done != 0;
    if (s->cfg.poll_chunk <= 0) chunk = std::min(128, chunk * 2);
  }
  if (!done) {  // residual of the last iterate (and the loop test for the iterates of the last pass)
    PMTRY(tiled_pass(s, in0 ^ (n & 1), m, 0, 0));
    PMTRY(read_state(s));
    done = s->h_state->done != 0;
  }
  int iters, buf;
  if (done) {
    iters = s->h_state->iters;
    const int nb = iters / T, mb = nb * T;
    buf = in0 ^ (nb & 1);
    if (iters > mb) {  // land on the exact iterate: replay the first iters-mb sweeps of that pass
      PMTRY(tiled_pass(s, buf, mb, iters - mb, 1));
      buf ^= 1;
    }
  } else {
    iters = K;
    buf = in0 ^ (n & 1);
  }
  s->tp_cur = buf;
  s->p_split = true;
  s->p_nat = false;
  if (iters >= 1) {
    PMTRY(publish_words(s, s->h_res, s->d_res + iters, sizeof(unsigned long long)));
    CK(cudaStreamSynchronize(s->stream));
    std::memcpy(res_out, s->h_res, 8);
  } else {
    *res_out = s->h_state->res_init;
  }
  if (s->trace.on) {
    CK(cudaStreamSynchronize(s->stream));
    trace_collect(s);
  }
  *iters_out = iters;
  return PM_OK;
}

// Reference ordering: one persistent CTA runs the whole solve (pm_kernels_lex.cuh).
static int lex_solve(pm_solver* s, int* iters_out, double* res_out) {
  const KP& k = s->kp;
  const size_t bytes = size_t(k.ny + 2) * size_t(k.nx + 2) * sizeof(double);
  const int use_smem = bytes <= size_t(200) * 1024 ? 1 : 0;
  const int threads = std::min(1024, ((k.nx + 31) / 32) * 32);
  const size_t smem = use_smem ? bytes : 0;
  const void* kern;
  if (k.case_id == PM_CASE_CAVITY) kern = reinterpret_cast<const void*>(&k_ppe_lex<0, false>);
  else if (k.has_mask) kern = reinterpret_cast<const void*>(&k_ppe_lex<1, true>);
  else kern = reinterpret_cast<const void*>(&k_ppe_lex<1, false>);
  if (use_smem && k.nx <= 1024 && std::getenv("PM_LEX_GENERIC") == nullptr) {
    // one column per thread: shuffled west value, register-carried south value, prefetched old operands
    const void* kc;
    if (k.case_id == PM_CASE_CAVITY) kc = reinterpret_cast<const void*>(&k_ppe_lex_cols<0, false>);
    else if (k.has_mask) kc = reinterpret_cast<const void*>(&k_ppe_lex_cols<1, true>);
    else kc = reinterpret_cast<const void*>(&k_ppe_lex_cols<1, false>);
    CK(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const int thr = std::min(1024, std::max(256, ((k.nx + 31) / 32) * 32));
    double* pgc = s->pl[s->p_cur];
    const double* fcc = s->pl[PL_F];
    const uint8_t* mc = s->mask;
    PpeState* stc = s->d_state;
    unsigned long long* rbc = s->d_res;
    void* cargs[] = {(void*)&k, (void*)&pgc, (void*)&fcc, (void*)&mc, (void*)&stc, (void*)&rbc};
    CK(cudaLaunchKernel(kc, dim3(1), dim3(thr), cargs, smem, s->stream));
  } else {
  if (use_smem) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  double* pg = s->pl[s->p_cur];
  const double* f = s->pl[PL_F];
  const uint8_t* m = s->mask;
  PpeState* st = s->d_state;
  unsigned long long* rb = s->d_res;
  int us = use_smem;
  void* args[] = {(void*)&k, (void*)&pg, (void*)&f, (void*)&m, (void*)&st, (void*)&rb, (void*)&us};
  CK(cudaLaunchKernel(kern, dim3(1), dim3(threads), args, smem, s->stream));
  }
  s->timing.kernel_launches++;
  s->timing.ppe_passes++;
  PMTRY(read_state(s));
  const int iters = s->h_state->iters;
  if (iters >= 1) {
    PMTRY(publish_words(s, s->h_res, s->d_res + iters, sizeof(unsigned long long)));
    CK(cudaStreamSynchronize(s->stream));
    std::memcpy(res_out, s->h_res, 8);
  } else {
    *res_out = s->h_state->res_init;
  }
  *iters_out = iters;
  return PM_OK;
}

template <class A>
static const void* small_kernel(const KP& k, int method) {
  const bool rb = method == PM_PPE_SOR_RB;
  if (k.case_id == PM_CASE_CAVITY)
    return rb ? reinterpret_cast<const void*>(&k_ppe_small<A, 0, false, PM_PPE_SOR_RB>) : reinterpret_cast<const void*>(&k_ppe_small<A, 0, false, PM_PPE_JACOBI>);
  if (k.has_mask)
    return rb ? reinterpret_cast<const void*>(&k_ppe_small<A, 1, true, PM_PPE_SOR_RB>) : reinterpret_cast<const void*>(&k_ppe_small<A, 1, true, PM_PPE_JACOBI>);
  return rb ? reinterpret_cast<const void*>(&k_ppe_small<A, 1, false, PM_PPE_SOR_RB>) : reinterpret_cast<const void*>(&k_ppe_small<A, 1, false, PM_PPE_JACOBI>);
}
template <class A>
static const void* cluster_kernel(const KP& k) {
  if (k.case_id == PM_CASE_CAVITY) return reinterpret_cast<const void*>(&k_ppe_cluster<A, 0, false>);
  if (k.has_mask) return reinterpret_cast<const void*>(&k_ppe_cluster<A, 1, true>);
  return reinterpret_cast<const void*>(&k_ppe_cluster<A, 1, false>);
}
// Small grids, red-black: the persistent solve on a cluster of PM_CLUSTER (16) CTAs with DSMEM halo rows (k_ppe_cluster).
static int cluster_solve(pm_solver* s) {
  const KP& k = s->kp;
  const int nr_max = k.ny / PM_CLUSTER + (k.ny % PM_CLUSTER ? 1 : 0);
  const size_t smem = size_t(nr_max + 2) * size_t(k.nx + 2) * sizeof(double);
  const void* kern = s->cfg.exact_arith ? cluster_kernel<Exact>(k) : cluster_kernel<Fast>(k);
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  if (PM_CLUSTER > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  // each thread keeps up to PM_CLUSTER_CPT cells per colour: enough threads for the band's half rows
  const int per_colour = nr_max * ((k.nx + 1) / 2);
  const int threads = std::min(512, std::max(64, (((per_colour + PM_CLUSTER_CPT - 1) / PM_CLUSTER_CPT + 31) / 32) * 32));
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(PM_CLUSTER);
  lc.blockDim = dim3(threads);
  lc.dynamicSmemBytes = smem;
  lc.stream = s->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = PM_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at;
  lc.numAttrs = 1;
  double* pg = s->pl[s->p_cur];
  const double* f = s->pl[PL_F];
  const uint8_t* m = s->mask;
  PpeState* st = s->d_state;
  unsigned long long* rb = s->d_res;
  void* args[] = {(void*)&k, (void*)&pg, (void*)&f, (void*)&m, (void*)&st, (void*)&rb};
  if (!s->cluster_checked) {  // a cluster of 16 is a non-portable size: ask once whether this device can place one
    int nclusters = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &lc);
    s->cluster_checked = true;
    if (e != cudaSuccess || nclusters < 1) {
      (void)cudaGetLastError();
      s->no_cluster = true;
      return PM_ERR_UNSUPPORTED;  // small_solve falls back to the single-CTA solve
    }
  }
  CK(cudaLaunchKernelExC(&lc, kern, args));
  s->timing.kernel_launches++;
  s->timing.ppe_passes++;
  return PM_OK;
}

// Small grids: Jacobi / red-black in one persistent CTA (pm_kernels_lex.cuh, k_ppe_small), or red-black on a
// cluster when the band of every CTA is at least two rows high and fits its shared memory.
static int small_solve(pm_solver* s, int* iters_out, double* res_out) {
  const KP& k = s->kp;
  {
    const int nr_max = k.ny / PM_CLUSTER + (k.ny % PM_CLUSTER ? 1 : 0);
    const size_t csmem = size_t(nr_max + 2) * size_t(k.nx + 2) * sizeof(double);
    const bool fits = nr_max * ((k.nx + 1) / 2) <= 512 * PM_CLUSTER_CPT;  // cells per colour and CTA held in registers
    int cst = PM_ERR_UNSUPPORTED;
    if (s->cfg.ppe_method == PM_PPE_SOR_RB && k.ny >= 2 * PM_CLUSTER && csmem <= size_t(200) * 1024 && fits && !s->no_cluster) {
      cst = cluster_solve(s);
      if (cst != PM_OK && !(cst == PM_ERR_UNSUPPORTED && s->no_cluster)) return cst;
    }
    if (cst == PM_OK) {
      PMTRY(read_state(s));
      const int iters = s->h_state->iters;
      if (iters >= 1) {
        PMTRY(publish_words(s, s->h_res, s->d_res + iters, sizeof(unsigned long long)));
        CK(cudaStreamSynchronize(s->stream));
        std::memcpy(res_out, s->h_res, 8);
      } else {
        *res_out = s->h_state->res_init;
      }
      *iters_out = iters;
      return PM_OK;
    }
  }
  const size_t smem = size_t(k.ny + 2) * size_t(k.nx + 2) * sizeof(double);
  const int cells = k.nx * k.ny;
  const int threads = std::min(1024, std::max(128, ((cells / 4 + 31) / 32) * 32));
  const void* kern = s->cfg.exact_arith ? small_kernel<Exact>(k, s->cfg.ppe_method) : small_kernel<Fast>(k, s->cfg.ppe_method);
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  double* pg = s->pl[s->p_cur];
  const double* f = s->pl[PL_F];
  const uint8_t* m = s->mask;
  PpeState* st = s->d_state;
  unsigned long long* rb = s->d_res;
  void* args[] = {(void*)&k, (void*)&pg, (void*)&f, (void*)&m, (void*)&st, (void*)&rb};
  CK(cudaLaunchKernel(kern, dim3(1), dim3(threads), args, smem, s->stream));
  s->timing.kernel_launches++;
  s->timing.ppe_passes++;
  PMTRY(read_state(s));
  const int iters = s->h_state->iters;
  if (iters >= 1) {
    PMTRY(publish_words(s, s->h_res, s->d_res + iters, sizeof(unsigned long long)));
    CK(cudaStreamSynchronize(s->stream));
    std::memcpy(res_out, s->h_res, 8);
  } else {
    *res_out = s->h_state->res_init;
  }
  *iters_out = iters;
  return PM_OK;
}

extern "C" int pm_ppe_solve(pm_solver* s, pm_ppe_result* out) {
  if (!s) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  const KP& k = s->kp;
  const pm_config& c = s->cfg;
  const bool cav = k.case_id == PM_CASE_CAVITY;
  if (c.nranks > 1 && s->nccl.comm == nullptr) return fail(s, PM_ERR_NCCL, "communicator missing");

  CK(cudaEventRecord(s->ev_a, s->stream));
  if (!s->f_max_valid) {
    CK(cudaMemsetAsync(&s->d_state->maxf2_bits, 0, sizeof(unsigned long long), s->stream));
    k_max_f<<<cell_grid(k), cell_block(), 0, s->stream>>>(k, s->pl[PL_F], s->mask, s->d_state);
    CKL(s);
    s->f_max_valid = true;
  }
  if (c.nranks > 1) {
    std::string e;
    if (!pm_nccl_allreduce_max_u64(&s->nccl, s->stream, &s->d_state->maxf2_bits, 1, &e)) return fail(s, PM_ERR_NCCL, "%s", e.c_str());
  }
  if (cav) {  // cold start from p == 0 (cavity-01.cpp:610-611)
    // Only the buffer the solve starts from needs clearing: the first pass overwrites every interior cell of the
    // other one, and in the cavity nothing ever writes a ghost or pad cell of either buffer (they stay 0 from
    // pm_create / pm_fill_zero; halo rows between slabs are re-exchanged every pass).
    if (s->use_tiled) {
      CK(cudaMemsetAsync(s->tp[s->tp_cur], 0, s->plane * sizeof(double), s->stream));
      s->p_split = true;
      s->p_nat = false;
    } else {
      CK(cudaMemsetAsync(s->pl[s->p_cur], 0, s->plane * sizeof(double), s->stream));
    }
  } else if (s->use_tiled) {
    PMTRY(ensure_p_split(s));  // warm start (channel-01.cpp:636)
  }
  if (!cav && !s->use_small && (c.ppe_method == PM_PPE_JACOBI || s->use_tiled)) {  // ping-pong solves: both buffers carry the corner ghosts
    if (s->use_tiled) k_copy_corners<<<1, 32, 0, s->stream>>>(k, s->tp[s->tp_cur], s->tp[s->tp_cur ^ 1], 1);
    else k_copy_corners<<<1, 32, 0, s->stream>>>(k, s->pl[s->p_cur], s->pl[s->p_cur == PL_P0 ? PL_P1 : PL_P0], 0);
    CKL(s);
  }
  CK(cudaMemsetAsync(s->d_res, 0, size_t(c.max_iters + 2) * sizeof(unsigned long long), s->stream));
  k_ppe_begin<<<1, 1, 0, s->stream>>>(k, s->d_state);
  CKL(s);

  int iters = 0;
  double res = 0.0;
  if (c.ppe_method == PM_PPE_SOR_LEX) {
    PMTRY(lex_solve(s, &iters, &res));
  } else if (s->use_tiled) {
    PMTRY(tiled_solve(s, &iters, &res));
    if (k.has_mask && iters >= 1 && std::getenv("PM_DEBUG_NO_FIXUP") == nullptr) {
      // The tiled solve leaves the solid cells one applyPressureGhosts behind (pm_kernels_tiled.cuh, masked_tile): apply the
      // last one to the final iterate, in the reference's order -- wall ghosts, then the solid cells (backwards_step-01.cpp:685-740).
      double* p = s->tp[s->tp_cur];
      k_pghost_walls<<<(std::max(k.nx, k.nyl) + 255) / 256, 256, 0, s->stream>>>(k, p, s->d_state, s->d_res, 0, 1, 0);
      CKL(s);
      PMTRY(exchange_halo1(s, p));
      k_pghost_solid<<<cell_grid(k), cell_block(), 0, s->stream>>>(k, p, s->mask, s->d_state, s->d_res, 0, 1, 0);
      CKL(s);
      PMTRY(exchange_halo1(s, p));
    }
  } else if (s->use_small) {
    PMTRY(small_solve(s, &iters, &res));
  } else {
    if (c.ppe_method != PM_PPE_JACOBI) PMTRY(exchange_halo1(s, s->pl[s->p_cur]));
    const int K = c.max_iters;
    int kdone = 0;
    int chunk = c.poll_chunk > 0 ? c.poll_chunk : std::max(4, std::min(256, s->last_iters / 8));
    bool done = false;
    while (kdone < K && !done) {
      const int n = std::min(chunk, K - kdone);
      for (int q = 1; q <= n; ++q) PMTRY(launch_iteration_simple(s, kdone + q, kdone + q));
      kdone += n;
      PMTRY(read_state(s));
      done = s->h_state->done != 0;
      if (c.poll_chunk <= 0) chunk = std::min(256, chunk * 2);
    }
    if (!done) {
      // the cap ended the loop; one more look at the flag logic is not needed: iterate K is final
      PMTRY(read_state(s));
    }
    iters = s->h_state->done ? s->h_state->iters : K;
    if (K == 0) iters = 0;
    unsigned long long bits = 0;
    if (iters >= 1) {
      PMTRY(publish_words(s, s->h_res, s->d_res + iters, sizeof(unsigned long long)));
      CK(cudaStreamSynchronize(s->stream));
      bits = s->h_res[0];
      std::memcpy(&res, &bits, 8);
    } else {
      res = s->h_state->res_init;
    }
    if (c.ppe_method == PM_PPE_JACOBI && (iters & 1)) s->p_cur = (s->p_cur == PL_P0) ? PL_P1 : PL_P0;
  }
  CK(cudaEventRecord(s->ev_b, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, s->ev_a, s->ev_b));
  s->timing.ppe_ms += ms;
  s->last_iters = iters;
  if (out) {
    out->iterations = iters;
    out->hit_cap = iters >= c.max_iters;
    out->residual = res;
    out->tolerance = s->h_state->tol;
    unsigned long long mb = s->h_state->maxf2_bits;
    std::memcpy(&out->max_source, &mb, 8);
  }
  return PM_OK;
}

static int step_impl(pm_solver* s, int nsteps, pm_ppe_result* last) {
  pm_ppe_result r{};
  for (int n = 0; n < nsteps; ++n) {
    const double ppe_before = s->timing.ppe_ms;
    CK(cudaEventRecord(s->ev_s0, s->stream));
    if (s->kp.case_id == PM_CASE_CAVITY) {  // cavity-01.cpp:387-390
      PMTRY(pm_apply_bc(s, 0));
      if (s->cfg.nranks == 1 || s->kp.nyl >= 2) {
        PMTRY(predict_source_fused(s));
      } else {
        PMTRY(pm_predict(s));
        PMTRY(pm_source(s));
      }
      PMTRY(pm_ppe_solve(s, &r));
      PMTRY(pm_correct(s));
    } else {  // channel-01.cpp:368-375
      if (s->kp.case_id == PM_CASE_CHANNEL && (s->cfg.nranks == 1 || s->kp.nyl >= 2) && std::getenv("PM_NO_FUSE") == nullptr) {
        PMTRY(predict_source_fused(s));  // predictor + the boundary values the source reads + source, then the BC kernel
      } else {
        PMTRY(pm_predict(s));
        PMTRY(pm_apply_bc(s, 1));
        PMTRY(pm_source(s));
      }
      PMTRY(pm_ppe_solve(s, &r));
      PMTRY(pm_correct(s));
      PMTRY(pm_apply_bc(s, 0));
    }
    CK(cudaEventRecord(s->ev_s1, s->stream));
    s->step_timed = true;
    s->last_step_ppe_ms = s->timing.ppe_ms - ppe_before;
  }
  if (last) *last = r;
  return PM_OK;
}

extern "C" int pm_step(pm_solver* s, int nsteps, pm_ppe_result* last) {
  if (!s || nsteps < 0) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  PMTRY(pipe_quiesce(s, "pm_step"));
  return step_impl(s, nsteps, last);
}

// ---------------------------------------------------------------------------
// host-resident steps, streamed
// ---------------------------------------------------------------------------
static int host_pipe_init(pm_solver* s) {
  HostPipe& h = s->hp;
  if (h.ready) return PM_OK;
  CK(cudaMalloc(&h.extra, s->plane * 5 * sizeof(double)));
  CK(cudaMemsetAsync(h.extra, 0, s->plane * 5 * sizeof(double), s->stream));  // pad cells of every plane are zero
  h.u[0] = s->pl[PL_U]; h.v[0] = s->pl[PL_V];
  h.u[1] = h.extra; h.u[2] = h.extra + s->plane;
  h.v[1] = h.extra + 2 * s->plane; h.v[2] = h.extra + 3 * s->plane;
  h.pout = h.extra + 4 * s->plane;
  CK(cudaStreamCreateWithFlags(&h.h2d, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&h.d2h, cudaStreamNonBlocking));
  for (int q = 0; q < 3; ++q) {
    CK(cudaEventCreateWithFlags(&h.ev_up[q], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h.ev_down[q], cudaEventDisableTiming));
  }
  CK(cudaEventCreateWithFlags(&h.ev_step, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&h.ev_pdown, cudaEventDisableTiming));
  CK(cudaStreamSynchronize(s->stream));
  h.ready = true;
  return PM_OK;
}
// This rank's rows of `field` between a dense host array (pm_upload_slab layout) and `plane`, asynchronously.
static int slab_copy_async(pm_solver* s, int field, double* plane, double* host, size_t count, bool to_device, cudaStream_t st) {
  int rows, cols, ja, jb;
  field_dims(s, field, &rows, &cols);
  local_row_span(s, rows, true, &ja, &jb);
  const size_t n = size_t(jb - ja + 1);
  if (count != n * cols) return fail(s, PM_ERR_INVALID_ARGUMENT, "slab of field %d expects %zu elements, got %zu", field, n * cols, count);
  const KP& k = s->kp;
  if (to_device)
    CK(cudaMemcpy2DAsync(plane + pm_idx(k, ja, 0), size_t(k.pitch) * 8, host, size_t(cols) * 8, size_t(cols) * 8, n, cudaMemcpyHostToDevice, st));
  else
    CK(cudaMemcpy2DAsync(host, size_t(cols) * 8, plane + pm_idx(k, ja, 0), size_t(k.pitch) * 8, size_t(cols) * 8, n, cudaMemcpyDeviceToHost, st));
  return PM_OK;
}

extern "C" int pm_host_step_submit(pm_solver* s, const double* u_in, size_t u_count, const double* v_in, size_t v_count,
                                   double* u_out, double* v_out, double* p_out, size_t p_count) {
  if (!s || !u_in || !v_in || !u_out || !v_out || !p_out) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  PMTRY(host_pipe_init(s));
  HostPipe& h = s->hp;
  if (h.submitted - h.run >= 2) return fail(s, PM_ERR_INVALID_ARGUMENT, "two steps are already submitted and not yet run");
  {
    int rows, cols, ja, jb;
    field_dims(s, PM_FIELD_P, &rows, &cols);
    local_row_span(s, rows, true, &ja, &jb);
    if (p_count != size_t(jb - ja + 1) * cols) return fail(s, PM_ERR_INVALID_ARGUMENT, "p slab expects %zu elements, got %zu", size_t(jb - ja + 1) * cols, p_count);
  }
  const int set = int(h.submitted % 3);
  // the planes of this set may still be the source of the download of the step three back
  if (h.down_pending[set]) CK(cudaStreamWaitEvent(h.h2d, h.ev_down[set], 0));
  PMTRY(slab_copy_async(s, PM_FIELD_U, h.u[set], const_cast<double*>(u_in), u_count, true, h.h2d));
  PMTRY(slab_copy_async(s, PM_FIELD_V, h.v[set], const_cast<double*>(v_in), v_count, true, h.h2d));
  CK(cudaEventRecord(h.ev_up[set], h.h2d));
  h.job[set] = {u_out, v_out, p_out};
  h.submitted++;
  return PM_OK;
}

extern "C" int pm_host_step_run(pm_solver* s, pm_ppe_result* r) {
  if (!s) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  HostPipe& h = s->hp;
  if (!h.ready || h.run >= h.submitted) return fail(s, PM_ERR_INVALID_ARGUMENT, "no submitted step to run");
  const int set = int(h.run % 3);
  CK(cudaStreamWaitEvent(s->stream, h.ev_up[set], 0));
  s->pl[PL_U] = h.u[set];
  s->pl[PL_V] = h.v[set];
  PMTRY(step_impl(s, 1, r));
  // results: p leaves through its own plane (the next solve overwrites both pressure buffers); u, v stay where
  // they are -- the next two steps use the other plane sets
  if (h.pdown_pending) CK(cudaStreamWaitEvent(s->stream, h.ev_pdown, 0));
  if (s->use_tiled && !s->p_nat) {
    PMTRY(convert_rows(s, s->tp[s->tp_cur], h.pout, 0));  // straight from the solve's split-row buffer
  } else {
    k_copy_plane<<<1184, 256, 0, s->stream>>>(reinterpret_cast<const double2*>(s->pl[s->p_cur]), reinterpret_cast<double2*>(h.pout), s->plane / 2);
    CKL(s);
  }
  CK(cudaEventRecord(h.ev_step, s->stream));
  CK(cudaStreamWaitEvent(h.d2h, h.ev_step, 0));
  const HostPipe::Job& j = h.job[set];
  int rows, cols, ja, jb;
  auto count_of = [&](int field) { field_dims(s, field, &rows, &cols); local_row_span(s, rows, true, &ja, &jb); return size_t(jb - ja + 1) * cols; };
  PMTRY(slab_copy_async(s, PM_FIELD_P, h.pout, j.p_out, count_of(PM_FIELD_P), false, h.d2h));
  CK(cudaEventRecord(h.ev_pdown, h.d2h));
  h.pdown_pending = true;
  PMTRY(slab_copy_async(s, PM_FIELD_U, h.u[set], j.u_out, count_of(PM_FIELD_U), false, h.d2h));
  PMTRY(slab_copy_async(s, PM_FIELD_V, h.v[set], j.v_out, count_of(PM_FIELD_V), false, h.d2h));
  CK(cudaEventRecord(h.ev_down[set], h.d2h));
  h.down_pending[set] = true;
  h.run++;
  return PM_OK;
}

extern "C" int pm_host_step_drain(pm_solver* s) {
  if (!s) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  if (!s->hp.ready) return PM_OK;
  CK(cudaStreamSynchronize(s->hp.h2d));
  CK(cudaStreamSynchronize(s->stream));
  CK(cudaStreamSynchronize(s->hp.d2h));
  return PM_OK;
}

extern "C" int pm_diagnostics(pm_solver* s, double* max_div, double* avg_ke) {
  if (!s || !max_div || !avg_ke) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  const KP& k = s->kp;
  CK(cudaMemsetAsync(&s->d_state->div_bits, 0, sizeof(unsigned long long), s->stream));
  PMTRY(exchange_halo1(s, s->pl[PL_V]));
  k_diag<<<cell_grid(k), cell_block(), 0, s->stream>>>(k, s->pl[PL_U], s->pl[PL_V], s->mask, s->d_state, s->d_partial);
  CKL(s);
  k_sum_partials<<<1, 1024, 0, s->stream>>>(s->d_partial, s->n_partial, s->d_state);
  CKL(s);
  if (s->cfg.nranks > 1) {
    std::string e;
    if (!pm_nccl_allreduce_max_u64(&s->nccl, s->stream, &s->d_state->div_bits, 1, &e) ||
        !pm_nccl_allreduce_sum_f64(&s->nccl, s->stream, &s->d_state->ke_sum, 1, &e))
      return fail(s, PM_ERR_NCCL, "%s", e.c_str());
  }
  PMTRY(read_state(s));
  unsigned long long b = s->h_state->div_bits;
  std::memcpy(max_div, &b, 8);
  const double ke = s->h_state->ke_sum;
  if (k.case_id == PM_CASE_STEP) *avg_ke = k.fluid_count_global > 0 ? ke / k.fluid_count_global : 0.0;
  else *avg_ke = ke / (k.nx * k.ny);
  return PM_OK;
}

// ---------------------------------------------------------------------------
// export: what the VTK writers print, formed on the device; the copy to the host runs on its own stream
// ---------------------------------------------------------------------------
// The staging buffers.  Allocation synchronises devices, so with several handles driven by threads of ONE process it must
// not run while another handle's NCCL kernel waits for this one's (the classic cudaMalloc / NCCL deadlock): such callers
// prepare every handle up front, behind a barrier; everyone else gets it lazily from pm_export_begin.
extern "C" int pm_export_prepare(pm_solver* s) {
  if (!s) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  auto& x = s->ex;
  if (x.dev) return PM_OK;
  const size_t n = size_t(s->kp.nx) * size_t(s->kp.nyl);
  CK(cudaMalloc(&x.dev, 5 * n * sizeof(double)));
  CK(cudaMallocHost(&x.host, 5 * n * sizeof(double)));
  CK(cudaStreamCreateWithFlags(&x.stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&x.ready, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&x.done, cudaEventDisableTiming));
  return PM_OK;
}

extern "C" int pm_export_begin(pm_solver* s) {
  if (!s) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  const KP& k = s->kp;
  auto& x = s->ex;
  if (x.pending) return fail(s, PM_ERR_INVALID_ARGUMENT, "pm_export_begin: the previous export has not been collected with pm_export_wait");
  const size_t n = size_t(k.nx) * size_t(k.nyl);
  PMTRY(pm_export_prepare(s));
  // the vorticity reads the cell centres one row up and down: one halo row of u and of v
  PMTRY(exchange_halo1(s, s->pl[PL_U]));
  PMTRY(exchange_halo1(s, s->pl[PL_V]));
  const int psplit = s->use_tiled && s->p_split && !s->p_nat;
  const double* p = psplit ? s->tp[s->tp_cur] : s->pl[s->p_cur];
  k_export<<<cell_grid(k), cell_block(), 0, s->stream>>>(k, s->pl[PL_U], s->pl[PL_V], p, psplit, s->mask, x.dev, n);
  CKL(s);
  CK(cudaEventRecord(x.ready, s->stream));
  CK(cudaStreamWaitEvent(x.stream, x.ready, 0));
  CK(cudaMemcpyAsync(x.host, x.dev, 5 * n * sizeof(double), cudaMemcpyDeviceToHost, x.stream));
  CK(cudaEventRecord(x.done, x.stream));
  x.pending = true;
  return PM_OK;
}

extern "C" int pm_export_wait(pm_solver* s, double* uc, double* vc, double* mag, double* p, double* vort, size_t count) {
  if (!s || !uc || !vc || !mag || !p || !vort) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  const KP& k = s->kp;
  auto& x = s->ex;
  if (!x.pending) return fail(s, PM_ERR_INVALID_ARGUMENT, "pm_export_wait without pm_export_begin");
  if (count != size_t(k.nx) * size_t(k.ny)) return fail(s, PM_ERR_INVALID_ARGUMENT, "pm_export_wait: count %zu != nx*ny", count);
  CK(cudaEventSynchronize(x.done));
  x.pending = false;
  const size_t n = size_t(k.nx) * size_t(k.nyl), off = size_t(k.j0) * size_t(k.nx);
  double* dst[5] = {uc, vc, mag, p, vort};
  for (int q = 0; q < 5; ++q) std::memcpy(dst[q] + off, x.host + size_t(q) * n, n * sizeof(double));
  return PM_OK;
}

extern "C" int pm_get_timing(pm_solver* s, pm_timing* t) {
  if (!s || !t) return PM_ERR_INVALID_ARGUMENT;
  if (s->step_timed) {  // the non-pressure phases of the most recent pm_step: its device time minus its pressure solve
    CK(cudaSetDevice(s->device));
    CK(cudaEventSynchronize(s->ev_s1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, s->ev_s0, s->ev_s1));
    s->timing.other_ms = std::max(0.0, double(ms) - s->last_step_ppe_ms);
  }
  *t = s->timing;
  return PM_OK;
}

extern "C" int pm_timer_start(pm_solver* s) {
  if (!s) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  CK(cudaEventRecord(s->ev_t0, s->stream));
  return PM_OK;
}
extern "C" int pm_timer_stop(pm_solver* s, double* elapsed_ms) {
  if (!s || !elapsed_ms) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  CK(cudaEventRecord(s->ev_t1, s->stream));
  CK(cudaEventSynchronize(s->ev_t1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, s->ev_t0, s->ev_t1));
  *elapsed_ms = ms;
  return PM_OK;
}

// Measurement helper (not part of pm.h): how many CTAs / clusters of the tiled pressure kernel the device keeps resident.
extern "C" int pm_debug_tiled_occupancy(pm_solver* s, int* ctas_per_sm, int* max_active_clusters, int* cluster_size) {
  if (!s || !s->use_tiled) return PM_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(s->device));
  const TiledPlan& pl = s->tiled;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, pl.kernel, pl.threads, size_t(pl.smem_bytes)));
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(pl.tiles_x, pl.tiles_y * pl.cs);
  lc.blockDim = dim3(pl.threads);
  lc.dynamicSmemBytes = size_t(pl.smem_bytes);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = unsigned(pl.cs); at[0].val.clusterDim.z = 1;
  lc.attrs = at;
  lc.numAttrs = 1;
  CK(cudaOccupancyMaxActiveClusters(max_active_clusters, pl.kernel, &lc));
  *cluster_size = pl.cs;
  return PM_OK;
}

#ifdef PM_TILE_PROFILE
// Debug builds only (make variant EXTRA=-DPM_TILE_PROFILE): cycles per phase of the tiled kernel, summed over CTAs.
extern "C" int pm_debug_tile_profile(unsigned long long out[8], int reset) {
  if (out && cudaMemcpyFromSymbol(out, g_tile_prof, 8 * sizeof(unsigned long long)) != cudaSuccess) return PM_ERR_CUDA;
  if (reset) {
    unsigned long long z[8] = {};
    if (cudaMemcpyToSymbol(g_tile_prof, z, sizeof z) != cudaSuccess) return PM_ERR_CUDA;
  }
  return PM_OK;
}
#endif
