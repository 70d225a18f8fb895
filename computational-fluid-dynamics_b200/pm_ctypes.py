"""ctypes binding of include/pm.h (the C-ABI of libpm.so).

Host-side mirror used by tests/ and bench.py.  It binds exactly the symbols
include/pm.h declares and nothing else; there is no fallback: if libpm.so is
missing or a symbol is absent, import-time loading raises.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PM_LIB") or os.path.join(HERE, "lib", "libpm.so")  # PM_LIB: experimental builds (make variant)

PM_OK = 0
CASE_CAVITY, CASE_CHANNEL, CASE_STEP = 0, 1, 2
PPE_JACOBI, PPE_SOR_RB, PPE_SOR_LEX, PPE_SOR_CHEBY = 0, 1, 2, 3
F_U, F_V, F_P, F_USTAR, F_VSTAR, F_F = range(6)
PATH_AUTO, PATH_SIMPLE, PATH_TILED, PATH_PERSISTENT = 0, 1, 2, 3


class PmConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("case_id", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32),
        ("dx", C.c_double), ("dy", C.c_double), ("nu", C.c_double), ("dt", C.c_double),
        ("u_ref", C.c_double), ("rho", C.c_double), ("omega", C.c_double),
        ("tol_factor", C.c_double), ("abs_tol", C.c_double),
        ("max_iters", C.c_int32), ("ppe_method", C.c_int32), ("exact_arith", C.c_int32),
        ("sweeps_per_pass", C.c_int32), ("kernel_path", C.c_int32),
        ("step_i_location", C.c_int32), ("inlet_j_max", C.c_int32),
        ("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32), ("poll_chunk", C.c_int32),
        ("nccl_id", C.c_uint8 * 128),
        ("lx", C.c_double), ("ly", C.c_double), ("re", C.c_double), ("cfl", C.c_double), ("final_time", C.c_double),
        ("total_steps", C.c_int32), ("print_interval", C.c_int32), ("save_interval", C.c_int32), ("reserved_", C.c_int32),
    ]

    def copy(self):
        c = PmConfig()
        C.memmove(C.byref(c), C.byref(self), C.sizeof(PmConfig))
        return c


class PmPpeResult(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("hit_cap", C.c_int32), ("residual", C.c_double),
                ("tolerance", C.c_double), ("max_source", C.c_double)]


class PmTiming(C.Structure):
    _fields_ = [("ppe_ms", C.c_double), ("other_ms", C.c_double),
                ("kernel_launches", C.c_int64), ("ppe_passes", C.c_int64)]


def field_shape(field, nx, ny):
    """Reference array shapes (cavity-01.cpp:433-441)."""
    if field in (F_U, F_USTAR):
        return (ny + 2, nx + 1)
    if field in (F_V, F_VSTAR):
        return (ny + 1, nx + 2)
    return (ny + 2, nx + 2)


EXPORTS = [
    "pm_config_init", "pm_slab_range", "pm_cheby_omega", "pm_omega_mixed_bc", "pm_stream_plan", "pm_create", "pm_destroy", "pm_last_error", "pm_status_string",
    "pm_abi_version", "pm_nccl_unique_id", "pm_upload", "pm_download", "pm_slab_rows", "pm_upload_slab", "pm_download_slab", "pm_upload_mask", "pm_download_mask",
    "pm_fill_random", "pm_fill_random_scaled", "pm_fill_zero", "pm_apply_bc", "pm_predict", "pm_source", "pm_ppe_solve", "pm_correct",
    "pm_step", "pm_host_step_submit", "pm_host_step_run", "pm_host_step_drain", "pm_diagnostics", "pm_export_prepare", "pm_export_begin", "pm_export_wait", "pm_sync", "pm_get_timing", "pm_timer_start", "pm_timer_stop",
]

_lib = None


def lib():
    """Load libpm.so (built by __graft_entry__.build()); raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name in EXPORTS:
        getattr(L, name)  # AttributeError if the library does not export what pm.h declares
    vp, dp = C.c_void_p, C.POINTER(C.c_double)
    L.pm_config_init.argtypes = [C.POINTER(PmConfig), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
    L.pm_slab_range.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.pm_cheby_omega.argtypes = [C.c_double, C.c_int]; L.pm_cheby_omega.restype = C.c_double
    L.pm_omega_mixed_bc.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]; L.pm_omega_mixed_bc.restype = C.c_double
    L.pm_stream_plan.argtypes = [C.c_int] * 7 + [C.POINTER(C.c_int)]
    L.pm_export_prepare.argtypes = [vp]
    L.pm_export_begin.argtypes = [vp]
    L.pm_export_wait.argtypes = [vp, dp, dp, dp, dp, dp, C.c_size_t]
    L.pm_create.argtypes = [C.POINTER(PmConfig), C.POINTER(vp)]
    L.pm_destroy.argtypes = [vp]
    L.pm_last_error.argtypes = [vp]; L.pm_last_error.restype = C.c_char_p
    L.pm_status_string.argtypes = [C.c_int]; L.pm_status_string.restype = C.c_char_p
    L.pm_nccl_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    L.pm_upload.argtypes = [vp, C.c_int, dp, C.c_size_t]
    L.pm_download.argtypes = [vp, C.c_int, dp, C.c_size_t]
    L.pm_slab_rows.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.pm_upload_slab.argtypes = [vp, C.c_int, dp, C.c_size_t]
    L.pm_download_slab.argtypes = [vp, C.c_int, dp, C.c_size_t]
    L.pm_upload_mask.argtypes = [vp, C.POINTER(C.c_uint8), C.c_size_t]
    L.pm_download_mask.argtypes = [vp, C.POINTER(C.c_uint8), C.c_size_t]
    L.pm_fill_random.argtypes = [vp, C.c_uint64]
    L.pm_fill_random_scaled.argtypes = [vp, C.c_uint64, C.c_double]
    L.pm_fill_zero.argtypes = [vp]
    L.pm_apply_bc.argtypes = [vp, C.c_int]
    for n in ("pm_predict", "pm_source", "pm_correct", "pm_sync"):
        getattr(L, n).argtypes = [vp]
    L.pm_ppe_solve.argtypes = [vp, C.POINTER(PmPpeResult)]
    L.pm_step.argtypes = [vp, C.c_int, C.POINTER(PmPpeResult)]
    L.pm_host_step_submit.argtypes = [vp, dp, C.c_size_t, dp, C.c_size_t, dp, dp, dp, C.c_size_t]
    L.pm_host_step_run.argtypes = [vp, C.POINTER(PmPpeResult)]
    L.pm_host_step_drain.argtypes = [vp]
    L.pm_diagnostics.argtypes = [vp, dp, dp]
    L.pm_get_timing.argtypes = [vp, C.POINTER(PmTiming)]
    L.pm_timer_start.argtypes = [vp]
    L.pm_timer_stop.argtypes = [vp, dp]
    _lib = L
    return L


class PmError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"pm status {status}: {msg}")
        self.status = status


def config_init(case_id, nx=0, ny=0, re=0.0, dt=0.0):
    cfg = PmConfig()
    st = lib().pm_config_init(C.byref(cfg), case_id, nx, ny, re, dt)
    if st != PM_OK:
        raise PmError(st, lib().pm_status_string(st).decode())
    return cfg


class Solver:
    """Thin object wrapper; method names follow the reference's member functions."""

    def __init__(self, cfg):
        self.cfg = cfg
        self._h = C.c_void_p()
        st = lib().pm_create(C.byref(cfg), C.byref(self._h))
        if st != PM_OK:
            raise PmError(st, (lib().pm_last_error(None) or b"").decode())

    def _ck(self, st):
        if st != PM_OK:
            raise PmError(st, (lib().pm_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            lib().pm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, field, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        assert a.shape == field_shape(field, self.cfg.nx, self.cfg.ny), (a.shape, field)
        self._ck(lib().pm_upload(self._h, field, a.ctypes.data_as(C.POINTER(C.c_double)), a.size))

    def download(self, field, out=None):
        if out is None:
            out = np.empty(field_shape(field, self.cfg.nx, self.cfg.ny), dtype=np.float64)
        self._ck(lib().pm_download(self._h, field, out.ctypes.data_as(C.POINTER(C.c_double)), out.size))
        return out

    def upload_mask(self, m):
        a = np.ascontiguousarray(m, dtype=np.uint8)
        self._ck(lib().pm_upload_mask(self._h, a.ctypes.data_as(C.POINTER(C.c_uint8)), a.size))

    def download_mask(self):
        out = np.empty((self.cfg.ny + 2, self.cfg.nx + 2), dtype=np.uint8)
        self._ck(lib().pm_download_mask(self._h, out.ctypes.data_as(C.POINTER(C.c_uint8)), out.size))
        return out

    def fill_random(self, seed, amplitude=1.0):
        self._ck(lib().pm_fill_random_scaled(self._h, seed, amplitude))

    def fill_zero(self):
        self._ck(lib().pm_fill_zero(self._h))

    def apply_bc(self, which=0):
        self._ck(lib().pm_apply_bc(self._h, which))

    def predict(self):
        self._ck(lib().pm_predict(self._h))

    def source(self):
        self._ck(lib().pm_source(self._h))

    def ppe_solve(self):
        r = PmPpeResult()
        self._ck(lib().pm_ppe_solve(self._h, C.byref(r)))
        return r

    def correct(self):
        self._ck(lib().pm_correct(self._h))

    def step(self, n=1):
        r = PmPpeResult()
        self._ck(lib().pm_step(self._h, n, C.byref(r)))
        return r

    def diagnostics(self):
        a, b = C.c_double(), C.c_double()
        self._ck(lib().pm_diagnostics(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def export_begin(self):
        self._ck(lib().pm_export_begin(self._h))

    def export_wait(self, out=None):
        """u_center, v_center, magnitude, pressure, vorticity as (ny, nx) arrays (this rank's rows filled)."""
        nx, ny = self.cfg.nx, self.cfg.ny
        out = out if out is not None else [np.zeros((ny, nx)) for _ in range(5)]
        ptrs = [a.ctypes.data_as(C.POINTER(C.c_double)) for a in out]
        self._ck(lib().pm_export_wait(self._h, *ptrs, nx * ny))
        return out

    def sync(self):
        self._ck(lib().pm_sync(self._h))

    def slab_rows(self, field):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._ck(lib().pm_slab_rows(self._h, field, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def upload_slab_ptr(self, field, ptr, count):
        """This rank's rows from a raw host pointer (e.g. a pinned torch tensor's data_ptr())."""
        self._ck(lib().pm_upload_slab(self._h, field, C.cast(ptr, C.POINTER(C.c_double)), count))

    def download_slab_ptr(self, field, ptr, count):
        self._ck(lib().pm_download_slab(self._h, field, C.cast(ptr, C.POINTER(C.c_double)), count))

    def host_step_submit(self, u_in, v_in, u_out, v_out, p_out):
        """(ptr, count) pairs for the inputs; ptrs for the outputs, p with its count (see pm_host_step_submit)."""
        dp = C.POINTER(C.c_double)
        self._ck(lib().pm_host_step_submit(self._h, C.cast(u_in[0], dp), u_in[1], C.cast(v_in[0], dp), v_in[1],
                                           C.cast(u_out, dp), C.cast(v_out, dp), C.cast(p_out[0], dp), p_out[1]))

    def host_step_run(self):
        r = PmPpeResult()
        self._ck(lib().pm_host_step_run(self._h, C.byref(r)))
        return r

    def host_step_drain(self):
        self._ck(lib().pm_host_step_drain(self._h))

    def timer_start(self):
        self._ck(lib().pm_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double()
        self._ck(lib().pm_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def timing(self):
        t = PmTiming()
        self._ck(lib().pm_get_timing(self._h, C.byref(t)))
        return t
