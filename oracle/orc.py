"""ctypes access to the CPU oracle (oracle/libref_cpu.so) and to the reference
itself (oracle/_ref/libref_*.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import importlib.util
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _pmct():
    spec = importlib.util.spec_from_file_location(
        "pm_ctypes", os.path.join(ROOT, "computational-fluid-dynamics_b200", "pm_ctypes.py"))
    import sys
    if "pm_ctypes" in sys.modules:
        return sys.modules["pm_ctypes"]
    m = importlib.util.module_from_spec(spec)
    sys.modules["pm_ctypes"] = m
    spec.loader.exec_module(m)
    return m


pmct = _pmct()
PmConfig, PmPpeResult, field_shape = pmct.PmConfig, pmct.PmPpeResult, pmct.field_shape

_orc = None


def orc_lib():
    global _orc
    if _orc is None:
        p = os.path.join(HERE, "libref_cpu.so")
        if not os.path.exists(p):
            raise RuntimeError(f"{p} not built (make -C oracle)")
        L = C.CDLL(p)
        vp = C.c_void_p
        L.orc_config_init.argtypes = [C.POINTER(PmConfig), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        L.orc_create.argtypes = [C.POINTER(PmConfig)]; L.orc_create.restype = vp
        L.orc_destroy.argtypes = [vp]
        L.orc_set_method.argtypes = [vp, C.c_int]
        L.orc_set_max_iters.argtypes = [vp, C.c_int]
        L.orc_set_omega.argtypes = [vp, C.c_double]
        L.orc_field.argtypes = [vp, C.c_int]; L.orc_field.restype = C.POINTER(C.c_double)
        L.orc_field_count.argtypes = [vp, C.c_int]; L.orc_field_count.restype = C.c_size_t
        L.orc_mask.argtypes = [vp]; L.orc_mask.restype = C.POINTER(C.c_uint8)
        L.orc_fill_random.argtypes = [vp, C.c_uint64]
        L.orc_fill_random_scaled.argtypes = [vp, C.c_uint64, C.c_double]
        L.orc_synth.argtypes = [C.c_uint64, C.c_int, C.c_uint64]; L.orc_synth.restype = C.c_double
        L.orc_apply_bc.argtypes = [vp, C.c_int]
        L.orc_predict.argtypes = [vp]
        L.orc_source.argtypes = [vp]; L.orc_source.restype = C.c_double
        L.orc_ppe_solve.argtypes = [vp, C.POINTER(PmPpeResult)]
        L.orc_correct.argtypes = [vp]
        L.orc_step.argtypes = [vp, C.c_int, C.POINTER(PmPpeResult)]
        L.orc_diagnostics.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_sweep_rows.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.orc_jacobi_rows.argtypes = [vp, C.POINTER(C.c_double), C.c_int, C.c_int]
        L.orc_pressure_ghosts.argtypes = [vp]
        L.orc_residual_rows.argtypes = [vp, C.c_int, C.c_int]; L.orc_residual_rows.restype = C.c_double
        _orc = L
    return _orc


def config_init(case_id, nx=0, ny=0, re=0.0, dt=0.0):
    cfg = PmConfig()
    st = orc_lib().orc_config_init(C.byref(cfg), case_id, nx, ny, re, dt)
    if st != 0:
        raise ValueError(f"orc_config_init status {st}")
    return cfg


class Oracle:
    """CPU oracle instance; fields are numpy views of the oracle's own storage."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.L = orc_lib()
        self.h = self.L.orc_create(C.byref(cfg))
        if not self.h:
            raise ValueError("orc_create failed")
        self.nx, self.ny = cfg.nx, cfg.ny

    def __del__(self):
        try:
            if self.h:
                self.L.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def field(self, fid):
        n = self.L.orc_field_count(self.h, fid)
        ptr = self.L.orc_field(self.h, fid)
        return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(field_shape(fid, self.nx, self.ny))

    def mask(self):
        return np.ctypeslib.as_array(self.L.orc_mask(self.h), shape=((self.ny + 2) * (self.nx + 2),)).reshape(self.ny + 2, self.nx + 2)

    def set_method(self, m): self.L.orc_set_method(self.h, m)
    def set_max_iters(self, k): self.L.orc_set_max_iters(self.h, k)
    def set_omega(self, w): self.L.orc_set_omega(self.h, w)
    def fill_random(self, seed, amplitude=1.0): self.L.orc_fill_random_scaled(self.h, seed, amplitude)
    def apply_bc(self, which=0): self.L.orc_apply_bc(self.h, which)
    def predict(self): self.L.orc_predict(self.h)
    def source(self): return self.L.orc_source(self.h)
    def correct(self): self.L.orc_correct(self.h)

    def ppe_solve(self):
        r = PmPpeResult()
        self.L.orc_ppe_solve(self.h, C.byref(r))
        return r

    def step(self, n=1):
        r = PmPpeResult()
        self.L.orc_step(self.h, n, C.byref(r))
        return r

    def diagnostics(self):
        a, b = C.c_double(), C.c_double()
        self.L.orc_diagnostics(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def sweep_rows(self, colour, ja, jb): self.L.orc_sweep_rows(self.h, colour, ja, jb)

    def jacobi_rows(self, src, ja, jb):
        s = np.ascontiguousarray(src, dtype=np.float64)
        self.L.orc_jacobi_rows(self.h, s.ctypes.data_as(C.POINTER(C.c_double)), ja, jb)

    def pressure_ghosts(self): self.L.orc_pressure_ghosts(self.h)
    def residual_rows(self, ja, jb): return self.L.orc_residual_rows(self.h, ja, jb)


def ref_available(name):
    return os.path.exists(os.path.join(HERE, "_ref", f"libref_{name}.so"))


class Reference:
    """The unmodified reference solver (oracle/_ref/libref_<name>.so), driven phase by phase."""

    def __init__(self, name):
        p = os.path.join(HERE, "_ref", f"libref_{name}.so")
        if not os.path.exists(p):
            raise FileNotFoundError(p)
        L = C.CDLL(p)
        vp = C.c_void_p
        L.ref_create.restype = vp
        for n in ("ref_destroy", "ref_predict", "ref_source", "ref_correct"):
            getattr(L, n).argtypes = [vp]
        for n in ("ref_nx", "ref_ny", "ref_total_steps", "ref_max_iters"):
            getattr(L, n).argtypes = [vp]; getattr(L, n).restype = C.c_int
        for n in ("ref_dt", "ref_omega", "ref_nu", "ref_dx", "ref_dy"):
            getattr(L, n).argtypes = [vp]; getattr(L, n).restype = C.c_double
        L.ref_field_count.argtypes = [vp, C.c_int]; L.ref_field_count.restype = C.c_size_t
        L.ref_get.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
        L.ref_set.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
        L.ref_get_mask.argtypes = [vp, C.POINTER(C.c_uint8)]
        L.ref_apply_bc.argtypes = [vp, C.c_int]
        L.ref_ppe.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.ref_step.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.ref_time_steps.argtypes = [vp, C.c_int]; L.ref_time_steps.restype = C.c_double
        self.L = L
        self.h = L.ref_create()
        if not self.h:
            raise RuntimeError("reference constructor failed")
        self.case_id = L.ref_case()
        self.nx, self.ny = L.ref_nx(self.h), L.ref_ny(self.h)

    def __del__(self):
        try:
            if self.h:
                self.L.ref_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def params(self):
        L, h = self.L, self.h
        return dict(nx=self.nx, ny=self.ny, dt=L.ref_dt(h), omega=L.ref_omega(h), nu=L.ref_nu(h),
                    dx=L.ref_dx(h), dy=L.ref_dy(h), total_steps=L.ref_total_steps(h), max_iters=L.ref_max_iters(h))

    def get(self, fid):
        out = np.empty(field_shape(fid, self.nx, self.ny), dtype=np.float64)
        assert self.L.ref_field_count(self.h, fid) == out.size
        self.L.ref_get(self.h, fid, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def set(self, fid, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        assert a.shape == field_shape(fid, self.nx, self.ny)
        self.L.ref_set(self.h, fid, a.ctypes.data_as(C.POINTER(C.c_double)))

    def mask(self):
        out = np.empty((self.ny + 2, self.nx + 2), dtype=np.uint8)
        self.L.ref_get_mask(self.h, out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def apply_bc(self, which=0): self.L.ref_apply_bc(self.h, which)
    def predict(self): self.L.ref_predict(self.h)
    def source(self): self.L.ref_source(self.h)
    def correct(self): self.L.ref_correct(self.h)

    def ppe(self):
        it, res = C.c_int(), C.c_double()
        self.L.ref_ppe(self.h, C.byref(it), C.byref(res))
        return it.value, res.value

    def step(self, n=1):
        it, res = C.c_int(), C.c_double()
        self.L.ref_step(self.h, n, C.byref(it), C.byref(res))
        return it.value, res.value

    def time_steps(self, n): return self.L.ref_time_steps(self.h, n)
