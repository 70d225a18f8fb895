"""GPU parity tests (run on the B200 box): libpm.so through its C-ABI against the CPU oracle on the
same seeded inputs, against the golden vectors recorded from the reference, and through
size-independent properties at sizes the oracle cannot reach.

Bars (BASELINE.json north_star): exact_arith=1 -> 0 ulp for every field of every phase (Jacobi and
red-black share the oracle's ordering); production arithmetic (FMA, reciprocal multiply) -> 1e-12
relative per phase; red-black SOR run to the reference tolerance vs the reference's lexicographic
result -> 1e-6 relative L2 for u, v, p.
"""
import os

import numpy as np
import pytest

from conftest import bits_equal, load_golden, max_ulp, rel_l2

pytestmark = pytest.mark.gpu

JAC, RB, LEX = 0, 1, 2
CASES = [(0, 48, 48), (0, 63, 63), (1, 93, 31), (1, 64, 40), (2, 64, 16), (2, 256, 32)]


def make_cfg(pm, case_id, nx, ny, method, exact, max_iters, omega=None, path=1):
    if case_id == 2 and path == 2 and os.environ.get("PM_LIB", "").endswith("cs4.so"):
        pytest.skip("the cluster build keeps the obstacle mask on the general path (masked tiles take no part in a cluster's exchange)")
    cfg = pm.config_init(case_id, nx, ny)
    if case_id == 2 and (nx, ny) != (256, 32):
        cfg.step_i_location, cfg.inlet_j_max = nx // 4, ny // 2
    cfg.ppe_method = method
    cfg.exact_arith = exact
    cfg.max_iters = max_iters
    cfg.kernel_path = path
    if omega is not None:
        cfg.omega = omega
    return cfg


def assert_fields_equal(S, O, fids, what):
    for fid in fids:
        a, b = S.download(fid), O.field(fid)
        assert bits_equal(a, b), f"{what}: field {fid} differs, max ulp {max_ulp(a, b)}, max abs {np.abs(a - b).max():.3e}"


def assert_fields_close(S, O, fids, tol, what):
    for fid in fids:
        a, b = S.download(fid), O.field(fid)
        scale = max(1.0, np.abs(b).max())
        assert np.abs(a - b).max() <= tol * scale, f"{what}: field {fid} off by {np.abs(a - b).max():.3e}"


@pytest.mark.parametrize("case_id,nx,ny", CASES)
def test_synthetic_state_matches_oracle(pm, orc, case_id, nx, ny):
    cfg = make_cfg(pm, case_id, nx, ny, JAC, 1, 10)
    S, O = pm.Solver(cfg), orc.Oracle(cfg)
    S.fill_random(42); O.fill_random(42)
    assert_fields_equal(S, O, range(6), "fill_random")
    if case_id == 2:
        assert np.array_equal(S.download_mask(), O.mask())


def test_upload_download_round_trip(pm):
    cfg = make_cfg(pm, 1, 37, 19, JAC, 1, 10)
    S = pm.Solver(cfg)
    rng = np.random.default_rng(0)
    for fid in range(6):
        a = rng.standard_normal(pm.field_shape(fid, 37, 19))
        S.upload(fid, a)
        assert bits_equal(S.download(fid), a)


@pytest.mark.parametrize("method,omega", [(JAC, 1.0), (JAC, 0.8), (RB, None), (LEX, None)])
@pytest.mark.parametrize("case_id,nx,ny", CASES)
def test_every_phase_bit_exact(pm, orc, case_id, nx, ny, method, omega):
    """Seeded random u, v, p; each phase compared right after it runs; K = 25 sweeps."""
    cfg = make_cfg(pm, case_id, nx, ny, method, 1, 25, omega)
    S, O = pm.Solver(cfg), orc.Oracle(cfg)
    S.fill_random(7); O.fill_random(7)
    S.apply_bc(0); O.apply_bc(0)
    assert_fields_equal(S, O, (0, 1), "apply_bc(u,v)")
    S.predict(); O.predict()
    assert_fields_equal(S, O, (3, 4), "predict")
    if case_id != 0:
        S.apply_bc(1); O.apply_bc(1)
        assert_fields_equal(S, O, (3, 4), "apply_bc(u*,v*)")
    S.source(); O.source()
    assert_fields_equal(S, O, (5,), "source")
    rs, ro = S.ppe_solve(), O.ppe_solve()
    assert (rs.iterations, rs.hit_cap) == (ro.iterations, ro.hit_cap)
    assert rs.tolerance == ro.tolerance and rs.max_source == ro.max_source
    assert rs.residual == ro.residual, (rs.residual, ro.residual)
    assert_fields_equal(S, O, (2,), "ppe")
    S.correct(); O.correct()
    assert_fields_equal(S, O, (0, 1), "correct")
    if case_id != 0:
        S.apply_bc(0); O.apply_bc(0)
        assert_fields_equal(S, O, (0, 1), "apply_bc after correct")
    ds, do = S.diagnostics(), O.diagnostics()
    assert ds[0] == do[0]
    assert abs(ds[1] - do[1]) <= 1e-12 * max(1.0, abs(do[1]))


@pytest.mark.parametrize("case_id,nx,ny,method", [(0, 96, 80, RB), (1, 93, 31, RB), (2, 64, 16, RB), (0, 300, 200, JAC)])
def test_streamed_host_steps_equal_plain_steps(pm, case_id, nx, ny, method):
    """pm_host_step_submit/run/drain (uploads, kernels and downloads of neighbouring steps overlapped through
    rotating plane sets) must return exactly what upload -> pm_step -> download returns, step by step."""
    rng = np.random.default_rng(5)
    cfg = make_cfg(pm, case_id, nx, ny, method, 1, 40, 0.9 if method == JAC else None, path=0)
    nsteps = 5
    ins = [(rng.uniform(-1, 1, pm.field_shape(0, nx, ny)), rng.uniform(-1, 1, pm.field_shape(1, nx, ny))) for _ in range(nsteps)]
    A = pm.Solver(cfg)
    want = []
    for u, v in ins:
        A.upload(0, u); A.upload(1, v)
        r = A.step(1)
        want.append((r.iterations, r.residual, A.download(0), A.download(1), A.download(2)))
    A.close()
    B = pm.Solver(cfg)
    outs = [tuple(np.full(pm.field_shape(f, nx, ny), np.nan) for f in (0, 1, 2)) for _ in range(nsteps)]
    res = []
    def submit(n):
        (u, v), (uo, vo, po) = ins[n], outs[n]
        B.host_step_submit((u.ctypes.data, u.size), (v.ctypes.data, v.size), uo.ctypes.data, vo.ctypes.data, (po.ctypes.data, po.size))
    submit(0)
    for n in range(nsteps):
        if n + 1 < nsteps:
            submit(n + 1)
        res.append(B.host_step_run())
    B.host_step_drain()
    for n in range(nsteps):
        assert (res[n].iterations, res[n].residual) == want[n][:2], f"step {n}"
        for q in range(3):
            assert bits_equal(outs[n][q], want[n][2 + q]), f"step {n} field {q}"
    # the handle's own state is that of the last step
    assert bits_equal(B.download(0), want[-1][2]) and bits_equal(B.download(2), want[-1][4])
    B.close()


@pytest.mark.parametrize("case_id,nx,ny", CASES)
def test_production_arithmetic_close_to_oracle(pm, orc, case_id, nx, ny):
    """exact_arith=0 (FMA contraction, reciprocal multiply, tree-summed mean): 1e-12 relative after one step."""
    cfg = make_cfg(pm, case_id, nx, ny, RB, 0, 25)
    S, O = pm.Solver(cfg), orc.Oracle(cfg)
    S.fill_random(11); O.fill_random(11)
    if case_id != 0:
        S.apply_bc(0); O.apply_bc(0)
    rs, ro = S.step(1), O.step(1)
    assert rs.iterations == ro.iterations
    # the random field makes |f| ~ 1e5..1e7, p ~ |f| h^2: compare relative to each field's own scale
    for fid in (0, 1, 2, 3, 4, 5):
        a, b = S.download(fid), O.field(fid)
        assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(b).max()), f"field {fid}: {np.abs(a - b).max():.3e} vs scale {np.abs(b).max():.3e}"


@pytest.mark.parametrize("case_id,nx,ny,steps", [(0, 32, 32, 6), (1, 93, 31, 3), (2, 64, 16, 3)])
def test_whole_steps_and_stopping_rule_bit_exact(pm, orc, case_id, nx, ny, steps):
    """From rest, run to the reference tolerance: same iteration counts, residuals and fields as the
    oracle's red-black restatement (the device-side stop flag must end the loop at the same k)."""
    cfg = make_cfg(pm, case_id, nx, ny, RB, 1, 10000)
    S, O = pm.Solver(cfg), orc.Oracle(cfg)
    S.apply_bc(0); O.apply_bc(0)
    for n in range(steps):
        rs, ro = S.step(1), O.step(1)
        assert (rs.iterations, rs.residual) == (ro.iterations, ro.residual), f"step {n}"
    assert_fields_equal(S, O, range(6), "whole steps")


def golden_cfg(pm, g, method, exact, max_iters=None):
    """pm_config carrying exactly the parameters the reference build of a golden fixture ran with."""
    cfg = pm.config_init(int(g["case_id"]), int(g["prm_nx"]), int(g["prm_ny"]))
    cfg.dt, cfg.omega, cfg.nu = float(g["prm_dt"]), float(g["prm_omega"]), float(g["prm_nu"])
    cfg.dx, cfg.dy = float(g["prm_dx"]), float(g["prm_dy"])
    cfg.max_iters = int(g["prm_max_iters"]) if max_iters is None else max_iters
    cfg.ppe_method, cfg.exact_arith = method, exact
    return cfg


# Production red-black (FMA arithmetic) against the reference's OWN fields (golden, lexicographic SOR, recorded from
# the unmodified reference) on its default programs and on BASELINE configs[0..2].  north_star's bar is 1e-6 relative
# L2.  It holds wherever the reference converges within its 10 000-iteration cap.  Where the reference itself stops
# at the cap unconverged (configs[1]: residual 47.8 against a tolerance of 0.0168 after step 1; configs[2]: 2.1
# against 6e-4) neither ordering has solved the Poisson problem and the two capped iterates differ by what is left
# of the error; the bounds below are the measured gaps (CPU oracle red-black vs golden: see the comments) times ~4.
# v is measured against the velocity scale |(u, v)|: in the channel's first step v is ~1e-10 of u.
#   name, steps, bound u, bound v / |(u,v)|, bound p
REFERENCE_GOLDEN_RUNS = [
    ("cavity_default", 4, 1e-6, 1e-6, 1e-6),      # measured 1.4e-10, 5.2e-8, 4.6e-8
    ("channel_default", 3, 1e-6, 1e-6, 1e-6),     # measured 1.5e-8, 4.8e-7 (v itself), 5.4e-7
    ("cavity_cfg0", 2, 1e-6, 1e-6, 1e-6),         # BASELINE configs[0]: measured 9.9e-10, 2.1e-7, 5.9e-8
    ("channel_cfg1", 1, 1e-3, 1e-3, 1e-3),        # BASELINE configs[1], reference capped unconverged: measured 2.1e-4, 2.3e-4, 2.5e-4
    ("step_default", 2, 1e-3, 1e-3, 5e-2),        # BASELINE configs[2], reference capped unconverged: measured 2.3e-4, 2.5e-4, 1.3e-2
    ("step_default_20", 20, 1e-5, 1e-5, 2e-5),    # same run 18 steps on (17 of 20 capped): the warm starts pull both together: 1.7e-6, 1.7e-6, 3.4e-6
]


@pytest.mark.parametrize("name,steps,bu,bv,bp", REFERENCE_GOLDEN_RUNS)
def test_red_black_to_tolerance_matches_reference_golden(pm, name, steps, bu, bv, bp):
    g = load_golden(name)
    assert int(g["steps"]) == steps
    cfg = golden_cfg(pm, g, RB, 0)
    chk = pm.config_init(int(g["case_id"]), int(g["prm_nx"]), int(g["prm_ny"]))
    if name in ("cavity_default", "channel_default", "step_default", "step_default_20"):
        assert chk.dt == cfg.dt and chk.omega == cfg.omega  # pm_config_init derives the reference's own constants
    S = pm.Solver(cfg)
    S.apply_bc(0)
    S.step(steps)
    u, v, p = (S.download(f) for f in (0, 1, 2))
    gu, gv, gp = (g[f"f{f}"] for f in (0, 1, 2))
    assert np.isfinite(u).all() and np.isfinite(v).all() and np.isfinite(p).all()
    vel = np.sqrt(np.linalg.norm(gu) ** 2 + np.linalg.norm(gv) ** 2)
    eu, ev, ep = rel_l2(u, gu), np.linalg.norm(v - gv) / vel, rel_l2(p, gp)
    assert eu < bu and ev < bv and ep < bp, (eu, ev, ep)


@pytest.mark.parametrize("name,steps", [("channel_cfg1", 1), ("step_default", 2)])
def test_red_black_converged_matches_reference_algorithm_on_capped_configs(pm, orc, name, steps):
    """BASELINE configs[1] and [2] with the cap lifted and the tolerance tightened on BOTH sides (tolerance_factor
    1e-10, the reference's 1e-7 leaves a 1e-5 relative gap in p between any two orderings that merely satisfy it):
    production red-black on the GPU against the reference's algorithm (oracle, lexicographic SOR, itself pinned bit for
    bit to the reference at the cap) -> 1e-6 relative L2 (measured on the CPU oracle: 6e-11 / 5e-11 abs / 4e-11 and
    9e-11 / 2e-10 / 4e-9).  ~40 000 sweeps per step."""
    g = load_golden(name)
    cfg = golden_cfg(pm, g, RB, 0, 2_000_000)
    cfg.tol_factor = 1e-10
    ocfg = cfg.copy()
    ocfg.ppe_method = LEX
    S, O = pm.Solver(cfg), orc.Oracle(ocfg)
    S.apply_bc(0); O.apply_bc(0)
    for n in range(steps):
        rs, ro = S.step(1), O.step(1)
        assert not rs.hit_cap and not ro.hit_cap and rs.residual <= rs.tolerance
    u, v, p = (S.download(f) for f in (0, 1, 2))
    ou, ov, op = (O.field(f) for f in (0, 1, 2))
    vel = np.sqrt(np.linalg.norm(ou) ** 2 + np.linalg.norm(ov) ** 2)
    assert rel_l2(u, ou) < 1e-6 and np.linalg.norm(v - ov) / vel < 1e-6 and rel_l2(p, op) < 1e-6


@pytest.mark.parametrize("name", ["cavity_k50_32", "channel_k50", "step_k50"])
def test_fixed_iteration_golden_is_reached_by_oracle_orderings(pm, orc, name):
    """Same inputs, same iteration count (50-sweep cap, 5 steps): GPU == oracle bitwise for the shared
    orderings; the golden (reference ordering) differs from them only by the ordering itself."""
    g = load_golden(name)
    case_id = int(g["case_id"])
    for method in (JAC, RB):
        cfg = make_cfg(pm, case_id, int(g["prm_nx"]), int(g["prm_ny"]), method, 1, 50, 1.0 if method == JAC else None)
        S, O = pm.Solver(cfg), orc.Oracle(cfg)
        S.apply_bc(0); O.apply_bc(0)
        rs, ro = S.step(5), O.step(5)
        assert rs.iterations == ro.iterations == 50
        assert_fields_equal(S, O, range(6), f"{name} method {method}")


@pytest.mark.parametrize("nx,ny", [(1024, 1024), (2048, 512)])
def test_large_grid_jacobi_bit_exact(pm, orc, nx, ny):
    cfg = make_cfg(pm, 0, nx, ny, JAC, 1, 3, 1.0)
    S, O = pm.Solver(cfg), orc.Oracle(cfg)
    S.fill_random(3); O.fill_random(3)
    rs, ro = S.step(1), O.step(1)
    assert rs.residual == ro.residual
    assert_fields_equal(S, O, range(6), "large jacobi")


TILED_CASES = [
    # case, nx, ny, method, T
    (0, 48, 48, RB, 1), (0, 48, 48, RB, 2), (0, 250, 131, RB, 2), (0, 250, 131, RB, 3), (0, 117, 61, RB, 3), (0, 250, 131, RB, 4), (0, 117, 61, RB, 4),
    (1, 93, 31, RB, 1), (1, 93, 31, RB, 2), (1, 300, 70, RB, 2), (1, 300, 70, RB, 3), (1, 121, 77, RB, 3), (1, 300, 70, RB, 4),
    (0, 48, 48, JAC, 1), (0, 250, 131, JAC, 2), (0, 250, 131, JAC, 4), (1, 93, 31, JAC, 1), (1, 300, 70, JAC, 2), (1, 300, 70, JAC, 4),
    # obstacle mask (backwards step): tiles with fluid cells only, with solid cells only, and with both
    (2, 64, 16, RB, 1), (2, 256, 32, RB, 2), (2, 300, 70, RB, 3), (2, 300, 70, RB, 4), (2, 250, 131, RB, 4), (2, 117, 61, RB, 3), (2, 640, 200, RB, 4),
]


@pytest.mark.parametrize("case_id,nx,ny,method,T", TILED_CASES)
@pytest.mark.parametrize("K", [7, 24])
def test_tiled_ppe_bit_exact(pm, orc, case_id, nx, ny, method, T, K):
    """The TMA-tiled, temporally blocked pressure solve (T sweeps per pass, K not a multiple of T)
    against the oracle: iterate, residual and iteration count, 0 ulp."""
    cfg = make_cfg(pm, case_id, nx, ny, method, 1, K, 0.9 if method == JAC else None, path=2)
    cfg.sweeps_per_pass = T
    S, O = pm.Solver(cfg), orc.Oracle(cfg)
    S.fill_random(17); O.fill_random(17)
    rs, ro = S.ppe_solve(), O.ppe_solve()
    assert (rs.iterations, rs.hit_cap) == (ro.iterations, ro.hit_cap)
    assert rs.residual == ro.residual, (rs.residual, ro.residual)
    assert_fields_equal(S, O, (2,), "tiled ppe")


@pytest.mark.parametrize("case_id,nx,ny,method,T", [(0, 400, 300, RB, 3), (0, 400, 300, RB, 2), (0, 400, 300, RB, 4), (1, 384, 200, RB, 3), (1, 384, 200, RB, 4), (0, 400, 300, JAC, 2),
                                                   (1, 384, 200, JAC, 4), (2, 640, 200, RB, 4), (2, 384, 200, RB, 3)])
def test_tiled_production_arithmetic_close_to_oracle(pm, orc, case_id, nx, ny, method, T):
    """Production arithmetic on the tiled path, grids with interior tiles (the residual-form relaxation with
    summed neighbours): iterate within 1e-12 of the oracle's relative to the field's scale, residual norm to
    1e-9 relative, same iteration count."""
    cfg = make_cfg(pm, case_id, nx, ny, method, 0, 23, 0.9 if method == JAC else None, path=2)
    cfg.sweeps_per_pass = T
    S, O = pm.Solver(cfg), orc.Oracle(cfg)
    S.fill_random(29); O.fill_random(29)
    rs, ro = S.ppe_solve(), O.ppe_solve()
    assert (rs.iterations, rs.hit_cap) == (ro.iterations, ro.hit_cap)
    assert abs(rs.residual - ro.residual) <= 1e-9 * ro.residual, (rs.residual, ro.residual)
    a, b = S.download(2), O.field(2)
    assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(b).max()), f"p off by {np.abs(a - b).max():.3e} vs scale {np.abs(b).max():.3e}"


@pytest.mark.parametrize("case_id,nx,ny,T,steps", [(0, 40, 40, 2, 4), (0, 40, 40, 3, 4), (0, 40, 40, 4, 4), (1, 93, 31, 2, 2), (1, 93, 31, 3, 2), (1, 93, 31, 4, 2),
                                                   (2, 64, 16, 2, 2), (2, 64, 16, 4, 2), (2, 96, 24, 3, 2)])
def test_tiled_stopping_rule_bit_exact(pm, orc, case_id, nx, ny, T, steps):
    """Run to the reference tolerance through the tiled path: the device-side loop test plus the
    partial replay pass must land on exactly the iterate the oracle stops at."""
    cfg = make_cfg(pm, case_id, nx, ny, RB, 1, 10000, path=2)
    cfg.sweeps_per_pass = T
    S, O = pm.Solver(cfg), orc.Oracle(cfg)
    S.apply_bc(0); O.apply_bc(0)
    for n in range(steps):
        rs, ro = S.step(1), O.step(1)
        assert (rs.iterations, rs.residual) == (ro.iterations, ro.residual), f"step {n}"
    assert_fields_equal(S, O, range(6), "tiled whole steps")


def test_cluster_build_bit_exact():
    """lib/libpm_cs4.so: the same sources with the red-black tiles stacked into thread-block clusters of four CTAs that
    push their edge rows into each other's shared memory after every colour half-sweep (st.async + mbarrier; DSMEM).
    The tiled parity tests above (0 ulp against the oracle, stopping rule, production arithmetic, ragged grids) are run
    again in a child process bound to that library."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "computational-fluid-dynamics_b200", "lib", "libpm_cs4.so")
    assert os.path.exists(lib), "lib/libpm_cs4.so not built (make -C computational-fluid-dynamics_b200)"
    env = dict(os.environ, PM_LIB=lib)
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k",
                          "tiled_ppe_bit_exact or tiled_production or tiled_stopping or ragged or auto_path"],
                         env=env, capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]


def blob_mask(nx, ny, seed):
    """An irregular obstacle field: rectangles, single solid cells, one-cell gaps, solids against every wall."""
    rng = np.random.default_rng(seed)
    m = np.zeros((ny + 2, nx + 2), dtype=np.uint8)
    m[1:ny + 1, 1:nx + 1] = 1
    for _ in range(14):
        h, w = int(rng.integers(1, max(2, ny // 4))), int(rng.integers(1, max(2, nx // 5)))
        j, i = int(rng.integers(1, ny - h + 2)), int(rng.integers(1, nx - w + 2))
        m[j:j + h, i:i + w] = 0
    for _ in range(40):
        m[int(rng.integers(1, ny + 1)), int(rng.integers(1, nx + 1))] = 0
    m[1:ny + 1, 1][::3] = 0  # solid cells on the inlet wall, the outlet wall and both plates
    m[ny, 1:nx + 1][::5] = 0
    m[1, 1:nx + 1][::7] = 0
    m[1:ny + 1, nx][::4] = 0
    return m


@pytest.mark.parametrize("nx,ny,T,seed", [(300, 131, 4, 1), (259, 97, 3, 2), (130, 60, 2, 3), (640, 200, 4, 4)])
@pytest.mark.parametrize("path", [1, 2])
def test_arbitrary_obstacle_mask_bit_exact(pm, orc, nx, ny, T, seed, path):
    """pm_upload_mask with an irregular obstacle field (the reference only ever builds one rectangle, but its loops are
    written for any is_fluid): whole steps through the general and the tiled path against the oracle, 0 ulp."""
    cfg = make_cfg(pm, 2, nx, ny, RB, 1, 21, path=path)
    cfg.sweeps_per_pass = T if path == 2 else 0
    m = blob_mask(nx, ny, seed)
    S, O = pm.Solver(cfg), orc.Oracle(cfg)
    S.upload_mask(m)
    O.mask()[:] = m
    S.fill_random(9, 2.0 ** -3); O.fill_random(9, 2.0 ** -3)
    S.apply_bc(0); O.apply_bc(0)
    for n in range(2):
        rs, ro = S.step(1), O.step(1)
        assert (rs.iterations, rs.residual) == (ro.iterations, ro.residual), f"step {n}"
    assert_fields_equal(S, O, range(6), "obstacle field")


@pytest.mark.parametrize("exact", [1, 0])
def test_large_step_grid_tiled_equals_general_path(pm, exact):
    """The obstacle mask at a size with thousands of tiles (2048 x 512: fluid-only, solid-only and mixed tiles, 8 mixed
    tile rows): tiled path == general path, bit for bit with exact arithmetic, 1e-12 with production arithmetic."""
    out = []
    for path, T in ((1, 0), (2, 4), (2, 3)):
        cfg = make_cfg(pm, 2, 2048, 512, RB, exact, 13, path=path)
        cfg.sweeps_per_pass = T
        S = pm.Solver(cfg)
        S.fill_random(5, 2.0 ** -6)
        S.apply_bc(0)
        r = S.step(2)
        assert r.iterations == 13 and r.hit_cap == 1
        out.append((r.residual, S.download(2), S.download(0), S.download(1)))
        S.close()
    for o in out[1:]:
        if exact:
            assert o[0] == out[0][0]
            assert all(bits_equal(a, b) for a, b in zip(o[1:], out[0][1:]))
        else:
            assert abs(o[0] - out[0][0]) <= 1e-9 * out[0][0]
            for a, b in zip(o[1:], out[0][1:]):
                assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(b).max())


@pytest.mark.parametrize("exact", [1, 0])
def test_large_grid_8192_tiled_equals_general_path(pm, exact):
    """BASELINE configs[3] size, where the oracle is too slow: with exact arithmetic the tiled path must give
    the same bits as the general path (itself pinned to the oracle at smaller sizes); with production
    arithmetic the two agree to 1e-12 relative."""
    n = 8192
    out = []
    for path, T in ((1, 0), (2, 2), (2, 3), (2, 4)):
        cfg = make_cfg(pm, 0, n, n, RB, exact, 9, path=path)
        cfg.sweeps_per_pass = T
        cfg.tol_factor = 1e-12  # at h = 1/8192 max|f| >= 2 nu U / h^3 = 1.1e9 and the reference's 1e-9 would skip the loop
        S = pm.Solver(cfg)
        S.fill_random(5, 2.0 ** -10)
        r = S.step(1)
        assert r.iterations == 9 and r.hit_cap == 1
        out.append((r.residual, S.download(2), S.download(0)))
        S.close()
    for o in out[1:]:
        if exact:
            assert o[0] == out[0][0]
            assert bits_equal(o[1], out[0][1]) and bits_equal(o[2], out[0][2])
        else:
            # production arithmetic: interior tiles relax in residual form (p += c*r), the general path
            # evaluates the reference tree with FMAs; both are roundings of the same real-number update
            assert abs(o[0] - out[0][0]) <= 1e-9 * out[0][0]
            for a, b in ((o[1], out[0][1]), (o[2], out[0][2])):
                assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max()
    assert np.isfinite(out[0][1]).all() and np.abs(out[0][1]).max() > 0


@pytest.mark.parametrize("exact", [1, 0])
@pytest.mark.parametrize("case_id,nx,ny", CASES + [(0, 300, 200), (1, 260, 180), (2, 300, 90)])
def test_cheby_sor_matches_oracle(pm, orc, case_id, nx, ny, exact):
    """PM_PPE_SOR_CHEBY (relaxation factor moving with every colour half-sweep, include/pm.h) towards pm_omega_mixed_bc:
    0 ulp against the oracle's restatement with exact arithmetic, whole steps; 1e-12 with production arithmetic."""
    cfg = make_cfg(pm, case_id, nx, ny, 3, exact, 37)
    cfg.omega = pm.lib().pm_omega_mixed_bc(case_id, nx, ny, cfg.dx, cfg.dy)
    S, O = pm.Solver(cfg), orc.Oracle(cfg)
    S.fill_random(13); O.fill_random(13)
    S.apply_bc(0); O.apply_bc(0)
    rs, ro = S.step(2), O.step(2)
    assert rs.iterations == ro.iterations == 37
    if exact:
        assert rs.residual == ro.residual
        assert_fields_equal(S, O, range(6), "cheby")
    else:
        assert abs(rs.residual - ro.residual) <= 1e-9 * ro.residual
        assert_fields_close(S, O, range(6), 1e-12, "cheby, production arithmetic")


@pytest.mark.parametrize("case_id,nx,ny", [(0, 64, 64), (1, 96, 32), (2, 128, 32)])
def test_omega_mixed_converges_on_every_path_with_the_oracles_count(pm, orc, case_id, nx, ny):
    """Red-black SOR with the mixed-BC factor run to the reference's tolerance: the persistent small-grid solve (AUTO), the
    general path and the Chebyshev schedule stop on the oracle's iterate (exact arithmetic), several times earlier than with
    the reference's factor."""
    w = pm.lib().pm_omega_mixed_bc(case_id, nx, ny, pm.config_init(case_id, nx, ny).dx, pm.config_init(case_id, nx, ny).dy)
    base = make_cfg(pm, case_id, nx, ny, RB, 1, 30000, path=0)
    Sb = pm.Solver(base); Sb.fill_random(3, 2.0 ** -6); Sb.apply_bc(0)
    rb = Sb.step(1)
    Sb.close()
    for method, path in ((RB, 0), (RB, 1), (3, 0)):
        cfg = make_cfg(pm, case_id, nx, ny, method, 1, 30000, w, path=path)
        S, O = pm.Solver(cfg), orc.Oracle(cfg)
        S.fill_random(3, 2.0 ** -6); O.fill_random(3, 2.0 ** -6)
        S.apply_bc(0); O.apply_bc(0)
        rs, ro = S.step(1), O.step(1)
        assert (rs.iterations, rs.residual, rs.hit_cap) == (ro.iterations, ro.residual, 0)
        assert 2 * rs.iterations <= rb.iterations
        assert_fields_equal(S, O, (0, 1, 2), f"omega mixed, method {method} path {path}")
        S.close()


def _stream_pair(pm, monkeypatch, cfg, seed, steps=1, prepare=None):
    """The same production red-black problem with the streaming pass (default) and with the tiled kernel alone."""
    out = []
    for no_stream in (True, False):
        if no_stream:
            monkeypatch.setenv("PM_NO_STREAM", "1")
        else:
            monkeypatch.delenv("PM_NO_STREAM", raising=False)
        S = pm.Solver(cfg)
        S.fill_random(seed)
        if cfg.case_id != 0:
            S.apply_bc(0)
        if prepare:
            prepare(S)
        r = S.step(steps)
        out.append((r, [S.download(f) for f in range(6)], S.timing().kernel_launches))
        S.close()
    return out


@pytest.mark.parametrize("case_id,nx,ny,K,steps", [(0, 1024, 1024, 23, 2), (0, 2000, 1500, 40, 1), (1, 2048, 1024, 24, 2), (1, 1500, 700, 9, 1), (0, 1400, 420, 100, 1)])
def test_streaming_pass_is_bit_identical_to_the_tiled_kernel(pm, monkeypatch, case_id, nx, ny, K, steps):
    """k_ppe_stream (interior tiles of production red-black solves, pm_kernels_stream.cuh) against k_ppe_tiled over the whole
    grid: every field, the iteration count and the residual, bit for bit -- also where K is not a multiple of the four
    sweeps of a pass (the last pass and the residual-only pass run on the tiled kernel alone)."""
    if os.environ.get("PM_LIB", "").endswith("cs4.so"):
        pytest.skip("the cluster build does not stream")
    cfg = make_cfg(pm, case_id, nx, ny, RB, 0, K, path=2)
    cfg.tol_factor = 1e-13  # large cavities: keep the reference's loop-entry rule (1.0 > tolerance) from skipping the solve
    (ra, fa, la), (rb, fb, lb) = _stream_pair(pm, monkeypatch, cfg, 31, steps)
    assert lb > la, "the streaming pass did not run (two launches per full pass instead of one)"
    assert (ra.iterations, ra.residual) == (rb.iterations, rb.residual) and ra.iterations == K
    for fid, (a, b) in enumerate(zip(fa, fb)):
        assert np.isfinite(b).all()
        assert bits_equal(a, b), f"field {fid}: max abs {np.abs(a - b).max():.3e}"


def test_streaming_pass_stops_on_the_same_iterate(pm, monkeypatch):
    """A solve that meets its tolerance in the middle of a pass: loop test on the device, replay of the partial pass."""
    if os.environ.get("PM_LIB", "").endswith("cs4.so"):
        pytest.skip("the cluster build does not stream")
    cfg = make_cfg(pm, 1, 1300, 420, RB, 0, 57, path=2)
    S = pm.Solver(cfg)
    S.fill_random(8); S.apply_bc(0)
    r0 = S.step(1)
    S.close()
    assert r0.iterations == 57
    cfg.max_iters = 10000
    cfg.tol_factor, cfg.abs_tol = 0.0, r0.residual * (1.0 + 1e-9)
    (ra, fa, _), (rb, fb, _) = _stream_pair(pm, monkeypatch, cfg, 8)
    assert 1 <= rb.iterations <= 57 and not rb.hit_cap
    assert (ra.iterations, ra.residual) == (rb.iterations, rb.residual)
    for a, b in zip(fa, fb):
        assert bits_equal(a, b)


@pytest.mark.parametrize("name", ["cavity_default", "channel_default", "step_default", "cavity_k50_32", "channel_k50", "step_k50",
                                  "cavity_cfg0", "channel_cfg1"])
def test_sor_lex_reproduces_the_reference_bit_for_bit(pm, name):
    """The reference's own ordering (wavefront-parallel lexicographic SOR) with exact arithmetic against the
    golden vectors recorded from the UNMODIFIED reference: same iteration counts, same residuals, and every
    field (u, v, p, u*, v*, f) identical to the last bit, for the three default programs, the two small
    BASELINE configs and the capped variants."""
    g = load_golden(name)
    case_id = int(g["case_id"])
    cfg = pm.config_init(case_id, int(g["prm_nx"]), int(g["prm_ny"]))
    cfg.dt, cfg.omega, cfg.nu = float(g["prm_dt"]), float(g["prm_omega"]), float(g["prm_nu"])
    cfg.dx, cfg.dy, cfg.max_iters = float(g["prm_dx"]), float(g["prm_dy"]), int(g["prm_max_iters"])
    cfg.ppe_method, cfg.exact_arith = LEX, 1
    S = pm.Solver(cfg)
    S.apply_bc(0)
    for n in range(int(g["steps"])):
        r = S.step(1)
        assert r.iterations == int(g["iters"][n]), f"step {n}"
        assert r.residual == float(g["res"][n]), f"step {n}"
    for fid in range(6):
        a, b = S.download(fid), g[f"f{fid}"]
        assert bits_equal(a, b), f"{name}: field {fid} max ulp {max_ulp(a, b)}"


def test_sor_lex_reports_unsupported_on_multirank(pm):
    cfg = make_cfg(pm, 0, 32, 32, LEX, 1, 10)
    cfg.nranks, cfg.rank = 2, 0
    with pytest.raises(pm.PmError) as e:
        pm.Solver(cfg)
    assert e.value.status == 5


def test_multi_gpu_slabs_match_single_gpu():
    """N > 1 on real GPUs (skipped on a single-GPU box; the CPU/gloo replay of the same schedule is
    tests/test_slab_gloo.py): torchrun tests/mgpu_check.py, slabs vs single GPU, bit for bit."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 4)}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tests", "mgpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]


# Ghia, Ghia & Shin (1982), Re = 100: u on the vertical centreline.
GHIA_Y = [1.0, 0.9766, 0.9688, 0.9609, 0.9531, 0.8516, 0.7344, 0.6172, 0.5, 0.4531, 0.2813, 0.1719, 0.1016, 0.0703, 0.0625, 0.0547, 0.0]
GHIA_U100 = [1.0, 0.84123, 0.78871, 0.73722, 0.68717, 0.23151, 0.00332, -0.13641, -0.20581, -0.21090, -0.15662, -0.10150, -0.06434,
             -0.04775, -0.04192, -0.03717, 0.0]


def centreline_u(u, n):
    """u at x = 0.5 on cell-centre heights, from the staggered field (u[j][i] = east face of cell i)."""
    uc = 0.5 * (u[1:n + 1, :-1] + u[1:n + 1, 1:])          # cell centres, columns 1..n
    col = 0.5 * (uc[:, n // 2 - 1] + uc[:, n // 2]) if n % 2 == 0 else uc[:, n // 2]
    y = (np.arange(n) + 0.5) / n
    return np.concatenate([[0.0], y, [1.0]]), np.concatenate([[0.0], col, [1.0]])


def test_cavity_centreline_matches_reference_algorithm_and_ghia(pm, orc):
    """Lid-driven cavity Re = 100 on 32x32 to t = 10 (near steady): the production GPU path (red-black,
    FMA) must give the reference algorithm's centreline (oracle, lexicographic SOR) to 1e-6, and both sit
    within the reference's own distance from Ghia et al. (max|du| = 0.071 at its default
    config; the Dirichlet-south quirk and the coarse grid are part of the reference)."""
    n = 32
    cfg = make_cfg(pm, 0, n, n, RB, 0, 10000)
    cfg.re, cfg.nu = 100.0, 1.0 / 100.0
    h = 1.0 / n
    cfg.dt = 0.5 * min(0.25 * h * h / cfg.nu, h / 1.0)
    steps = int(10.0 / cfg.dt)
    S = pm.Solver(cfg)
    S.apply_bc(0)
    S.step(steps)
    ocfg = cfg.copy()
    ocfg.ppe_method = LEX
    O = orc.Oracle(ocfg)
    O.apply_bc(0)
    O.step(steps)
    for fid in (0, 1, 2):
        assert rel_l2(S.download(fid), O.field(fid)) < 1e-6, f"field {fid}"
    y, ug = centreline_u(S.download(0), n)
    _, uo = centreline_u(O.field(0), n)
    assert np.abs(ug - uo).max() < 1e-6
    ghia = np.interp(GHIA_Y[::-1], y, ug)[::-1]
    assert np.abs(ghia - np.array(GHIA_U100)).max() < 0.02  # measured 0.0046 for the reference algorithm itself


def test_channel_profile_matches_reference_algorithm(pm, orc):
    """Channel 60x20, Re = 20, to t = 4: outlet profile of the production GPU path equals the reference
    algorithm's (oracle) to 1e-6 relative L2; it is parabola-like but not 6y(1-y) — the reference loses mass
    flux through its uncorrected outlet, so the analytic profile is NOT the bar."""
    cfg = make_cfg(pm, 1, 60, 20, RB, 0, 10000)
    cfg.re, cfg.nu = 20.0, 1.0 / 20.0
    m = min(cfg.dx, cfg.dy)
    cfg.dt = 0.25 * min(0.25 * m * m / cfg.nu, m / 1.0)
    steps = int(4.0 / cfg.dt)
    S = pm.Solver(cfg)
    S.apply_bc(0)
    S.step(steps)
    ocfg = cfg.copy()
    ocfg.ppe_method = LEX
    O = orc.Oracle(ocfg)
    O.apply_bc(0)
    O.step(steps)
    for fid in (0, 1, 2):
        assert rel_l2(S.download(fid), O.field(fid)) < 1e-6, f"field {fid}"
    prof = S.download(0)[1:21, 59]
    assert prof.max() > 1.2 and prof.max() < 1.5 and prof.argmax() in (9, 10)
    yc = (np.arange(20) + 0.5) / 20
    assert np.abs(prof - 6 * yc * (1 - yc)).max() < 0.25


@pytest.mark.parametrize("method,omega,one_sm", [(JAC, 0.9, False), (RB, None, False), (RB, None, True)])
@pytest.mark.parametrize("exact", [1, 0])
@pytest.mark.parametrize("case_id,nx,ny", CASES + [(0, 128, 128), (1, 256, 64), (0, 37, 19)])
def test_persistent_small_grid_solve(pm, orc, monkeypatch, case_id, nx, ny, method, omega, exact, one_sm):
    """The persistent solves used for small grids (kernel_path = persistent; red-black: a cluster of 8 CTAs with
    distributed-shared-memory halo rows, or one CTA; Jacobi: one CTA): whole steps run to the reference
    tolerance (cap 400) — iteration counts, residuals and all fields against the oracle,
    0 ulp with exact arithmetic, 1e-12 relative with production arithmetic."""
    if one_sm:
        monkeypatch.setenv("PM_NO_CLUSTER", "1")  # red-black otherwise runs on a cluster of 8 CTAs with DSMEM halos
    cfg = make_cfg(pm, case_id, nx, ny, method, exact, 400, omega, path=3)
    S, O = pm.Solver(cfg), orc.Oracle(cfg)
    S.fill_random(23, 2.0 ** -6); O.fill_random(23, 2.0 ** -6)
    S.apply_bc(0); O.apply_bc(0)
    for n in range(2):
        rs, ro = S.step(1), O.step(1)
        if exact:
            assert (rs.iterations, rs.residual) == (ro.iterations, ro.residual), f"step {n}"
        else:
            assert abs(rs.iterations - ro.iterations) <= 1
    if exact:
        assert_fields_equal(S, O, range(6), "persistent solve")
    else:
        for fid in range(6):
            a, b = S.download(fid), O.field(fid)
            assert np.abs(a - b).max() <= 1e-10 * max(1.0, np.abs(b).max()), f"field {fid}"


@pytest.mark.parametrize("case_id,nx,ny,method", [(0, 300, 200, RB), (1, 260, 180, RB), (0, 300, 200, JAC), (2, 300, 90, RB)])
def test_auto_path_mid_size_bit_exact(pm, orc, case_id, nx, ny, method):
    """kernel_path = auto on grids between the small-grid limit and the benchmark sizes (tiled kernel with few
    tiles; general kernels for the masked step case): one whole step against the oracle, 0 ulp."""
    cfg = make_cfg(pm, case_id, nx, ny, method, 1, 20, 0.9 if method == JAC else None, path=0)
    S, O = pm.Solver(cfg), orc.Oracle(cfg)
    S.fill_random(31, 2.0 ** -5); O.fill_random(31, 2.0 ** -5)
    S.apply_bc(0); O.apply_bc(0)
    rs, ro = S.step(1), O.step(1)
    assert (rs.iterations, rs.residual) == (ro.iterations, ro.residual)
    assert_fields_equal(S, O, range(6), "auto path")


@pytest.mark.parametrize("path", [1, 3, 2])
@pytest.mark.parametrize("case_id,nx,ny", [(0, 2, 2), (0, 3, 5), (1, 2, 3), (1, 5, 2), (2, 6, 4), (0, 116, 28), (0, 117, 29), (1, 129, 45), (0, 127, 17)])
def test_ragged_and_tiny_grids_bit_exact(pm, orc, case_id, nx, ny, path):
    """Edge sizes: the smallest legal grids, odd extents (column pairs straddling the east wall), grids of exactly
    one output tile and one cell more — through the general, persistent and tiled paths."""
    cfg = make_cfg(pm, case_id, nx, ny, RB, 1, 12, path=path)
    if case_id == 2:
        cfg.step_i_location, cfg.inlet_j_max = 2, 2
    cfg.sweeps_per_pass = 3 if path == 2 else 0
    S, O = pm.Solver(cfg), orc.Oracle(cfg)
    S.fill_random(77); O.fill_random(77)
    S.apply_bc(0); O.apply_bc(0)
    rs, ro = S.step(2), O.step(2)
    assert (rs.iterations, rs.residual) == (ro.iterations, ro.residual)
    assert_fields_equal(S, O, range(6), "ragged/tiny")


@pytest.mark.parametrize("path", [1, 2, 3])
def test_loop_entry_quirks(pm, orc, path):
    """max_iters = 0 and the cavity's loop-entry rule: the reference starts its residual at
    1.0, so a source with 1e-9*max|f| >= 1 makes it skip the solve and return p == 0 after 0 iterations."""
    cfg = make_cfg(pm, 0, 130, 40, RB, 1, 0, path=path)
    cfg.sweeps_per_pass = 2 if path == 2 else 0
    S, O = pm.Solver(cfg), orc.Oracle(cfg)
    S.fill_random(3); O.fill_random(3)
    rs, ro = S.step(1), O.step(1)
    assert rs.iterations == ro.iterations == 0 and rs.residual == ro.residual == 1.0
    assert_fields_equal(S, O, range(6), "max_iters=0")
    cfg = make_cfg(pm, 0, 130, 40, RB, 1, 50, path=path)
    cfg.sweeps_per_pass = 2 if path == 2 else 0
    S, O = pm.Solver(cfg), orc.Oracle(cfg)
    S.fill_random(3, 2.0 ** 14); O.fill_random(3, 2.0 ** 14)   # max|f| ~ 1e12 -> tolerance ~ 1e3 >= 1
    rs, ro = S.step(1), O.step(1)
    assert ro.tolerance >= 1.0 and rs.iterations == ro.iterations == 0
    assert not S.download(2).any()
    assert_fields_equal(S, O, range(6), "loop skipped")


def _export_reference(case_id, cfg, u, v, p, mask):
    """The writers' host loops (driver_main.cpp before the device-side export; cavity-01.cpp:717-733,187-223 and siblings) in numpy:
    every operation individually rounded."""
    nx, ny = cfg.nx, cfg.ny
    idx, idy = 1.0 / cfg.dx, 1.0 / cfg.dy
    F = mask.astype(bool)
    uc = np.zeros((ny + 2, nx + 2)); vc = np.zeros((ny + 2, nx + 2))
    J, I = np.meshgrid(np.arange(1, ny + 1), np.arange(1, nx + 1), indexing="ij")
    uc[1:ny + 1, 1:nx + 1] = np.where(F[1:ny + 1, 1:nx + 1], 0.5 * (u[1:ny + 1, 0:nx] + u[1:ny + 1, 1:nx + 1]), 0.0)
    vc[1:ny + 1, 1:nx + 1] = np.where(F[1:ny + 1, 1:nx + 1], 0.5 * (v[0:ny, 1:nx + 1] + v[1:ny + 1, 1:nx + 1]), 0.0)
    c = (slice(1, ny + 1), slice(1, nx + 1))
    e, w = (slice(1, ny + 1), slice(2, nx + 2)), (slice(1, ny + 1), slice(0, nx))
    n, s_ = (slice(2, ny + 2), slice(1, nx + 1)), (slice(0, ny), slice(1, nx + 1))
    if case_id == 0:
        dvdx = np.where(I == 1, (vc[e] - vc[c]) * idx, np.where(I == nx, (vc[c] - vc[w]) * idx, (vc[e] - vc[w]) * idx * 0.5))
        dudy = np.where(J == 1, (uc[n] - uc[c]) * idx, np.where(J == ny, (uc[c] - uc[s_]) * idx, (uc[n] - uc[s_]) * idx * 0.5))
        vort = dvdx - dudy
    elif case_id == 1:
        dvdx = np.where(I == 1, (vc[e] - vc[c]) * idx, np.where(I == nx, (vc[c] - vc[w]) * idx, 0.5 * (vc[e] - vc[w]) * idx))
        dudy = np.where(J == 1, (uc[n] - uc[c]) * idy, np.where(J == ny, (uc[c] - uc[s_]) * idy, 0.5 * (uc[n] - uc[s_]) * idy))
        vort = dvdx - dudy
    else:
        ok = F[c] & ~((I == 1) | (I == nx) | (J == 1) | (J == ny)) & F[e] & F[w] & F[n] & F[s_]
        vort = np.where(ok, 0.5 * (vc[e] - vc[w]) * idx - 0.5 * (uc[n] - uc[s_]) * idy, 0.0)
    mag = np.where(F[c], np.sqrt(uc[c] * uc[c] + vc[c] * vc[c]), 0.0)
    return [uc[c], vc[c], mag, np.where(F[c], p[c], 0.0), vort]


@pytest.mark.parametrize("case_id,nx,ny,path", [(0, 63, 63, 0), (1, 93, 31, 0), (2, 256, 32, 0), (2, 64, 16, 1), (0, 300, 200, 2), (1, 384, 200, 2)])
def test_device_side_export_matches_the_writers_host_loops(pm, case_id, nx, ny, path):
    """pm_export_begin / pm_export_wait: cell-centre velocities, magnitude, pressure and vorticity bit for bit what the VTK
    writers' loops compute from the downloaded fields (incl. p straight from the split-row buffer of a tiled solve)."""
    cfg = make_cfg(pm, case_id, nx, ny, RB, 1, 25, path=path)
    S = pm.Solver(cfg)
    S.fill_random(21, 0.25)
    S.apply_bc(0)
    S.step(2)
    S.export_begin()
    S.step(1)  # the export was taken before this step; the copy overlaps it
    got = S.export_wait()
    T = pm.Solver(cfg)
    T.fill_random(21, 0.25)
    T.apply_bc(0)
    T.step(2)
    want = _export_reference(case_id, cfg, T.download(0), T.download(1), T.download(2), T.download_mask())
    for name, a, b in zip(("u_center", "v_center", "magnitude", "pressure", "vorticity"), got, want):
        assert bits_equal(a, np.ascontiguousarray(b)), f"{name}: max abs {np.abs(a - b).max():.3e}"
    S.close(); T.close()
