"""CPU tests of the drop-in boundary: libpm.so loads, exports every symbol include/pm.h declares,
its host-side logic agrees with the oracle, and it refuses to run without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pm):
    L = pm.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"libpm.so does not export {s}"
    assert set(syms) == set(pm.EXPORTS), "pm_ctypes.EXPORTS out of sync with include/pm.h"
    assert L.pm_abi_version() == 1


def test_struct_layout_matches_header(pm):
    # pm_create rejects a struct of the wrong size: a cheap ABI check that needs no GPU
    cfg = pm.config_init(pm.CASE_CAVITY)
    assert cfg.struct_size == C.sizeof(pm.PmConfig)
    bad = cfg.copy()
    bad.struct_size = 8
    h = C.c_void_p()
    assert pm.lib().pm_create(C.byref(bad), C.byref(h)) == 1
    assert b"ABI mismatch" in pm.lib().pm_last_error(None)


@pytest.mark.parametrize("case_id", [0, 1, 2])
@pytest.mark.parametrize("args", [(0, 0, 0.0, 0.0), (128, 128, 100.0, 1e-3), (256, 64, 1000.0, 5e-4)])
def test_config_init_matches_oracle(pm, orc, case_id, args):
    a = pm.config_init(case_id, *args)
    b = orc.config_init(case_id, *args)
    for k in ("nx", "ny", "total_steps", "step_i_location", "inlet_j_max", "max_iters", "print_interval", "save_interval"):
        assert getattr(a, k) == getattr(b, k), k
    for k in ("dx", "dy", "nu", "dt", "u_ref", "rho", "omega", "tol_factor", "abs_tol", "lx", "ly", "re", "cfl", "final_time"):
        assert getattr(a, k) == getattr(b, k), k


def test_error_behaviour_mirrors_reference(pm):
    L = pm.lib()
    h = C.c_void_p()
    cfg = pm.config_init(pm.CASE_CAVITY)
    bad = cfg.copy(); bad.nx = 0
    assert L.pm_create(C.byref(bad), C.byref(h)) == 1  # invalid_argument, cavity-01.cpp:57-59
    assert b"Field dimensions must be positive" in L.pm_last_error(None)
    bad = cfg.copy(); bad.dt = 0.0
    assert L.pm_create(C.byref(bad), C.byref(h)) == 2  # runtime_error, cavity-01.cpp:423-425
    assert b"non-positive" in L.pm_last_error(None)
    st = pm.config_init(pm.CASE_STEP)
    bad = st.copy(); bad.step_i_location = st.nx
    assert L.pm_create(C.byref(bad), C.byref(h)) == 2  # backwards_step-01.cpp:459-461
    assert b"outside computational domain" in L.pm_last_error(None)
    assert L.pm_config_init(None, 0, 0, 0, 0.0, 0.0) == 1
    assert L.pm_config_init(C.byref(cfg), 7, 0, 0, 0.0, 0.0) == 1


def test_slab_ranges_tile_the_grid(pm):
    L = pm.lib()
    for ny, n in [(16384 * 8, 8), (31, 2), (63, 4), (10, 3)]:
        nxt = 0
        for r in range(n):
            j0, nyl = C.c_int(), C.c_int()
            assert L.pm_slab_range(ny, n, r, C.byref(j0), C.byref(nyl)) == 0
            assert j0.value == nxt and nyl.value >= ny // n
            nxt += nyl.value
        assert nxt == ny
    assert L.pm_slab_range(4, 8, 0, None, None) == 1


def test_no_cpu_fallback(pm):
    """Without a CUDA device the product must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = pm.config_init(pm.CASE_CAVITY, 16, 16)
    h = C.c_void_p()
    assert pm.lib().pm_create(C.byref(cfg), C.byref(h)) == 3  # PM_ERR_CUDA
    assert b"no CPU path" in pm.lib().pm_last_error(None)
    with pytest.raises(pm.PmError):
        pm.Solver(cfg)


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference: the reference's CPU code on a bounded sample; exactly one JSON line on stdout
    with the contract's keys (the driver computes the GPU/CPU ratio from it)."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-n", "512"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mcell-updates/s per projection step" and d["unit"] == "Mcell-updates/s"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["higher_is_better"] is True


def test_split_row_layout_properties(tmp_path):
    """The split-row layout of the tiled solve (pm_split_col, TileCfg::PSH): bijection per row, even/odd halves, and a
    16-byte aligned, in-range TMA box start for every tile of every tile shape.  Host-only program built with nvcc."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path / "split_check"
    src = os.path.join(ROOT, "tests", "helpers", "split_layout_check.cu")
    subprocess.run([nvcc, "-std=c++17", "-o", str(exe), src], check=True, capture_output=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "split layout ok" in out.stdout, out.stdout + out.stderr


# ---- host logic of the streaming pressure pass (pm_kernels_stream.cuh, stream_shape) -------------------------------
TX, TY, SH, H, SW = 112, 32, 48, 8, 128  # tile geometry of the red-black plan at T = 4 (pm_tile_cfg.cuh)


def _interior(nx, ny, nyl, j0, bx, by):
    """k_ppe_tiled's `interior` for an independent tile: every updatable cell strictly inside the domain, data for all neighbours."""
    ib, jb = 1 + bx * TX - H, 1 + by * TY - H
    return (ib + 1 >= 2 and ib + SW - 2 <= nx - 1 and j0 + jb + 1 >= 2 and j0 + jb + SH - 2 <= ny - 1
            and jb + SH - 1 <= nyl + H and jb >= 1 - H)


@pytest.mark.parametrize("nx,ny,nranks,slots", [(8192, 8192, 1, 1776), (16384, 16384, 1, 1776), (1400, 420, 1, 1776), (2000, 1500, 1, 1776),
                                                 (16384, 16384 * 8, 8, 1776), (1400, 1408, 2, 1776), (1400, 1410, 2, 1184), (900, 300, 1, 1776)])
def test_streaming_plan_arithmetic(pm, nx, ny, nranks, slots):
    """The rectangle the streaming pass takes holds interior tiles only and cannot be widened; its chunks tile the rows; the
    chunk height is the cheapest under the wave model the plan states (whole waves of `slots` warps, rows + 24 ticks each)."""
    L = pm.lib()
    for rank in sorted({0, nranks // 2, nranks - 1}):
        j0, nyl = C.c_int(), C.c_int()
        assert L.pm_slab_range(ny, nranks, rank, C.byref(j0), C.byref(nyl)) == 0
        j0, nyl = j0.value, nyl.value
        tiles_x, tiles_y = -(-nx // TX), -(-nyl // TY)
        row_lo, row_hi = 0, tiles_y
        if nranks > 1:  # edge tile rows go ahead of the rest (slab_edge_rows in pm_capi.cu)
            top = 1
            while top < tiles_y and nyl - (tiles_y - top) * TY < H:
                top += 1
            row_lo, row_hi = 1, tiles_y - top
        out = (C.c_int * 8)()
        on = L.pm_stream_plan(nx, ny, nyl, j0, row_lo, row_hi, slots, out)
        ok = [[_interior(nx, ny, nyl, j0, bx, by) for bx in range(tiles_x)] for by in range(tiles_y)]
        if not on:
            assert sum(ok[by][bx] for by in range(row_lo, row_hi) for bx in range(tiles_x)) < 64 or nx < 400
            continue
        bx0, nbx, by0, nby, rows, nch, items, nframe = list(out)
        assert all(ok[by][bx] for by in range(by0, by0 + nby) for bx in range(bx0, bx0 + nbx))
        assert by0 >= row_lo and by0 + nby <= row_hi
        assert (bx0 == 0 or not ok[by0][bx0 - 1]) and (bx0 + nbx == tiles_x or not ok[by0][bx0 + nbx])
        assert (by0 == row_lo or not ok[by0 - 1][bx0]) and (by0 + nby == row_hi or not ok[by0 + nby][bx0])
        assert nframe == (row_hi - row_lo) * tiles_x - nbx * nby
        total = nby * TY
        assert rows % TY == 0 and nch * rows >= total > (nch - 1) * rows and items == nbx * nch
        cost = lambda R: -(-(nbx * -(-total // R)) // slots) * (R + 2 * H + 8)
        assert cost(rows) == min(cost(R) for R in range(TY, total + 1, TY))
    if (nx, ny, nranks) == (8192, 8192, 1):
        assert (bx0, nbx, by0, nby, rows, items) == (1, 72, 1, 254, 352, 1728)  # the measured launch shape
