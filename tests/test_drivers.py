"""GPU tests of the drop-in drivers (bin/cavity, bin/channel): run with the reference's ordering and exact
arithmetic they must write the SAME BYTES as the unmodified reference programs — stdout (ANSI stripped),
stderr warnings, and every VTK frame (md5) — recorded in tests/golden/drivers.json by
tests/golden/make_vtk_golden.py from the reference's own run.  The production ordering (red-black) is checked
on the printed diagnostics, which agree to the printed precision."""
import hashlib
import json
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "computational-fluid-dynamics_b200", "bin")
ANSI = re.compile(r"\x1b\[[0-9;]*m")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "drivers.json")))


def run(exe, args, cwd, timeout=600):
    p = subprocess.run([os.path.join(BIN, exe)] + args, cwd=cwd, capture_output=True, text=True, timeout=timeout)
    return p.returncode, ANSI.sub("", p.stdout).splitlines(), ANSI.sub("", p.stderr).splitlines()


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


@pytest.mark.parametrize("exe,case,stop", [("cavity", "cavity", 200), ("channel", "channel", 200)])
def test_reference_ordering_writes_identical_bytes(tmp_path, exe, case, stop):
    rc, out, err = run(exe, ["--ppe", "sor-lex", "--exact", "1", "--stop-after", str(stop)], tmp_path)
    assert rc == 0, err
    gold = GOLD[case]
    # stdout up to and including the export of frame `stop` is the reference's, line for line
    last = f"Exported VTK file: {case}_flow_{stop:06d}.vtk"
    n = gold["stdout"].index(last) + 1
    assert out[:n] == gold["stdout"][:n]
    for step in range(0, stop + 1, 100):
        name = f"{case}_flow_{step:06d}.vtk"
        assert md5(tmp_path / "vtk_output" / name) == gold["md5"][name], name
    if case == "channel":  # step 2 of the reference run stops at the 10000-iteration cap and says so on stderr
        assert err[:1] == gold["stderr_first"][:1]
    else:
        assert err == []


def test_backwards_step_reference_ordering_writes_identical_bytes(tmp_path):
    """backwards_step-01.cpp (VTKWriter with the FluidMask block and the literal "0.0" of solid cells, :87-311; log lines
    :1054-1060): the reference needs 40 minutes for its 3072 steps, so its record is a run cut after frame 20
    (make_vtk_golden.py --cases backwards_step --cut-s 45).  Frames 0, 10, 20, the stdout up to the export of frame 20
    and the cap warnings of the first 20 steps (16 of them stop at 10 000 iterations: tests/golden/step_default_20.npz)
    must be the reference's, byte for byte."""
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "step_default_20.npz"))
    capped = int((g["iters"] == 10000).sum())
    rc, out, err = run("backwards_step", ["--ppe", "sor-lex", "--exact", "1", "--stop-after", "20"], tmp_path)
    assert rc == 0, err
    gold = GOLD["backwards_step"]
    n = gold["stdout"].index("Exported VTK file: backwards_step_000020.vtk") + 1
    assert out[:n] == gold["stdout"][:n]
    for step in (0, 10, 20):
        name = f"backwards_step_{step:06d}.vtk"
        assert md5(tmp_path / "vtk_output" / name) == gold["md5"][name], name
    assert capped == 16 and err[:capped] == gold["stderr_first"][:capped]


def test_readme_flags_and_error_exit(tmp_path):
    rc, out, err = run("cavity", ["--Re", "100", "--Nx", "128", "--Ny", "128", "--dt", "1e-3", "--stop-after", "2", "--no-vtk"], tmp_path)
    assert rc == 0
    assert "Grid: 128x128 (spacing=0.007812)" in out and "Time: dt=0.001000, steps=20000, final_time=20.000000" in out
    rc, out, err = run("channel", ["--Re", "1000", "--Nx", "256", "--Ny", "64", "--dt", "5e-4", "--stop-after", "1", "--no-vtk"], tmp_path)
    assert rc == 0 and "Grid: 256x64 (dx=0.011719, dy=0.015625)" in out
    # STEP_LOCATION / dx = nx / 4 truncates to 0 for nx = 2: the reference's constructor check fires
    rc, out, err = run("backwards_step", ["--Nx", "2", "--Ny", "8", "--stop-after", "1", "--no-vtk"], tmp_path)
    assert rc == 1 and err[-1].startswith("Error: Step location is outside computational domain!")


def test_production_ordering_prints_the_reference_diagnostics(tmp_path):
    """Red-black SOR (default) needs more sweeps than the lexicographic order, but max(div) and avg_KE agree
    with the reference's log to every printed digit over the first 300 steps of the default cavity run."""
    rc, out, err = run("cavity", ["--stop-after", "300", "--no-vtk"], tmp_path)
    assert rc == 0
    ours = [l for l in out if l.startswith("Step")]
    ref = [l for l in GOLD["cavity"]["stdout"] if l.startswith("Step")][:3]
    assert len(ours) == 3
    for a, b in zip(ours, ref):
        assert a.split("| SOR_iters")[0] == b.split("| SOR_iters")[0]


@pytest.mark.parametrize("exe", ["channel", "backwards_step"])
def test_two_gpu_driver_writes_the_single_gpu_frames(tmp_path, exe):
    """bin/<case> --gpus 2 (one host thread and one handle per GPU, NCCL between them) against --gpus 1 with exact arithmetic:
    the same VTK bytes.  The default channel (31 rows -> slabs of 16 + 15) is the case where the kernel path once depended on
    the rank's own slab height and the ranks' collective sequences diverged."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    args = ["--steps", "6", "--save-interval", "2", "--max-iters", "40", "--exact", "1"]
    d1, d2 = tmp_path / "one", tmp_path / "two"
    rc1, _, err1 = run(exe, args + ["--outdir", str(d1)], tmp_path, timeout=90)
    rc2, _, err2 = run(exe, args + ["--gpus", "2", "--outdir", str(d2)], tmp_path, timeout=90)
    assert rc1 == 0 and rc2 == 0, (err1, err2)
    names = sorted(os.listdir(d1))
    assert names == sorted(os.listdir(d2)) and len(names) >= 4
    for n in names:
        assert md5(d1 / n) == md5(d2 / n), n
