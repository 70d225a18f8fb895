"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference
(oracle/_ref/libref_*.so, built by oracle/build_ref.sh from /root/reference).

The reference ships no golden vectors, so these files pin parity: each holds the
reference's own u, v, p (+ u*, v*, f) after a few projection steps from its own initial state,
with the per-step SOR iteration counts and final residuals.  Run in the build container only:

    make -C oracle && python tests/golden/make_golden.py
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orc  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# name -> (reference build, steps)
CASES = {
    "cavity_default": ("cavity_default", 4),     # 63x63 Re=1000 (cavity-01.cpp defaults)
    "channel_default": ("channel_default", 3),   # 93x31 Re=100
    "step_default": ("step_default", 2),         # 256x32 Re=100, obstacle mask
    "cavity_cfg0": ("cavity_cfg0", 2),           # BASELINE configs[0]: 128x128 Re=100 dt=1e-3
    "channel_cfg1": ("channel_cfg1", 1),         # BASELINE configs[1]: 256x64 Re=1000 dt=5e-4
    "cavity_k50_32": ("cavity_k50_32", 5),       # 32x32, 50-iteration cap
    "channel_k50": ("channel_k50", 5),           # 93x31, 50-iteration cap
    "step_k50": ("step_k50", 5),                 # 256x32, 50-iteration cap
    "step_default_20": ("step_default", 20),     # longer horizon: 20 steps, every one of them ends at the 10 000 cap
}


def main():
    only = sys.argv[1:]
    for name, (build, steps) in CASES.items():
        if only and name not in only:
            continue
        R = orc.Reference(build)
        prm = R.params()
        iters, res = [], []
        for _ in range(steps):
            it, r = R.step(1)
            iters.append(it)
            res.append(r)
        data = {f"f{fid}": R.get(fid) for fid in range(6)}
        data["mask"] = R.mask()
        data["iters"] = np.array(iters, dtype=np.int32)
        data["res"] = np.array(res, dtype=np.float64)
        data["case_id"] = np.int32(R.case_id)
        data["steps"] = np.int32(steps)
        for k, v in prm.items():
            data[f"prm_{k}"] = np.float64(v) if isinstance(v, float) else np.int32(v)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: steps={steps} iters={iters} res={res[-1]:.6e} -> {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
