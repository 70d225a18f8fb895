"""Wall time per projection step of BASELINE configs[0..2] (the reference's real, latency-bound cases):
GPU through the C-ABI (production red-black, persistent single-CTA solve; and the bit-exact sor-lex mode)
next to the unmodified reference on one host core (oracle/_ref).  Prints one JSON object.

    python tools/small_configs_timing.py [steps]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "computational-fluid-dynamics_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pm_ctypes as pm  # noqa: E402
import orc  # noqa: E402

CONFIGS = [
    ("configs[0] cavity Re=100 128x128 dt=1e-3", pm.CASE_CAVITY, (128, 128, 100.0, 1e-3), "cavity_cfg0"),
    ("configs[1] channel Re=1000 256x64 dt=5e-4", pm.CASE_CHANNEL, (256, 64, 1000.0, 5e-4), "channel_cfg1"),
    ("configs[2] backwards step Re=100 256x32 (mask)", pm.CASE_STEP, (0, 0, 0.0, 0.0), "step_default"),
]


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    out = []
    for name, case, args, refname in CONFIGS:
        row = {"config": name, "steps": steps}
        for label, method, exact, path in (("gpu_sor_rb", pm.PPE_SOR_RB, 0, pm.PATH_AUTO), ("gpu_sor_rb_general_path", pm.PPE_SOR_RB, 0, pm.PATH_SIMPLE),
                                           ("gpu_sor_lex_exact", pm.PPE_SOR_LEX, 1, pm.PATH_AUTO)):
            cfg = pm.config_init(case, *args)
            cfg.ppe_method, cfg.exact_arith, cfg.kernel_path = method, exact, path
            S = pm.Solver(cfg)
            S.apply_bc(0)
            S.step(2)
            S.sync()
            t0 = time.perf_counter()
            iters = 0
            for _ in range(steps):
                iters += S.step(1).iterations
            S.sync()
            dt = time.perf_counter() - t0
            row[label] = {"ms_per_step": 1e3 * dt / steps, "iters_per_step": iters / steps, "us_per_iteration": 1e6 * dt / max(iters, 1)}
            S.close()
        if orc.ref_available(refname):
            R = orc.Reference(refname)
            R.step(2)
            n = max(2, steps // 4)
            it = 0
            t0 = time.perf_counter()
            for _ in range(n):
                it += R.step(1)[0]
            dt = time.perf_counter() - t0
            row["cpu_reference_1core"] = {"ms_per_step": 1e3 * dt / n, "iters_per_step": it / n, "us_per_iteration": 1e6 * dt / max(it, 1)}
        out.append(row)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
