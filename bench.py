#!/usr/bin/env python
"""bench.py — Mcell-updates/s per projection step (BASELINE.json metric) on N B200s.

A "step" is one full projection step (BC fill, predictor, divergence source, K pressure-Poisson
iterations each with its residual norm, velocity correction) over one synthetic grid.
N = 1: BASELINE configs[3], lid-driven cavity Re=1000 at 8192x8192.  N > 1: configs[4], a
16384 x (16384*N) cavity slab-decomposed in j, 16384^2 cells per GPU (weak scaling).
K = 100 iterations per step via max_iters (the reference semantics with max_sor_iterations
lowered); the reference tolerance stays in place and is never met at this size.

One JSON line on stdout (rank 0).  `--impl reference` times the reference's own CPU code
(oracle/_ref, built from the unmodified sources) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "computational-fluid-dynamics_b200"))

METRIC = "Mcell-updates/s per projection step"
UNIT = "Mcell-updates/s"
K_ITERS = 100  # headline; --k-iters overrides (K = 1 exposes the non-pressure passes)


def bytes_per_cell_step(k):
    """Algorithmic HBM bytes per cell per projection step, SURVEY §8d: 96 + 24 K."""
    return 96 + 24 * k


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (one streaming nvidia-smi process,
    a row every 50 ms; rows are timestamped on arrival and only those inside [start, stop] are kept)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None
        self.t_start = self.t_stop = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))
        except Exception:
            pass

    def mark_start(self):
        self.t_start = time.perf_counter()

    def stop(self):
        self.t_stop = time.perf_counter()
        time.sleep(0.12)  # let the last rows of the region arrive
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=5)
        sm, mx, reasons = [], 0.0, set()
        rows = [r for t, r in self.rows if self.t_start is None or self.t_start <= t <= self.t_stop + 0.1] or [r for _, r in self.rows]
        for r in rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(n, steps, warmup):
    """Reference CPU code (oracle/_ref/libref_cavity_k100_<n>.so), single thread as the reference is."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    name = f"cavity_k100_{n}"
    if orc.ref_available(name) and K_ITERS == 100:
        R = orc.Reference(name)
        assert R.params()["max_iters"] == K_ITERS
        if warmup:
            R.time_steps(warmup)
        secs = R.time_steps(steps)
        kind = "reference"
    else:  # the oracle port (same algorithm, runtime parameters)
        cfg = orc.config_init(0, n, n)
        cfg.max_iters, cfg.ppe_method = K_ITERS, 2
        O = orc.Oracle(cfg)
        O.apply_bc(0)
        if warmup:
            O.step(warmup)
        t0 = time.perf_counter(); O.step(steps); secs = time.perf_counter() - t0
        kind = "port"
    return n * n * steps / secs / 1e6, secs, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.ref_n
    rate, secs, kind = cpu_reference_rate(n, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (solver's own initial state, lid at rest start)",
        "config": {"workload": f"lid-driven cavity Re=1000, K={K_ITERS} lexicographic SOR iterations/step (reference ordering)",
                   "sample": f"{n}x{n} sub-grid of the 8192x8192 workload; rate is per cell so it carries over"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": f"cavity {n}x{n}, K={K_ITERS}, {args.steps} steps, g++ -O2 -ffp-contract=off, the reference is single-threaded"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_ours(args):
    import torch
    import pm_ctypes as pm
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    if world == 1:
        nx = ny_local = args.n or 8192
        workload = f"lid-driven cavity Re=1000 {nx}x{nx} on 1 B200 (BASELINE configs[3])"
    else:
        nx = ny_local = args.n or 16384
        workload = f"lid-driven cavity Re=1000 {nx}x{ny_local * world}, {nx}x{ny_local} per GPU, {world} j-slabs (BASELINE configs[4])"
    ny = ny_local * world
    cfg = pm.config_init(pm.CASE_CAVITY, nx, ny, 1000.0, 0.0)
    cfg.max_iters = K_ITERS
    # The reference enters its loop only if 1.0 > tolerance_factor * max|f| (cavity-01.cpp:618,632,635).  At
    # h = 1/8192 the lid corners alone give max|f| = 2*nu*U/h^3 = 1.1e9, so with the compiled-in 1e-9 the
    # solver would not sweep at all; 1e-12 keeps the rule in force and lets the K-iteration cap end the loop.
    cfg.tol_factor = 1e-12
    cfg.ppe_method = {"rb": pm.PPE_SOR_RB, "jacobi": pm.PPE_JACOBI}[args.ppe]
    if args.ppe == "jacobi":
        cfg.omega = 1.0
    cfg.exact_arith = args.exact
    cfg.kernel_path = args.path
    cfg.sweeps_per_pass = args.sweeps
    cfg.device = local
    cfg.rank, cfg.nranks = rank, world
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            import ctypes as C
            buf = (C.c_uint8 * 128)()
            assert pm.lib().pm_nccl_unique_id(buf) == 0
            idt = torch.tensor(list(buf), dtype=torch.uint8)
        idt = idt.cuda()
        dist.broadcast(idt, 0)
        for q, b in enumerate(idt.cpu().tolist()):
            cfg.nccl_id[q] = b
    S = pm.Solver(cfg)
    # u, v ~ 2^-10 * U(-1,1) by global flat index; p cold-starts in the cavity.  The amplitude keeps
    # max|f| < 1e9: with U(-1,1) at h = 1/8192 the reference's own loop test (res = 1.0 > 1e-9*max|f|,
    # cavity-01.cpp:618,632,635) is false before the first sweep and the solver would do no work at all.
    S.fill_random(42, 2.0 ** -10)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        S.sync()

    for _ in range(args.warmup):
        S.step(1)
    t_before = S.timing()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)  # nvidia-smi needs a moment to start streaming
    barrier()
    if sampler:
        sampler.mark_start()
    S.timer_start()
    for _ in range(args.steps):
        r = S.step(1)
    ms = S.timer_stop()
    barrier()
    clocks = sampler.stop() if sampler else None
    t_after = S.timing()
    assert r.iterations == K_ITERS, f"PPE stopped after {r.iterations} iterations (expected the K={K_ITERS} cap)"
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    cells = nx * ny  # all ranks
    value = cells * args.steps / (ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (the pressure sweep + fused residual) ----
    peak, peak_src = measured_peaks()
    ppe_ms = t_after.ppe_ms - t_before.ppe_ms
    passes = t_after.ppe_passes - t_before.ppe_passes
    sweeps_per_pass = (K_ITERS * args.steps) / passes if passes else 0
    bytes_per_pass = 24.0 * nx * ny_local * sweeps_per_pass  # this rank's cells
    ms_per_pass = ppe_ms / passes if passes else float("nan")
    achieved = bytes_per_pass / (ms_per_pass * 1e-3) / 1e9
    step_bytes = bytes_per_cell_step(K_ITERS) * nx * ny_local
    traffic = None  # measured DRAM bytes per launch (ncu), when this exact workload was profiled
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath) and world == 1 and args.ppe == "rb" and not args.exact:
        key = f"k_ppe_tiled<Fast,0,1,{int(round(sweeps_per_pass + 0.4))},0> {nx}x{ny}"  # same loads and stores since the first capture
        traffic = json.load(open(tpath)).get(key, {}).get("dram_bytes_per_launch")
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "kernel": "pressure sweep pass (sweep(s) + fused inf-norm residual)", "peak_source": peak_src,
        "algorithmic_bytes_per_launch": bytes_per_pass, "ms_per_launch": ms_per_pass, "sweeps_per_launch": sweeps_per_pass,
        "whole_step": {"algorithmic_GBps": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                       "frac_of_measured": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak,
                       "frac_of_nominal_8TBps": step_bytes / (ms / args.steps * 1e-3) / 1e9 / 8000.0,
                       "bytes_per_cell_step": bytes_per_cell_step(K_ITERS)},
    }

    # ---- e2e: the same step through the C-ABI with HOST buffers (pinned), every step's copies inside the timed region ----
    # Headline: pm_host_step_submit/run/drain, which overlap the upload of step n+1 and the download of step n-1 with
    # the kernels of step n (three rotating plane sets, one stream per copy direction).  `serial` is the plain sequence
    # pm_upload_slab -> pm_step -> pm_download_slab with nothing overlapped.
    e2e = None
    if not args.no_e2e:
        shp = {f: S.slab_rows(f)[1:] for f in (pm.F_U, pm.F_V, pm.F_P)}  # this rank's rows, reference row layout
        host = {f: torch.empty(shp[f], dtype=torch.float64).pin_memory() for f in shp}
        for f in (pm.F_U, pm.F_V):
            S.download_slab_ptr(f, host[f].data_ptr(), host[f].numel())
        src = {f: host[f].clone().pin_memory() for f in (pm.F_U, pm.F_V)}  # the steps' inputs (the cavity cold-starts p, cavity-01.cpp:610-611)
        h2d = 8 * (host[pm.F_U].numel() + host[pm.F_V].numel())
        d2h = 8 * (host[pm.F_U].numel() + host[pm.F_V].numel() + host[pm.F_P].numel())

        def wall_ms(fn, n):
            barrier()
            t0 = time.perf_counter()
            fn(n)
            barrier()
            ms_ = (time.perf_counter() - t0) * 1e3
            if dist is not None:
                t = torch.tensor([ms_], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_ = float(t.item())
            return ms_

        def serial(n):
            for _ in range(n):
                for f in (pm.F_U, pm.F_V):
                    S.upload_slab_ptr(f, src[f].data_ptr(), src[f].numel())
                S.step(1)
                for f in (pm.F_U, pm.F_V, pm.F_P):
                    S.download_slab_ptr(f, host[f].data_ptr(), host[f].numel())

        def streamed(n):
            def submit():
                S.host_step_submit((src[pm.F_U].data_ptr(), src[pm.F_U].numel()), (src[pm.F_V].data_ptr(), src[pm.F_V].numel()),
                                   host[pm.F_U].data_ptr(), host[pm.F_V].data_ptr(), (host[pm.F_P].data_ptr(), host[pm.F_P].numel()))
            submit()
            for q in range(n):
                if q + 1 < n:
                    submit()
                S.host_step_run()
            S.host_step_drain()

        s_steps = max(1, min(args.steps, 2 if world == 1 else 1))
        s_ms = wall_ms(serial, s_steps)
        e_steps = max(2, min(args.steps, 20 if world == 1 else 8))  # 16384^2 slabs move 4x the bytes per step
        streamed(2)  # first use allocates the extra plane sets
        e_ms = wall_ms(streamed, e_steps)
        e2e = {"value": cells * e_steps / (e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "steps": e_steps, "ms_per_step": e_ms / e_steps,
               "what": "pm_host_step_submit/run/drain: u, v from pinned host memory in, u, v, p to pinned host memory out, every step; "
                       "copies of neighbouring steps overlap the kernels; wall clock from the first submit to the end of the last download",
               "serial": {"value": cells * s_steps / (s_ms * 1e-3) / 1e6, "ms_per_step": s_ms / s_steps, "steps": s_steps,
                          "what": "pm_upload_slab(u,v) + pm_step + pm_download_slab(u,v,p), nothing overlapped"}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, secs, kind = cpu_reference_rate(args.cpu_n, args.cpu_steps, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": f"cavity {args.cpu_n}x{args.cpu_n} Re=1000 K={K_ITERS}, {args.cpu_steps} steps in {secs:.1f} s, unmodified reference built g++ -O2 -ffp-contract=off (single-threaded by design), host has {os.cpu_count()} cores"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic (splitmix64 2^-10*U(-1,1) u,v by global flat index, seed 42)",
            "config": {"workload": workload, "ppe": f"{args.ppe}, K={K_ITERS} iterations/step (max_iters cap), residual every iteration, tolerance_factor 1e-12",
                       "arith": "exact (no FMA)" if args.exact else "production (FMA)",
                       "kernel_path": {0: "auto", 1: "simple", 2: "tiled"}[args.path],
                       "l2": "inputs larger than L2 (>= 537 MB per field vs 126 MB L2); no flush needed",
                       "parallelism": f"{world} j-slab(s)"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(t_after.kernel_launches - t_before.kernel_launches), "clocks": clocks,
        }
        emit(line)
    S.close()
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The one JSON line goes to the process's original stdout; everything else any library prints
    (NCCL's version banner, torchrun notices) was redirected to stderr in main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # fd 1 -> stderr for the rest of the run (C libraries included)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=0, help="override the grid edge (per GPU)")
    ap.add_argument("--ppe", default="rb", choices=["rb", "jacobi"])
    ap.add_argument("--exact", type=int, default=0)
    ap.add_argument("--path", type=int, default=0)
    ap.add_argument("--sweeps", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-n", type=int, default=1024)
    ap.add_argument("--cpu-steps", type=int, default=12)
    ap.add_argument("--ref-n", type=int, default=2048)
    ap.add_argument("--k-iters", type=int, default=100, help="pressure iterations per step (max_iters cap)")
    args = ap.parse_args()
    global K_ITERS
    K_ITERS = args.k_iters
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
