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

Beside the contract keys the line carries
  parity_check   the benchmarked (tiled, production-arithmetic) run compared with a general-path run of the same
                 steps on the same inputs (rel. L-inf of u, v, p; all finite), and for N > 1 a small exact-arithmetic
                 slab run compared bit for bit with the same problem on one GPU;
  secondary      K = 1 (the non-pressure passes), and the channel / backwards-step forms on a large grid;
  small_configs  BASELINE configs[0..2] at their real tolerances: ms/step on the GPU and of the reference on one core;
  weak_base      N > 1 only: the same 16384^2 slab timed on ONE GPU in this very run (the base of weak-scaling efficiency).
"""
import argparse
import hashlib
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


def kernel_source_id():
    """Identity of the build the ncu traffic figure belongs to: md5 over the sources of the pressure-pass kernels."""
    h = hashlib.md5()
    for f in ("pm_kernels_stream.cuh", "pm_kernels_tiled.cuh", "pm_tile_cfg.cuh", "pm_common.cuh"):
        h.update(open(os.path.join(ROOT, "computational-fluid-dynamics_b200", "csrc", f), "rb").read())
    return h.hexdigest()


def measured_traffic(key):
    """DRAM bytes per launch of the dominant kernel from the ncu capture of THIS build (profiles/r02_traffic.json,
    written by tools/ncu_traffic.py from `ncu --set full`); None when the kernel sources changed since that capture."""
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(tpath):
        return None, "no ncu capture on record"
    t = json.load(open(tpath))
    if t.get("kernel_source_md5") != kernel_source_id():
        return None, "kernel sources changed since the ncu capture on record"
    e = t.get("entries", {}).get(key)
    if not e:
        return None, f"no ncu capture for {key}"
    return e["dram_bytes_per_launch"], e.get("source")


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


def nccl_id_bytes(pm, dist, rank):
    import ctypes as C
    import torch
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_uint8 * 128)()
        assert pm.lib().pm_nccl_unique_id(buf) == 0
        idt = torch.tensor(list(buf), dtype=torch.uint8)
    idt = idt.cuda()
    dist.broadcast(idt, 0)
    return idt.cpu().tolist()


def bench_cfg(pm, case, nx, ny, k_iters, args, local, path=None):
    """The synthetic benchmark configuration of one case on an nx x ny grid with a K-iteration cap."""
    if case == "cavity":
        cfg = pm.config_init(pm.CASE_CAVITY, nx, ny, 1000.0, 0.0)
        # The reference enters its loop only if 1.0 > tolerance_factor * max|f| (cavity-01.cpp:618,632,635).  At
        # h = 1/8192 the lid corners alone give max|f| = 2*nu*U/h^3 = 1.1e9, so with the compiled-in 1e-9 the
        # solver would not sweep at all; 1e-12 keeps the rule in force and lets the K-iteration cap end the loop.
        cfg.tol_factor = 1e-12
    else:
        cfg = pm.config_init(pm.CASE_CHANNEL if case == "channel" else pm.CASE_STEP, nx, ny, 1000.0 if case == "channel" else 100.0, 0.0)
    cfg.max_iters = k_iters
    cfg.ppe_method = {"rb": pm.PPE_SOR_RB, "jacobi": pm.PPE_JACOBI}[args.ppe]
    if args.ppe == "jacobi":
        cfg.omega = 1.0
    cfg.exact_arith = args.exact
    cfg.kernel_path = args.path if path is None else path
    cfg.sweeps_per_pass = args.sweeps
    cfg.device = local
    return cfg


def init_state(S, case):
    # u, v ~ 2^-10 * U(-1,1) by global flat index; p cold-starts in the cavity.  The amplitude keeps
    # max|f| < 1e9: with U(-1,1) at h = 1/8192 the reference's own loop test (res = 1.0 > 1e-9*max|f|,
    # cavity-01.cpp:618,632,635) is false before the first sweep and the solver would do no work at all.
    S.fill_random(42, 2.0 ** -10)
    if case != "cavity":
        S.apply_bc(0)  # the channel / step constructors apply the BCs once before the loop (channel-01.cpp:352)


def device_timed(S, steps, warmup):
    """ms per step with the handle's own CUDA events, after `warmup` untimed steps."""
    for _ in range(warmup):
        S.step(1)
    S.sync()
    S.timer_start()
    for _ in range(steps):
        r = S.step(1)
    return S.timer_stop() / steps, r


def secondary_lines(pm, args, local):
    """Single-GPU side measurements: K = 1 (what is left when the pressure loop is one sweep) and the other two
    forms of the path (channel: dx != dy, in-tile wall ghosts, source mean; step: obstacle mask) on 8192 x 2048."""
    out = {}
    n = args.n or 8192
    cfg = bench_cfg(pm, "cavity", n, n, 1, args, local)
    S = pm.Solver(cfg)
    init_state(S, "cavity")
    ms, r = device_timed(S, 10, 3)
    S.close()
    out["k1_cavity"] = {"workload": f"cavity {n}x{n}, K=1", "ms_per_step": ms, "value": n * n / ms / 1e3, "unit": UNIT,
                        "algorithmic_GBps": bytes_per_cell_step(1) * n * n / ms / 1e6}
    for case in ("channel", "step"):
        nx, ny = (n, n // 4)
        cfg = bench_cfg(pm, case, nx, ny, K_ITERS, args, local)
        S = pm.Solver(cfg)
        init_state(S, case)
        t0 = S.timing()
        ms, r = device_timed(S, 6, 2)
        t1 = S.timing()
        S.close()
        out[f"{case}_{nx}x{ny}"] = {"workload": f"{case} {nx}x{ny}, K={K_ITERS} red-black iterations/step, residual every iteration",
                                    "ms_per_step": ms, "value": nx * ny / ms / 1e3, "unit": UNIT, "iterations": r.iterations,
                                    "algorithmic_GBps": bytes_per_cell_step(K_ITERS) * nx * ny / ms / 1e6,
                                    "launches_per_step": (t1.kernel_launches - t0.kernel_launches) / 8}
    # the weak-scaling slab (BASELINE configs[4]: 16384^2 per GPU) on this one GPU
    if not args.n:
        n2 = 16384
        cfg = bench_cfg(pm, "cavity", n2, n2, K_ITERS, args, local)
        S = pm.Solver(cfg)
        init_state(S, "cavity")
        ms, r = device_timed(S, 4, 2)
        S.close()
        out[f"cavity_{n2}x{n2}"] = {"workload": f"cavity {n2}x{n2}, K={K_ITERS} (one slab of the N > 1 runs)", "ms_per_step": ms, "value": n2 * n2 / ms / 1e3,
                                    "unit": UNIT, "iterations": r.iterations, "algorithmic_GBps": bytes_per_cell_step(K_ITERS) * n2 * n2 / ms / 1e6}
    return out


SMALL_CONFIGS = [
    ("configs[0] cavity Re=100 128x128 dt=1e-3", "cavity", (128, 128, 100.0, 1e-3), "cavity_cfg0"),
    ("configs[1] channel Re=1000 256x64 dt=5e-4", "channel", (256, 64, 1000.0, 5e-4), "channel_cfg1"),
    ("configs[2] backwards step Re=100 256x32 (mask)", "step", (0, 0, 0.0, 0.0), "step_default"),
]


def small_configs(pm, local, with_cpu):
    """BASELINE configs[0..2] exactly as the reference runs them (real tolerances, 10 000 cap): wall ms per step of the
    production path (red-black, persistent cluster solve) and of the unmodified reference on one host core."""
    rows = []
    for name, case, a, refname in SMALL_CONFIGS:
        cid = {"cavity": pm.CASE_CAVITY, "channel": pm.CASE_CHANNEL, "step": pm.CASE_STEP}[case]
        cfg = pm.config_init(cid, *a)
        cfg.ppe_method, cfg.device = pm.PPE_SOR_RB, local
        S = pm.Solver(cfg)
        S.apply_bc(0)
        S.step(2)
        S.sync()
        t0 = time.perf_counter()
        iters, steps = 0, 10
        for _ in range(steps):
            iters += S.step(1).iterations
        S.sync()
        dt = time.perf_counter() - t0
        S.close()
        row = {"config": name, "gpu_ms_per_step": 1e3 * dt / steps, "gpu_iterations_per_step": iters / steps,
               "gpu_us_per_iteration": 1e6 * dt / max(iters, 1), "ordering": "red-black (production); the reference's is lexicographic"}
        # beyond the reference (non-default, `--omega mixed` of the drivers): the relaxation factor of the mixed-BC operator
        cfg2 = pm.config_init(cid, *a)
        cfg2.ppe_method, cfg2.device = pm.PPE_SOR_RB, local
        cfg2.omega = pm.lib().pm_omega_mixed_bc(cid, cfg2.nx, cfg2.ny, cfg2.dx, cfg2.dy)
        S = pm.Solver(cfg2)
        S.apply_bc(0)
        S.step(2)
        S.sync()
        t0 = time.perf_counter()
        iters2 = 0
        for _ in range(steps):
            iters2 += S.step(1).iterations
        S.sync()
        dt2 = time.perf_counter() - t0
        S.close()
        row["omega_mixed"] = {"omega": cfg2.omega, "reference_omega": cfg.omega, "gpu_ms_per_step": 1e3 * dt2 / steps,
                              "gpu_iterations_per_step": iters2 / steps, "what": "same solver, same tolerance, omega = pm_omega_mixed_bc (include/pm.h)"}
        if with_cpu:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import orc
            if orc.ref_available(refname):
                R = orc.Reference(refname)
                R.step(2)
                it, n = 0, 3
                t0 = time.perf_counter()
                for _ in range(n):
                    it += R.step(1)[0]
                dt = time.perf_counter() - t0
                row.update({"reference_cpu_ms_per_step": 1e3 * dt / n, "reference_iterations_per_step": it / n, "reference_kind": "oracle/_ref (unmodified reference), 1 core"})
        rows.append(row)
    return rows


def slabs_bit_equal(pm, dist, rank, world, local):
    """One small exact-arithmetic problem (tiled path, T = 4, several tile rows per slab) solved on `world` slabs and,
    on rank 0, on one GPU: the union of the slabs must equal the single-GPU u, v, p bit for bit."""
    import numpy as np
    import torch
    nx, nyr, K, steps = 1100, 600, 22, 1
    ny = nyr * world
    cfg = pm.config_init(pm.CASE_CAVITY, nx, ny)
    cfg.ppe_method, cfg.exact_arith, cfg.kernel_path, cfg.sweeps_per_pass, cfg.max_iters, cfg.device = pm.PPE_SOR_RB, 1, pm.PATH_TILED, 4, K, local
    single = None
    if rank == 0:
        S1 = pm.Solver(cfg)
        S1.fill_random(5, 2.0 ** -4)
        r1 = S1.step(steps)
        single = [S1.download(f) for f in (pm.F_U, pm.F_V, pm.F_P)]
        S1.close()
    mcfg = cfg.copy()
    mcfg.rank, mcfg.nranks = rank, world
    for q, b in enumerate(nccl_id_bytes(pm, dist, rank)):
        mcfg.nccl_id[q] = b
    S = pm.Solver(mcfg)
    S.fill_random(5, 2.0 ** -4)
    r = S.step(steps)
    ok = True
    for q, f in enumerate((pm.F_U, pm.F_V, pm.F_P)):
        a = np.zeros(pm.field_shape(f, nx, ny))
        S.download(f, a)
        t = torch.from_numpy(a.view(np.int64).copy()).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)  # rows are disjoint over ranks and zero elsewhere
        if rank == 0:
            ok = ok and np.array_equal(t.cpu().numpy(), single[q].view(np.int64))
    S.close()
    if rank == 0:
        ok = ok and (r.iterations, r.residual) == (r1.iterations, r1.residual)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(t, 0)
    return bool(t.item()), f"cavity {nx}x{ny} on {world} slabs vs 1 GPU, exact arithmetic, tiled T=4, K={K}: u, v, p, iterations, residual"


def run_ours(args):
    import numpy as np
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

    case = args.case
    if case == "cavity":
        if world == 1:
            nx = ny_local = args.n or 8192
            workload = f"lid-driven cavity Re=1000 {nx}x{nx} on 1 B200 (BASELINE configs[3])"
        else:
            nx = ny_local = args.n or 16384
            workload = f"lid-driven cavity Re=1000 {nx}x{ny_local * world}, {nx}x{ny_local} per GPU, {world} j-slabs (BASELINE configs[4])"
    else:
        nx = args.n or 8192
        ny_local = nx // 4
        workload = f"{case} flow {nx}x{ny_local * world} ({nx}x{ny_local} per GPU), large-grid form of BASELINE configs[{1 if case == 'channel' else 2}]"
    ny = ny_local * world

    def make_solver(path=None):
        cfg = bench_cfg(pm, case, nx, ny, K_ITERS, args, local, path)
        cfg.rank, cfg.nranks = rank, world
        if world > 1:
            for q, b in enumerate(nccl_id_bytes(pm, dist, rank)):
                cfg.nccl_id[q] = b
        S_ = pm.Solver(cfg)
        init_state(S_, case)
        return S_

    S = make_solver()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        S.sync()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
    ms = max_over_ranks(ms)
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
    traffic, traffic_src = None, None
    if world == 1 and args.ppe == "rb" and not args.exact:
        traffic, traffic_src = measured_traffic(f"{case} {nx}x{ny} T={int(round(sweeps_per_pass + 0.4))}")
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
        "frac_real_traffic": (traffic / (ms_per_pass * 1e-3) / 1e9 / peak) if traffic else None,
        "kernel": "pressure pass: k_ppe_stream (interior) + k_ppe_tiled (wall tiles), 4 sweeps + the inf-norm residual of every iterate", "peak_source": peak_src,
        "algorithmic_bytes_per_launch": bytes_per_pass, "ms_per_launch": ms_per_pass, "sweeps_per_launch": sweeps_per_pass,
        "whole_step": {"algorithmic_GBps": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                       "frac_of_measured": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak,
                       "frac_of_nominal_8TBps": step_bytes / (ms / args.steps * 1e-3) / 1e9 / 8000.0,
                       "bytes_per_cell_step": bytes_per_cell_step(K_ITERS)},
    }

    # ---- parity of the benchmarked run: the same steps through the general path (one kernel per colour + a residual
    # pass, itself pinned to the oracle bit for bit at oracle sizes), same inputs; every rank compares its own slab ----
    parity = None
    if not args.no_parity:
        nsteps = args.warmup + args.steps
        mine = {}
        for f in (pm.F_U, pm.F_V, pm.F_P):
            shp = S.slab_rows(f)[1:]
            a = np.empty(shp, dtype=np.float64)
            S.download_slab_ptr(f, a.ctypes.data, a.size)
            mine[f] = a
        G = make_solver(path=pm.PATH_SIMPLE)
        t0 = time.perf_counter()
        rg = G.step(nsteps)
        G.sync()
        g_ms = (time.perf_counter() - t0) * 1e3 / nsteps
        worst, finite = {}, True
        for name, f in (("u", pm.F_U), ("v", pm.F_V), ("p", pm.F_P)):
            b = np.empty_like(mine[f])
            G.download_slab_ptr(f, b.ctypes.data, b.size)
            finite = finite and bool(np.isfinite(mine[f]).all()) and bool(np.isfinite(b).all())
            scale = max_over_ranks(float(np.abs(b).max()))
            worst[name] = max_over_ranks(float(np.abs(mine[f] - b).max())) / (scale if scale > 0 else 1.0)
            del b
        G.close()
        del mine
        finite = max_over_ranks(0.0 if finite else 1.0) == 0.0
        parity = {"against": "general path (kernel_path=1) run of the same steps on the same inputs", "steps": nsteps,
                  "rel_linf": worst, "finite": finite, "iterations_equal": rg.iterations == r.iterations,
                  "residual_rel_diff": abs(rg.residual - r.residual) / rg.residual if rg.residual else 0.0,
                  "general_path_ms_per_step": g_ms, "ok": bool(finite and max(worst.values()) <= (0.0 if args.exact else 1e-9))}
        if world > 1:
            ok, what = slabs_bit_equal(pm, dist, rank, world, local)
            parity["slabs_bit_equal"] = ok
            parity["slabs_case"] = what
            parity["ok"] = bool(parity["ok"] and ok)

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
            return max_over_ranks((time.perf_counter() - t0) * 1e3)

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
        del host, src
    gpu_launches = int(t_after.kernel_launches - t_before.kernel_launches)
    S.close()

    # ---- weak-scaling base: the same per-GPU slab on ONE GPU, timed in this very run (rank 0; the others wait) ----
    weak_base = None
    if world > 1 and not args.no_weak_base:
        if rank == 0:
            cfg1 = bench_cfg(pm, case, nx, ny_local, K_ITERS, args, local)
            S1 = pm.Solver(cfg1)
            init_state(S1, case)
            b_ms, _ = device_timed(S1, max(3, min(args.steps, 20)), max(3, args.warmup))  # as many steps as the run it is the base of (power state)
            S1.close()
            weak_base = {"workload": f"{case} {nx}x{ny_local} on 1 GPU (one slab of the above), same K, same kernel", "ms_per_step": b_ms,
                         "value": nx * ny_local / b_ms / 1e3, "unit": UNIT,
                         "efficiency_vs_this_base": (value / world) / (nx * ny_local / b_ms / 1e3)}
        dist.barrier()

    secondary = small = None
    if rank == 0 and world == 1 and case == "cavity" and not args.no_secondary:
        secondary = secondary_lines(pm, args, local)
        small = small_configs(pm, local, not args.no_cpu)

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
            "config": {"workload": workload, "ppe": f"{args.ppe}, K={K_ITERS} iterations/step (max_iters cap), residual every iteration" + (", tolerance_factor 1e-12" if case == "cavity" else ""),
                       "arith": "exact (no FMA)" if args.exact else "production (FMA)",
                       "kernel_path": {0: "auto", 1: "simple", 2: "tiled"}[args.path],
                       "l2": "inputs larger than L2 (>= 537 MB per field vs 126 MB L2); no flush needed",
                       "parallelism": f"{world} j-slab(s)"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "parity_check": parity,
            "gpu_launches": gpu_launches, "clocks": clocks,
        }
        if weak_base is not None:
            line["weak_base"] = weak_base
        if secondary is not None:
            line["secondary"] = secondary
            line["small_configs"] = small
        emit(line)
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
    ap.add_argument("--case", default="cavity", choices=["cavity", "channel", "step"], help="which reference solver's form of the step (default: the BASELINE cavity)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-weak-base", action="store_true")
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
