"""Multi-GPU parity check, one process per GPU (run under torchrun on a box with >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py

Every rank owns a j-slab of the same global problem (NCCL halo rows + residual all-reduce inside libpm.so);
rank 0 also solves the whole problem on its own GPU.  With exact_arith=1 the union of the slabs must equal the
single-GPU fields bit for bit, and iteration counts / residuals must agree (SURVEY 8e "exactness").
tests/test_gpu_parity.py::test_multi_gpu_slabs_match_single_gpu launches this when 2 GPUs are visible.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "computational-fluid-dynamics_b200"))
import pm_ctypes as pm  # noqa: E402


def nccl_id(rank):
    t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = (C.c_uint8 * 128)()
        assert pm.lib().pm_nccl_unique_id(buf) == 0, pm.lib().pm_last_error(None)
        t = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    return t.cpu().tolist()


def union(arr):
    """Rows are disjoint across ranks and zero elsewhere: summing the int64 bit patterns rebuilds the field."""
    t = torch.from_numpy(arr.view(np.int64).copy()).cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().view(np.float64)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = [
        # case, nx, ny_per_rank, method, exact, path, T, K, steps
        (pm.CASE_CAVITY, 300, 70, pm.PPE_SOR_RB, 1, pm.PATH_SIMPLE, 0, 30, 2),
        (pm.CASE_CAVITY, 300, 70, pm.PPE_JACOBI, 1, pm.PATH_SIMPLE, 0, 30, 2),
        (pm.CASE_CAVITY, 300, 70, pm.PPE_SOR_RB, 1, pm.PATH_TILED, 2, 31, 2),
        (pm.CASE_CAVITY, 300, 71, pm.PPE_SOR_RB, 1, pm.PATH_TILED, 3, 31, 2),   # odd slab height: odd row parity on rank 1
        (pm.CASE_CAVITY, 300, 70, pm.PPE_JACOBI, 1, pm.PATH_TILED, 2, 31, 2),
        (pm.CASE_CHANNEL, 300, 50, pm.PPE_SOR_RB, 0, pm.PATH_SIMPLE, 0, 30, 2),  # fast policy: the mean is a tree sum
        (pm.CASE_CHANNEL, 300, 50, pm.PPE_SOR_RB, 0, pm.PATH_TILED, 2, 30, 2),
        (pm.CASE_CAVITY, 1100, 600, pm.PPE_SOR_RB, 1, pm.PATH_TILED, 3, 20, 1),   # several tile rows per slab: edge/interior launches
        (pm.CASE_CAVITY, 1100, 600, pm.PPE_SOR_RB, 1, pm.PATH_TILED, 4, 22, 1),   # deepest halo (8 rows = all pad rows)
        (pm.CASE_CHANNEL, 600, 333, pm.PPE_SOR_RB, 0, pm.PATH_AUTO, 0, 21, 2),    # production path as the drivers run it: T = 4, split-row buffers, row kernels
        (pm.CASE_CAVITY, 64, 64, pm.PPE_SOR_RB, 1, pm.PATH_SIMPLE, 0, 10000, 2),  # run to tolerance: same stopping iterate
        (pm.CASE_CAVITY, 200, 70, pm.PPE_SOR_RB, 1, pm.PATH_TILED, 3, 20000, 1),  # ... on the tiled path: device-side loop test over slabs, replay of the partial pass
        (pm.CASE_CHANNEL, 200, 60, pm.PPE_SOR_RB, 0, pm.PATH_TILED, 4, 30000, 1),
        # exact arithmetic for channel/step: the reference's serial source mean is continued from slab to slab
        (pm.CASE_CHANNEL, 300, 50, pm.PPE_SOR_RB, 1, pm.PATH_SIMPLE, 0, 30, 2),
        (pm.CASE_CHANNEL, 300, 50, pm.PPE_SOR_RB, 1, pm.PATH_TILED, 4, 30, 2),
        # obstacle mask; with 2 ranks and even ny the slab cut falls exactly on inlet_j_max (the top wall of the inlet channel)
        (pm.CASE_STEP, 256, 16, pm.PPE_SOR_RB, 1, pm.PATH_SIMPLE, 0, 30, 3),
        (pm.CASE_STEP, 256, 16, pm.PPE_JACOBI, 1, pm.PATH_SIMPLE, 0, 30, 3),
        (pm.CASE_STEP, 320, 45, pm.PPE_SOR_RB, 0, pm.PATH_AUTO, 0, 25, 2),
        (pm.CASE_STEP, 256, 16, pm.PPE_SOR_RB, 1, pm.PATH_SIMPLE, 0, 10000, 2),
        # obstacle mask on the tiled path: fluid-only, solid-only and mixed tiles, several tile rows per slab, 8-row halos
        (pm.CASE_STEP, 1100, 300, pm.PPE_SOR_RB, 1, pm.PATH_TILED, 4, 22, 2),
        (pm.CASE_STEP, 1100, 301, pm.PPE_SOR_RB, 1, pm.PATH_TILED, 3, 20, 2),
        (pm.CASE_STEP, 900, 260, pm.PPE_SOR_RB, 0, pm.PATH_AUTO, 0, 25, 2),
        # production path with the streaming pass on every slab (edge tile rows and the frame on k_ppe_tiled, the rest streamed)
        (pm.CASE_CAVITY, 1400, 704, pm.PPE_SOR_RB, 0, pm.PATH_AUTO, 0, 22, 2),
        (pm.CASE_CHANNEL, 1400, 705, pm.PPE_SOR_RB, 0, pm.PATH_AUTO, 0, 21, 1),   # odd slab height: odd row parity on rank 1
    ]
    failures = 0
    for (case, nx, nyr, method, exact, path, T, K, steps) in cases:
        ny = nyr * world
        cfg = pm.config_init(case, nx, ny)
        cfg.ppe_method, cfg.exact_arith, cfg.kernel_path, cfg.sweeps_per_pass, cfg.max_iters = method, exact, path, T, K
        if method == pm.PPE_JACOBI:
            cfg.omega = 0.9
        if nx >= 1400:
            cfg.tol_factor = 1e-13  # large grids: keep the reference's loop-entry rule (1.0 > tolerance) from skipping the solve
        cfg.device = local
        single = None
        if rank == 0:
            S1 = pm.Solver(cfg)
            S1.fill_random(5, 2.0 ** -4)
            S1.apply_bc(0)
            r1 = S1.step(steps)
            single = [S1.download(f) for f in (pm.F_U, pm.F_V, pm.F_P)]
            S1.close()
        mcfg = cfg.copy()
        mcfg.rank, mcfg.nranks = rank, world
        for q, b in enumerate(nccl_id(rank)):
            mcfg.nccl_id[q] = b
        S = pm.Solver(mcfg)
        S.fill_random(5, 2.0 ** -4)
        S.apply_bc(0)
        r = S.step(steps)
        md, ke = S.diagnostics()
        got = []
        for f in (pm.F_U, pm.F_V, pm.F_P):
            a = np.zeros(pm.field_shape(f, nx, ny))
            S.download(f, a)
            got.append(union(a))
        S.close()
        if rank == 0:
            ok = (r.iterations, r.residual) == (r1.iterations, r1.residual) if exact else r.iterations == r1.iterations
            worst = 0.0
            for a, b in zip(got, single):
                if exact:
                    ok = ok and np.array_equal(a.view(np.uint64), b.view(np.uint64))
                else:
                    scale = max(1.0, np.abs(b).max())
                    worst = max(worst, np.abs(a - b).max() / scale)
                    ok = ok and np.abs(a - b).max() <= 1e-12 * scale
            print(f"case={case} {nx}x{ny} method={method} exact={exact} path={path} T={T} K={K}: "
                  f"iters {r.iterations}/{r1.iterations} res {r.residual:.6e}/{r1.residual:.6e} worst_rel={worst:.2e} -> {'OK' if ok else 'MISMATCH'}", flush=True)
            failures += 0 if ok else 1
    t = torch.tensor([failures], device="cuda")
    dist.broadcast(t, 0)
    dist.destroy_process_group()
    if int(t.item()):
        sys.exit(1)
    if rank == 0:
        print("mgpu_check: all slab runs match the single-GPU fields")


if __name__ == "__main__":
    main()
