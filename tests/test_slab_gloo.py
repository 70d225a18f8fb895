"""CPU tests of the N > 1 path (world_size 2 and 3, gloo): the slab decomposition in j, the halo-row
exchange schedule and the max-allreduce of the residual that libpm.so performs with NCCL, replayed
with the oracle's row-range kernels and torch.distributed send/recv.  The union of the slabs must equal
the single-domain oracle bit for bit (SURVEY §8e "exactness"), for

  * the general path: one halo row after every colour half-sweep, residual allreduced every iteration;
  * the tiled path: a halo H = 2T rows deep exchanged once per pass of T red-black sweeps, every rank
    recomputing the rows of its neighbours that fall inside its shrinking valid region.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup():
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "computational-fluid-dynamics_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import orc
    import pm_ctypes
    return orc, pm_ctypes


def _exchange(p, ja, jb, depth, rank, world):
    """Rows jb-depth+1..jb go up into the upper rank's rows below its ja; rows ja..ja+depth-1 go down."""
    reqs, bufs = [], []
    if rank + 1 < world:
        up = torch.from_numpy(p[jb - depth + 1:jb + 1].copy())
        rcv = torch.empty_like(up)
        reqs += [dist.isend(up, rank + 1), dist.irecv(rcv, rank + 1)]
        bufs.append(("up", rcv))
    if rank > 0:
        dn = torch.from_numpy(p[ja:ja + depth].copy())
        rcv = torch.empty_like(dn)
        reqs += [dist.isend(dn, rank - 1), dist.irecv(rcv, rank - 1)]
        bufs.append(("dn", rcv))
    for r in reqs:
        r.wait()
    for where, rcv in bufs:
        if where == "up":
            p[jb + 1:jb + 1 + depth] = rcv.numpy()
        else:
            p[ja - depth:ja] = rcv.numpy()


def _allreduce_max(x):
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _worker(rank, world, port, case_id, nx, ny, K, T, mode, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc, pm = _setup()
    import ctypes as C
    cfg = orc.config_init(case_id, nx, ny)
    cfg.ppe_method, cfg.max_iters = 1, K
    j0, nyl = C.c_int(), C.c_int()
    assert pm.lib().pm_slab_range(ny, world, rank, C.byref(j0), C.byref(nyl)) == 0  # the product's own partition
    ja, jb = j0.value + 1, j0.value + nyl.value
    O = orc.Oracle(cfg)
    O.fill_random(99)
    p = O.field(2)
    if case_id == 0:
        p[:] = 0.0  # cavity cold start
    # every rank starts from the same global field but may only trust its own rows +- halo afterwards:
    # poison the rest so that any read outside the exchanged halos shows up as a mismatch
    H = 2 * T if mode == "tiled" else 1
    lo, hi = max(ja - H, 0), min(jb + H, ny + 1)
    if rank > 0:
        p[:lo] = np.nan
    if rank + 1 < world:
        p[hi + 1:] = np.nan
    res_hist = []
    if mode == "general":
        for k in range(K):
            O.sweep_rows(0, ja, jb)
            _exchange(p, ja, jb, 1, rank, world)
            O.sweep_rows(1, ja, jb)
            if case_id != 0:
                _ghosts(O, p, nx, ny, ja, jb, rank, world)
            if case_id == 2:
                # the solid cells of my edge rows extrapolate from the neighbour slab's fresh fluid values
                # (pm_capi.cu launch_iteration_simple: exchange, k_pghost_solid, exchange)
                _exchange(p, ja, jb, 1, rank, world)
                _solid_ghosts(O.mask(), p, nx, ny, ja, jb)
            _exchange(p, ja, jb, 1, rank, world)
            res_hist.append(_allreduce_max(O.residual_rows(ja, jb)))
    else:
        k = 0
        while k < K:
            nsw = min(T, K - k)
            for t in range(nsw):
                for colour in (0, 1):
                    s = 2 * t + colour + 1  # half-sweeps done after this one; valid region shrinks by one row each
                    a = 1 if rank == 0 else ja - H + s
                    b = ny if rank + 1 == world else jb + H - s
                    O.sweep_rows(colour, max(a, 1), min(b, ny))
                if case_id != 0:
                    _ghosts(O, p, nx, ny, max(ja - H + 2 * t + 2, 1), min(jb + H - 2 * t - 2, ny), rank, world)
                if t < nsw - 1:  # rows ja-1 / jb+1 are still inside the valid region
                    res_hist.append(_allreduce_max(O.residual_rows(ja, jb)))
            k += nsw
            _exchange(p, ja, jb, H, rank, world)
            # the last iterate of the pass: its neighbours' rows arrive with the exchange (in the kernel the
            # black part is taken right after the update, the red part by the next pass before its first update)
            res_hist.append(_allreduce_max(O.residual_rows(ja, jb)))
    np.save(os.path.join(out_dir, f"p_{rank}.npy"), p[ja - (1 if rank == 0 else 0):jb + 1 + (1 if rank + 1 == world else 0)].copy())
    if rank == 0:
        np.save(os.path.join(out_dir, "res.npy"), np.array(res_hist))
    dist.destroy_process_group()


def _ghosts(O, p, nx, ny, a, b, rank, world):
    """applyPressureGhosts (channel-01.cpp:531-541) restricted to the rows a rank holds."""
    p[a:b + 1, 0] = p[a:b + 1, 1]
    p[a:b + 1, nx + 1] = 0.0
    if rank == 0:
        p[0, 1:nx + 1] = p[1, 1:nx + 1]
    if rank + 1 == world:
        p[ny + 1, 1:nx + 1] = p[ny, 1:nx + 1]


def _solid_ghosts(mask, p, nx, ny, a, b):
    """Solid-cell extrapolation of applyPressureGhosts (backwards_step-01.cpp:709-739) on rows a..b: only solid
    cells are written and only fluid cells are read, so the visiting order does not matter."""
    for j in range(a, b + 1):
        for i in range(1, nx + 1):
            if mask[j, i]:
                continue
            s, n = 0.0, 0
            if i > 1 and mask[j, i - 1]:
                s += p[j, i - 1]; n += 1
            if i < nx and mask[j, i + 1]:
                s += p[j, i + 1]; n += 1
            if j > 1 and mask[j - 1, i]:
                s += p[j - 1, i]; n += 1
            if j < ny and mask[j + 1, i]:
                s += p[j + 1, i]; n += 1
            if n > 0:
                p[j, i] = s / n


def test_step_case_slabs_equal_single_domain(tmp_path):
    """The obstacle mask with the slab cut exactly on inlet_j_max (the top wall of the inlet channel; 2 slabs, even ny)
    and off it (3 slabs): general-path schedule with the extra exchange around the solid-cell extrapolation."""
    orc, _ = _setup()
    for world in (2, 3):
        case_id, nx, ny, K = 2, 32, 24, 6
        out = tmp_path / f"w{world}"
        out.mkdir()
        port = 29500 + (os.getpid() * 11 + world * 17) % 2000
        mp.spawn(_worker, args=(world, port, case_id, nx, ny, K, 1, "general", str(out)), nprocs=world, join=True)
        cfg = orc.config_init(case_id, nx, ny)
        assert cfg.inlet_j_max == ny // 2
        cfg.ppe_method, cfg.max_iters = 1, K
        O = orc.Oracle(cfg)
        O.fill_random(99)
        p = O.field(2)
        res = []
        for k in range(K):
            O.sweep_rows(0, 1, ny)
            O.sweep_rows(1, 1, ny)
            O.pressure_ghosts()
            res.append(O.residual_rows(1, ny))
        got = np.concatenate([np.load(out / f"p_{r}.npy") for r in range(world)], axis=0)
        a, b = got.copy(), p.copy()
        for arr in (a, b):
            arr[0, 0] = arr[0, -1] = arr[-1, 0] = arr[-1, -1] = 0.0
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
        assert np.array_equal(np.load(out / "res.npy"), np.array(res))


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case_id,nx,ny", [(0, 20, 26), (1, 30, 23)])
@pytest.mark.parametrize("mode,T", [("general", 1), ("tiled", 2), ("tiled", 3), ("tiled", 4)])
def test_slabs_equal_single_domain(tmp_path, world, case_id, nx, ny, mode, T):
    orc, _ = _setup()
    K = 7
    port = 29500 + (os.getpid() * 7 + world * 13 + case_id * 5 + T) % 2000
    mp.spawn(_worker, args=(world, port, case_id, nx, ny, K, T, mode, str(tmp_path)), nprocs=world, join=True)
    # single-domain oracle
    cfg = orc.config_init(case_id, nx, ny)
    cfg.ppe_method, cfg.max_iters = 1, K
    O = orc.Oracle(cfg)
    O.fill_random(99)
    p = O.field(2)
    if case_id == 0:
        p[:] = 0.0
    res = []
    for k in range(K):
        O.sweep_rows(0, 1, ny)
        O.sweep_rows(1, 1, ny)
        if case_id != 0:
            O.pressure_ghosts()
        res.append(O.residual_rows(1, ny))
    got = np.concatenate([np.load(tmp_path / f"p_{r}.npy") for r in range(world)], axis=0)
    assert got.shape == p.shape
    # corners are never written by any sweep; the slabs poison what they do not own
    a, b = got.copy(), p.copy()
    for arr in (a, b):
        arr[0, 0] = arr[0, -1] = arr[-1, 0] = arr[-1, -1] = 0.0
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    assert np.array_equal(np.load(tmp_path / "res.npy"), np.array(res))
