"""CPU model of the streaming pressure pass's schedule (csrc/pm_kernels_stream.cuh): the 2T = 8 colour half-sweeps
of a pass as a pipeline over the rows of a strip -- at tick tau row tau enters, half-sweep h works on row tau - h (h ascending,
north neighbour as half-sweep h - 1 left it in this tick, south neighbour as half-sweep h + 1 left it in the previous one) --
must give, on the strip's output cells (H = 8 columns in from either side, H rows in from either end of the chunk), exactly what
four red-black sweeps over the whole grid give, although the strip only ever sees its own 128 columns and R + 16 rows.
numpy, every operation individually rounded: equality is bitwise.  No GPU, no library."""
import numpy as np
import pytest

H, W, NST = 8, 128, 8


def relax(p, e, w, n, s, f, idx2, cw):
    """Interior cell of the cavity form in residual form, neighbours summed in a fixed association (rb_half_lean)."""
    r = idx2 * (((e + n) + (w + s)) - 4.0 * p) - f
    return p + cw * r


def global_sweeps(P, F, idx2, cw, nhalf):
    """nhalf colour half-sweeps over the interior of P (its outer ring is boundary data), colour 0 = (i + j) even first."""
    P = P.copy()
    J, I = np.meshgrid(np.arange(P.shape[0]), np.arange(P.shape[1]), indexing="ij")
    inner = np.zeros_like(P, dtype=bool)
    inner[1:-1, 1:-1] = True
    for h in range(nhalf):
        m = inner & (((I + J) & 1) == (h & 1))
        new = relax(P, np.roll(P, -1, 1), np.roll(P, 1, 1), np.roll(P, -1, 0), np.roll(P, 1, 0), F, idx2, cw)
        P[m] = new[m]
    return P


def stream_strip(P0, F, j0, i0, R, idx2, cw):
    """The pipeline over rows j0 - H .. j0 + R + H - 1 and columns i0 - H .. i0 - H + W - 1 of the initial field P0.
    Returns the R x (W - 2H) output block."""
    nrows = R + 2 * H
    rows = {}                        # rows in flight: index -> current values of the strip's W columns
    out = np.zeros((R, W - 2 * H))
    cols = np.arange(i0 - H, i0 - H + W)
    garbage = np.full(W, 1.0e300)    # whatever lies outside the strip / the chunk: must never reach an output cell

    def load(t):
        return P0[j0 - H + t, cols].copy() if 0 <= t < nrows else garbage.copy()

    for tau in range(nrows - 1 + NST):          # (the kernel stops once the last output row has left: the rest is idle here)
        rows[tau] = load(tau)
        north_raw = load(tau + 1)
        for h in range(NST):                    # ascending: half-sweep h reads what half-sweep h - 1 produced in this tick
            t = tau - h
            if t < 0 or t >= nrows:
                continue
            j = j0 - H + t
            cur = rows[t]
            north = north_raw if h == 0 else rows.get(t + 1, garbage)
            south = rows.get(t - 1, garbage)
            east = np.append(cur[1:], garbage[0])   # beyond the strip: the neighbour lane does not exist
            west = np.append(garbage[0], cur[:-1])
            tgt = ((cols + j) & 1) == (h & 1)
            new = relax(cur, east, west, north, south, F[j, cols], idx2, cw)
            with np.errstate(all="ignore"):
                cur[tgt] = new[tgt]
        t = tau - (NST - 1)                     # this row has passed all half-sweeps
        if H <= t < H + R:
            out[t - H] = rows[t][H:W - H]
        rows.pop(tau - NST, None)               # only 8 rows (+ the one below half-sweep 7's) are ever held
    return out


@pytest.mark.parametrize("R,j0,i0", [(32, 9, 9), (64, 40, 23), (96, 12, 140)])
def test_streaming_schedule_equals_four_global_sweeps(R, j0, i0):
    rng = np.random.default_rng(R)
    ny, nx = j0 + R + 3 * H, i0 + W + 2 * H
    P0 = rng.uniform(-1, 1, (ny, nx))
    F = rng.uniform(-1, 1, (ny, nx))
    idx2, cw = 64.0 ** 2, 1.9 / (4 * 64.0 ** 2)
    want = global_sweeps(P0, F, idx2, cw, NST)[j0:j0 + R, i0:i0 + W - 2 * H]
    got = stream_strip(P0, F, j0, i0, R, idx2, cw)
    assert np.isfinite(got).all(), "something outside the strip reached an output cell"
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


def test_halo_of_seven_is_not_enough():
    """The same pipeline with the output block widened by one ring must differ somewhere: H = 2T is tight."""
    global H
    R, j0, i0 = 32, 9, 9
    rng = np.random.default_rng(1)
    ny, nx = j0 + R + 4 * H, i0 + W + 3 * H
    P0 = rng.uniform(-1, 1, (ny, nx))
    F = rng.uniform(-1, 1, (ny, nx))
    idx2, cw = 64.0 ** 2, 1.9 / (4 * 64.0 ** 2)
    full = global_sweeps(P0, F, idx2, cw, NST)
    old = H
    try:
        H = 7
        got = stream_strip(P0, F, j0, i0, R, idx2, cw)
        want = full[j0:j0 + R, i0:i0 + W - 2 * H]
        with np.errstate(all="ignore"):
            assert not np.array_equal(got, want)
    finally:
        H = old
