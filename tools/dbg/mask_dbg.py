import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "computational-fluid-dynamics_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pm_ctypes as pm, orc
from test_gpu_parity import blob_mask, make_cfg
nx, ny, T, seed = 300, 131, 4, 1
m = blob_mask(nx, ny, seed)
for K in (1, 2, 4, 5, 8, 9, 21):
    cfg = make_cfg(pm, 2, nx, ny, 1, 1, K, path=2); cfg.sweeps_per_pass = T
    S, O = pm.Solver(cfg), orc.Oracle(cfg)
    S.upload_mask(m); O.mask()[:] = m
    S.fill_random(9, 2.0 ** -3); O.fill_random(9, 2.0 ** -3)
    rs, ro = S.ppe_solve(), O.ppe_solve()
    a, b = S.download(2), O.field(2)
    d = np.abs(a - b)
    bad = np.argwhere(d > 0)
    print(f"K={K}: res {rs.residual!r} vs {ro.residual!r}  ndiff={len(bad)}")
    for (j, i) in bad[:12]:
        kind = "ghost" if (j in (0, ny + 1) or i in (0, nx + 1)) else ("fluid" if m[j, i] else "solid")
        nb = "".join(str(int(m[jj, ii])) for jj, ii in ((j, i - 1), (j, i + 1), (j - 1, i), (j + 1, i)) if 0 <= jj <= ny + 1 and 0 <= ii <= nx + 1)
        print(f"   (j={j}, i={i}) {kind} nbWESN={nb} gpu={a[j,i]!r} orc={b[j,i]!r}  tile bx={(i-1)//112} by={(j-1)//32} col-in-block={(i-1)%112} row-in-block={(j-1)%32}")
    S.close()
