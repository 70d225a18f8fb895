import sys, os
os.environ["PM_DEBUG_NO_FIXUP"] = "1"
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "computational-fluid-dynamics_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pm_ctypes as pm, orc
from test_gpu_parity import blob_mask, make_cfg
nx, ny, T, seed = 300, 131, 4, 1
m = blob_mask(nx, ny, seed)
interior = np.zeros_like(m, dtype=bool); interior[1:ny+1, 1:nx+1] = True
for K in (4, 5, 6, 7, 8):
    cfg = make_cfg(pm, 2, nx, ny, 1, 1, K, path=2); cfg.sweeps_per_pass = T
    S = pm.Solver(cfg); S.upload_mask(m); S.fill_random(9, 2.0 ** -3)
    S.ppe_solve(); a = S.download(2); S.close()
    # expected: fluid cells at iterate K, solid cells + nothing else lagging: oracle after K-1 full iterations gives solids; after K gives fluid
    ocfg = make_cfg(pm, 2, nx, ny, 1, 1, K - 1, path=2)
    O1 = orc.Oracle(ocfg); O1.mask()[:] = m; O1.fill_random(9, 2.0 ** -3); O1.ppe_solve(); lag = O1.field(2).copy()
    ocfg2 = make_cfg(pm, 2, nx, ny, 1, 1, K, path=2)
    O2 = orc.Oracle(ocfg2); O2.mask()[:] = m; O2.fill_random(9, 2.0 ** -3); O2.ppe_solve(); cur = O2.field(2).copy()
    fluid = (m == 1) & interior; solid = (m == 0) & interior
    bf = np.argwhere(fluid & (a != cur)); bs = np.argwhere(solid & (a != lag))
    print(f"K={K}: fluid cells wrong {len(bf)}, solid cells not equal to the lagged ghost pass {len(bs)} (of {solid.sum()})")
    for (j, i) in list(bf[:6]) + list(bs[:10]):
        kind = "fluid" if m[j, i] else "solid"
        nb = "".join(str(int(m[jj, ii])) for jj, ii in ((j, i - 1), (j, i + 1), (j - 1, i), (j + 1, i)))
        print(f"   (j={j}, i={i}) {kind} nbWESN={nb} gpu={a[j,i]!r} lag={lag[j,i]!r} cur={cur[j,i]!r} col-in-block={(i-1)%112} row-in-block={(j-1)%32}")
