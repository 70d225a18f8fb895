"""One projection step of configs[0] (cavity 128x128 Re=100 dt=1e-3) through the cluster solve, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "computational-fluid-dynamics_b200"))
import pm_ctypes as pm
cfg = pm.config_init(pm.CASE_CAVITY, 128, 128, 100.0, 1e-3)
cfg.ppe_method = pm.PPE_SOR_RB
S = pm.Solver(cfg)
S.apply_bc(0)
r = S.step(2)
print("iterations", r.iterations, "residual", r.residual)
