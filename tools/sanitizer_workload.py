import sys
sys.path.insert(0,'oracle'); sys.path.insert(0,'computational-fluid-dynamics_b200')
import numpy as np, pm_ctypes as pm
# small cases through every kernel family: general path (all three cases), tiled (boundary + interior tiles, T=2,3), lex
for case,nx,ny in ((0,40,40),(1,50,20),(2,64,16)):
    for meth in (0,1,2):
        cfg=pm.config_init(case,nx,ny)
        if case==2: cfg.step_i_location, cfg.inlet_j_max = 16, 8
        cfg.ppe_method=meth; cfg.max_iters=6; cfg.exact_arith=1; cfg.kernel_path=1
        if meth==0: cfg.omega=1.0
        S=pm.Solver(cfg); S.fill_random(1); S.apply_bc(0); S.step(2); S.diagnostics(); S.download(2); S.close()
for case,nx,ny,T in ((0,300,150,2),(0,300,150,3),(1,300,150,3),(0,131,77,3)):
    for meth in (0,1):
        cfg=pm.config_init(case,nx,ny); cfg.ppe_method=meth; cfg.max_iters=7; cfg.exact_arith=0; cfg.kernel_path=2; cfg.sweeps_per_pass=T if meth==1 else 2
        if meth==0: cfg.omega=1.0
        S=pm.Solver(cfg); S.fill_random(1); S.apply_bc(0); S.step(2); S.download(2); S.close()
print("sanitizer workload done")
