import os, sys, time
sys.path.insert(0, os.path.join(os.getcwd(), "computational-fluid-dynamics_b200"))
import pm_ctypes as pm
CONFIGS = [("cfg0 cavity 128^2", pm.CASE_CAVITY, (128, 128, 100.0, 1e-3)), ("cfg1 channel 256x64", pm.CASE_CHANNEL, (256, 64, 1000.0, 5e-4)), ("cfg2 step 256x32", pm.CASE_STEP, (0, 0, 0.0, 0.0))]
for name, case, a in CONFIGS:
    cfg = pm.config_init(case, *a); cfg.ppe_method = pm.PPE_SOR_RB
    S = pm.Solver(cfg); S.apply_bc(0); S.step(2); S.sync()
    t0 = time.perf_counter(); it = 0
    for _ in range(10): it += S.step(1).iterations
    S.sync(); dt = time.perf_counter() - t0
    print(f"{name}: {1e3*dt/10:.2f} ms/step, {it/10:.0f} it, {1e6*dt/it:.2f} us/it", flush=True)
    S.close()
