"""Phase breakdown of the tiled pressure kernel (debug build):
   cd computational-fluid-dynamics_b200 && make variant NAME=prof NSEG=5 RPT=8 MINB=2 EXTRA=-DPM_TILE_PROFILE
   PM_LIB=computational-fluid-dynamics_b200/lib/libpm_prof.so python tools/tile_profile.py [n] [sweeps]
Prints the mean cycles a CTA spends in each phase (thread 0's clock)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "computational-fluid-dynamics_b200"))
import pm_ctypes as pm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cfg = pm.config_init(pm.CASE_CAVITY, n, n, 1000.0, 0.0)
cfg.max_iters, cfg.tol_factor, cfg.ppe_method, cfg.sweeps_per_pass = 99, 1e-12, pm.PPE_SOR_RB, T
S = pm.Solver(cfg)
S.fill_random(42, 2.0 ** -10)
S.step(1)
S.sync()
a, b, c = C.c_int(), C.c_int(), C.c_int()
if pm.lib().pm_debug_tiled_occupancy(S._h, C.byref(a), C.byref(b), C.byref(c)) == 0:
    print(f"occupancy: {a.value} CTAs/SM, max active clusters {b.value} of size {c.value} ({b.value * c.value} CTAs = {b.value * c.value / 148:.2f} per SM)")
buf = (C.c_ulonglong * 8)()
pm.lib().pm_debug_tile_profile(buf, 1)
S.timer_start()
S.step(2)
ms = S.timer_stop()
pm.lib().pm_debug_tile_profile(buf, 0)
ctas = buf[7]
names = ["masks + f loads issued", "wait for the TMA tile", "own cells out of the tile", "sweeps", "write-out + residual atomics"]
names.append("kernel entry: mbarrier, TMA issue, loop-test loads, barrier")
tot = sum(buf[q] for q in range(6))
print(f"{n}x{n}: {ms / 2:.2f} ms/step, {ctas} tile CTAs, {tot / ctas:.0f} cycles per CTA")
print(f"  (inside the sweeps) waiting for the neighbour CTAs' edge rows, lead threads of the two edge segments together: {buf[6] / ctas:9.0f} cycles per CTA")
for q, nm in enumerate(names):
    print(f"  {nm:45s} {buf[q] / ctas:9.0f} cycles  {100.0 * buf[q] / tot:5.1f} %")
