"""C3 on the GPU: timing of the thread-per-instance tier on the MPC problem."""
import sys, time; sys.path.insert(0, ".")
import numpy as np, torch, optas_b200
from optas_b200 import problems
prob = problems.point_mass_mpc()
t = time.time(); s = optas_b200.B200Solver(prob.opt).setup("ipopt", timing=True); print("setup", time.time() - t, s.kernel_info(), flush=True)
for B in (1024, 16384):
    P, X0 = prob.sample(B, 1)
    t = time.time(); r = s.solve_arrays(P, X0); dt = time.time() - t
    ms, n = s._handle.kernel_time()
    ok = r["status"] <= 1
    print(f"C3 B={B}: wall {dt*1e3:.1f} ms kernel {ms/n:.1f} ms conv {ok.mean():.4f} iters mean {r['iters'][ok].mean():.1f} -> {ok.sum()/(ms/n)*1e3:.3e} inst/s", flush=True)
