"""Where does the large-tier kernel spend its time?  Repeat launches, vary the trip budget."""
import sys, time; sys.path.insert(0, ".")
import numpy as np, torch, optas_b200
from optas_b200 import problems
prob = problems.dual_arm()
P, X0 = prob.sample(1024)
for mt in (1, 2, 4, 8, 250):
    s = optas_b200.B200Solver(prob.opt).setup("ipopt", {"max_trips": mt}, timing=True)
    for rep in range(3):
        t = time.time(); r = s.solve_arrays(P, X0); dt = time.time() - t
        ms, n = s._handle.kernel_time()
        print(f"max_trips {mt:3d} rep {rep}: wall {dt*1e3:8.1f} ms kernel {ms/n:8.1f} ms trips-done status {np.bincount(r['status'], minlength=5)}", flush=True)
