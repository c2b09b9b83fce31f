"""One launch of the cooperative tier (for ncu).  usage: python tools/coop_one.py c3|c4|c5 B [tpb]"""
import sys; sys.path.insert(0, ".")
import numpy as np, optas_b200
from optas_b200 import problems
CASES = {"c3": (problems.point_mass_mpc, {}), "c5": (problems.dual_arm, {}),
         "c4": (problems.figure_eight, {"max_iter": 400, "max_trips": 2500})}
mk, opts = CASES[sys.argv[1]]
B = int(sys.argv[2]); tpb = int(sys.argv[3]) if len(sys.argv) > 3 else 0
prob = mk()
s = optas_b200.B200Solver(prob.opt).setup("ipopt", opts, coop=True, threads_per_block=tpb, timing=True)
P, X0 = prob.sample(B, seed=1)
r = s.solve_arrays(P, X0)
ms, n = s._handle.kernel_time()
print(sys.argv[1], B, "kernel ms", ms / n, "status", np.bincount(r["status"], minlength=5), "iters", r["iters"].mean())
