"""Convergence of the problems where the feasibility-restoration phase matters, status counts per problem."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optas_b200
from optas_b200 import problems

out = {}
for name, B, opts, zero in (("point_mass_mpc", 16384, {}, True), ("joint_space_planner", 8192, {}, False), ("lwr_axis_ik", 65536, {}, False),
                            ("sphere_collision_avoidance", 256, {"max_iter": 300, "max_trips": 1500}, False), ("lwr_ik", 65536, {}, False)):
    prob = getattr(problems, name)()
    P, X0 = prob.sample(B)
    if zero:
        X0 = np.zeros_like(X0)
    s = optas_b200.B200Solver(prob.opt).setup("ipopt", opts, timing=True)
    r = s.solve_arrays(P, X0)
    ms, n = s._handle.kernel_time()
    out[name + (" (zero seed)" if zero else "")] = {"B": B, "tier": s.tier_info()["tier"], "status_counts": [int((r["status"] == k).sum()) for k in range(5)],
                                                    "converged_fraction": float((r["status"] == 0).mean()), "mean_iters": float(r["iters"].mean()), "kernel_ms": ms / n}
print(json.dumps(out))
