"""Status / iteration statistics of the C2 batch (bench workload), for tail analysis."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import optas_b200
from optas_b200 import problems

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
prob = problems.lwr_ik()
P, X0 = prob.sample(B, seed=0)
s = optas_b200.B200Solver(prob.opt).setup("ipopt", timing=True)
r = s.solve_arrays(P, X0)
r = s.solve_arrays(P, X0)
ms, n = s._handle.kernel_time()
st, it = r["status"], r["iters"]
out = {"B": B, "ms_per_launch": ms / n, "status_counts": {int(k): int((st == k).sum()) for k in range(5)},
       "iters_percentiles": {str(q): float(np.percentile(it, q)) for q in (50, 90, 99, 99.9, 100)},
       "iters_mean": float(it.mean()), "tier": s.tier_info()["tier"], "kernel": s.kernel_info()}
bad = np.where(st >= 2)[0]
out["bad_kkt"] = [float(v) for v in r["kkt"][bad][:20]]
out["bad_iters"] = [int(v) for v in it[bad][:20]]
print(json.dumps(out))
np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "c2_bad.npz"), P=P[bad], X0=X0[bad], X=r["x"][bad], st=st[bad], it=it[bad], kkt=r["kkt"][bad])
