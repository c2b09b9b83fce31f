"""Team tier vs the round-1 thread-per-instance kernel on the C2 batch: same results, kernel time of each."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optas_b200
from optas_b200 import problems

name = sys.argv[1] if len(sys.argv) > 1 else "lwr_ik"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
kw = {}
if len(sys.argv) > 3:
    kw["threads_per_block"] = int(sys.argv[3])
prob = getattr(problems, name)()
P, X0 = prob.sample(B, seed=0)
out = {"problem": name, "B": B}
res = {}
for label, team in (("team", None), ("thread", False)):
    s = optas_b200.B200Solver(prob.opt).setup("ipopt", timing=True, team=team, **(kw if team is None else {}))
    r = s.solve_arrays(P, X0)
    s._handle.kernel_time()
    for _ in range(5):
        r = s.solve_arrays(P, X0)
    ms, n = s._handle.kernel_time()
    res[label] = r
    out[label] = {"tier": s.tier_info()["tier"], "ms_per_launch": ms / n, "kernel": s.kernel_info(),
                  "blocks_per_sm": s.tier_info()["blocks_per_sm"], "smem": s.tier_info()["smem_dynamic"],
                  "status_counts": [int((r["status"] == k).sum()) for k in range(5)], "mean_iters": float(r["iters"].mean())}
a, b = res["team"], res["thread"]
ok = (a["status"] == 0) & (b["status"] == 0)
out["same_status"] = float((a["status"] == b["status"]).mean())
out["same_iters"] = float((a["iters"] == b["iters"]).mean())
out["max_dx_both_converged"] = float(np.abs(a["x"][ok] - b["x"][ok]).max()) if ok.any() else None
print(json.dumps(out))
