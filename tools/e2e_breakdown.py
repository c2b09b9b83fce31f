"""Where the end-to-end time of the dict API goes (C2, 65536 instances): python tools/e2e_breakdown.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import optas_b200
from optas_b200 import problems
from optas_b200.solver import host_array

B = 65536
prob = problems.lwr_ik()
P, X0 = prob.sample(B, seed=0)
s = optas_b200.B200Solver(prob.opt).setup("ipopt", timing=True)
pd, xd = prob.param_dict(P), prob.seed_dict(X0)
for _ in range(3):
    s.reset_parameters(pd); s.reset_initial_seed(xd); sol = s.solve()
s._handle.kernel_time()
t = {"reset_parameters": 0.0, "reset_initial_seed": 0.0, "solve": 0.0}
n = 10
for _ in range(n):
    t0 = time.perf_counter(); s.reset_parameters(pd)
    t1 = time.perf_counter(); s.reset_initial_seed(xd)
    t2 = time.perf_counter(); sol = s.solve(); st = s.stats()["n_converged"]
    t3 = time.perf_counter()
    t["reset_parameters"] += t1 - t0; t["reset_initial_seed"] += t2 - t1; t["solve"] += t3 - t2
ms, k = s._handle.kernel_time()
out = {k_: 1e3 * v / n for k_, v in t.items()}
out["kernel_ms"] = ms / k
# raw pieces
Pp, Xp = host_array(P.shape), host_array(X0.shape); Pp[...] = P; Xp[...] = X0
t0 = time.perf_counter()
for _ in range(n): r = s.solve_arrays(Pp, Xp)
out["solve_arrays_pinned_ms"] = 1e3 * (time.perf_counter() - t0) / n
Pd, Xd = torch.from_numpy(P).cuda(), torch.from_numpy(X0).cuda()
Xo = torch.empty_like(Xd); st = torch.empty(B, dtype=torch.int32, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(n): s.solve_raw(Pd, Xd, Xo, None, None, st, None, None, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize(); out["solve_raw_device_ms"] = 1e3 * (time.perf_counter() - t0) / n
print(json.dumps(out))
