"""Finer breakdown of one host-buffer bo_solve call (C2, 65536): allocation of page-locked results vs copies vs kernel."""
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
lo = s._lowered
Pp, Xp = host_array(P.shape), host_array(X0.shape); Pp[...] = P; Xp[...] = X0
out = {}
def alloc():
    return {"x": host_array((B, lo.nx)), "lam": host_array((B, lo.n_eq + lo.n_ineq)), "f": host_array((B,)),
            "status": host_array((B,), np.int32), "iters": host_array((B,), np.int32), "kkt": host_array((B,))}
for _ in range(3): r = alloc(); s._handle.solve(B, Pp, Xp, r["x"], r["lam"], r["f"], r["status"], r["iters"], r["kkt"])
n = 10
t0 = time.perf_counter()
for _ in range(n): r2 = alloc()
out["alloc_results_ms"] = 1e3 * (time.perf_counter() - t0) / n
s._handle.kernel_time()
t0 = time.perf_counter()
for _ in range(n): s._handle.solve(B, Pp, Xp, r["x"], r["lam"], r["f"], r["status"], r["iters"], r["kkt"])
out["solve_prealloc_all_outputs_ms"] = 1e3 * (time.perf_counter() - t0) / n
ms, k = s._handle.kernel_time(); out["kernel_ms"] = ms / k
t0 = time.perf_counter()
for _ in range(n): s._handle.solve(B, Pp, Xp, r["x"], None, None, r["status"], r["iters"], None)
out["solve_prealloc_x_status_only_ms"] = 1e3 * (time.perf_counter() - t0) / n
# raw copies
d = torch.empty((B, 17), dtype=torch.float64, device="cuda"); hp = torch.from_numpy(r["lam"])
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(n): hp.copy_(d, non_blocking=True); torch.cuda.synchronize()
out["d2h_8.9MB_ms"] = 1e3 * (time.perf_counter() - t0) / n
hx = torch.from_numpy(Pp); dd = torch.empty((B, 10), dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(n): dd.copy_(hx, non_blocking=True); torch.cuda.synchronize()
out["h2d_5.2MB_ms"] = 1e3 * (time.perf_counter() - t0) / n
print(json.dumps(out))
