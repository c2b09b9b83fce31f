"""C2 kernel time with the batch in its natural order vs sorted by the infeasibility of the seed (most infeasible first):
the launch ends with its slowest instance, and slow instances are mostly those whose seed is far from feasible.
usage: python tools/order_probe.py [B]"""
import sys; sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import torch, optas_b200, numpy as np
from optas_b200 import problems
import fk_ref
prob = problems.lwr_ik()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
P, X0 = prob.sample(B, 0)
p0, _ = fk_ref.lwr_position_and_jacobian(X0)
theta = np.linalg.norm(P[:, 7:10] - p0, axis=1)
s = optas_b200.B200Solver(prob.opt).setup("ipopt", {}, timing=True)
dev = "cuda"
def run(order, label, reps=5, full=True):
    Pd, X0d = torch.from_numpy(np.ascontiguousarray(P[order])).to(dev), torch.from_numpy(np.ascontiguousarray(X0[order])).to(dev)
    Xd = torch.empty_like(X0d)
    st = torch.empty(B, dtype=torch.int32, device=dev); it = torch.empty(B, dtype=torch.int32, device=dev)
    for _ in range(2): s.solve_raw(Pd, X0d, Xd, None, None, st, it, None)
    torch.cuda.synchronize(); s._handle.kernel_time()
    for _ in range(reps): s.solve_raw(Pd, X0d, Xd, None, None, st, it, None)
    torch.cuda.synchronize(); ms, n = s._handle.kernel_time()
    print(f"{label:45s} {ms/n:7.3f} ms  {B/(ms/n)*1e3:.3e} inst/s  status {np.bincount(st.cpu().numpy(), minlength=5)}", flush=True)
d = P[:, 7:10] - p0
run(np.arange(B), "natural order")
run(np.argsort(-np.abs(d).sum(axis=1)), "1-norm of the position error, descending")
run(np.argsort(-np.abs(d).max(axis=1)), "max-norm of the position error, descending")
run(np.argsort(-theta), "seed infeasibility, descending")
run(np.argsort(theta), "seed infeasibility, ascending (worst case)")
rng = np.random.default_rng(0); run(rng.permutation(B), "random permutation")
# coarse version: 16 buckets only
q = np.quantile(theta, np.linspace(0, 1, 17)[1:-1]); bucket = np.searchsorted(q, theta)
run(np.argsort(-bucket, kind="stable"), "16 quantile buckets, descending")
