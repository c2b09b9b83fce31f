"""C2 kernel: time per trip as a function of how full the machine is (guides the tail work: is a lone lane's trip bound by
dependent-instruction latency / instruction fetch, or does it speed up when the SMs are empty?).  The SAME instance is
replicated B times, so every lane takes the same number of trips and  kernel time / trips = time per trip."""
import sys; sys.path.insert(0, ".")
import torch, optas_b200, numpy as np
from optas_b200 import problems
prob = problems.lwr_ik()
dev = "cuda"
P1, X1 = prob.sample(4096, 0)
s = optas_b200.B200Solver(prob.opt).setup("ipopt", timing=True)
# pick an instance with a typical and one with a long iteration count
Pd, X0d = torch.from_numpy(P1).to(dev), torch.from_numpy(X1).to(dev); Xd = torch.empty_like(X0d)
st = torch.empty(4096, dtype=torch.int32, device=dev); it = torch.empty(4096, dtype=torch.int32, device=dev)
s.solve_raw(Pd, X0d, Xd, None, None, st, it, None); torch.cuda.synchronize()
itc = it.cpu().numpy(); ok = st.cpu().numpy() <= 1
order = np.argsort(itc)
picks = {"median": int(order[len(order) // 2]), "p99": int(order[int(0.99 * len(order))])}
print("tier", s.tier_info().get("tier"), "threads/CTA", s.tier_info().get("threads_per_block"), "CTAs/SM", s.tier_info().get("blocks_per_sm"))
for name, i in picks.items():
    for B in (1, 32, 148, 148 * 32, 148 * 256, 148 * 256 * 4):
        P = np.ascontiguousarray(np.tile(P1[i], (B, 1))); X0 = np.ascontiguousarray(np.tile(X1[i], (B, 1)))
        Pd, X0d = torch.from_numpy(P).to(dev), torch.from_numpy(X0).to(dev); Xd = torch.empty_like(X0d)
        st = torch.empty(B, dtype=torch.int32, device=dev); it = torch.empty(B, dtype=torch.int32, device=dev)
        for _ in range(2): s.solve_raw(Pd, X0d, Xd, None, None, st, it, None)
        torch.cuda.synchronize(); s._handle.kernel_time()
        for _ in range(5): s.solve_raw(Pd, X0d, Xd, None, None, st, it, None)
        torch.cuda.synchronize(); ms, n = s._handle.kernel_time()
        print(f"{name:6s} instance {i:5d} iters {int(itc[i]):3d}  B {B:7d}: {ms/n:8.4f} ms per launch -> {ms/n/max(1,int(itc[i]))*1e3:8.2f} us per iteration", flush=True)
