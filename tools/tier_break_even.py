"""Where the cooperative tier starts to beat the thread-per-instance sparse tier: MPC ticks of growing horizon."""
import sys; sys.path.insert(0, ".")
import torch, optas_b200, numpy as np
from optas_b200 import problems
dev = "cuda"
B = 16384
for T in (6, 8, 10, 12, 14, 16):
    prob = problems.point_mass_mpc(T=T)
    P, X0 = prob.sample(B)
    rows = prob.opt.nx + prob.opt.na + prob.opt.nh
    for coop in (True, False):
        s = optas_b200.B200Solver(prob.opt).setup("ipopt", timing=True, coop=coop)
        Pd, X0d = torch.from_numpy(P).to(dev), torch.from_numpy(X0).to(dev); Xd = torch.empty_like(X0d)
        st = torch.empty(B, dtype=torch.int32, device=dev); it = torch.empty(B, dtype=torch.int32, device=dev)
        s.solve_raw(Pd, X0d, Xd, None, None, st, it, None)
        torch.cuda.synchronize(); s._handle.kernel_time()
        for _ in range(2): s.solve_raw(Pd, X0d, Xd, None, None, st, it, None)
        torch.cuda.synchronize(); ms, n = s._handle.kernel_time()
        ti = s.tier_info()
        print(f"T {T:3d} KKT rows {rows:4d} coop={coop!s:5s} tier {ti['tier']:6s}: {ms/n:9.3f} ms -> {B/(ms/n)*1e3:.3e} inst/s conv {float((st<=1).float().mean()):.4f}", flush=True)
