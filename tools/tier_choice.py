"""Thread-per-instance (sparse tier) vs one-instance-per-CTA (cooperative tier) on the small / medium 8f-3 problems."""
import sys; sys.path.insert(0, ".")
import torch, optas_b200, numpy as np
from optas_b200 import problems
dev = "cuda"
for name, B in (("lwr_axis_ik", 65536), ("point_mass_mpc", 16384), ("joint_space_planner", 8192)):
    prob = getattr(problems, name)()
    P, X0 = prob.sample(B)
    for coop in (True, False):
        s = optas_b200.B200Solver(prob.opt).setup("ipopt", timing=True, coop=coop)
        Pd, X0d = torch.from_numpy(P).to(dev), torch.from_numpy(X0).to(dev); Xd = torch.empty_like(X0d)
        st = torch.empty(B, dtype=torch.int32, device=dev); it = torch.empty(B, dtype=torch.int32, device=dev)
        s.solve_raw(Pd, X0d, Xd, None, None, st, it, None)
        torch.cuda.synchronize(); s._handle.kernel_time()
        for _ in range(2): s.solve_raw(Pd, X0d, Xd, None, None, st, it, None)
        torch.cuda.synchronize(); ms, n = s._handle.kernel_time()
        ti = s.tier_info()
        print(f"{name:22s} B {B:6d} coop={coop!s:5s} tier {ti['tier']:6s} tpb {ti['threads_per_block']:4d}: {ms/n:9.3f} ms -> {B/(ms/n)*1e3:.3e} inst/s conv {float((st<=1).float().mean()):.4f} iters {float(it.float().mean()):.2f}", flush=True)
