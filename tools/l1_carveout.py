"""C2 solver kernel: L1 / shared-memory split (B200OPTAS_L1_CARVEOUT = -1 driver heuristic | 0 max L1 | 100 max smem) and
resident CTAs per SM, at the bench batch and at a 1 M batch."""
import os, sys; sys.path.insert(0, ".")
import torch, optas_b200, numpy as np
from optas_b200 import problems
prob = problems.lwr_ik()
dev = "cuda"
def run(B, carve, bps):
    os.environ["B200OPTAS_L1_CARVEOUT"] = str(carve)
    P, X0 = prob.sample(B, 0)
    s = optas_b200.B200Solver(prob.opt).setup("ipopt", timing=True, blocks_per_sm=bps)
    Pd, X0d = torch.from_numpy(P).to(dev), torch.from_numpy(X0).to(dev); Xd = torch.empty_like(X0d)
    st = torch.empty(B, dtype=torch.int32, device=dev); it = torch.empty(B, dtype=torch.int32, device=dev)
    for _ in range(2): s.solve_raw(Pd, X0d, Xd, None, None, st, it, None)
    torch.cuda.synchronize(); s._handle.kernel_time()
    for _ in range(4): s.solve_raw(Pd, X0d, Xd, None, None, st, it, None)
    torch.cuda.synchronize(); ms, n = s._handle.kernel_time()
    print(f"B {B:8d} carveout {carve:4d} blocks_per_sm {s.tier_info().get('blocks_per_sm')}: {ms/n:8.3f} ms -> {B/(ms/n)*1e3:.3e} inst/s conv {float((st<=1).float().mean()):.5f}", flush=True)
for B in (65536, 1 << 20):
    for carve in (-1, 0, 100):
        run(B, carve, 0)
    for bps in (1, 2, 3):
        run(B, 0, bps)
