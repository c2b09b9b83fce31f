"""Kernel time vs trip budget / batch size (separates the throughput-bound bulk from the tail)."""
import sys; sys.path.insert(0, ".")
import torch, optas_b200, numpy as np
from optas_b200 import problems
prob = problems.lwr_ik()
dev = "cuda"
def run(B, tpb, max_trips, reps=4, bps=0):
    P, X0 = prob.sample(B, 0)
    s = optas_b200.B200Solver(prob.opt).setup("ipopt", {"max_trips": max_trips}, timing=True, threads_per_block=tpb, blocks_per_sm=bps)
    Pd, X0d = torch.from_numpy(P).to(dev), torch.from_numpy(X0).to(dev); Xd = torch.empty_like(X0d)
    st = torch.empty(B, dtype=torch.int32, device=dev)
    for _ in range(2): s.solve_raw(Pd, X0d, Xd, None, None, st, None, None)
    torch.cuda.synchronize(); s._handle.kernel_time()
    for _ in range(reps): s.solve_raw(Pd, X0d, Xd, None, None, st, None, None)
    torch.cuda.synchronize(); ms, n = s._handle.kernel_time()
    print(f"B {B:7d} tpb {tpb:3d} bps {bps} max_trips {max_trips:5d}: {ms/n:8.3f} ms  conv {float((st<=1).float().mean()):.4f}  -> {B/(ms/n)*1e3:.3e} inst/s", flush=True)
import os
for defs in ("", "-DBO_INNER_ROUNDS=1"):
    os.environ["B200OPTAS_JIT_DEFINES"] = defs
    print("defines:", defs or "(none)", flush=True)
    run(65536, 64, 250)
    run(1048576, 64, 250)
