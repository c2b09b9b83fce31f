"""One C2 launch with B copies of one instance (default 148: one lane per SM) -- the regime the tail of a batch runs in.
Run under ncu to see what a nearly empty machine stalls on:
  ncu --set full --clock-control none -k bo_solve_kernel -c 1 -o gpurun_out/lone_lane python tools/lone_lane.py 148"""
import sys; sys.path.insert(0, ".")
import torch, optas_b200, numpy as np
from optas_b200 import problems
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
prob = problems.lwr_ik()
P1, X1 = prob.sample(4096, 0)
i = 1805  # a p99 instance of this sample (35 iterations; tools/trip_latency.py)
P = np.ascontiguousarray(np.tile(P1[i], (B, 1))); X0 = np.ascontiguousarray(np.tile(X1[i], (B, 1)))
s = optas_b200.B200Solver(prob.opt).setup("ipopt")
dev = "cuda"
Pd, X0d = torch.from_numpy(P).to(dev), torch.from_numpy(X0).to(dev); Xd = torch.empty_like(X0d)
st = torch.empty(B, dtype=torch.int32, device=dev); it = torch.empty(B, dtype=torch.int32, device=dev)
s.solve_raw(Pd, X0d, Xd, None, None, st, it, None); torch.cuda.synchronize()
print("B", B, "iters", int(it[0]), "status", int(st[0]))
