"""C2 kernel time vs trip budget / batch size (separates the throughput-bound bulk from the tail of slow instances)."""
import sys; sys.path.insert(0, ".")
import torch, optas_b200, numpy as np
from optas_b200 import problems
prob = problems.lwr_ik()
dev = "cuda"
def run(B, max_trips, max_iter=0, reps=4):
    P, X0 = prob.sample(B, 0)
    opts = {"max_trips": max_trips}
    if max_iter: opts["max_iter"] = max_iter
    s = optas_b200.B200Solver(prob.opt).setup("ipopt", opts, timing=True)
    Pd, X0d = torch.from_numpy(P).to(dev), torch.from_numpy(X0).to(dev); Xd = torch.empty_like(X0d)
    st = torch.empty(B, dtype=torch.int32, device=dev); it = torch.empty(B, dtype=torch.int32, device=dev)
    for _ in range(2): s.solve_raw(Pd, X0d, Xd, None, None, st, it, None)
    torch.cuda.synchronize(); s._handle.kernel_time()
    for _ in range(reps): s.solve_raw(Pd, X0d, Xd, None, None, st, it, None)
    torch.cuda.synchronize(); ms, n = s._handle.kernel_time()
    itc = it.cpu().numpy(); stc = st.cpu().numpy()
    print(f"B {B:7d} max_trips {max_trips:4d} max_iter {max_iter or 100:3d}: {ms/n:8.3f} ms  conv {float((stc<=1).mean()):.5f}  -> {B/(ms/n)*1e3:.3e} inst/s"
          f"  iters p50/p99/p99.9/max {np.percentile(itc,50):.0f}/{np.percentile(itc,99):.0f}/{np.percentile(itc,99.9):.0f}/{itc.max()} status {np.bincount(stc, minlength=5)}", flush=True)
for mt in (40, 60, 80, 120, 250):
    run(65536, mt)
run(65536, 250, 60)
run(1048576, 250)
run(1048576, 80)
