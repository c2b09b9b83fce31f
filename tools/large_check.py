"""C4 / C5 on the GPU through the table-driven large tier: timing and convergence."""
import sys, time; sys.path.insert(0, ".")
import numpy as np, torch, optas_b200
from optas_b200 import problems
which = sys.argv[1] if len(sys.argv) > 1 else "dual_arm"
sizes = [int(a) for a in sys.argv[2:] if not a.startswith("-")] or [256, 2048]
bps = int(next((a[5:] for a in sys.argv if a.startswith("-bps=")), "0"))
tpb = int(next((a[5:] for a in sys.argv if a.startswith("-tpb=")), "0"))
prob = getattr(problems, which)()
opts = {"max_iter": 400, "max_trips": 2500} if which == "figure_eight" else {}
t = time.time(); s = optas_b200.B200Solver(prob.opt).setup("ipopt", opts, timing=True, blocks_per_sm=bps, threads_per_block=tpb); print(which, "setup", round(time.time() - t, 1), s.kernel_info(), flush=True)
for B in sizes:
    P, X0 = prob.sample(B)
    t = time.time(); r = s.solve_arrays(P, X0); dt = time.time() - t
    ms, n = s._handle.kernel_time()
    ok = r["status"] <= 1
    print(f"{which} B={B}: wall {dt*1e3:.0f} ms kernel {ms/n:.0f} ms conv {ok.mean():.4f} status {np.bincount(r['status'], minlength=5)} iters mean {r['iters'][ok].mean():.1f} "
          f"kkt max {r['kkt'][ok].max():.2e} -> {ok.sum()/(ms/n)*1e3:.3e} inst/s", flush=True)
