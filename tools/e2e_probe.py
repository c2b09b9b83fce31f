"""Wall time of solve_arrays (host arrays in, host arrays out) against the kernel time, per call."""
import sys, time; sys.path.insert(0, ".")
import numpy as np, torch, optas_b200
from optas_b200 import problems
from optas_b200.solver import host_array
def pinned(a):
    out = host_array(a.shape); out[...] = a; return out
for name, B in (("lwr_axis_ik", 65536), ("lwr_ik", 65536)):
    prob = getattr(problems, name)()
    P, X0 = prob.sample(B); P, X0 = pinned(P), pinned(X0)
    for timing in (True, False):
        s = optas_b200.B200Solver(prob.opt).setup("ipopt", timing=timing)
        walls = []
        for k in range(6):
            t0 = time.perf_counter(); r = s.solve_arrays(P, X0); walls.append((time.perf_counter() - t0) * 1e3)
        km = s._handle.kernel_time() if timing else (float("nan"), 1)
        print(f"{name:12s} timing={timing!s:5s} tier {s.tier_info()['tier']:6s} wall ms per call {[round(w, 2) for w in walls]} kernel ms {km[0] / max(1, km[1]):.3f}", flush=True)
        del r
