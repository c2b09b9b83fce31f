"""Cooperative tier (one instance per CTA) on the GPU: timing, convergence, determinism, agreement with the
thread-per-instance tiers.  usage: python tools/coop_check.py [c3|c4|c5 ...]"""
import sys, time; sys.path.insert(0, ".")
import numpy as np, optas_b200
from optas_b200 import problems

CASES = {
    "c3": (problems.point_mass_mpc, 16384, {}, dict(tpbs=(32, 64), ref=True)),
    "c5": (problems.dual_arm, 4096, {}, dict(tpbs=(256,), ref=False)),
    "c4": (problems.figure_eight, 1024, {"max_iter": 400, "max_trips": 2500}, dict(tpbs=(256,), ref=False)),
}
for name in (sys.argv[1:] or ["c3", "c5", "c4"]):
    mk, B, opts, cfg = CASES[name]
    prob = mk()
    P, X0 = prob.sample(B, seed=1)
    ref = None
    if cfg["ref"]:
        s = optas_b200.B200Solver(prob.opt).setup("ipopt", opts, timing=True, coop=False)
        s.solve_arrays(P[:256], X0[:256])
        s._handle.kernel_time()
        ref = s.solve_arrays(P, X0)
        ms, n = s._handle.kernel_time()
        print(f"{name} per-thread tier {s.tier_info()['tier']}: {ms/n:9.2f} ms for {B} -> {B/(ms/n)*1e3:10.0f} inst/s, status {np.bincount(ref['status'], minlength=5)}", flush=True)
    for tpb in cfg["tpbs"]:
        s = optas_b200.B200Solver(prob.opt).setup("ipopt", opts, timing=True, coop=True, threads_per_block=tpb)
        ti = s.tier_info()
        s.solve_arrays(P[:256], X0[:256])
        s._handle.kernel_time()
        t = time.time(); r = s.solve_arrays(P, X0); wall = time.time() - t
        ms, n = s._handle.kernel_time()
        r2 = s.solve_arrays(P[: B // 2], X0[: B // 2])
        same = all(np.array_equal(r[k][: B // 2], r2[k]) for k in ("x", "lam", "f", "status", "iters"))
        ok = r["status"] <= 1
        print(f"{name} coop tpb {tpb:3d} ctas/sm {ti['blocks_per_sm']} smem {ti['smem_dynamic']}: {ms/n:9.2f} ms for {B} -> {B/(ms/n)*1e3:10.0f} inst/s "
              f"(wall {wall*1e3:.0f} ms) status {np.bincount(r['status'], minlength=5)} mean iters {r['iters'].mean():.1f} "
              f"max kkt(ok) {r['kkt'][ok].max() if ok.any() else float('nan'):.2e} bitwise-same-on-half-batch {same}", flush=True)
        if ref is not None:
            both = ok & (ref["status"] <= 1)
            print(f"     vs per-thread: max |x diff| {np.abs(r['x'][both] - ref['x'][both]).max():.2e}, |f diff| {np.abs(r['f'][both] - ref['f'][both]).max():.2e}", flush=True)
