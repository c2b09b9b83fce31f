"""Device-resident step time of the horizon configs in the caller's order vs scheduled most-infeasible-seed first.
usage: python tools/order_probe_configs.py c3 c3_zero_seed c4 c5_zero_seed"""
import sys; sys.path.insert(0, ".")
import torch, numpy as np, optas_b200, bench
for name in sys.argv[1:]:
    cfg = bench.CONFIGS[name]
    prob, P, X0 = bench.config_inputs(name, 0, 1)
    if cfg["seed"] == "zeros":
        X0 = np.zeros_like(X0)
    B = X0.shape[0]
    for sched in (None, "seed_infeasibility"):
        s = optas_b200.B200Solver(prob.opt).setup("ipopt", cfg["opts"], schedule=sched, **cfg.get("setup", {}))
        Pd, X0d = torch.from_numpy(P).cuda(), torch.from_numpy(X0).cuda()
        Xd = torch.empty_like(X0d); st = torch.empty(B, dtype=torch.int32, device="cuda"); it = torch.empty(B, dtype=torch.int32, device="cuda")
        s.solve_raw(Pd, X0d, Xd, None, None, st, it, None)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = cfg["steps"]
        a.record()
        for _ in range(reps): s.solve_raw(Pd, X0d, Xd, None, None, st, it, None)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        print(f"{name:14s} schedule {str(sched):20s} {ms:9.3f} ms  {B / ms * 1e3:.4e} inst/s  status {np.bincount(st.cpu().numpy(), minlength=5)}", flush=True)
