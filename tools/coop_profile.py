"""Phase breakdown of the cooperative tier: a -DBO_PROFILE build of the kernel returns clock64() counters
(thread 0 of each CTA) in place of the first 8 solution entries.  usage: python tools/coop_profile.py c3|c4|c5 [tpb]"""
import os, sys; sys.path.insert(0, ".")
os.environ["B200OPTAS_JIT_DEFINES"] = ("-DBO_PROFILE=1 " + os.environ.get("EXTRA_DEFINES", "")).strip()
import numpy as np, optas_b200
from optas_b200 import problems

CASES = {"c3": (problems.point_mass_mpc, {}), "c5": (problems.dual_arm, {}),
         "c4": (problems.figure_eight, {"max_iter": 400, "max_trips": 2500})}
name = sys.argv[1]
tpb = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mk, opts = CASES[name]
prob = mk()
compile_only = len(sys.argv) > 3 and sys.argv[3] == "compile"
s = optas_b200.B200Solver(prob.opt).setup("ipopt", opts, coop=True, threads_per_block=tpb, compile_only=compile_only)
ti = s.tier_info()
print(name, ti)
if compile_only:
    sys.exit(0)
B = ti["n_sm"] * ti["blocks_per_sm"]
P, X0 = prob.sample(B, seed=1)
r = s.solve_arrays(P, X0)
c = r["x"][:, :8]
it = np.maximum(r["iters"], 1)[:, None]
names = ["kkt tape", "f/c tape", "assembly", "factor", "solves", "-", "-", "total"]
tot = c[:, 7].mean()
print(f"{name}: {B} instances (one per CTA), mean iterations {r['iters'].mean():.1f}, status {np.bincount(r['status'], minlength=5)}")
for k in (0, 1, 2, 3, 4, 7):
    print(f"  {names[k]:10s} {c[:, k].mean()/1e3:10.1f} kcycles per instance ({100*c[:, k].mean()/tot:5.1f} %)  {(c[:, k:k+1]/it).mean()/1e3:8.2f} kcycles per iteration")
print(f"  factorisations per iteration {(c[:, 5:6]/it).mean():.2f}, solves per iteration {(c[:, 6:7]/it).mean():.2f}, "
      f"kcycles per factorisation {c[:, 3].sum()/max(c[:, 5].sum(), 1)/1e3:.1f}, per solve {c[:, 4].sum()/max(c[:, 6].sum(), 1)/1e3:.1f}")
rest = tot - c[:, :5].sum(axis=1).mean()
print(f"  {'other':10s} {rest/1e3:10.1f} kcycles per instance ({100*rest/tot:5.1f} %)")
