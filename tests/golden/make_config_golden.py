"""Golden vectors for the BASELINE configs C3 / C4 / C5, built by the UNMODIFIED reference example scripts.

The scripts' own classes (`point_mass_mpc.Controller`, `figure_eight_plan.Planner`, `dual_arm.DualKukaPlanner`) are
imported from /root/reference/example and instantiated as they are; only `CasADiSolver.setup` (the `cs.nlpsol` call --
IPOPT is absent) is replaced by a no-op, and pybullet / matplotlib are inert stubs (tests/golden/ref_shim.py).  From
the `Optimization` each script builds, this records class, dimensions, label order and -- at seeded random (x, p) --
f, v = [k; g; a; -a; h; -h], df, and two projections of the constraint Jacobian (dv @ u, dv.T @ w with
u, w = cos / sin ramps, see `probe_vectors`) so the files stay small.  Run in the build container only:

    python tests/golden/make_config_golden.py    ->  tests/golden/config_golden.json
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

optas = ref_shim.install()
optas.CasADiSolver.setup = lambda self, *a, **k: self  # no IPOPT here; everything before this call is the reference's
for stub in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "pybullet_api"):
    sys.modules.setdefault(stub, ref_shim._Anything(stub))
sys.path.insert(0, "/root/reference/example")


def _xacro_process(filename):
    """xacro is absent: figure_eight_plan.py asks for med7.urdf.xacro; the reference ships its expansion med7.urdf in
    the same directory (the file torque_control_example.py loads), which is what this returns."""
    assert filename.endswith(".urdf.xacro"), filename
    return open(filename[:-len(".xacro")]).read()


xacro_stub = types.ModuleType("xacro")
xacro_stub.process = _xacro_process
sys.modules["xacro"] = xacro_stub
import optas.models  # noqa: E402

optas.models.xacro = xacro_stub


def probe_vectors(nv, nx):
    return np.cos(0.37 * np.arange(nx) + 0.1), np.sin(0.23 * np.arange(nv) + 0.2)


def record(opt, seed, n_cases=2):
    rng = np.random.default_rng(seed)
    arr = lambda m: np.asarray(optas.DM(m).toarray()).flatten().tolist()
    u, w = probe_vectors(opt.nv, opt.nx)
    cases = []
    for _ in range(n_cases):
        x, p = rng.uniform(-1.0, 1.0, opt.nx), rng.uniform(-1.0, 1.0, opt.np)
        dv = np.asarray(optas.DM(opt.dv(x, p)).toarray())
        cases.append({"x": x.tolist(), "p": p.tolist(), "f": arr(opt.f(x, p))[0], "v": arr(opt.v(x, p)), "df": arr(opt.df(x, p)),
                      "dv_u": (dv @ u).tolist(), "dvT_w": (dv.T @ w).tolist()})
    return {"class": type(opt).__name__, "dims": [opt.nx, opt.np, opt.nk, opt.na, opt.ng, opt.nh, opt.nv],
            "decision_variables": list(opt.decision_variables.keys()), "parameters": list(opt.parameters.keys()),
            "cases": cases}


if __name__ == "__main__":
    out = {}
    import point_mass_mpc

    out["c3_point_mass_mpc"] = record(point_mass_mpc.Controller().solver.opt, 31)
    print("c3", out["c3_point_mass_mpc"]["class"], out["c3_point_mass_mpc"]["dims"])
    import figure_eight_plan

    out["c4_figure_eight_no_limits"] = record(figure_eight_plan.Planner().solver.opt, 41)
    print("c4", out["c4_figure_eight_no_limits"]["class"], out["c4_figure_eight_no_limits"]["dims"])
    import dual_arm

    out["c5_dual_arm"] = record(dual_arm.DualKukaPlanner().solver.opt, 51)
    print("c5", out["c5_dual_arm"]["class"], out["c5_dual_arm"]["dims"])
    json.dump(out, open(os.path.join(HERE, "config_golden.json"), "w"))
    print(os.path.getsize(os.path.join(HERE, "config_golden.json")), "bytes")
