"""north_star: "plugs in behind optas.Solver ... so example/ scripts run unchanged".

The reference's example scripts are imported UNMODIFIED from /root/reference/example with this package registered as
`optas` (and inert stubs for pybullet / matplotlib, which only the scripts' `main()` animation code touches).  Their own
planner / controller classes build their problems through this package's models / builder, construct
`optas.CasADiSolver(...).setup("ipopt")` -- here in compile-only mode, since this machine has no GPU: tapes are lowered,
CUDA is generated and compiled for sm_100a -- and the scripts' own `reset(...)` methods feed parameters and seeds.
Runs only where the reference checkout exists (the build container)."""
import os
import subprocess
import sys
import textwrap

import pytest

REF_EXAMPLES = "/root/reference/example"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PRELUDE = f"""
import sys, types
sys.path.insert(0, {ROOT!r})
import numpy as np
import optas_b200
sys.modules["optas"] = optas_b200
sys.path.insert(0, {os.path.join(ROOT, "tests")!r})
import shim_templates  # tests/ only: the reference's Manager glue (out of scope for the package, SURVEY.md 2 #13)
optas_b200.templates = shim_templates
sys.modules["optas.templates"] = shim_templates
sys.modules["optas.spatialmath"] = optas_b200.spatialmath


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {{"__init__": lambda self, *a, **k: None}})


for stub in ("pybullet_api", "pybullet", "pybullet_data", "matplotlib", "matplotlib.pyplot", "matplotlib.animation"):
    sys.modules[stub] = _Anything(stub)
_setup = optas_b200.B200Solver.setup
optas_b200.B200Solver.setup = lambda self, *a, **k: _setup(self, *a, **dict(k, compile_only=True))  # no GPU here
sys.path.insert(0, {REF_EXAMPLES!r})
"""

CASES = {
    "point_mass_mpc": """
        import point_mass_mpc
        c = point_mass_mpc.Controller()
        s = c.solver
        assert type(s).__name__ == "CasADiSolver" and isinstance(s, optas_b200.B200Solver)
        assert (s.opt.nx, s.opt.nv) == (80, 264) and s.tier_info()["tier"] == "coop"
        s.reset_parameters({"curr": [0.1, 0.2], "dcurr": [0, 0], "goal": np.ones((2, 20)), "obs": np.zeros((2, 20))})
        assert float(s.p.toarray()[0]) == 0.1 and s.p.shape == (84, 1)
    """,
    "dual_arm": """
        import dual_arm
        planner = dual_arm.DualKukaPlanner()
        s = planner.solver
        assert isinstance(s, optas_b200.B200Solver) and (s.opt.nx, s.opt.na) == (1386, 700)
        q = optas_b200.deg2rad([0, -30, 0, 90, 0, 30, 0])
        planner.reset(q, q)                      # the script's own reset(): parameters + seed dictionaries
        assert abs(float(s.p.toarray()[1]) + np.pi / 6) < 1e-15
        assert planner.is_ready() and planner.is_first_solve()
    """,
    "figure_eight_plan": """
        import figure_eight_plan
        planner = figure_eight_plan.Planner()
        s = planner.solver
        assert isinstance(s, optas_b200.B200Solver) and (s.opt.nx, s.opt.na, s.opt.nh) == (693, 357, 200)
        qc = optas_b200.deg2rad([0, 30, 0, -90, 0, -30, 0])
        planner.reset(qc)
        x0 = s.x0.toarray().flatten()
        assert np.allclose(x0[:7], qc.toarray().flatten()) and np.allclose(x0[343:350], qc.toarray().flatten()) and not x0[350:].any()
    """,
    "simple_joint_space_planner": """
        import simple_joint_space_planner as sjp
        from optas_b200 import problems
        planner = sjp.SimpleJointSpacePlanner(open(problems.MED7_URDF).read(), "lbr_link_ee", 4.0)
        s = planner.solver
        assert isinstance(s, optas_b200.B200Solver)
        assert (s.opt.nx, s.opt.np, s.opt.nk, s.opt.na, s.opt.ng, s.opt.nh) == (280, 21, 0, 147, 40, 7)
        q0 = np.deg2rad([0, 45, 0, -90, 0, -45, 0])
        planner.reset(q0, [0.4, 0.3, 0.4], [0, 1, 0, 0], q0)
        assert s.p.shape == (21, 1)
    """,
    "other_scripts": """
        # every other script whose planner / controller class can be constructed without a simulator
        import point_mass_planner, figure_eight_plan_6dof, pushing, TrackingBall, torque_control_example
        got = {}
        for label, make in [("point_mass_planner", lambda: point_mass_planner.Planner()),
                            ("figure_eight_plan_6dof", lambda: figure_eight_plan_6dof.Planner()),
                            ("pushing.TOMPCCPlanner", lambda: pushing.TOMPCCPlanner(0.1, 0.2, 0.1)),
                            ("pushing.IK", lambda: pushing.IK(0.02, 0.5)),
                            ("TrackingBall", lambda: TrackingBall.TrackingController(0.02)),          # setup("sqpmethod")
                            ("torque_control", lambda: torque_control_example.TrackingController(0.02))]:
            s = make().solver
            assert isinstance(s, optas_b200.B200Solver), label
            got[label] = (type(s).__name__, s.opt.nx, s.opt.nv, type(s.opt).__name__)
        assert got["point_mass_planner"] == ("CasADiSolver", 180, 593, "QuadraticCostNonlinearConstraints"), got
        assert got["figure_eight_plan_6dof"][1] == 594 and got["pushing.TOMPCCPlanner"][0] == "ScipyMinimizeSolver", got
        assert got["TrackingBall"][1:3] == (7, 3) and got["torque_control"][1:3] == (7, 3), got
    """,
    "example": """
        # example/example.py is a flat script ending in solver.solve() + a VTK window: run its lines up to the solve
        src = open(sys.path[0] + "/example.py").read()
        head = src.split("solution = solver.solve()")[0]
        ns = {"__file__": sys.path[0] + "/example.py", "__name__": "__main__"}
        exec(compile(head, "example.py", "exec"), ns)
        s = ns["solver"]
        assert isinstance(s, optas_b200.B200Solver) and (s.opt.nx, s.opt.np, s.opt.nk, s.opt.nh) == (7, 10, 14, 3)
        assert abs(float(s.p.toarray()[8]) - 0.3) < 1e-12      # p_goal = p(q_nominal) + [0, 0.3, -0.2]
        assert not s.x0.toarray().any()                         # SURVEY.md 3.4-1: the script's seed key is ignored -> zeros
    """,
}


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="reference checkout not present on this machine")
@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_example_builds_and_sets_up_on_this_package(name):
    code = PRELUDE + textwrap.dedent(CASES[name]) + "\nprint('OK')\n"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-1500:] + r.stderr[-2500:]
