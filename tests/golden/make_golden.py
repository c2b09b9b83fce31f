"""Golden vectors produced by the UNMODIFIED reference package (/root/reference/optas: models.py, spatialmath.py,
builder.py, optimization.py) imported through tests/golden/ref_shim.py.  The reference's Python runs as is -- chain
walking, frame conventions, constraint sorting and stacking are the reference's; only the scalar arithmetic
underneath (`casadi`) is this repo's stand-in, since the casadi wheel is absent.  Run in the build container only:

    python tests/golden/make_golden.py      ->  tests/golden/kinematics_golden.json, tests/golden/ik_problem_golden.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

optas = ref_shim.install()

REF_ROBOTS = "/root/reference/example/robots"
KIN = {  # file under optas_b200/robots (or tests/golden) -> (reference URDF, links to sample)
    "kuka_lwr.urdf": (os.path.join(REF_ROBOTS, "kuka_lwr", "kuka_lwr.urdf"), ["end_effector_ball", "lwr_arm_5_link"]),
    "med7.urdf": (os.path.join(REF_ROBOTS, "kuka_lbr", "med7.urdf"), ["lbr_link_ee", "lbr_link_3"]),
    "tester_robot.urdf": ("/root/reference/tests/tester_robot.urdf", ["eff"]),
}


def kinematics():
    rng = np.random.default_rng(10)
    out = {}
    for name, (path, links) in KIN.items():
        model = optas.RobotModel(urdf_filename=path)
        lo = model.lower_actuated_joint_limits.toarray().flatten()
        up = model.upper_actuated_joint_limits.toarray().flatten()
        cases = []
        for link in links:
            for _ in range(4):
                q = rng.uniform(np.maximum(lo, -3.0), np.minimum(up, 3.0))
                arr = lambda m: np.asarray(m.toarray()).tolist()
                cases.append({
                    "link": link, "q": q.tolist(),
                    "transform": arr(model.get_global_link_transform(link, q)),
                    "position": arr(model.get_global_link_position(link, q)),
                    "quaternion": arr(model.get_global_link_quaternion(link, q)),
                    "geometric_jacobian": arr(model.get_global_link_geometric_jacobian(link, q)),
                    "rpy": arr(model.get_global_link_rpy(link, q)),
                })
        out[name] = cases
        print(name, len(cases), "cases")
    json.dump(out, open(os.path.join(HERE, "kinematics_golden.json"), "w"), indent=1)


def ik_problem():
    """C1 / C2 exactly as example/example.py:12-37 builds it, then f, v = [k; g; a; -a; h; -h], df, dv of the built
    Optimization (optimization.py:27-51) at random (x, p)."""
    urdf = os.path.join(REF_ROBOTS, "kuka_lwr", "kuka_lwr.urdf")
    robot = optas.RobotModel(urdf_filename=urdf, time_derivs=[0])
    name = robot.get_name()
    builder = optas.OptimizationBuilder(T=1, robots=robot)
    qn = builder.add_parameter("q_nominal", robot.ndof)
    pg = builder.add_parameter("p_goal", 3)
    q = builder.get_model_state(name, 0)
    end_effector_name = "end_effector_ball"
    p = robot.get_global_link_position(end_effector_name, q)
    builder.add_equality_constraint("end_goal", p, pg)
    builder.add_cost_term("nominal", optas.sumsqr(q - qn))
    builder.enforce_model_limits(name)
    opt = builder.build()
    rng = np.random.default_rng(11)
    cases = []
    for _ in range(6):
        x, pp = rng.uniform(-2.0, 2.0, opt.nx), rng.uniform(-1.0, 1.0, opt.np)
        arr = lambda m: np.asarray(optas.DM(m).toarray()).tolist()
        cases.append({"x": x.tolist(), "p": pp.tolist(), "f": arr(opt.f(x, pp)), "v": arr(opt.v(x, pp)),
                      "df": arr(opt.df(x, pp)), "dv": arr(opt.dv(x, pp)), "k": arr(opt.k(x, pp)), "h": arr(opt.h(x, pp))})
    out = {"class": type(opt).__name__, "dims": [opt.nx, opt.np, opt.nk, opt.na, opt.ng, opt.nh, opt.nv],
           "decision_variables": list(opt.decision_variables.keys()), "parameters": list(opt.parameters.keys()),
           "cases": cases}
    json.dump(out, open(os.path.join(HERE, "ik_problem_golden.json"), "w"), indent=1)
    print("ik problem", out["class"], out["dims"], out["decision_variables"], out["parameters"])


if __name__ == "__main__":
    kinematics()
    ik_problem()
