"""Golden vectors for inverse dynamics, produced by the UNMODIFIED reference's own `RobotModel.rnea`
(/root/reference/optas/models.py:1731-1884) imported through tests/golden/ref_shim.py (the reference's Python
runs as is; its casadi / urdf_parser_py imports resolve to this repo's stand-ins -- the recursion, the indexing and
every quirk are the reference's).  Run in the build container only:

    python tests/golden/make_rnea_golden.py      ->  tests/golden/rnea_golden.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

optas = ref_shim.install()

ROBOTS = {
    "tester_robot_revolute.urdf": "/root/reference/tests/tester_robot_revolute.urdf",
    "med7.urdf": "/root/reference/example/robots/kuka_lbr/med7.urdf",
}

if __name__ == "__main__":
    out = {}
    rng = np.random.default_rng(20)
    for name, path in ROBOTS.items():
        model = optas.RobotModel(urdf_filename=path)
        cases = []
        for _ in range(8):
            q, qd, qdd = rng.uniform(-2.0, 2.0, (3, model.ndof))
            tau = model.rnea(q, qd, qdd).toarray().flatten()
            cases.append({"q": q.tolist(), "qd": qd.tolist(), "qdd": qdd.tolist(), "tau": tau.tolist()})
        out[name] = cases
        print(name, model.ndof, "dof", cases[0]["tau"])
    with open(os.path.join(HERE, "rnea_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)
