"""Inverse dynamics (reference optas/models.py:1731-1884, tests/test_models.py:1012-1052).

The reference pins `rnea` only against pybullet (absent here) at atol 8e-2.  Here: the numpy oracle
(oracle/rnea_ref.py) is pinned by physics that does not go through the recursion, and the package's symbolic `rnea`
must agree with that oracle to round-off; on the GPU the same expression graph runs through the streaming kernel."""
import os

import numpy as np
import pytest

import rnea_ref

from optas_b200 import problems, sym as cs
from optas_b200.models import JointTypeNotSupported, RobotModel

HERE = os.path.dirname(os.path.abspath(__file__))
REVOLUTE = os.path.join(HERE, "golden", "tester_robot_revolute.urdf")  # the reference's own RNEA test robot
ROBOTS = [problems.MED7_URDF, REVOLUTE]


@pytest.mark.parametrize("path", [problems.MED7_URDF, REVOLUTE])
def test_oracle_is_pinned_by_physics(path):
    n = RobotModel(urdf_filename=path).ndof
    rng = np.random.default_rng(1)
    z, I = np.zeros(n), np.eye(n)
    for _ in range(4):
        q, qd = rng.uniform(-1.5, 1.5, (2, n))
        g = rnea_ref.rnea(path, q, z, z)
        h = 1e-6
        dV = np.array([(rnea_ref.potential_energy(path, q + h * I[j]) - rnea_ref.potential_energy(path, q - h * I[j])) / (2 * h)
                       for j in range(n)])
        assert np.abs(g - dV).max() < 1e-7          # gravity torques = gradient of the potential energy
        M = np.stack([rnea_ref.rnea(path, q, z, I[j]) - g for j in range(n)], axis=1)
        assert np.abs(M - M.T).max() < 1e-12 and np.linalg.eigvalsh(0.5 * (M + M.T)).min() > 0.0
        c1, c2 = rnea_ref.rnea(path, q, qd, z) - g, rnea_ref.rnea(path, q, 2 * qd, z) - g
        assert np.abs(c2 - 4 * c1).max() < 1e-11     # Coriolis / centrifugal terms are quadratic in qd
    if path == problems.MED7_URDF:                   # not a vacuous check: gravity loads the shoulder
        assert abs(g[1]) > 1.0


@pytest.mark.parametrize("path", ROBOTS)
def test_model_rnea_matches_oracle(path):
    robot = RobotModel(urdf_filename=path)
    rng = np.random.default_rng(2)
    for _ in range(10):  # the reference draws all three from get_random_joint_positions (tests/test_models.py:1041-1045)
        q, qd, qdd = (robot.get_random_joint_positions().toarray().flatten() for _ in range(3))
        tau = robot.rnea(q, qd, qdd)
        assert isinstance(tau, cs.DM) and tau.shape == (robot.ndof, 1)
        assert np.abs(tau.toarray().flatten() - rnea_ref.rnea(path, q, qd, qdd)).max() < 1e-11
    del rng


def test_rnea_symbolic_and_unsupported_joints():
    robot = RobotModel(urdf_filename=REVOLUTE)
    q, qd, qdd = (cs.SX.sym(s, robot.ndof) for s in ("q", "qd", "qdd"))
    assert isinstance(robot.rnea(q, qd, qdd), cs.SX)               # tests/test_models.py:1033-1038
    prismatic = RobotModel(urdf_filename=os.path.join(HERE, "golden", "tester_robot.urdf"))
    with pytest.raises(JointTypeNotSupported):                      # models.py:1743-1746
        prismatic.rnea(np.zeros(3), np.zeros(3), np.zeros(3))
    lwr = RobotModel(urdf_filename=problems.LWR_URDF)               # its first joint is actuated: models.py:1748-1749
    with pytest.raises(JointTypeNotSupported):
        lwr.rnea(np.zeros(7), np.zeros(7), np.zeros(7))


@pytest.mark.gpu
def test_rnea_batch_on_the_streaming_kernel():
    """The RNEA graph of the LBR Med7 (7 torques from q, qd, qdd: 168 B in, 56 B out per evaluation) through
    bo_eval_kernel for a batch, against the numpy oracle."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from optas_b200.function import B200Function

    robot = RobotModel(urdf_filename=problems.MED7_URDF)
    q, qd, qdd = (cs.SX.sym(s, robot.ndof) for s in ("q", "qd", "qdd"))
    fun = B200Function(cs.Function("rnea", [q, qd, qdd], [robot.rnea(q, qd, qdd)]))
    rng = np.random.default_rng(3)
    B = 100_003  # ragged tail tile
    Q, QD, QDD = rng.uniform(-2.0, 2.0, (3, B, robot.ndof))
    tau = fun(Q, QD, QDD)
    assert tau.shape == (B, 7)
    for b in list(range(64)) + [B - 1]:
        ref = rnea_ref.rnea(problems.MED7_URDF, Q[b], QD[b], QDD[b])
        assert np.abs(tau[b] - ref).max() < 1e-10 * max(1.0, np.abs(ref).max())
    # size-independent property at full batch: linear in qdd for fixed (q, qd)
    tau0 = fun(Q, QD, np.zeros_like(QDD))
    tau2 = fun(Q, QD, 2.0 * QDD)
    assert np.abs((tau2 - tau0) - 2.0 * (tau - tau0)).max() < 1e-9


@pytest.mark.parametrize("name,path", [("tester_robot_revolute.urdf", REVOLUTE), ("med7.urdf", problems.MED7_URDF)])
def test_oracle_and_model_match_the_reference_golden_vectors(name, path):
    """tests/golden/rnea_golden.json: outputs of the UNMODIFIED reference `RobotModel.rnea` (run through
    tests/golden/ref_shim.py by tests/golden/make_rnea_golden.py).  The oracle and this package's rnea must both
    reproduce them to round-off."""
    import json

    cases = json.load(open(os.path.join(HERE, "golden", "rnea_golden.json")))[name]
    robot = RobotModel(urdf_filename=path)
    assert len(cases) >= 8
    for c in cases:
        tau = np.array(c["tau"])
        scale = max(1.0, np.abs(tau).max())
        assert np.abs(rnea_ref.rnea(path, c["q"], c["qd"], c["qdd"]) - tau).max() < 1e-12 * scale
        assert np.abs(robot.rnea(c["q"], c["qd"], c["qdd"]).toarray().flatten() - tau).max() < 1e-12 * scale
