"""The workload problems (BASELINE.json `configs`), built through this package's own builder the
way the reference's example scripts build them, plus their synthetic batch inputs.

Each factory returns a ``Problem``: the built ``Optimization``, samplers for the batched
parameter / seed matrices in ``vec()`` layout, and the model functions evaluated by the streaming
kernel.  Robot descriptions are the kinematics-only URDFs under optas_b200/robots/ (derived from
the reference's assets by tests/golden/make_robot_assets.py).
"""

from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Callable, Dict, Tuple

import numpy as np

from . import sym as cs
from .builder import OptimizationBuilder
from .models import RobotModel, TaskModel
from .solver import unpack_batch

ROBOTS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "robots")
LWR_URDF = os.path.join(ROBOTS, "kuka_lwr.urdf")
MED7_URDF = os.path.join(ROBOTS, "med7.urdf")


@dataclass
class Problem:
    name: str
    opt: object
    sample: Callable[[int, int], Tuple[np.ndarray, np.ndarray]]  # (B, seed) -> P [B, np], X0 [B, nx]
    functions: Dict[str, object] = field(default_factory=dict)
    models: Dict[str, object] = field(default_factory=dict)

    def param_dict(self, P: np.ndarray) -> Dict[str, np.ndarray]:
        return unpack_batch(self.opt.parameters, np.atleast_2d(P))

    def seed_dict(self, X0: np.ndarray) -> Dict[str, np.ndarray]:
        return unpack_batch(self.opt.decision_variables, np.atleast_2d(X0))


# ----------------------------------------------------------------------------------------------
# C1 / C2: KUKA LWR 7-DoF position IK (reference: example/example.py:12-37)
# ----------------------------------------------------------------------------------------------

LWR_EE = "end_effector_ball"
LWR_Q_NOMINAL = np.deg2rad([0.0, 45.0, 0.0, -90.0, 0.0, -45.0, 0.0])


def lwr_ik() -> Problem:
    robot = RobotModel(urdf_filename=LWR_URDF, time_derivs=[0])
    name = robot.get_name()
    builder = OptimizationBuilder(T=1, robots=robot)
    qn = builder.add_parameter("q_nominal", robot.ndof)
    pg = builder.add_parameter("p_goal", 3)
    q = builder.get_model_state(name, 0)
    p = robot.get_global_link_position(LWR_EE, q)
    builder.add_equality_constraint("end_goal", p, pg)
    builder.add_cost_term("nominal", cs.sumsqr(q - qn))
    builder.enforce_model_limits(name)
    opt = builder.build()

    qs = cs.SX.sym("q", robot.ndof)
    pos = robot.get_global_link_position(LWR_EE, qs)
    fk = cs.Function("fk", [qs], [pos])
    fk_jac = cs.Function("fk_jac", [qs], [pos, cs.jacobian(pos, qs)])
    lo = robot.lower_actuated_joint_limits.toarray().flatten()
    up = robot.upper_actuated_joint_limits.toarray().flatten()

    def sample(B: int, seed: int = 0):
        """C2 inputs: q_rand ~ U(0.8 lo, 0.8 up), p_goal = FK(q_rand) (reachable by
        construction), q_nominal fixed, seed x0 = q_nominal for every instance."""
        rng = np.random.default_rng(seed)
        q_rand = rng.uniform(0.8 * lo, 0.8 * up, size=(B, robot.ndof))
        goals = fk._tape_eval_batch(q_rand) if hasattr(fk, "_tape_eval_batch") else _eval_rows(fk, q_rand)
        P = np.concatenate([np.tile(LWR_Q_NOMINAL, (B, 1)), goals], axis=1)
        X0 = np.tile(LWR_Q_NOMINAL, (B, 1))
        return np.ascontiguousarray(P), np.ascontiguousarray(X0)

    return Problem("lwr_ik", opt, sample, {"fk_jac": fk_jac}, {"robot": robot})


def _eval_rows(fun, rows: np.ndarray) -> np.ndarray:
    """Evaluate a one-input, one-output Function on every row (host, numpy tape interpreter; used
    only to *generate* synthetic inputs such as reachable goals)."""
    from .tape import Tape

    tape = Tape.from_function(fun)
    return np.ascontiguousarray(tape.eval_numpy([np.asarray(rows, dtype=float).T])[0].T)


def lwr_ik_example_instance() -> Tuple[np.ndarray, np.ndarray]:
    """C1: exactly example/example.py:44-56 -- p_goal = p(q_nominal) + [0, 0.3, -0.2]; returns (p, x0=q_nominal)."""
    prob = lwr_ik()
    fk = prob.functions["fk_jac"]
    p_nominal = _eval_rows(cs.Function("fk", fk.sx_in(), [fk.sx_out(0)]), LWR_Q_NOMINAL[None, :])[0]
    p_goal = p_nominal + np.array([0.0, 0.3, -0.2])
    return np.concatenate([LWR_Q_NOMINAL, p_goal]), LWR_Q_NOMINAL.copy()


# ----------------------------------------------------------------------------------------------
# Booth function: the one solver known-answer test the reference pins (tests/test_solver.py:19-54)
# ----------------------------------------------------------------------------------------------


def booth() -> Problem:
    task = TaskModel("booth", dim=2, time_derivs=[0])
    builder = OptimizationBuilder(T=1, tasks=task)
    X = builder.get_model_state("booth", 0)
    x, y = X[0], X[1]
    a = builder.add_parameter("a")
    b = builder.add_parameter("b")
    builder.add_cost_term("booth", (x + a * y - b) ** 2 + (2 * x + y - 5) ** 2)
    opt = builder.build()

    def sample(B: int, seed: int = 0):
        P = np.tile(np.array([2.0, 7.0]), (B, 1))
        return P, np.zeros((B, 2))

    return Problem("booth", opt, sample)


ALL_BUILDERS = [lwr_ik, booth]
