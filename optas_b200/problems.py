"""The workload problems (BASELINE.json `configs`), built through this package's own builder the
way the reference's example scripts build them, plus their synthetic batch inputs.

Each factory returns a ``Problem``: the built ``Optimization``, samplers for the batched
parameter / seed matrices in ``vec()`` layout, and the model functions evaluated by the streaming
kernel.  Robot descriptions are the kinematics + inertial URDFs under optas_b200/robots/ (derived from
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
        goals = _eval_rows(fk, q_rand)
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




# ----------------------------------------------------------------------------------------------
# C3: point-mass MPC tick (reference: example/point_mass_mpc.py:88-154, class Controller)
# ----------------------------------------------------------------------------------------------


def point_mass_mpc(T: int = 20) -> Problem:
    dt, obs_rad, pm_radius = 0.05, 0.2, 0.1
    point_mass = TaskModel("point_mass", 2, time_derivs=[0, 1], dlim={0: [-1.5, 1.5], 1: [-1, 1]})
    name = point_mass.get_name()
    builder = OptimizationBuilder(T, tasks=point_mass, derivs_align=True)
    curr = builder.add_parameter("curr", 2)
    dcurr = builder.add_parameter("dcurr", 2)
    goal = builder.add_parameter("goal", 2, T)
    obs = builder.add_parameter("obs", 2, T)
    builder.enforce_model_limits(name, time_deriv=0)
    builder.enforce_model_limits(name, time_deriv=1)
    builder.integrate_model_states(name, time_deriv=1, dt=dt)
    builder.fix_configuration(name, config=curr)
    builder.fix_configuration(name, config=dcurr, time_deriv=1)
    X = builder.get_model_states(name)
    safe_dist_sq = (obs_rad + pm_radius) ** 2
    for i in range(T):
        builder.add_geq_inequality_constraint(f"obs_avoid_{i}", cs.sumsqr(obs[:, i] - X[:, i]), safe_dist_sq)
    builder.add_cost_term("optimal_path", cs.sumsqr(goal - X))
    dX = builder.get_model_states(name, time_deriv=1)
    w = 0.0025 / float(T)
    ddX = (dX[:, 1:] - dX[:, :-1]) / dt
    builder.add_cost_term("minimize_acceleration", w * cs.sumsqr(ddX))
    opt = builder.build()

    def sample(B: int, seed: int = 1):
        """C3 inputs: random current state away from the obstacle, obstacle on the
        script's circular path (point_mass_mpc.py:293-306), goal ramp toward (1, 1); seed = hold-position trajectory."""
        rng = np.random.default_rng(seed)
        t = rng.uniform(0.0, 4.4, size=B)
        k = np.arange(T)
        phase = np.pi * (t[:, None] + dt * k[None, :]) - np.pi
        obs_path = 0.15 * np.stack([np.sin(phase), np.cos(phase) + 1.0], axis=1)  # [B, 2, T]
        cur = np.empty((B, 2))
        todo = np.ones(B, dtype=bool)
        while todo.any():
            cand = rng.uniform(-1.2, 1.2, size=(int(todo.sum()), 2))
            far = np.linalg.norm(cand - obs_path[todo, :, 0], axis=1) >= 0.35
            idx = np.where(todo)[0]
            cur[idx[far]] = cand[far]
            todo[idx[far]] = False
        dcur = rng.uniform(-0.5, 0.5, size=(B, 2))
        ramp = np.minimum(1.0, dt * k / 4.4)
        goal_path = cur[:, :, None] + (1.0 - cur[:, :, None]) * ramp[None, None, :]
        P = np.concatenate([cur, dcur, goal_path.transpose(0, 2, 1).reshape(B, -1), obs_path.transpose(0, 2, 1).reshape(B, -1)],
                           axis=1)
        # seed: hold the current position with zero velocity (layout: y/x 2 x T column-major, then dy/x).
        # The all-zero seed of the script's first tick puts every knot at the origin, which lies ON the
        # obstacle's path: the obstacle constraints start violated with (near-)zero gradient.
        X0 = np.concatenate([np.tile(cur, (1, T)), np.zeros((B, 2 * T))], axis=1)
        return np.ascontiguousarray(P), np.ascontiguousarray(X0)

    return Problem("point_mass_mpc", opt, sample, {}, {"point_mass": point_mass})




# ----------------------------------------------------------------------------------------------
# C5: dual-arm planner (reference: example/dual_arm.py:17-128 and :130-144)
# ----------------------------------------------------------------------------------------------

DUAL_ARM_Q0 = np.deg2rad([0.0, -30.0, 0.0, 90.0, 0.0, 30.0, 0.0])


def dual_arm(T: int = 50) -> Problem:
    Tmax, link_ee = 10.0, LWR_EE
    dt = Tmax / float(T - 1)

    def arm(name, base_position):
        model = RobotModel(urdf_filename=LWR_URDF, name=name, time_derivs=[0, 1])
        model.add_base_frame("global_world", xyz=base_position)
        return model

    kukal, kukar = arm("kukal", [0.0, -0.25, 0.0]), arm("kukar", [0.0, 0.25, 0.0])
    builder = OptimizationBuilder(T=T, robots=[kukal, kukar])
    qcl = builder.add_parameter("qcl", kukal.ndof)
    qcr = builder.add_parameter("qcr", kukar.ndof)
    builder.fix_configuration("kukal", qcl)
    builder.fix_configuration("kukar", qcr)
    builder.integrate_model_states("kukal", time_deriv=1, dt=dt)
    builder.integrate_model_states("kukar", time_deriv=1, dt=dt)
    posl_ee = kukal.get_global_link_position_function(link_ee, n=T)
    posr_ee = kukar.get_global_link_position_function(link_ee, n=T)
    Ql, Qr = builder.get_model_states("kukal"), builder.get_model_states("kukar")
    ee_pos_pathl, ee_pos_pathr = posl_ee(Ql), posr_ee(Qr)
    dQl, dQr = builder.get_model_states("kukal", time_deriv=1), builder.get_model_states("kukar", time_deriv=1)
    w_dq = 0.01
    builder.add_cost_term("kukal_min_join_vel", w_dq * cs.sumsqr(dQl))
    builder.add_cost_term("kukar_min_join_vel", w_dq * cs.sumsqr(dQr))
    pos0l = kukal.get_global_link_position(link_ee, qcl)
    pos0r = kukar.get_global_link_position(link_ee, qcr)
    pos1l, pos1r = pos0l + cs.DM([-0.1, 0.1, -0.2]), pos0r + cs.DM([-0.1, -0.1, -0.2])
    pos2l, pos2r = pos1l + cs.DM([0.0, 0.0, 0.3]), pos1r + cs.DM([0.0, 0.0, 0.3])
    path_eel, path_eer = cs.SX.zeros(3, T), cs.SX.zeros(3, T)
    for i in range(T):
        alpha_ = float(i) / float(T - 1)
        if alpha_ < 0.4:
            alpha = alpha_ / 0.4
            path_eel[:, i] = alpha * pos1l + (1.0 - alpha) * pos0l
            path_eer[:, i] = alpha * pos1r + (1.0 - alpha) * pos0r
        elif alpha_ < 0.5:
            path_eel[:, i], path_eer[:, i] = pos1l, pos1r
        else:
            alpha = (alpha_ - 0.5) / 0.5
            path_eel[:, i] = alpha * pos2l + (1.0 - alpha) * pos1l
            path_eer[:, i] = alpha * pos2r + (1.0 - alpha) * pos1r
    builder.add_cost_term("ee_pos_pathl", cs.sumsqr(ee_pos_pathl - path_eel))
    builder.add_cost_term("ee_pos_pathr", cs.sumsqr(ee_pos_pathr - path_eer))
    opt = builder.build()

    def sample(B: int, seed: int = 3):
        """C5 inputs: qcl, qcr = deg2rad[0,-30,0,90,0,30,0] + N(0, 0.1^2) each; seed: both arms
        hold their start configuration with zero velocity."""
        rng = np.random.default_rng(seed)
        qcl_ = DUAL_ARM_Q0 + 0.1 * rng.standard_normal((B, 7))
        qcr_ = DUAL_ARM_Q0 + 0.1 * rng.standard_normal((B, 7))
        P = np.concatenate([qcl_, qcr_], axis=1)
        X0 = np.concatenate([np.tile(qcl_, (1, T)), np.zeros((B, 7 * (T - 1))), np.tile(qcr_, (1, T)),
                             np.zeros((B, 7 * (T - 1)))], axis=1)
        return np.ascontiguousarray(P), np.ascontiguousarray(X0)

    return Problem("dual_arm", opt, sample, {}, {"kukal": kukal, "kukar": kukar})


# ----------------------------------------------------------------------------------------------
# C4: figure-of-eight trajectory optimisation (reference: example/figure_eight_plan.py:16-113)
# ----------------------------------------------------------------------------------------------

MED7_EE = "lbr_link_ee"
FIG8_Q0 = np.deg2rad([0.0, 30.0, 0.0, -90.0, 0.0, -30.0, 0.0])


def figure_eight(T: int = 50, joint_limits: bool = True) -> Problem:
    """``joint_limits=True`` adds ``enforce_model_limits`` on q: BASELINE.json's C4 says "with joint-limit
    ineq"; the reference script itself has none."""
    Tmax = 10.0
    t = cs.linspace(0, Tmax, T)
    dt = Tmax / float(T - 1)
    kuka = RobotModel(urdf_filename=MED7_URDF, time_derivs=[0, 1])
    name = kuka.get_name()
    builder = OptimizationBuilder(T=T, robots=[kuka])
    qc = builder.add_parameter("qc", kuka.ndof)
    builder.fix_configuration(name, config=qc)
    builder.fix_configuration(name, time_deriv=1)
    builder.integrate_model_states(name, time_deriv=1, dt=dt)
    Q = builder.get_model_states(name)
    pos_ee = kuka.get_global_link_position_function(MED7_EE, n=T)(Q)
    pc = kuka.get_global_link_position(MED7_EE, qc)
    Rc = kuka.get_global_link_rotation(MED7_EE, qc)
    quatc = kuka.get_global_link_quaternion(MED7_EE, qc)
    path = cs.SX.zeros(3, T)
    path[0, :] = 0.2 * cs.sin(t * cs.pi * 0.5).T
    path[1, :] = 0.1 * cs.sin(t * cs.pi).T
    for k in range(T):
        path[:, k] = pc + Rc @ path[:, k]
    builder.add_cost_term("ee_path", 1000.0 * cs.sumsqr(path - pos_ee))
    dQ = builder.get_model_states(name, time_deriv=1)
    builder.add_cost_term("min_join_vel", 0.01 * cs.sumsqr(dQ))
    quat = kuka.get_global_link_quaternion_function(MED7_EE, n=T)
    builder.add_equality_constraint("no_eff_rot", quat(Q), quatc)
    if joint_limits:
        builder.enforce_model_limits(name)
    opt = builder.build()
    lo = kuka.lower_actuated_joint_limits.toarray().flatten()
    up = kuka.upper_actuated_joint_limits.toarray().flatten()

    def sample(B: int, seed: int = 2):
        """C4 inputs: qc = deg2rad[0,30,0,-90,0,-30,0] + N(0, 0.1^2) clipped to 0.95 limits;
        seed q/x = qc tiled over the horizon, dq/x = 0 (figure_eight_plan.py:123-124)."""
        rng = np.random.default_rng(seed)
        qc_ = np.clip(FIG8_Q0 + 0.1 * rng.standard_normal((B, 7)), 0.95 * lo, 0.95 * up)
        X0 = np.concatenate([np.tile(qc_, (1, T)), np.zeros((B, 7 * (T - 1)))], axis=1)
        return np.ascontiguousarray(qc_), np.ascontiguousarray(X0)

    return Problem("figure_eight", opt, sample, {}, {"robot": kuka})


# ----------------------------------------------------------------------------------------------
# Differential IK as a QP (reference: example/planar_idk.py:12-70) -- the "next" row 8f-1
# ----------------------------------------------------------------------------------------------

PLANAR_URDF = os.path.join(ROBOTS, "planar_3dof.urdf")


def planar_idk() -> Problem:
    robot = RobotModel(urdf_filename=PLANAR_URDF, time_derivs=[1])
    name, link_ee = robot.get_name(), "end"
    builder = OptimizationBuilder(T=1, robots=[robot], derivs_align=True)
    dq = builder.get_model_states(name, time_deriv=1)
    q = builder.add_parameter("q", robot.ndof)
    dx = builder.add_parameter("dx", 2)  # the script hard-codes dx = [0.01, 0]; a parameter here so it can be batched
    J = robot.get_global_link_linear_jacobian_function(link_ee)
    quat = robot.get_global_link_quaternion_function(link_ee)
    phi = lambda q_: 2.0 * cs.atan2(quat(q_)[2], quat(q_)[3])
    J_phi = robot.get_global_link_angular_geometric_jacobian_function(link=link_ee)
    dt, lim = 0.01, 0.1
    builder.add_cost_term("cost", cs.sumsqr(dq))
    builder.add_equality_constraint("FDK", (J(q)[0:2, :]) @ dq, dx)
    builder.add_bound_inequality_constraint("joint", [-lim] * 3, dq, [lim] * 3)
    builder.add_bound_inequality_constraint("task", -70 * (cs.pi / 180.0), phi(q) + dt * (J_phi(q)[2, :]) @ dq, 0.0)
    opt = builder.build()

    def sample(B: int, seed: int = 5):
        rng = np.random.default_rng(seed)
        q_ = np.array([2.39, -2.55, -0.46]) + 0.05 * rng.standard_normal((B, 3))
        dx_ = np.array([0.01, 0.0]) + 0.002 * rng.standard_normal((B, 2))
        return np.ascontiguousarray(np.concatenate([q_, dx_], axis=1)), np.zeros((B, 3))

    return Problem("planar_idk", opt, sample, {"J": J}, {"robot": robot})


def lwr_diff_ik_qp(dt: float = 0.01) -> Problem:
    """Differential IK as a QP (reference: example/experiment1.py:14-128, class ExprIK / IK1 -- there on the OSQP interface,
    3000 ticks): min |dq|^2 + 1000 |J(qc) dq - v_goal|^2  s.t.  joint limits on qc + dt dq, z-band on the end effector.
    nx 7, no equalities, 16 inequality rows; J, the limits and the z-band depend on the parameters (qc, xydir) only."""
    robot = RobotModel(urdf_filename=LWR_URDF, time_derivs=[1])
    name = robot.get_name()
    builder = OptimizationBuilder(T=1, robots=[robot], derivs_align=True)
    qd = builder.get_model_state(name, 0, time_deriv=1)
    qc = builder.add_parameter("qc", robot.ndof)
    xydir = builder.add_parameter("xydir", 2)
    J = robot.get_global_link_geometric_jacobian(LWR_EE, qc)
    veff = J @ qd
    builder.add_cost_term("min_qd", cs.sumsqr(qd))
    vg = cs.vertcat(0.1 * xydir, cs.DM.zeros(4))
    builder.add_cost_term("eff_motion", 1000.0 * cs.sumsqr(veff - vg))
    lo, up = robot.lower_actuated_joint_limits, robot.upper_actuated_joint_limits
    qn = qc + dt * qd
    builder.add_leq_inequality_constraint("lower_qlim", lo, qn)
    builder.add_leq_inequality_constraint("upper_qlim", qn, up)
    pn = robot.get_global_link_position(LWR_EE, qc) + dt * veff[:3]
    builder.add_leq_inequality_constraint("lower_zlim", 0.025, pn[2])
    builder.add_leq_inequality_constraint("upper_zlim", pn[2], 1.5)
    opt = builder.build()

    def sample(B: int, seed: int = 6):
        rng = np.random.default_rng(seed)
        qc_ = LWR_Q_NOMINAL + 0.3 * rng.standard_normal((B, 7))
        ang = rng.uniform(0.0, 2.0 * np.pi, B)
        return np.ascontiguousarray(np.concatenate([qc_, np.cos(ang)[:, None], np.sin(ang)[:, None]], axis=1)), np.zeros((B, 7))

    return Problem("lwr_diff_ik_qp", opt, sample, {}, {"robot": robot})


def box_qp(n: int = 8, n_eq: int = 3, seed: int = 0) -> Problem:
    """A family of strictly convex QPs with active bounds (test workload of the QP path): min x'Px + q'x  s.t.  A x = b,
    -1 <= x <= 1, with P, A fixed and (q, b) the parameters -- about n/3 of the bounds are active at the solution."""
    rng = np.random.default_rng(seed)
    L = rng.standard_normal((n, n))
    Pm = L @ L.T / n + 0.05 * np.eye(n)
    A = rng.standard_normal((n_eq, n))
    task = TaskModel("v", n, time_derivs=[0])
    builder = OptimizationBuilder(T=1, tasks=[task])
    x = builder.get_model_state("v", 0)
    q = builder.add_parameter("q", n)
    builder.add_cost_term("quad", x.T @ cs.DM(Pm) @ x + q.T @ x)
    if n_eq > 0:
        rhs = builder.add_parameter("rhs", n_eq)
        builder.add_equality_constraint("lin", cs.DM(A) @ x, rhs)
    builder.add_bound_inequality_constraint("box", [-1.0] * n, x, [1.0] * n)
    opt = builder.build()

    def sample(B: int, seed: int = 7):
        r = np.random.default_rng(seed)
        return (np.ascontiguousarray(np.concatenate([3.0 * r.standard_normal((B, n)), 0.3 * r.standard_normal((B, n_eq))], axis=1)),
                np.zeros((B, n)))

    return Problem(f"box_qp_{n}_{n_eq}", opt, sample, {}, {"P": Pm, "A": A})


# ----------------------------------------------------------------------------------------------
# "Next" row 8f-3: the rest of the RobotModel surface through the same solver
#   joint-space planner to an end-effector POSE with height constraints on two links
#     (reference: example/simple_joint_space_planner.py:14-71)
#   trajectory optimisation with sphere collision avoidance, three derivative orders
#     (reference: example/sphere_collision_avoidance.py:54-98, builder.py:366-417)
# ----------------------------------------------------------------------------------------------


def joint_space_planner(T: int = 20) -> Problem:
    duration, zpad = 4.0, 0.05
    dt = duration / float(T - 1)
    robot = RobotModel(urdf_filename=MED7_URDF, time_derivs=[0, 1])
    name = robot.get_name()
    builder = OptimizationBuilder(T=T, robots=robot, derivs_align=True)
    qn = builder.add_parameter("nominal_joint_state", robot.ndof)
    qc = builder.add_parameter("current_joint_state", robot.ndof)
    pg = builder.add_parameter("position_goal", 3)
    og = builder.add_parameter("orientation_goal", 4)
    builder.fix_configuration(name, config=qc)
    qF = builder.get_model_state(name, -1)
    builder.add_equality_constraint("final_position", robot.get_global_link_position(MED7_EE, qF), pg)
    builder.add_equality_constraint("final_orientation", robot.get_global_link_quaternion(MED7_EE, qF), og)
    builder.integrate_model_states(name, time_deriv=1, dt=dt)
    for t in range(T):
        q = builder.get_model_state(name, t)
        builder.add_cost_term(f"nominal_{t}", 0.1 * cs.sumsqr(q - qn))
        builder.add_geq_inequality_constraint(f"eff_safe_{t}", robot.get_global_link_position(MED7_EE, q)[2] + zpad)
        builder.add_geq_inequality_constraint(f"elbow_safe_{t}", robot.get_global_link_position("lbr_link_3", q)[2] + zpad)
    dQ = builder.get_model_states(name, time_deriv=1)
    builder.add_cost_term("minimize_velocity", 0.1 * cs.sumsqr(dQ))
    ddQ = (dQ[:, 1:] - dQ[:, :-1]) / dt
    builder.add_cost_term("minimize_acceleration", 10 * cs.sumsqr(ddQ))
    builder.fix_configuration(name, t=-1, time_deriv=1)
    opt = builder.build()

    qs = cs.SX.sym("q", robot.ndof)
    pose = cs.Function("pose", [qs], [cs.vertcat(robot.get_global_link_position(MED7_EE, qs),
                                                  robot.get_global_link_quaternion(MED7_EE, qs))])
    lo = robot.lower_actuated_joint_limits.toarray().flatten()
    up = robot.upper_actuated_joint_limits.toarray().flatten()

    def sample(B: int, seed: int = 6):
        """Current configuration = the script's q0 = deg2rad[0,45,0,-90,0,-45,0] + N(0, 0.05^2); the goal POSE is
        the forward kinematics of a second configuration q0 + N(0, 0.3^2) (reachable by construction; the script's
        single goal is p = [0.4, 0.3, 0.4], quat = [0, 1, 0, 0]); seed: hold the current configuration."""
        rng = np.random.default_rng(seed)
        q0 = np.deg2rad([0.0, 45.0, 0.0, -90.0, 0.0, -45.0, 0.0])
        qc_ = np.clip(q0 + 0.05 * rng.standard_normal((B, 7)), 0.95 * lo, 0.95 * up)
        qg_ = np.clip(q0 + 0.3 * rng.standard_normal((B, 7)), 0.95 * lo, 0.95 * up)
        goal = _eval_rows(pose, qg_)
        P = np.concatenate([np.tile(q0, (B, 1)), qc_, goal], axis=1)
        X0 = np.concatenate([np.tile(qc_, (1, T)), np.zeros((B, 7 * T))], axis=1)
        return np.ascontiguousarray(P), np.ascontiguousarray(X0)

    return Problem("joint_space_planner", opt, sample, {"pose": pose}, {"robot": robot})


SPHERE_LINKS = ["end_effector_ball", "lwr_arm_7_link", "lwr_arm_5_link", "lwr_arm_6_link"]
SPHERE_Q_NOMINAL = np.deg2rad([0.0, -45.0, 0.0, 90.0, 0.0, 45.0, 0.0])
# solution of the script's first stage (`lwr_axis_ik` at start_eff_position = [0.825, -0.35, 0.2], seed = nominal);
# tests/test_solver_logic.py::test_sphere_collision_first_stage_ik re-derives it
SPHERE_Q_START = np.array([-0.38934800247392193, -1.1451497139998776, -0.3501660198135529, 1.2868149461781317,
                           -0.5796474583683217, 1.0978271686956158, 0.0])


def lwr_axis_ik() -> Problem:
    """First stage of example/sphere_collision_avoidance.py (:20-42, `compute_initial_configuration`): IK to an
    end-effector position with the tool z axis kept at its nominal direction, joint limits, nominal-posture cost."""
    robot = RobotModel(urdf_filename=LWR_URDF, time_derivs=[0, 1, 2])
    name = robot.get_name()
    z0 = robot.get_global_link_transform(LWR_EE, SPHERE_Q_NOMINAL)[:3, 2]
    builder = OptimizationBuilder(1, robots=robot, derivs_align=True)
    start = builder.add_parameter("start_eff_position", 3)  # the constant [0.825, -0.35, 0.2] in the script
    q = builder.get_model_state(name, 0)
    Tq = robot.get_global_link_transform(LWR_EE, q)
    builder.enforce_model_limits(name)
    builder.add_equality_constraint("eff_pos", Tq[:3, 3], start)
    builder.add_equality_constraint("eff_ori", Tq[:3, 2], z0)
    builder.initial_configuration(name, time_deriv=1)
    builder.initial_configuration(name, time_deriv=2)
    builder.add_cost_term("nominal", cs.sumsqr(q - SPHERE_Q_NOMINAL))
    opt = builder.build()

    def sample(B: int, seed: int = 8):
        rng = np.random.default_rng(seed)
        P = np.array([0.825, -0.35, 0.2]) + 0.02 * rng.standard_normal((B, 3))
        X0 = np.concatenate([np.tile(SPHERE_Q_NOMINAL, (B, 1)), np.zeros((B, 14))], axis=1)
        return np.ascontiguousarray(P), np.ascontiguousarray(X0)

    return Problem("lwr_axis_ik", opt, sample, {}, {"robot": robot})



def sphere_collision_avoidance(T: int = 20, n_obstacles: int = 6, reduce_constraint: bool = True) -> Problem:
    duration = 10.0
    dt = duration / float(T - 1)
    robot = RobotModel(urdf_filename=LWR_URDF, time_derivs=[0, 1, 2])
    name = robot.get_name()
    qnom = SPHERE_Q_NOMINAL
    z0 = robot.get_global_link_transform(LWR_EE, qnom)[:3, 2]
    builder = OptimizationBuilder(T, robots=robot, derivs_align=True)
    q0 = builder.add_parameter("q0", robot.ndof)  # the script computes it with a first IK solve and bakes it in
    pgoal = builder.add_parameter("goal_eff_position", 3)  # a constant [0.825, 0.35, 0.2] in the script
    builder.enforce_model_limits(name)
    builder.integrate_model_states(name, 2, dt)
    builder.integrate_model_states(name, 1, dt)
    builder.initial_configuration(name, init=q0)
    builder.initial_configuration(name, time_deriv=1)
    builder.initial_configuration(name, time_deriv=2)
    qf = builder.get_model_state(name, t=-1)
    Tf = robot.get_global_link_transform(LWR_EE, qf)
    builder.add_equality_constraint("eff_pos", Tf[:3, 3], pgoal, reduce_constraint=reduce_constraint)
    builder.add_equality_constraint("eff_ori", Tf[:3, 2], z0, reduce_constraint=reduce_constraint)
    builder.add_cost_term("min_vel", cs.sumsqr(builder.get_model_states(name, time_deriv=1)))
    builder.add_cost_term("min_acc", 100 * cs.sumsqr(builder.get_model_states(name, time_deriv=2)))
    builder.add_cost_term("nominal", 1e2 * cs.sumsqr(qf - qnom))
    obstacle_names = [f"obs{i}" for i in range(n_obstacles)]
    builder.sphere_collision_avoidance_constraints(name, obstacle_names, link_names=SPHERE_LINKS)
    opt = builder.build()

    def sample(B: int, seed: int = 7):
        """The script's scene: a column of six spheres (radius 0.1) at x = 0.55, y = 0, z = 0.1 .. 0.6, link radius
        0.15, start at the first stage's IK solution + N(0, 0.02^2), goal position [0.825, 0.35, 0.2] + N(0, 0.02^2)."""
        rng = np.random.default_rng(seed)
        d: Dict[str, np.ndarray] = {"q0": SPHERE_Q_START + 0.02 * rng.standard_normal((B, 7)),
                                    "goal_eff_position": np.array([0.825, 0.35, 0.2]) + 0.02 * rng.standard_normal((B, 3))}
        for ln in SPHERE_LINKS:
            d[ln + "_radii"] = np.full((B, 1), 0.15)
        for i, on in enumerate(obstacle_names):
            d[on + "_position"] = np.tile(np.array([0.55, 0.0, 0.1 * (i + 1)]), (B, 1))
            d[on + "_radii"] = np.full((B, 1), 0.1)
        from .solver import pack_batch

        P, _ = pack_batch(opt.parameters, d)
        X0d = {f"{name}/q/x": np.repeat(d["q0"][:, :, None], T, axis=2)}
        X0, _ = pack_batch(opt.decision_variables, X0d)
        return np.ascontiguousarray(P), np.ascontiguousarray(X0)

    return Problem("sphere_collision_avoidance", opt, sample, {}, {"robot": robot})


ALL_BUILDERS = [lwr_ik, booth, point_mass_mpc, figure_eight, dual_arm, planar_idk, joint_space_planner, lwr_axis_ik, lwr_diff_ik_qp]
