"""The host-side mirror of the reference interfaces (optas_b200.{sym,spatialmath,sx_container,models,
builder,optimization}) against the semantics the reference's own tests pin.  Each test names the
reference test it replays.  (The UNMODIFIED reference package also runs on the expression layer here:
tests/golden/ref_shim.py; that check needs /root/reference and is not part of this suite.)"""
import os

import numpy as np
import pytest
from scipy.spatial.transform import Rotation as Rot

import optas_b200 as optas
from optas_b200 import problems
from optas_b200.sx_container import SXContainer

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NUM_RANDOM = 50


def isclose(A, B):
    return np.isclose(np.asarray(A, dtype=float), np.asarray(B, dtype=float)).all()


# ---- tests/test_optas_utils.py ------------------------------------------------------------------
def test_deg2rad_rad2deg_clip():
    x = np.random.default_rng(0).uniform(-400, 400, 10)
    assert isclose(optas.deg2rad(x).toarray().flatten(), np.deg2rad(x))
    assert isclose(optas.rad2deg(x).toarray().flatten(), np.rad2deg(x))
    assert isclose(optas.clip(x, -1.0, 2.0).toarray().flatten(), np.clip(x, -1.0, 2.0))
    s = optas.SX.sym("s", 3)
    assert isinstance(optas.deg2rad(s), optas.SX)


# ---- tests/test_sx_container.py -----------------------------------------------------------------
def test_sx_container_semantics():
    a, b = SXContainer(), SXContainer()
    a["x"] = optas.SX.sym("x", 2, 3)
    b["y"] = optas.SX.sym("y", 4)
    c = a + b
    assert list(c.keys()) == ["x", "y"] and c.numel() == 10
    with pytest.raises(KeyError):
        c["x"] = optas.SX.sym("x", 2)
    # dict2vec: missing labels -> zeros, unknown labels ignored, column-major flattening (test :41-51)
    v = c.dict2vec({"x": np.arange(6).reshape(2, 3), "z": np.ones(5)}).toarray().flatten()
    assert isclose(v, [0, 3, 1, 4, 2, 5, 0, 0, 0, 0])
    d = c.vec2dict(np.arange(10.0))
    assert isclose(d["x"].toarray(), np.arange(6.0).reshape(3, 2).T) and d["y"].shape == (4, 1)
    assert all(isclose(z.toarray(), 0) for z in c.zero().values())


# ---- tests/test_spatialmath.py -------------------------------------------------------------------
def test_spatialmath_against_scipy():
    rng = np.random.default_rng(1)
    for _ in range(NUM_RANDOM):
        th = rng.uniform(-2 * np.pi, 2 * np.pi)
        assert isclose(optas.rotx(th).toarray(), Rot.from_euler("x", th).as_matrix())
        assert isclose(optas.roty(th).toarray(), Rot.from_euler("y", th).as_matrix())
        assert isclose(optas.rotz(th).toarray(), Rot.from_euler("z", th).as_matrix())
        v = rng.uniform(-1, 1, 3)
        v /= np.linalg.norm(v)
        assert isclose(optas.angvec2r(th, v).toarray(), Rot.from_rotvec(th * v).as_matrix())
        rpy = rng.uniform(-np.pi, np.pi, 3)
        assert isclose(optas.rpy2r(rpy).toarray(), Rot.from_euler("xyz", rpy).as_matrix())
        T = optas.rt2tr(optas.rpy2r(rpy), v)
        assert isclose((optas.invt(T) @ T).toarray(), np.eye(4))
        q = optas.Quaternion.fromrpy(rpy).getquat().toarray().flatten()
        assert isclose(q, Rot.from_euler("xyz", rpy).as_quat()) or isclose(-q, Rot.from_euler("xyz", rpy).as_quat())
    # list / tuple / 1-D arguments become column vectors (arrayify_args, test :75-81)
    assert optas.skew([1.0, 2.0, 3.0]).shape == (3, 3)
    assert isinstance(optas.rotz(optas.SX.sym("t")), optas.SX)


def test_quaternion_product_order():
    """q0 * q1 composes like Rot(q1) * Rot(q0) (tests/test_spatialmath.py:415-425)."""
    rng = np.random.default_rng(2)
    for _ in range(NUM_RANDOM):
        a, b = Rot.random(random_state=int(rng.integers(1 << 30))).as_quat(), Rot.random(random_state=int(rng.integers(1 << 30))).as_quat()
        q = (optas.Quaternion(*a) * optas.Quaternion(*b)).getquat().toarray().flatten()
        ref = (Rot.from_quat(b) * Rot.from_quat(a)).as_quat()
        assert isclose(q, ref) or isclose(-q, ref)


# ---- tests/test_optimization.py -------------------------------------------------------------------
def test_derive_jacobian_and_hessian_functions():
    """tests/test_optimization.py:21-41."""
    from optas_b200.optimization import derive_jacobian_and_hessian_functions

    x, p = optas.SX.sym("x", 2), optas.SX.sym("p", 1)
    fun = optas.Function("fun", [x, p], [p * x[0] ** 2 + x[0] * x[1] ** 3])
    J, H = derive_jacobian_and_hessian_functions("fun", fun, x, p)
    xv, pv = np.array([1.5, -0.7]), np.array([2.0])
    assert isclose(J(xv, pv).toarray().flatten(), [2 * pv[0] * xv[0] + xv[1] ** 3, 3 * xv[0] * xv[1] ** 2])
    assert isclose(H(xv, pv).toarray(), [[2 * pv[0], 3 * xv[1] ** 2], [3 * xv[1] ** 2, 6 * xv[0] * xv[1]]])


def test_vertcon_ordering_and_ir_dimensions():
    """v = [k; g; a; -a; h; -h], lbv = 0, ubv = 1e10 (tests/test_optimization.py:44-70, :289-292)."""
    b = optas.OptimizationBuilder(T=1, tasks=optas.TaskModel("t", 3, time_derivs=[0]))
    x = b.get_model_state("t", 0)
    p = b.add_parameter("p", 3)
    b.add_cost_term("c", optas.sumsqr(x - p))
    b.add_leq_inequality_constraint("k", x[0], 2.0)                  # linear inequality: 2 - x0 >= 0
    b.add_geq_inequality_constraint("g", optas.cos(x[1]), -0.5)      # nonlinear inequality
    b.add_equality_constraint("a", x[0] + x[2], p[0])                # linear equality
    b.add_equality_constraint("h", x[1] * x[2], 1.0)                 # nonlinear equality
    opt = b.build()
    assert type(opt).__name__ == "QuadraticCostNonlinearConstraints"
    assert (opt.nx, opt.np, opt.nk, opt.ng, opt.na, opt.nh, opt.nv) == (3, 3, 1, 1, 1, 1, 6)
    xv, pv = np.array([0.3, -0.4, 1.7]), np.array([0.5, 0.1, -0.2])
    k, g = 2.0 - xv[0], np.cos(xv[1]) + 0.5
    a, h = pv[0] - (xv[0] + xv[2]), 1.0 - xv[1] * xv[2]
    assert isclose(opt.v(xv, pv).toarray().flatten(), [k, g, a, -a, h, -h])
    assert isclose(opt.lbv.toarray(), 0.0) and isclose(opt.ubv.toarray(), 1e10)
    # quadratic cost: f = x'Px + q'x + const, P = 0.5 ddf, q = df(0) (optimization.py:219-223)
    P, q = opt.P(pv).toarray(), opt.q(pv).toarray().flatten()
    assert isclose(P, np.eye(3)) and isclose(q, -2 * pv)
    # linear families: k = Mx + c, a = Ax + b
    assert isclose(opt.M(pv).toarray() @ xv + opt.c(pv).toarray().flatten(), [k])
    assert isclose(opt.A(pv).toarray() @ xv + opt.b(pv).toarray().flatten(), [a])


# ---- tests/test_builder.py -------------------------------------------------------------------------
def test_builder_counts_and_dispatch():
    T = 10
    task = optas.TaskModel("test", 3, time_derivs=[0, 1], dlim={0: [[-1, -2, -3], [1, 2, 3]]})
    for align, n_dx in ((False, T - 1), (True, T)):
        b = optas.OptimizationBuilder(T, tasks=task, derivs_align=align)
        assert b.get_model_states("test").shape == (3, T) and b.get_model_states("test", 1).shape == (3, n_dx)
        b.integrate_model_states("test", 1, 0.1)
        assert b._lin_eq_constraints.numel() == 3 * (T - 1)          # tests/test_builder.py:300-311
        b.enforce_model_limits("test")
        assert b._lin_ineq_constraints.numel() == 2 * 3 * T          # :313-326 (3x1 limits broadcast over 3xT)
    b = optas.OptimizationBuilder(T, tasks=task)
    X = b.get_model_states("test")
    b.add_equality_constraint("lin", X - 2.0)                        # affine -> linear family (:248-298)
    b.add_equality_constraint("nonlin", optas.cos(X[0, 0]) + 1.0)
    assert b._lin_eq_constraints.numel() == 3 * T and b._eq_constraints.numel() == 1
    # class dispatch (:358-419)
    def build(cost, con=None):
        bb = optas.OptimizationBuilder(T, tasks=optas.TaskModel("m", 2, time_derivs=[0]))
        Y = bb.get_model_states("m")
        bb.add_cost_term("c", cost(Y))
        if con == "lin":
            bb.add_bound_inequality_constraint("lim", -100, Y, 100)
        elif con == "nonlin":
            bb.add_equality_constraint("e", optas.sumsqr(Y), 1.0)
        return type(bb.build()).__name__
    quad, nonl = (lambda Y: optas.sumsqr(Y)), (lambda Y: optas.sumsqr(Y) + optas.cos(Y[0, 0] * Y[1, 0]))
    assert build(quad) == "QuadraticCostUnconstrained" and build(quad, "lin") == "QuadraticCostLinearConstraints"
    assert build(quad, "nonlin") == "QuadraticCostNonlinearConstraints" and build(nonl) == "NonlinearCostUnconstrained"
    assert build(nonl, "lin") == "NonlinearCostLinearConstraints" and build(nonl, "nonlin") == "NonlinearCostNonlinearConstraints"
    with pytest.raises(AssertionError):
        optas.OptimizationBuilder(1, tasks=optas.TaskModel("m", 2, time_derivs=[0, 1]))  # T too low


# ---- tests/test_models.py (closed form instead of roboticstoolbox) -------------------------------------
def test_robot_model_naming_limits_and_fk():
    robot = optas.RobotModel(urdf_filename=os.path.join(GOLDEN, "tester_robot.urdf"), time_derivs=[0, 1])
    assert robot.get_name() == "test_robot" and robot.ndof == 3
    assert robot.state_name(0) == "test_robot/q" and robot.state_optimized_name(1) == "test_robot/dq/x"
    assert robot.state_parameter_name(0) == "test_robot/q/p"
    lo, up = robot.get_limits(0)
    assert isclose(lo.toarray().flatten()[1:], [-1.0, 0.0]) and isclose(up.toarray().flatten()[1:], [1.0, 1.0])
    assert lo.toarray().flatten()[0] <= -1e9 and up.toarray().flatten()[0] >= 1e9      # continuous joint (models.py:444-466)
    fk = robot.get_global_link_position_function("eff")
    rng = np.random.default_rng(3)
    for _ in range(NUM_RANDOM):
        q = rng.uniform(-1.5, 1.5, 3)
        ref = [2 * np.cos(q[0]) + np.cos(q[0] + q[1]), 2 * np.sin(q[0]) + np.sin(q[0] + q[1]), q[2] + 0.5]
        assert isclose(fk(q).toarray().flatten(), ref)
    # relative frames: T_link_base = T_link_world @ inv(T_base_world) (models.py:898; test_models.py:505-511)
    q = rng.uniform(-1, 1, 3)
    Tl, Tb = robot.get_global_link_transform("eff", q), robot.get_global_link_transform("link2", q)
    assert isclose(robot.get_link_transform("eff", q, "link2").toarray(), (Tl @ optas.invt(Tb)).toarray())
    # trajectory functions: n columns in, n columns out (Function.map semantics, models.py:786-787)
    Q = rng.uniform(-1, 1, (3, 5))
    P = robot.get_global_link_position_function("eff", n=5)(Q).toarray()
    assert P.shape == (3, 5) and isclose(P[:, 2], fk(Q[:, 2]).toarray().flatten())


def test_geometric_jacobian_matches_finite_differences():
    prob = problems.lwr_ik()
    robot = prob.models["robot"]
    q = np.random.default_rng(4).uniform(-2, 2, 7)
    J = robot.get_global_link_geometric_jacobian(problems.LWR_EE, q).toarray()
    assert J.shape == (6, 7)
    h = 1e-6
    for j in range(7):
        dq = np.zeros(7)
        dq[j] = h
        num = (robot.get_global_link_position(problems.LWR_EE, q + dq) - robot.get_global_link_position(problems.LWR_EE, q - dq)).toarray().flatten() / (2 * h)
        assert np.abs(J[:3, j] - num).max() < 1e-8
    Jl = robot.get_global_link_linear_jacobian(problems.LWR_EE, q).toarray()
    assert isclose(Jl, J[:3])
