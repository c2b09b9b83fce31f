"""Solver LOGIC on a GPU-less machine: the generated solver source (the exact text NVRTC compiles
for sm_100a) is compiled for the host by tests/hostsim.py and checked against the CPU oracle.
This is a harness for CI without a GPU; it is not a product path (see tests/hostsim.py)."""
import numpy as np
import pytest

import kkt_check
import slsqp_driver
from hostsim import HostSim

import optas_b200
from optas_b200 import problems


def _sim(prob, **kw):
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt", compile_only=True, **kw)
    lo = solver._lowered
    return HostSim(solver.kernel_source(), lo.nx, lo.np_, lo.n_eq, lo.n_ineq, ldl_table=solver.ldl_table(),
                   dtable=solver.dtable()), lo


def test_booth_known_answer():
    """Reference tests/test_solver.py:45-54: Booth, a=2, b=7, seed (0,0) -> x=1, y=3, did_solve."""
    prob = problems.booth()
    sim, _ = _sim(prob)
    r = sim.solve(np.array([[2.0, 7.0]]), np.zeros((1, 2)))
    assert r["status"][0] == 0
    assert np.isclose(r["x"][0, 0], 1.0) and np.isclose(r["x"][0, 1], 3.0)


def test_c1_example_instance_matches_golden():
    prob = problems.lwr_ik()
    sim, _ = _sim(prob)
    p, x0 = problems.lwr_ik_example_instance()
    r = sim.solve(p[None, :], x0[None, :])
    golden = np.array([-0.24904555, 1.14296583, -0.12385623, -1.29959808, 0.03860145, -0.66073211, 0.0])
    assert r["status"][0] == 0
    assert np.abs(r["x"][0] - golden).max() < 1e-7
    assert abs(r["f"][0] - 0.29579887518014) < 1e-8


def test_c2_batch_converges_and_passes_the_parity_protocol():
    """(i) oracle KKT residual, (ii) polish check -- the oracle seeded at the result
    stays there, (iii) fraction of instances landing where the oracle lands from the same seed."""
    prob = problems.lwr_ik()
    sim, lo = _sim(prob)
    B = 512
    P, X0 = prob.sample(B, seed=3)
    r = sim.solve(P, X0)
    ok = r["status"] <= 1
    assert ok.mean() >= 0.99
    res = kkt_check.kkt_residual(prob, r["x"][ok], P[ok], r["lam"][ok][:, :lo.n_eq], r["lam"][ok][:, lo.n_eq:])
    assert res.max() < 1e-6, res.max()
    assert np.median(res) < 1e-8
    # polish
    op = slsqp_driver.OracleProblem(prob.opt)
    idx = np.where(ok)[0][:64]
    worst = 0.0
    for i in idx:
        pol = slsqp_driver.solve_slsqp(op, P[i], r["x"][i], form="split", options={"ftol": 1e-15, "maxiter": 200})
        worst = max(worst, np.abs(pol.x - r["x"][i]).max() / max(1.0, np.abs(pol.x).max()))
    assert worst < 1e-6, worst
    # same seed, same basin (informational threshold: the oracle itself fails on ~6 % of these)
    same = 0
    tried = 0
    for i in idx[:32]:
        o = slsqp_driver.solve_slsqp(op, P[i], X0[i], form="split", options={"ftol": 1e-15, "maxiter": 500})
        if o.success:
            tried += 1
            same += np.abs(o.x - r["x"][i]).max() < 1e-5
    assert tried > 0 and same / tried > 0.5


def test_joint_limits_are_respected():
    prob = problems.lwr_ik()
    sim, _ = _sim(prob)
    P, X0 = prob.sample(256, seed=5)
    r = sim.solve(P, X0)
    robot = prob.models["robot"]
    lo = robot.lower_actuated_joint_limits.toarray().flatten()
    up = robot.upper_actuated_joint_limits.toarray().flatten()
    ok = r["status"] <= 1
    assert (r["x"][ok] >= lo - 1e-9).all() and (r["x"][ok] <= up + 1e-9).all()


def test_unreachable_goal_is_reported_not_converged():
    prob = problems.lwr_ik()
    sim, _ = _sim(prob)
    p, x0 = problems.lwr_ik_example_instance()
    p = p.copy()
    p[7:] = [5.0, 5.0, 5.0]  # far outside the workspace
    r = sim.solve(p[None, :], x0[None, :])
    assert r["status"][0] >= 2


def test_pack_unpack_roundtrip():
    from optas_b200.solver import pack_batch, unpack_batch

    prob = problems.lwr_ik()
    P, _ = prob.sample(5, seed=0)
    d = unpack_batch(prob.opt.parameters, P)
    assert d["q_nominal"].shape == (5, 7, 1) and d["p_goal"].shape == (5, 3, 1)
    M, B = pack_batch(prob.opt.parameters, d)
    assert B == 5 and np.array_equal(M, P)
    # [B, m*n] flat form, missing labels -> zeros, unknown labels ignored (sx_container.py:113-123)
    M2, B2 = pack_batch(prob.opt.parameters, {"p_goal": P[:, 7:], "nonsense": np.ones(3)})
    assert B2 == 5 and np.array_equal(M2[:, 7:], P[:, 7:]) and not M2[:, :7].any()
    M3, B3 = pack_batch(prob.opt.parameters, {"q_nominal": problems.LWR_Q_NOMINAL})
    assert B3 is None and M3.shape == (1, 10)


@pytest.mark.parametrize("coop", [True, False])
def test_c3_point_mass_mpc_tick(coop):
    """C3 (example/point_mass_mpc.py Controller, T=20): nx=80, 42 linear equalities, 160 bounds, 20
    obstacle inequalities.  Dimensions as in SURVEY.md 8a; solutions checked by the oracle."""
    prob = problems.point_mass_mpc()
    opt = prob.opt
    assert (opt.nx, opt.np, opt.nk, opt.na, opt.ng, opt.nh, opt.nv) == (80, 84, 160, 42, 20, 0, 264)
    assert type(opt).__name__ == "QuadraticCostNonlinearConstraints"
    sim, lo = _sim(prob, coop=coop)
    P, X0 = prob.sample(48, seed=1)
    r = sim.solve(P, X0)
    ok = r["status"] <= 1
    assert ok.mean() >= 0.97
    res = kkt_check.kkt_residual(prob, r["x"][ok][:16], P[ok][:16], r["lam"][ok][:16, :lo.n_eq], r["lam"][ok][:16, lo.n_eq:])
    assert res.max() < 1e-6
    # polish: the oracle seeded at the result stays there
    op = slsqp_driver.OracleProblem(opt)
    i = int(np.where(ok)[0][0])
    pol = slsqp_driver.solve_slsqp(op, P[i], r["x"][i], form="split", options={"ftol": 1e-15, "maxiter": 100})
    assert np.abs(pol.x - r["x"][i]).max() / max(1.0, np.abs(pol.x).max()) < 1e-6


@pytest.mark.parametrize("coop", [True, False])
def test_c5_dual_arm_large_tier(coop):
    """C5 (example/dual_arm.py): 1386 variables, 700 linear equalities, T = 50 -- the cooperative tier (one
    instance per CTA; default) and the table-driven thread-per-instance tier (interpreted tapes + sparse LDL').
    Checked against the independent numpy closed form."""
    import problems_ref

    prob = problems.dual_arm()
    opt = prob.opt
    assert (opt.nx, opt.np, opt.nk, opt.na, opt.ng, opt.nh, opt.nv) == (1386, 14, 0, 700, 0, 0, 1400)
    assert type(opt).__name__ == "NonlinearCostLinearConstraints"
    sim, lo = _sim(prob, coop=coop)
    P, X0 = prob.sample(3, seed=3)
    # the tapes against the closed form at a random (infeasible) point
    import tape_vm

    rng = np.random.default_rng(0)
    x = X0[0] + 0.05 * rng.standard_normal(opt.nx)
    f, ce, _ = tape_vm.CTape(lo.fc)(x[None, :], P[:1])
    assert abs(f[0, 0] - problems_ref.dual_arm_cost(x, P[0])) < 1e-12
    assert np.abs(ce[0] - problems_ref.dual_arm_constraints(x, P[0])).max() < 1e-14
    r = sim.solve(P, X0)
    assert (r["status"] == 0).all()
    for i in range(3):
        k = problems_ref.sparse_kkt_residual(lo, r["x"][i], P[i], r["lam"][i, :lo.n_eq], r["lam"][i, lo.n_eq:])
        assert max(k["stationarity"], k["eq"]) < 1e-7
        assert abs(r["f"][i] - problems_ref.dual_arm_cost(r["x"][i], P[i])) < 1e-12
        assert np.abs(problems_ref.dual_arm_constraints(r["x"][i], P[i])).max() < 1e-9


def test_c4_figure_eight_large_tier():
    """C4 (example/figure_eight_plan.py + joint limits): 693 variables, 357 linear + 200 quaternion
    equalities (rank deficient by construction), 700 bounds."""
    import problems_ref

    prob = problems.figure_eight()
    opt = prob.opt
    assert (opt.nx, opt.np, opt.nk, opt.na, opt.ng, opt.nh) == (693, 7, 700, 357, 0, 200)
    assert type(opt).__name__ == "NonlinearCostNonlinearConstraints"
    sim, lo = _sim(prob)
    P, X0 = prob.sample(2, seed=2)
    r = sim.solve(P, X0, max_iter=400, max_trips=2500)
    # converged to tol = 1e-8, not merely "acceptable": the least-squares multipliers are refined towards the
    # unregularised solution, so the dc regularisation no longer leaves a 2.7e-8 floor in the dual infeasibility
    assert (r["status"] == 0).all(), r["status"]
    for i in range(2):
        k = problems_ref.sparse_kkt_residual(lo, r["x"][i], P[i], r["lam"][i, :lo.n_eq], r["lam"][i, lo.n_eq:])
        assert k["eq"] < 1e-9 and k["ineq"] < 1e-9 and k["stationarity"] < 1e-8 and k["complementarity"] < 1e-8


def test_planar_differential_ik_qp():
    """example/planar_idk.py: a QuadraticCostLinearConstraints problem; inside the bounds the QP solution
    equals the pseudo-inverse solution (as the script itself prints for comparison)."""
    prob = problems.planar_idk()
    assert type(prob.opt).__name__ == "QuadraticCostLinearConstraints"
    sim, lo = _sim(prob)
    q_t, dx = np.array([2.39, -2.55, -0.46]), np.array([0.01, 0.0])
    r = sim.solve(np.concatenate([q_t, dx])[None, :], np.zeros((1, 3)))
    assert r["status"][0] == 0 and r["iters"][0] <= 10
    J = prob.functions["J"](q_t).toarray()[0:2, :]
    assert np.abs(r["x"][0] - np.linalg.pinv(J) @ dx).max() < 1e-7
    P, X0 = prob.sample(128)
    rb = sim.solve(P, X0)
    assert (rb["status"] == 0).all()
    res = kkt_check.kkt_residual(prob, rb["x"][:32], P[:32], rb["lam"][:32, :lo.n_eq], rb["lam"][:32, lo.n_eq:])
    assert res.max() < 1e-7


# ---- the rest of the RobotModel surface through the same solver --------------------------------


@pytest.mark.parametrize("coop", [True, False])
def test_joint_space_planner_to_a_pose(coop):
    """example/simple_joint_space_planner.py:14-71 (med7, T = 20, derivs_align): 280 variables, 147 linear
    equalities, 7 pose equalities on the last knot (position + quaternion), 40 link-height inequalities."""
    prob = problems.joint_space_planner()
    opt = prob.opt
    assert (opt.nx, opt.np, opt.nk, opt.na, opt.ng, opt.nh, opt.nv) == (280, 21, 0, 147, 40, 7, 348)
    assert type(opt).__name__ == "QuadraticCostNonlinearConstraints"
    sim, lo = _sim(prob, coop=coop)
    P, X0 = prob.sample(6)
    r = sim.solve(P, X0)
    assert (r["status"] == 0).all(), r["status"]
    res = kkt_check.kkt_residual(prob, r["x"], P, r["lam"][:, :lo.n_eq], r["lam"][:, lo.n_eq:])
    assert res.max() < 1e-7
    # the last knot reaches the goal pose (the quaternion up to sign is fixed by the equality itself), the first is
    # the current configuration, the last velocity is zero, and both links stay above z = -0.05
    sol = prob.seed_dict(r["x"])
    Q, dQ = sol["med7/q/x"], sol["med7/dq/x"]
    import fk_ref

    chain_ee, chain_3 = fk_ref.Chain(problems.MED7_URDF, problems.MED7_EE), fk_ref.Chain(problems.MED7_URDF, "lbr_link_3")
    qF = Q[:, :, -1]
    assert np.abs(chain_ee.fk(qF)[1] - P[:, 14:17]).max() < 1e-8          # independent numpy kinematics (oracle/fk_ref.py)
    assert np.abs(chain_ee.quaternion(qF) - P[:, 17:21]).max() < 1e-8
    for t in range(Q.shape[2]):
        assert chain_ee.fk(Q[:, :, t])[1][:, 2].min() > -0.05 - 1e-8 and chain_3.fk(Q[:, :, t])[1][:, 2].min() > -0.05 - 1e-8
    assert np.abs(Q[:, :, 0] - P[:, 7:14]).max() < 1e-9 and np.abs(dQ[:, :, -1]).max() < 1e-9
    # (no SLSQP polish here: position + unit quaternion are 7 equalities of rank 6, on which scipy's SLSQP stops with
    #  "inequality constraints incompatible"; the oracle KKT residual and the independent kinematics stand in)


def test_sphere_collision_first_stage_ik():
    """example/sphere_collision_avoidance.py:20-42: IK to a position with the tool z axis held (three equalities of
    rank two -- a unit vector), joint limits on q, zero velocity / acceleration knots; its solution at the script's
    start position is the constant the second stage starts from."""
    prob = problems.lwr_axis_ik()
    opt = prob.opt
    assert (opt.nx, opt.np, opt.nk, opt.na, opt.ng, opt.nh) == (21, 3, 14, 14, 0, 6)
    sim, lo = _sim(prob)
    P, X0 = prob.sample(16)
    P[0] = [0.825, -0.35, 0.2]
    r = sim.solve(P, X0)
    assert (r["status"] == 0).all()
    assert np.abs(r["x"][0, :7] - problems.SPHERE_Q_START).max() < 1e-7
    res = kkt_check.kkt_residual(prob, r["x"], P, r["lam"][:, :lo.n_eq], r["lam"][:, lo.n_eq:])
    assert res.max() < 1e-7
    robot = prob.models["robot"]
    Tq = robot.get_global_link_transform(problems.LWR_EE, r["x"][3, :7]).toarray()
    T0 = robot.get_global_link_transform(problems.LWR_EE, problems.SPHERE_Q_NOMINAL).toarray()
    assert np.abs(Tq[:3, 3] - P[3]).max() < 1e-8 and np.abs(Tq[:3, 2] - T0[:3, 2]).max() < 1e-8


def test_sphere_collision_problem_dimensions_and_seed():
    """example/sphere_collision_avoidance.py:54-98 builds through `sphere_collision_avoidance_constraints`
    (builder.py:366-417): 20 knots x 4 link spheres x 6 obstacles = 480 inequalities, two reduced (sum-of-squares)
    equalities; the seed (hold the start configuration) is collision free."""
    prob = problems.sphere_collision_avoidance()
    opt = prob.opt
    assert (opt.nx, opt.np, opt.nk, opt.na, opt.ng, opt.nh, opt.nv) == (420, 38, 280, 287, 480, 2, 1338)
    P, X0 = prob.sample(4)
    op = slsqp_driver.OracleProblem(opt)
    g = np.array([np.asarray(opt.g(X0[b], P[b])).flatten() for b in range(4)])
    assert g.shape == (4, 480) and g.min() > 0.0
    assert op is not None


def test_tier_choice_follows_the_measured_rule():
    """bo_problem_create picks the kernel tier from the problem sizes alone (measured on B200 in
    profiles/r01_tier_choice.txt / r01_tier_break_even.txt, r02_team_vs_thread.txt): up to 14 KKT rows the team tier (state in
    shared memory) when variables + constraints >= 20, else the thread-per-instance dense tier; thread-per-instance sparse for
    mid-size problems; one instance per CTA once nx + n_eq + n_ineq > 64 and the tapes split into stages."""
    expect = [(problems.lwr_ik(), "team"), (problems.planar_idk(), "qp"), (problems.booth(), "qp"), (problems.lwr_axis_ik(), "sparse"),
              (problems.point_mass_mpc(T=6), "coop"), (problems.point_mass_mpc(), "coop"), (problems.joint_space_planner(), "coop")]
    for prob, tier in expect:
        s = optas_b200.B200Solver(prob.opt).setup("ipopt", compile_only=True)
        assert s.tier_info()["tier"] == tier, (prob.name, s.tier_info()["tier"])
    forced = optas_b200.B200Solver(problems.lwr_axis_ik().opt).setup("ipopt", compile_only=True, coop=True)
    assert forced.tier_info()["tier"] == "coop"   # BO_FLAG_COOP still overrides
    old = optas_b200.B200Solver(problems.lwr_ik().opt).setup("ipopt", compile_only=True, team=False)
    assert old.tier_info()["tier"] == "dense"     # BO_FLAG_NO_TEAM: the round-1 thread-per-instance kernel


def test_team_tier_takes_the_same_iterations_as_the_thread_tier():
    """The team tier (bo_ipm_team.cuh) is the same algorithm as the thread-per-instance dense tier: on the host build
    of both generated sources every instance ends with the same status after the same number of iterations and trips."""
    from optas_b200 import _capi
    from optas_b200.lowering import lower_problem

    for prob, B in ((problems.lwr_ik(), 1024), (problems.planar_idk(), 256), (problems.booth(), 16)):
        lo = lower_problem(prob.opt)
        P, X0 = prob.sample(B, seed=3)
        res = {}
        for label, flag in (("team", _capi.BO_FLAG_TEAM), ("thread", _capi.BO_FLAG_NO_TEAM | _capi.BO_FLAG_NO_QP)):
            h = _capi.ProblemHandle(lo, flags=_capi.BO_FLAG_COMPILE_ONLY | flag)
            assert h.tier_info()["tier"] == ("team" if label == "team" else "dense")
            res[label] = HostSim(h.source(), lo.nx, lo.np_, lo.n_eq, lo.n_ineq).solve(P, X0, max_step=h.options()["max_step"])
        a, b = res["team"], res["thread"]
        assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["iters"], b["iters"])
        assert np.array_equal(a["trips"], b["trips"])
        ok = a["status"] == 0
        assert ok.mean() > 0.99 and np.abs(a["x"][ok] - b["x"][ok]).max() < 1e-9


# ---- the QP path (csrc/jit/bo_qp_reg.cuh) ------------------------------------------------------------
def _sim_of(prob, **kw):
    s = optas_b200.B200Solver(prob.opt).setup("ipopt", compile_only=True, **kw)
    lo = s._lowered
    return s, lo, HostSim(s.kernel_source(), lo.nx, lo.np_, lo.n_eq, lo.n_ineq, ldl_table=s.ldl_table())


def test_qp_path_is_chosen_from_the_tape_not_from_the_class_name():
    """bo_problem_create runs the QP iteration exactly when no Jacobian / Hessian output of the kkt tape depends on x, y, z
    (reference: OSQPSolver / CVXOPTSolver take P, q, M, c, A, b as functions of p only, optas/solver.py:455-467, 540-551)."""
    for prob, want in ((problems.planar_idk(), "qp"), (problems.lwr_diff_ik_qp(), "qp"), (problems.box_qp(), "qp"),
                       (problems.booth(), "qp"), (problems.lwr_ik(), "team"), (problems.lwr_axis_ik(), "sparse")):
        s = optas_b200.B200Solver(prob.opt).setup("ipopt", compile_only=True)
        assert s.tier_info()["tier"] == want, (prob.name, s.tier_info())
    off = optas_b200.B200Solver(problems.planar_idk().opt).setup("ipopt", compile_only=True, qp=False)
    assert off.tier_info()["tier"] == "dense"  # BO_FLAG_NO_QP: the general interior-point kernel
    mid = optas_b200.B200Solver(problems.box_qp(24, 4).opt).setup("ipopt", compile_only=True, coop=False)
    assert mid.tier_info()["tier"] == "qp_sparse"


@pytest.mark.parametrize("make", [problems.planar_idk, problems.lwr_diff_ik_qp, problems.box_qp, lambda: problems.box_qp(12, 0, seed=3)])
def test_qp_path_matches_the_oracle_and_the_interior_point_kernel(make):
    """Same minimiser as the general kernel and as scipy SLSQP on the numpy closed form of the QP, the oracle's KKT residual
    at tol, fewer iterations, and f / kkt reported at the returned x (re-evaluated by the tape, not carried along)."""
    from scipy.optimize import minimize

    prob = make()
    P, X0 = prob.sample(192)
    s, lo, sim = _sim_of(prob)
    assert s.tier_info()["tier"] == "qp"
    r = sim.solve(P, X0, max_step=s._handle.options()["max_step"])
    assert (r["status"] == 0).all() and r["kkt"].max() <= 1e-8
    # IPOPT's scaled error (s_d, s_c: instances that start outside their limits carry large multipliers)
    res = kkt_check.kkt_residual(prob, r["x"][:48], P[:48], r["lam"][:48, :lo.n_eq], r["lam"][:48, lo.n_eq:], scaled=True)
    assert res.max() < 2e-8
    _, _, gen = _sim_of(prob, qp=False)
    g = gen.solve(P, X0, max_step=s._handle.options()["max_step"])
    both = g["status"] == 0
    assert both.mean() > 0.9 and r["iters"][both].mean() < g["iters"][both].mean()
    assert np.abs(r["f"] - g["f"])[both].max() < 1e-6 * max(1.0, np.abs(g["f"][both]).max())
    # f is evaluated AT x: re-evaluate the objective with the oracle's tape interpreter
    import tape_vm
    f_at_x = tape_vm.CTape(lo.fc)(r["x"][:8], P[:8])[0].ravel()
    assert np.abs(f_at_x - r["f"][:8]).max() <= 1e-12 * max(1.0, np.abs(f_at_x).max())
    if "P" in prob.models:  # the box family has a closed form: independent active-set solve (scipy SLSQP)
        Pm, A = prob.models["P"], prob.models["A"]
        n, me = Pm.shape[0], A.shape[0]
        for i in range(6):
            q, b = P[i, :n], P[i, n:]
            cons = [{"type": "eq", "fun": lambda v: A @ v - b, "jac": lambda v: A}] if me else []
            ref = minimize(lambda v: v @ Pm @ v + q @ v, np.zeros(n), jac=lambda v: 2 * Pm @ v + q, method="SLSQP",
                           bounds=[(-1, 1)] * n, constraints=cons, options={"ftol": 1e-14, "maxiter": 500})
            assert ref.status in (0, 8) and np.abs(ref.x - r["x"][i]).max() < 1e-6  # 8: no further descent at the minimiser


def test_qp_path_edge_cases():
    """Equality-only and unconstrained QPs take ONE Newton step; a semidefinite Hessian and linearly dependent equality rows
    are regularised; an infeasible QP ends with a failure status instead of spinning."""
    import optas_b200.sym as cs
    from optas_b200.builder import OptimizationBuilder
    from optas_b200.models import TaskModel

    rng = np.random.default_rng(1)
    n = 6

    def build(Pm, A, box):
        b = OptimizationBuilder(T=1, tasks=[TaskModel("v", n, time_derivs=[0])])
        x = b.get_model_state("v", 0)
        q = b.add_parameter("q", n)
        b.add_cost_term("quad", x.T @ cs.DM(Pm) @ x + q.T @ x)
        if A is not None:
            b.add_equality_constraint("lin", cs.DM(A) @ x, b.add_parameter("rhs", A.shape[0]))
        if box:
            b.add_bound_inequality_constraint("box", [-1.0] * n, x, [1.0] * n)
        return b.build()

    def run(opt, P, X0, **kw):
        s = optas_b200.B200Solver(opt).setup("ipopt", compile_only=True, **kw)
        lo = s._lowered
        sim = HostSim(s.kernel_source(), lo.nx, lo.np_, lo.n_eq, lo.n_ineq)
        return s.tier_info()["tier"], sim.solve(P, X0, max_step=s._handle.options()["max_step"])

    Ppd = np.eye(n) + 0.1 * np.ones((n, n))
    # unconstrained: x = -(2P)^-1 q in one step, whatever the seed
    Q = rng.standard_normal((16, n))
    tier, r = run(build(Ppd, None, False), Q, rng.standard_normal((16, n)))
    assert tier == "qp" and (r["status"] == 0).all() and (r["iters"] == 1).all()
    assert np.abs(r["x"] - np.linalg.solve(2 * Ppd, -Q.T).T).max() < 1e-12
    # equality-only
    A = rng.standard_normal((2, n))
    tier, r = run(build(Ppd, A, False), rng.standard_normal((16, n + 2)), rng.standard_normal((16, n)))
    assert (r["status"] == 0).all() and (r["iters"] == 1).all()
    # rank-2 Hessian inside a box
    L = rng.standard_normal((n, 2))
    opt = build(L @ L.T, None, True)
    Q = rng.standard_normal((32, n))
    _, r = run(opt, Q, np.zeros((32, n)))
    _, g = run(opt, Q, np.zeros((32, n)), qp=False)
    assert (r["status"] == 0).all() and np.abs(r["f"] - g["f"]).max() < 1e-6
    # dependent equality rows (consistent right-hand sides)
    a = rng.standard_normal((1, n))
    rhs = rng.standard_normal((16, 1))
    _, r = run(build(Ppd, np.concatenate([a, 2 * a]), True), np.concatenate([rng.standard_normal((16, n)), rhs, 2 * rhs], axis=1), np.zeros((16, n)))
    assert (r["status"] <= 1).all()
    # infeasible: x0 + x1 = 5 inside [-1, 1]^n
    A = np.zeros((1, n))
    A[0, :2] = 1.0
    _, r = run(build(Ppd, A, True), np.concatenate([rng.standard_normal((4, n)), 5 * np.ones((4, 1))], axis=1), np.zeros((4, n)))
    assert (r["status"] >= 2).all() and (r["trips"] <= 250).all()


# ---- option defaults (round-1 advisor findings) ----------------------------------------------------------------------
def _far_qp(g=100.0):
    """min ||x - g||^2 s.t. x0 + x1 = 2 g, x >= -1: a convex QP whose optimum lies ~g away from the zero seed."""
    import optas_b200.sym as cs
    from optas_b200.lowering import lower_nlp

    x = cs.SX.sym("x", 2)
    p = cs.SX.sym("p", 1)
    f = cs.sumsqr(x - p[0])
    return lower_nlp(x, p, f, cs.vertcat(x[0] + x[1] - 2.0 * p[0]), x + 1.0)


def test_step_cap_is_off_for_problems_without_trigonometry():
    """The default step cap (0.5, absolute) made any optimum further than max_iter/2 from the seed unreachable.  It now
    applies only when the tapes contain sin/cos/tan; an explicit max_step is always honoured."""
    from optas_b200 import _capi

    lo = _far_qp()
    h = _capi.ProblemHandle(lo, flags=_capi.BO_FLAG_COMPILE_ONLY)
    o = h.options()
    assert o["max_step"] < 0 and o["max_iter"] == 100 and o["max_trips"] == 250
    sim = HostSim(h.source(), lo.nx, lo.np_, lo.n_eq, lo.n_ineq)
    r = sim.solve(np.array([[100.0]]), np.zeros((1, 2)), max_step=o["max_step"])
    assert r["status"][0] == 0 and r["iters"][0] <= 12
    assert np.allclose(r["x"][0], [100.0, 100.0], atol=1e-6)
    # the old default reproduces the advisor's failure: 100 iterations x 0.5 = 50
    r_old = sim.solve(np.array([[100.0]]), np.zeros((1, 2)), max_step=0.5)
    assert r_old["status"][0] == 2 and np.abs(r_old["x"][0]).max() <= 50.0 + 1e-9
    assert _capi.ProblemHandle(lo, flags=_capi.BO_FLAG_COMPILE_ONLY, max_step=0.25).options()["max_step"] == 0.25
    # kinematics keep the cap
    ik = optas_b200.B200Solver(problems.lwr_ik().opt).setup("ipopt", compile_only=True)
    assert ik._handle.options()["max_step"] == 0.5


def test_trip_budget_follows_max_iter():
    """`ipopt.max_iter` / scipy `maxiter` above the default trip budget used to be silently ineffective."""
    prob = problems.booth()
    s = optas_b200.B200Solver(prob.opt).setup("ipopt", {"ipopt.max_iter": 1000}, compile_only=True)
    o = s._handle.options()
    assert o["max_iter"] == 1000 and o["max_trips"] == 2500
    s = optas_b200.ScipyMinimizeSolver(prob.opt).setup(method="SLSQP", options={"maxiter": 40}, compile_only=True)
    assert s._handle.options()["max_trips"] == 250
    s = optas_b200.B200Solver(prob.opt).setup("ipopt", {"ipopt.max_iter": 1000, "max_trips": 77}, compile_only=True)
    assert s._handle.options()["max_trips"] == 77
    # whole warps only (the kernels use full-mask warp votes): 48 -> 64 on the thread-per-instance tier, 32 G on the team tier
    assert optas_b200.B200Solver(prob.opt).setup("ipopt", compile_only=True, threads_per_block=48, team=False)._handle.tier_info()["threads_per_block"] == 64
    assert optas_b200.B200Solver(prob.opt).setup("ipopt", compile_only=True, threads_per_block=80, team=True)._handle.tier_info()["threads_per_block"] == 64


def test_raw_buffers_are_validated_before_the_library_sees_them():
    prob = problems.lwr_ik()
    s = optas_b200.B200Solver(prob.opt).setup("ipopt", compile_only=True)
    P, X0 = prob.sample(4, seed=0)
    X = np.empty_like(X0)
    with pytest.raises(ValueError, match="float64"):
        s.solve_raw(P.astype(np.float32), X0, X)
    with pytest.raises(ValueError, match="contiguous"):
        s.solve_raw(np.asfortranarray(P), X0, X)
    with pytest.raises(ValueError, match="rows"):
        s.solve_raw(P[:3], X0, X)
    with pytest.raises(ValueError, match="int32"):
        s.solve_raw(P, X0, X, status=np.zeros(4, dtype=np.int64))


def test_result_buffers_are_recycled_only_when_nobody_holds_them():
    """solve_arrays() recycles its page-locked result arrays (allocating them costs more than the PCIe copies), but never a
    set the caller can still see: a live array or a live VIEW of one (what solve() hands out) keeps its set out of the pool."""
    prob = problems.booth()
    s = optas_b200.B200Solver(prob.opt).setup("ipopt", compile_only=True)
    a = s._result_buffers(100)
    ida = id(a["x"])
    b = s._result_buffers(100)
    assert id(b["x"]) != ida                      # first set still held
    view = a["x"][:, 0]
    del a
    c = s._result_buffers(100)
    assert id(c["x"]) != ida                      # a view of the first set is alive
    del view, b, c
    d = s._result_buffers(100)
    assert id(d["x"]) in [id(st["arrays"]["x"]) for st in s._pool] and len(s._pool) == 3
    e = s._result_buffers(7)
    assert e["x"].shape == (7, 2)                 # another batch size gets its own set


def test_seed_infeasibility_predictor_is_the_constraint_violation_of_the_seed():
    """The function behind ``setup(schedule="seed_infeasibility")`` (which instance goes to the GPU first), evaluated by the
    oracle's tape interpreter: theta(x0, p) = sum max(0, -v(x0, p))^2 with v the reference's stacked constraint vector
    (optas/optimization.py:27-51); zero at a feasible seed, growing with the distance of the goal."""
    import tape_vm
    from optas_b200 import sym as cs
    from optas_b200.tape import Tape

    prob = problems.lwr_ik()
    opt = prob.opt
    theta = cs.sumsqr(cs.fmax(0.0, -opt.v(opt.x, opt.p)))
    tape = Tape.from_function(cs.Function("seed_infeasibility", [opt.x, opt.p], [theta]))
    assert list(tape.in_sizes) == [opt.nx, opt.np] and list(tape.out_sizes) == [1]
    P, X0 = prob.sample(16, seed=3)
    got = tape_vm.CTape(tape)(X0, P)[0].ravel()
    for i in range(16):
        v = np.asarray(opt.v(X0[i], P[i])).ravel()
        assert abs(got[i] - (np.maximum(0.0, -v) ** 2).sum()) < 1e-12
    assert (got > 0).all()  # the sampled goals are not at the seed
    # scheduling order = decreasing theta: a permutation of the batch
    order = np.argsort(-got)
    assert sorted(order.tolist()) == list(range(16))
