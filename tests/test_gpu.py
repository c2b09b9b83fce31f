"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI
(libb200optas.so via optas_b200._capi); the oracle (oracle/) is only the checker.

Tolerances: all arithmetic is IEEE binary64.  FK / Jacobian values must agree with the numpy
restatement to 1e-12 absolute (a few ulp of O(1) quantities: FMA contraction and the device sincos
differ from libm in the last bits).  Converged decision variables must agree with the oracle's
polished solution to 1e-6 relative (north_star: "1e-6 rel-tol on decision variables and KKT residual"); the oracle KKT
residual of an instance reported CONVERGED (status 0) must be <= 1e-8 in IPOPT's scaling, of one
reported ACCEPTABLE (status 1) <= 1e-6.  The per-config parity protocol lives in tests/test_gpu_parity.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FK_TOL = 1e-12
X_RTOL = 1e-6
KKT_TOL = 1e-8            # status 0
KKT_TOL_ACCEPTABLE = 1e-6  # status 1


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch


@pytest.fixture(scope="module")
def ik(torch_cuda):
    import optas_b200
    from optas_b200 import problems

    prob = problems.lwr_ik()
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt")
    return prob, solver


def _solve_host(solver, P, X0):
    B = X0.shape[0]
    lo = solver._lowered
    out = dict(x=np.empty((B, lo.nx)), lam=np.empty((B, lo.n_eq + lo.n_ineq)), f=np.empty(B),
               status=np.empty(B, dtype=np.int32), iters=np.empty(B, dtype=np.int32), kkt=np.empty(B))
    solver.solve_raw(np.ascontiguousarray(P), np.ascontiguousarray(X0), out["x"], out["lam"], out["f"], out["status"],
                     out["iters"], out["kkt"])
    return out


def _assert_kkt(prob, lo, r, P, idx):
    """Oracle KKT residual (IPOPT scaling): <= 1e-8 for status 0, <= 1e-6 for status 1 (counted separately)."""
    import kkt_check

    idx = np.asarray(idx)
    res = kkt_check.kkt_residual(prob, r["x"][idx], P[idx], r["lam"][idx][:, :lo.n_eq], r["lam"][idx][:, lo.n_eq:], scaled=True)
    conv = r["status"][idx] == 0
    assert conv.any()
    assert res[conv].max() <= KKT_TOL, res[conv].max()
    if (~conv).any():
        assert res[~conv].max() <= KKT_TOL_ACCEPTABLE, res[~conv].max()


def test_fk_jacobian_matches_oracle_full_batch(ik):
    import fk_ref
    from optas_b200.function import B200Function

    prob, _ = ik
    fk = B200Function(prob.functions["fk_jac"])
    rng = np.random.default_rng(0)
    for B in (65536, 1000, 129, 1):  # full tiles, ragged tail, single instance
        q = rng.uniform(-2.9, 2.9, (B, 7))
        p, J = fk(q)
        p_ref, J_ref = fk_ref.lwr_position_and_jacobian(q)
        assert np.abs(p - p_ref).max() < FK_TOL
        assert np.abs(J - J_ref).max() < FK_TOL


def test_fk_device_pointers_and_unaligned_views(ik, torch_cuda):
    import fk_ref
    from optas_b200.function import B200Function

    torch = torch_cuda
    prob, _ = ik
    fk = B200Function(prob.functions["fk_jac"])
    B = 4096 + 37
    q = torch.rand((B + 1, 7), dtype=torch.float64, device="cuda") * 4 - 2
    p = torch.empty((B, 3), dtype=torch.float64, device="cuda")
    J = torch.empty((B, 21), dtype=torch.float64, device="cuda")
    qv = q[1:]  # starts 56 bytes into the allocation: only 8-byte aligned -> non-bulk path
    fk.eval_raw(B, [qv], [p, J], stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    p_ref, J_ref = fk_ref.lwr_position_and_jacobian(qv.cpu().numpy())
    assert np.abs(p.cpu().numpy() - p_ref).max() < FK_TOL
    assert np.abs(J.cpu().numpy() - J_ref).max() < FK_TOL


def test_booth_known_answer(torch_cuda):
    """Reference tests/test_solver.py:45-54, through the drop-in classes and both setup spellings."""
    import optas_b200
    from optas_b200 import problems

    prob = problems.booth()
    for cls, setup in ((optas_b200.CasADiSolver, lambda s: s.setup("ipopt")),
                       (optas_b200.CasADiSolver, lambda s: s.setup("qpoases")),
                       (optas_b200.ScipyMinimizeSolver, lambda s: s.setup(method="SLSQP", tol=1e-6))):
        solver = setup(cls(prob.opt))
        solver.reset_parameters({"a": 2.0, "b": 7.0})
        solution = solver.solve()
        assert solver.did_solve()
        xy = solution["booth/y"].toarray().flatten()
        assert np.isclose(xy[0], 1.0) and np.isclose(xy[1], 3.0)
        assert solver.number_of_iterations() >= 1 and solver.stats()["success"]


def test_c1_example_script_flow(torch_cuda):
    """example/example.py:40-60 with the reference's seed-key quirk and with the
    correct key; the latter must reproduce the golden q*."""
    import optas_b200 as optas
    from optas_b200 import problems

    prob = problems.lwr_ik()
    robot = prob.models["robot"]
    name = robot.get_name()
    solver = optas.CasADiSolver(prob.opt).setup("ipopt")
    q_nominal = optas.deg2rad([0, 45, 0, -90, 0, -45, 0])
    p_goal = robot.get_global_link_position(problems.LWR_EE, q_nominal) + optas.DM([0.0, 0.3, -0.2])
    solver.reset_parameters({"q_nominal": q_nominal, "p_goal": p_goal})
    solver.reset_initial_seed({f"{name}/q/x": q_nominal})
    solution = solver.solve()
    assert solver.did_solve()
    q = solution[f"{name}/q"].toarray().flatten()
    golden = np.array([-0.24904555, 1.14296583, -0.12385623, -1.29959808, 0.03860145, -0.66073211, 0.0])
    assert np.abs(q - golden).max() < 1e-7
    assert isinstance(solution[f"{name}/q"], optas.DM) and solution[f"{name}/q"].shape == (7, 1)
    # the script's own (wrong) key is ignored -> seed zeros, exactly as in the reference
    solver.reset_initial_seed({f"{name}/q": q_nominal})
    assert not np.asarray(solver._X0).any()


def test_c2_batch_parity_protocol(ik):
    import kkt_check
    import slsqp_driver

    prob, solver = ik
    B = 4096
    P, X0 = prob.sample(B, seed=11)
    r = _solve_host(solver, P, X0)
    ok = r["status"] <= 1
    assert ok.mean() >= 0.995, ok.mean()
    lo = solver._lowered
    idx = np.where(ok)[0][:512]
    _assert_kkt(prob, lo, r, P, idx)
    op = slsqp_driver.OracleProblem(prob.opt)
    worst = 0.0
    for i in idx[:96]:
        pol = slsqp_driver.solve_slsqp(op, P[i], r["x"][i], form="split", options={"ftol": 1e-15, "maxiter": 200})
        worst = max(worst, np.abs(pol.x - r["x"][i]).max() / max(1.0, np.abs(pol.x).max()))
    assert worst < X_RTOL, worst


def test_c2_full_size_properties(ik, torch_cuda):
    """65536 instances: size-independent properties -- FK(q*) == p_goal (round trip through the
    streaming kernel), joint limits, finite outputs, f == ||q* - q_nominal||^2."""
    from optas_b200.function import B200Function

    prob, solver = ik
    B = 65536
    P, X0 = prob.sample(B, seed=0)
    r = _solve_host(solver, P, X0)
    ok = r["status"] <= 1
    assert ok.mean() >= 0.995
    assert np.isfinite(r["x"]).all()
    fk = B200Function(prob.functions["fk_jac"])
    p, _ = fk(r["x"])
    assert np.abs(p[ok] - P[ok, 7:]).max() < 1e-7
    robot = prob.models["robot"]
    lo = robot.lower_actuated_joint_limits.toarray().flatten()
    up = robot.upper_actuated_joint_limits.toarray().flatten()
    assert (r["x"][ok] >= lo - 1e-9).all() and (r["x"][ok] <= up + 1e-9).all()
    assert np.abs(r["f"][ok] - ((r["x"][ok] - P[ok, :7]) ** 2).sum(1)).max() < 1e-9
    assert (r["iters"][ok] <= 200).all() and (r["kkt"][ok] <= 1e-6).all()


def test_shard_equivalence_and_device_pointers(ik, torch_cuda):
    """Per-instance results do not depend on how the batch is sharded or where the buffers live
    (no cross-instance arithmetic): bitwise equality."""
    torch = torch_cuda
    prob, solver = ik
    B = 3000  # not a multiple of the block size
    P, X0 = prob.sample(B, seed=21)
    whole = _solve_host(solver, P, X0)
    parts = [_solve_host(solver, P[a:b], X0[a:b]) for a, b in ((0, 1), (1, 1501), (1501, 3000))]
    for key in ("x", "lam", "f", "status", "iters", "kkt"):
        glued = np.concatenate([p[key] for p in parts])
        assert np.array_equal(glued, whole[key]), key
    Pd, X0d = torch.from_numpy(P).cuda(), torch.from_numpy(X0).cuda()
    Xd = torch.empty_like(X0d)
    st = torch.empty(B, dtype=torch.int32, device="cuda")
    solver.solve_raw(Pd, X0d, Xd, None, None, st, None, None, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(Xd.cpu().numpy(), whole["x"]) and np.array_equal(st.cpu().numpy(), whole["status"])


def test_batched_dict_api(ik):
    prob, solver = ik
    B = 257
    P, X0 = prob.sample(B, seed=4)
    solver.reset_parameters(prob.param_dict(P))
    solver.reset_initial_seed(prob.seed_dict(X0))
    sol = solver.solve()
    name = prob.models["robot"].get_name()
    assert sol[f"{name}/q"].shape == (B, 7, 1) and sol[f"{name}/q/x"].shape == (B, 7, 1)
    st = solver.stats()
    assert st["status"].shape == (B,) and st["n_converged"] >= 0.99 * B
    ref = _solve_host(solver, P, X0)
    assert np.array_equal(sol[f"{name}/q"][:, :, 0], ref["x"])
    # zero seed default (solver.py:76) and empty batch
    solver.reset_initial_seed({})
    solver.reset_parameters({"q_nominal": P[0, :7], "p_goal": P[0, 7:]})
    one = solver.solve()
    assert one[f"{name}/q"].shape == (7, 1)


def test_nlpsol_factory_matches_solver(ik):
    """The CasADi factory signature (optas/solver.py:346-398): problem = {x, p, f, g = v(x, p)},
    lbg = 0, ubg = 1e10 -- the +- pairs are merged back into equalities and the result equals the
    Solver-API result."""
    import optas_b200 as optas
    from optas_b200 import problems

    prob, solver = ik
    opt = prob.opt
    x, p = opt.decision_variables.vec(), opt.parameters.vec()
    nlp = optas.nlpsol("solver", "ipopt", {"x": x, "p": p, "f": opt.f(x, p), "g": opt.v(x, p)}, {})
    pv, x0 = problems.lwr_ik_example_instance()
    sol = nlp(x0=x0, p=pv, lbg=opt.lbv, ubg=opt.ubv)
    assert nlp.stats()["success"] and nlp.stats()["iter_count"] >= 1
    ref = _solve_host(solver, pv[None, :], x0[None, :])
    assert np.array_equal(sol["x"].toarray().flatten(), ref["x"][0])
    assert sol["lam_g"].shape == (20, 1) and sol["g"].shape == (20, 1)
    # stationarity in CasADi's sign convention: grad f + J_g' lam_g = 0
    import slsqp_driver

    op = slsqp_driver.OracleProblem(opt)
    xs = sol["x"].toarray().flatten()
    resid = op.df(xs, pv) + op.dv(xs, pv).T @ sol["lam_g"].toarray().flatten()
    assert np.abs(resid).max() < 1e-6


@pytest.mark.parametrize("coop", [True, False])
def test_c3_point_mass_mpc_batch(torch_cuda, coop):
    """C3: MPC tick, 80 variables / 42 equalities / 180 inequalities, checked by the oracle -- on the cooperative
    tier (one instance per CTA, the default for horizon problems) and the thread-per-instance sparse tier."""
    import kkt_check
    import optas_b200
    from optas_b200 import problems

    prob = problems.point_mass_mpc()
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt", coop=coop)
    assert solver.tier_info()["tier"] == ("coop" if coop else "sparse")
    B = 512
    P, X0 = prob.sample(B, seed=1)
    r = _solve_host(solver, P, X0)
    ok = r["status"] <= 1
    assert ok.mean() >= 0.97, ok.mean()
    lo = solver._lowered
    idx = np.where(ok)[0][:24]
    _assert_kkt(prob, lo, r, P, idx)
    # dynamics x_{t+1} = x_t + dt v_t and the initial state hold to round-off (linear equalities)
    sol = prob.seed_dict(r["x"][ok])
    Y, dY = sol["point_mass/y/x"], sol["point_mass/dy/x"]
    assert np.abs(Y[:, :, 1:] - Y[:, :, :-1] - 0.05 * dY[:, :, :-1]).max() < 1e-9
    assert np.abs(Y[:, :, 0] - P[ok, 0:2]).max() < 1e-9
    assert (np.abs(Y) <= 1.5 + 1e-9).all() and (np.abs(dY) <= 1.0 + 1e-9).all()


@pytest.mark.parametrize("coop", [True, False])
def test_c5_dual_arm_batch(torch_cuda, coop):
    """C5 through the cooperative tier (default) and the table-driven thread-per-instance tier, checked against the
    numpy closed form of the problem."""
    import optas_b200
    import problems_ref
    from optas_b200 import problems

    prob = problems.dual_arm()
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt", coop=coop)
    assert solver.tier_info()["tier"] == ("coop" if coop else "large")
    B = 96
    P, X0 = prob.sample(B, seed=3)
    r = _solve_host(solver, P, X0)
    assert (r["status"] <= 1).all(), np.bincount(r["status"])
    lo = solver._lowered
    for i in (0, 17, 95):
        k = problems_ref.sparse_kkt_residual(lo, r["x"][i], P[i], r["lam"][i, :lo.n_eq], r["lam"][i, lo.n_eq:], scaled=True)
        assert max(k["stationarity"], k["eq"], k["complementarity"]) <= (KKT_TOL if r["status"][i] == 0 else KKT_TOL_ACCEPTABLE)
        assert abs(r["f"][i] - problems_ref.dual_arm_cost(r["x"][i], P[i])) < 1e-10
        assert np.abs(problems_ref.dual_arm_constraints(r["x"][i], P[i])).max() < 1e-8
    # dict API: trajectories come back as [B, 7, T] and the first knot is the commanded configuration
    solver.reset_parameters(prob.param_dict(P))
    solver.reset_initial_seed(prob.seed_dict(X0))
    sol = solver.solve()
    assert sol["kukal/q"].shape == (B, 7, 50) and sol["kukar/dq"].shape == (B, 7, 49)
    assert np.abs(sol["kukal/q"][:, :, 0] - P[:, :7]).max() < 1e-8


@pytest.mark.parametrize("T", [10, 50])
def test_c4_figure_eight(torch_cuda, T):
    """C4 (nonlinear cost, quaternion equalities, joint-limit bounds): a short horizon and the full T = 50 instance
    of BASELINE.json (1250-row KKT system, factor in shared memory, one instance per CTA)."""
    import optas_b200
    import problems_ref
    from optas_b200 import problems

    prob = problems.figure_eight(T=T)
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt", {"max_iter": 400, "max_trips": 2500})
    assert solver.tier_info()["tier"] == "coop"
    B = 64
    P, X0 = prob.sample(B, seed=2)
    r = _solve_host(solver, P, X0)
    ok = r["status"] <= 1
    assert ok.mean() >= 0.9, np.bincount(r["status"])
    lo = solver._lowered
    i = int(np.where(ok)[0][0])
    k = problems_ref.sparse_kkt_residual(lo, r["x"][i], P[i], r["lam"][i, :lo.n_eq], r["lam"][i, lo.n_eq:])
    assert k["eq"] < 1e-6 and k["ineq"] < 1e-9


def test_coop_tier_is_bitwise_reproducible_across_batch_splits(torch_cuda):
    """No cross-instance arithmetic and fixed reduction trees inside a CTA: any split of the batch (= any number of
    GPUs) returns the same bits, also with device-resident buffers."""
    import optas_b200
    from optas_b200 import problems

    torch = torch_cuda
    prob = problems.point_mass_mpc()
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt")
    assert solver.tier_info()["tier"] == "coop"
    B = 1000
    P, X0 = prob.sample(B, seed=7)
    full = _solve_host(solver, P, X0)
    parts = [_solve_host(solver, P[a:b], X0[a:b]) for a, b in ((0, 333), (333, 1000))]
    for key in ("x", "lam", "f", "status", "iters"):
        assert np.array_equal(full[key], np.concatenate([p[key] for p in parts])), key
    Pd, Xd = torch.from_numpy(P).cuda(), torch.from_numpy(X0).cuda()
    out = torch.empty((B, prob.opt.nx), dtype=torch.float64, device="cuda")
    solver.solve_raw(Pd, Xd, out, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), full["x"])


def test_mpc_warm_start_device_resident(torch_cuda):
    """Closed-loop MPC usage (point_mass_mpc.py:156-175: seed the next tick with the previous solution)
    with everything resident on the device: parameters, seed and solution are CUDA tensors, the
    solution buffer of tick k is the seed buffer of tick k+1, nothing crosses PCIe between ticks."""
    import optas_b200
    from optas_b200 import problems

    torch = torch_cuda
    prob = problems.point_mass_mpc()
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt")
    B = 256
    P, X0 = prob.sample(B, seed=1)
    Pd = torch.from_numpy(P).cuda()
    Xa, Xb = torch.from_numpy(X0).cuda(), torch.empty((B, prob.opt.nx), dtype=torch.float64, device="cuda")
    st = torch.empty(B, dtype=torch.int32, device="cuda")
    it = torch.empty(B, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    iters = []
    for tick in range(3):
        solver.solve_raw(Pd, Xa, Xb, None, None, st, it, None, stream=stream)
        torch.cuda.synchronize()
        assert (st <= 1).float().mean().item() >= 0.97
        iters.append(it.float().mean().item())
        Xa, Xb = Xb, Xa                      # previous solution becomes the seed
        Pd[:, 0:2] += 0.05 * Pd[:, 2:4] * 0  # (parameters could be advanced here; kept fixed for the assertion)
    assert iters[1] <= iters[0] and iters[2] <= iters[0]   # warm start never needs more iterations than the cold tick


def test_mpc_tick_as_cuda_graph(torch_cuda):
    """the whole tick (counter reset + solver kernel) captured once, replayed per tick with the
    solution buffer as the next seed; the replay must give the bits of a plain call."""
    import optas_b200
    from optas_b200 import problems

    torch = torch_cuda
    prob = problems.point_mass_mpc()
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt")
    B = 128
    P, X0 = prob.sample(B, seed=2)
    Pd = torch.from_numpy(P).cuda()
    X = torch.from_numpy(X0).cuda()          # seed and solution share one buffer: warm start from the last tick
    st = torch.empty(B, dtype=torch.int32, device="cuda")
    it = torch.empty(B, dtype=torch.int32, device="cuda")
    ref = torch.from_numpy(X0).cuda()
    solver.solve_raw(Pd, ref, ref, None, None, st, it, None, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    first_iters = it.clone()
    graph = solver.capture_tick(Pd, X, X, status=st, iters=it)   # (its warm-up call already solved tick 0 in place)
    X.copy_(torch.from_numpy(X0).cuda())
    graph.replay()
    torch.cuda.synchronize()
    assert (st <= 1).all()
    assert torch.equal(X, ref) and torch.equal(it, first_iters)
    # next tick: the plant moved a little; warm-started replay needs fewer iterations than the cold tick
    Pd[:, 0:2] += 0.01
    graph.replay()
    torch.cuda.synchronize()
    assert (st <= 1).float().mean() > 0.95
    assert it.float().mean() < first_iters.float().mean()


def test_batched_diagnostics(torch_cuda):
    """evaluate_cost_terms / violated_constraints for a whole batch through the streaming kernel,
    against the reference's one-instance-at-a-time evaluation on the host graph layer."""
    import optas_b200
    from optas_b200 import problems

    prob = problems.point_mass_mpc()
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt")
    B = 64
    P, X0 = prob.sample(B, seed=3)
    r = _solve_host(solver, P, X0)
    xd, pd = prob.seed_dict(r["x"]), prob.param_dict(P)
    terms = solver.evaluate_cost_terms(xd, pd)
    assert len(terms) == len(prob.opt.cost_terms) and all(t.shape == (B,) for t in terms)
    assert np.abs(sum(terms) - r["f"]).max() < 1e-10
    assert np.abs(solver.evaluate_cost(xd, pd) - r["f"]).max() < 1e-10
    lin_eq, eq, lin_ineq, ineq = solver.violated_constraints(xd, pd)
    assert [c.label for c in ineq] == list(prob.opt.ineq_constraints.keys())
    ok = r["status"] <= 1
    for c in lin_eq + eq:
        assert np.abs(c.diff[ok]).max() < 1e-7
    for c in lin_ineq + ineq:
        assert c.diff[ok].min() > -1e-7 and c.pattern.shape == c.diff.shape
    # one instance, reference path (host): same numbers
    i = 5
    x1 = {k: v[i] for k, v in xd.items()}
    p1 = {k: v[i] for k, v in pd.items()}
    for tb, t1 in zip(terms, solver.evaluate_cost_terms(x1, p1)):
        assert abs(tb[i] - float(t1)) < 1e-12
    one = solver.violated_constraints(x1, p1)
    for fam_b, fam_1 in zip((lin_eq, eq, lin_ineq, ineq), one):
        for cb, c1 in zip(fam_b, fam_1):
            assert cb.label == c1.label
            assert np.abs(cb.diff[i] - c1.diff.toarray()).max() < 1e-12


def test_c3_graphs_match_the_reference_script_golden(torch_cuda):
    """C3's cost and stacked constraint vector v = [k; g; a; -a; h; -h], evaluated by the streaming kernel from this
    package's problem, against tests/golden/config_golden.json -- values produced by the unmodified reference script's
    `point_mass_mpc.Controller` (tests/golden/make_config_golden.py)."""
    import json
    import os
    from optas_b200 import problems, sym as cs
    from optas_b200.function import B200Function

    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "config_golden.json")))["c3_point_mass_mpc"]
    opt = problems.point_mass_mpc().opt
    x, p = cs.SX.sym("x", opt.nx), cs.SX.sym("p", opt.np)
    fun = B200Function(cs.Function("fv", [x, p], [opt.f(x, p), opt.v(x, p)]))
    X = np.array([c["x"] for c in g["cases"]])
    P = np.array([c["p"] for c in g["cases"]])
    f_gpu, v_gpu = fun(X, P)
    for i, c in enumerate(g["cases"]):
        assert abs(f_gpu[i, 0] - c["f"]) < 1e-12 * max(1.0, abs(c["f"]))
        assert np.abs(v_gpu[i] - np.array(c["v"])).max() < 1e-12


def test_qp_drop_in_classes(torch_cuda):
    """OSQPSolver / CVXOPTSolver spellings (optas/solver.py:426-580) on the differential-IK QP of
    example/planar_idk.py; the reference asserts QP-only for both."""
    import optas_b200 as optas
    from optas_b200 import problems

    prob = problems.planar_idk()
    q_t, dx = np.array([2.39, -2.55, -0.46]), np.array([0.01, 0.0])
    J = prob.functions["J"](q_t).toarray()[0:2, :]
    for make in (lambda: optas.CVXOPTSolver(prob.opt).setup(), lambda: optas.OSQPSolver(prob.opt).setup(use_warm_start=True),
                 lambda: optas.CasADiSolver(prob.opt).setup("qpoases")):
        solver = make()
        solver.reset_parameters({"q": q_t, "dx": dx})
        sol = solver.solve()
        assert solver.did_solve()
        key = [k for k in sol if k.endswith("/dq")][0]
        assert np.abs(sol[key].toarray().flatten() - np.linalg.pinv(J) @ dx).max() < 1e-7
    with pytest.raises(AssertionError):
        optas.OSQPSolver(problems.dual_arm().opt).setup(use_warm_start=False)


@pytest.mark.parametrize("name", ["planar_idk", "lwr_diff_ik_qp", "box_qp"])
def test_qp_kernel_batch(torch_cuda, name):
    """SURVEY.md 8f-1 -- the dedicated QP iteration (csrc/jit/bo_qp_reg.cuh; reference role: OSQPSolver / CVXOPTSolver,
    optas/solver.py:426-580, example/experiment1.py:100-128) on a full batch through the C ABI: every instance converged to
    tol, the oracle's scaled KKT residual at the returned points, the same minimisers as the general interior-point
    kernel, fewer iterations; the box family additionally against scipy SLSQP on its numpy closed form."""
    import time
    import kkt_check
    import optas_b200
    from optas_b200 import problems
    from scipy.optimize import minimize

    prob = getattr(problems, name)()
    B = 65536
    P, X0 = prob.sample(B)
    qp = optas_b200.B200Solver(prob.opt).setup("ipopt")
    gen = optas_b200.B200Solver(prob.opt).setup("ipopt", qp=False)
    assert qp.tier_info()["tier"] == "qp" and gen.tier_info()["tier"] in ("dense", "team")
    lo = qp._lowered
    r = qp.solve_arrays(P, X0)
    g = gen.solve_arrays(P, X0)
    ok = r["status"] == 0
    assert ok.mean() >= 0.9999 and r["kkt"][ok].max() <= 1e-8  # diff-IK: 3 of 65536 start 30+ rad/s outside their limits
    both = ok & (g["status"] == 0)
    assert both.mean() > 0.9
    assert r["iters"][both].mean() < g["iters"][both].mean()
    assert np.abs(r["f"] - g["f"])[both].max() < 1e-6 * max(1.0, np.abs(g["f"][both]).max())
    idx = np.nonzero(ok)[0][np.linspace(0, ok.sum() - 1, 64).astype(int)]
    res = kkt_check.kkt_residual(prob, r["x"][idx], P[idx], r["lam"][idx, :lo.n_eq], r["lam"][idx, lo.n_eq:], scaled=True)
    assert res.max() < 2e-8
    if "P" in prob.models:
        Pm, A = prob.models["P"], prob.models["A"]
        n = Pm.shape[0]
        for i in idx[:8]:
            q, b = P[i, :n], P[i, n:]
            ref = minimize(lambda v: v @ Pm @ v + q @ v, np.zeros(n), jac=lambda v: 2 * Pm @ v + q, method="SLSQP", bounds=[(-1, 1)] * n,
                           constraints=[{"type": "eq", "fun": lambda v: A @ v - b, "jac": lambda v: A}], options={"ftol": 1e-14, "maxiter": 500})
            assert ref.status in (0, 8) and np.abs(ref.x - r["x"][i]).max() < 1e-6
    rates = {}
    for label, s in (("qp", qp), ("general", gen)):
        s.solve_arrays(P, X0)
        t0 = time.perf_counter()
        s.solve_arrays(P, X0)
        rates[label] = B / (time.perf_counter() - t0)
    print(f"[qp] {name}: iterations {r['iters'].mean():.2f} (general kernel {g['iters'][both].mean():.2f}, converged {both.mean():.4f}); "
          f"host-to-host {rates['qp'] / 1e6:.2f} M QP/s vs {rates['general'] / 1e6:.2f} M/s")


def test_joint_space_planner_batch(torch_cuda):
    """SURVEY.md 8f-3 -- example/simple_joint_space_planner.py (pose goal = position + quaternion equalities on the
    last knot, link-height inequalities on every knot) as a batch on the cooperative tier; checked by the oracle's KKT
    residual and the independent numpy kinematics."""
    import fk_ref
    import kkt_check
    import optas_b200
    from optas_b200 import problems

    prob = problems.joint_space_planner()
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt")
    assert solver.tier_info()["tier"] == "coop"
    B = 1024
    P, X0 = prob.sample(B)
    r = _solve_host(solver, P, X0)
    ok = r["status"] <= 1
    assert ok.mean() >= 0.98, np.bincount(r["status"])  # measured on B200: 1011 of 1024 converge, 13 end in a line-search failure
    lo = solver._lowered
    idx = np.where(ok)[0][:16]
    _assert_kkt(prob, lo, r, P, idx)
    sol = prob.seed_dict(r["x"][ok])
    Q, dQ = sol["med7/q/x"], sol["med7/dq/x"]
    chain = fk_ref.Chain(problems.MED7_URDF, problems.MED7_EE)
    assert np.abs(chain.fk(Q[:, :, -1])[1] - P[ok, 14:17]).max() < 1e-7
    assert np.abs(chain.quaternion(Q[:, :, -1]) - P[ok, 17:21]).max() < 1e-7
    assert np.abs(Q[:, :, 0] - P[ok, 7:14]).max() < 1e-9 and np.abs(dQ[:, :, -1]).max() < 1e-9
    assert np.abs(Q[:, :, 1:] - Q[:, :, :-1] - (4.0 / 19.0) * dQ[:, :, :-1]).max() < 1e-9


def test_axis_ik_first_stage_batch(torch_cuda):
    """SURVEY.md 8f-3 -- first stage of example/sphere_collision_avoidance.py (:20-42): position + tool-axis IK through
    `get_global_link_transform`; a rank-deficient equality block (unit vector)."""
    import kkt_check
    import optas_b200
    from optas_b200 import problems

    prob = problems.lwr_axis_ik()
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt")
    B = 4096
    P, X0 = prob.sample(B)
    P[0] = [0.825, -0.35, 0.2]
    r = _solve_host(solver, P, X0)
    ok = r["status"] <= 1
    assert ok.mean() >= 0.99, np.bincount(r["status"])
    assert np.abs(r["x"][0, :7] - problems.SPHERE_Q_START).max() < 1e-7
    lo = solver._lowered
    idx = np.where(ok)[0][:32]
    _assert_kkt(prob, lo, r, P, idx)


def test_fk_jacobian_matches_reference_golden_vectors(ik):
    """The streaming kernel against tests/golden/kinematics_golden.json -- outputs of the unmodified reference's
    RobotModel (generated by tests/golden/make_golden.py): position and the linear half of the geometric Jacobian."""
    import json
    import os
    from optas_b200.function import B200Function

    prob, _ = ik
    cases = [c for c in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kinematics_golden.json")))["kuka_lwr.urdf"]
             if c["link"] == "end_effector_ball"]
    assert len(cases) >= 4
    q = np.array([c["q"] for c in cases])
    p_gpu, J_gpu = B200Function(prob.functions["fk_jac"])(q)
    for i, c in enumerate(cases):
        assert np.abs(p_gpu[i] - np.array(c["position"]).flatten()).max() < FK_TOL
        J = np.array(c["geometric_jacobian"])[:3]
        assert np.abs(J_gpu[i].reshape(7, 3).T - J).max() < FK_TOL  # outputs are column-major flattenings (sx_container.py:83-89)


def test_ik_solution_against_the_reference_built_problem(ik):
    """Solve C2 instances on the GPU, then evaluate the REFERENCE-built problem's stacked constraint vector on the
    result via the golden file's layout: v = [k; g; a; -a; h; -h] >= -1e-8 is the feasibility statement the
    reference hands to IPOPT (solver.py:333-398, lbg = 0)."""
    prob, solver = ik
    P, X0 = prob.sample(512, seed=4)
    r = _solve_host(solver, P, X0)
    ok = r["status"] <= 1
    assert ok.mean() > 0.97
    opt = prob.opt
    for b in np.where(ok)[0][:32]:
        v = np.asarray(opt.v(r["x"][b], P[b]).toarray()).flatten()
        assert v.shape == (20,) and v.min() > -1e-8


def test_error_on_fail(torch_cuda):
    import optas_b200
    from optas_b200 import problems

    prob = problems.lwr_ik()
    solver = optas_b200.B200Solver(prob.opt, error_on_fail=True).setup("ipopt")
    p, x0 = problems.lwr_ik_example_instance()
    solver.reset_parameters({"q_nominal": p[:7], "p_goal": [5.0, 5.0, 5.0]})
    solver.reset_initial_seed({"kuka/q/x": x0})
    with pytest.raises(RuntimeError, match="Solver failed!"):
        solver.solve()


def test_smoke_entry_point(torch_cuda):
    import __graft_entry__

    __graft_entry__.smoke()


def test_seed_infeasibility_schedule_changes_the_order_not_the_results(torch_cuda):
    """``setup(schedule="seed_infeasibility")``: the device-resident batch goes to the kernel most-infeasible-seed first
    (longest-processing-time-first with theta(x0) as the predictor); instances are independent, so every output must be
    bitwise what the natural order gives."""
    import optas_b200
    from optas_b200 import problems

    torch = torch_cuda
    prob = problems.lwr_ik()
    B = 4096
    P, X0 = prob.sample(B, seed=5)
    outs = {}
    for sched in (None, "seed_infeasibility"):
        solver = optas_b200.B200Solver(prob.opt).setup("ipopt", schedule=sched)
        lo = solver._lowered
        Pd, X0d = torch.from_numpy(P).cuda(), torch.from_numpy(X0).cuda()
        X = torch.empty_like(X0d)
        lam = torch.empty((B, lo.n_eq + lo.n_ineq), dtype=torch.float64, device="cuda")
        f, kkt = torch.empty(B, dtype=torch.float64, device="cuda"), torch.empty(B, dtype=torch.float64, device="cuda")
        st, it = torch.empty(B, dtype=torch.int32, device="cuda"), torch.empty(B, dtype=torch.int32, device="cuda")
        solver.solve_raw(Pd, X0d, X, lam, f, st, it, kkt)
        torch.cuda.synchronize()
        outs[sched] = [t.cpu().numpy() for t in (X, lam, f, st, it, kkt)]
        if sched:
            theta = solver._sched["theta"].cpu().numpy().ravel()
            v = np.asarray(prob.opt.v(X0[7], P[7])).ravel()
            assert abs(theta[7] - (np.maximum(0.0, -v) ** 2).sum()) < 1e-12  # the predictor is the seed's constraint violation
    for a, b in zip(outs[None], outs["seed_infeasibility"]):
        assert np.array_equal(a, b)
    assert (outs[None][3] == 0).mean() > 0.99
