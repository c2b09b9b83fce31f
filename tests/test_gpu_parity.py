"""Parity protocol of SURVEY.md 8c on the CUDA path, one test per BASELINE.json config, at the BASELINE batch where the
oracle side stays within seconds (the GPU solves the full batch; the CPU oracle checks a sample of it).

(i)   oracle KKT residual -- stationarity, equality / inequality feasibility, dual sign, complementarity, computed on the
      CPU from the problem's tapes -- <= 1e-8 in IPOPT's scaling for every instance the GPU reports CONVERGED (status 0);
      instances reported ACCEPTABLE (status 1) are counted separately and held to 1e-6;
(ii)  polish: an independent CPU solver seeded at the GPU result stays there (1e-6 relative on x, 1e-9 relative on f) --
      scipy SLSQP for C2 (the reference's runnable formulation), the oracle's sparse interior point (oracle/ipm_ref.py,
      a restatement of the reference's nlpsol("ipopt") call) for the horizon configs where dense SLSQP is O(n^3);
(iii) same seed, same basin: the fraction of instances on which the oracle, started from the SAME seed, lands on the
      same minimiser, printed and compared with the oracle's own success rate.
IPOPT itself cannot be installed in this image: parity with its iterates is unpinned, this protocol
is what stands in.  Everything on the GPU side goes through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KKT_TOL = 1e-8          # status 0 ("converged"): the reference's ipopt tol
KKT_TOL_ACCEPTABLE = 1e-6  # status 1 ("acceptable"): ipopt's acceptable_tol
X_RTOL = 1e-6
F_RTOL = 1e-9


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch


def _solve(solver, P, X0):
    return solver.solve_arrays(np.ascontiguousarray(P), np.ascontiguousarray(X0))


def _check_kkt(lo, r, P, idx, tol):
    import problems_ref

    worst = {}
    for i in idx:
        k = problems_ref.sparse_kkt_residual(lo, r["x"][i], P[i], r["lam"][i, :lo.n_eq], r["lam"][i, lo.n_eq:], scaled=True)
        for name in ("stationarity", "eq", "ineq", "dual_sign", "complementarity"):
            worst[name] = max(worst.get(name, 0.0), k[name])
    assert max(worst.values()) <= tol, worst
    return worst


def _polish(ipm, lo, r, P, idx, max_step):
    moves_x, moves_f = [], []
    for i in idx:
        # The oracle refines the GPU's point to 1e-10 -- two digits beyond the GPU's tolerance -- on the SAME central-path point (barrier parameter = the GPU's final mean s z; an interior-point
        # solution at tol 1e-8, IPOPT's included, sits on the mu ~ 2.5e-9 point of the path, O(1e-6) from the mu -> 0 limit
        # in the directions of weakly active bounds) and must not move it.
        z = r["lam"][i, lo.n_eq:]
        mu_gpu = float(np.mean(z * ipm.eval_fc(r["x"][i], P[i])[2])) if lo.n_ineq else 1e-9
        q = ipm.solve(P[i], r["x"][i], y0=r["lam"][i, :lo.n_eq], z0=z, mu0=max(mu_gpu, 1e-12), fixed_mu=True, tol=1e-10,
                      max_iter=8, max_step=max_step)
        assert q["kkt"] <= 1e-9, (i, q["status"], q["kkt"], q["iters"])
        moves_x.append(np.abs(q["x"] - r["x"][i]).max() / max(1.0, np.abs(q["x"]).max()))
        moves_f.append(abs(q["f"] - r["f"][i]) / max(1.0, abs(q["f"])))
    # Every instance but at most one within X_RTOL / F_RTOL, none beyond 100 x that.  Why "but one": a KKT error of 1e-8 pins x
    # to 1e-6 only while the reduced KKT system's condition number is below ~100; about 1 in 300 instances of C3 from the
    # zero seed (a knot whose obstacle constraint is weakly active) is worse conditioned -- the host build of the same
    # kernel shows one instance of 300 moving 2.3e-5 under this polish, the other 299 at most 5.7e-7 ("Oracle and
    # parity status") -- and which 8 instances are drawn depends on which ones converged.
    moves_x, moves_f = np.sort(moves_x), np.sort(moves_f)
    assert moves_x[:-1].max(initial=0.0) <= X_RTOL and moves_f[:-1].max(initial=0.0) <= F_RTOL, (moves_x, moves_f)
    assert moves_x[-1] <= 100 * X_RTOL and moves_f[-1] <= 100 * F_RTOL, (moves_x, moves_f)
    if len(moves_x) > 1 and moves_x[-1] > X_RTOL:
        print(f"   polish: one ill-conditioned instance moved {moves_x[-1]:.1e} (the others at most {moves_x[-2]:.1e})")
    worst_x = moves_x[-2] if len(moves_x) > 1 and moves_x[-1] > X_RTOL else moves_x[-1]
    worst_f = moves_f[-2] if len(moves_f) > 1 and moves_f[-1] > F_RTOL else moves_f[-1]
    return worst_x, worst_f


def _horizon_config(factory, B, opts, n_kkt=32, n_polish=8, min_converged=0.99, zero_seed=False):
    import ipm_ref
    import optas_b200
    from optas_b200 import problems

    prob = getattr(problems, factory)()
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt", opts)
    assert solver.tier_info()["tier"] == "coop"
    lo = solver._lowered
    P, X0 = prob.sample(B)
    if zero_seed:
        X0 = np.zeros_like(X0)
    r = _solve(solver, P, X0)
    conv, acc = r["status"] == 0, r["status"] == 1
    print(f"\n[{factory}{' zero seed' if zero_seed else ''}] B={B}: converged (1e-8) {int(conv.sum())}, acceptable (1e-6) {int(acc.sum())}, "
          f"failed {int((r['status'] >= 2).sum())}, mean iterations {r['iters'].mean():.1f}")
    assert conv.mean() >= min_converged, np.bincount(r["status"], minlength=5)
    rng = np.random.default_rng(0)
    idx = rng.choice(np.where(conv)[0], size=min(n_kkt, int(conv.sum())), replace=False)
    worst = _check_kkt(lo, r, P, idx, KKT_TOL)
    print(f"   oracle KKT (scaled) on {len(idx)} converged instances: " + ", ".join(f"{k} {v:.1e}" for k, v in worst.items()))
    if acc.any():
        _check_kkt(lo, r, P, np.where(acc)[0][:n_kkt], KKT_TOL_ACCEPTABLE)
    wx, wf = _polish(ipm_ref.SparseIPM(lo), lo, r, P, idx[:n_polish], solver._handle.options()["max_step"])
    print(f"   polish (oracle interior point seeded at the GPU result) on {min(n_polish, len(idx))}: max rel move {wx:.1e}, rel df {wf:.1e}")
    return prob, solver, P, X0, r


def test_c3_parity_protocol_full_batch(torch_cuda):
    """BASELINE config 3: point_mass_mpc.py T = 20, batch 16384 (example/point_mass_mpc.py:88-175)."""
    prob, solver, P, X0, r = _horizon_config("point_mass_mpc", 16384, {})
    ok = r["status"] == 0
    sol = prob.seed_dict(r["x"][ok])
    Y, dY = sol["point_mass/y/x"], sol["point_mass/dy/x"]
    assert np.abs(Y[:, :, 1:] - Y[:, :, :-1] - 0.05 * dY[:, :, :-1]).max() < 1e-9  # dynamics, size-independent property


def test_c3_from_the_zero_seed(torch_cuda):
    """The workload as SURVEY.md 8d defines it: x0 = 0, the script's cold first tick (point_mass_mpc.py:157-161).  Every
    knot starts on the obstacle's path; the interior-point iteration needs its feasibility-restoration phase here
    (62 % converged without it in round 1)."""
    _horizon_config("point_mass_mpc", 16384, {}, zero_seed=True, min_converged=0.99)


def test_c4_parity_protocol(torch_cuda):
    """BASELINE config 4: figure_eight_plan.py T = 50 with joint-limit bounds (nx 693, 557 eq, 700 ineq), batch 4096 on the
    GPU; stationarity AND complementarity asserted (round 1 checked one instance's feasibility only)."""
    _horizon_config("figure_eight", 4096, {"max_iter": 400, "max_trips": 2500}, min_converged=0.99)


def test_c5_parity_protocol(torch_cuda):
    """BASELINE config 5: dual_arm.py T = 50 (nx 1386, 700 eq): one GPU's shard of the 32768 batch when sharded over 8."""
    import problems_ref

    prob, solver, P, X0, r = _horizon_config("dual_arm", 4096, {})
    for i in np.where(r["status"] == 0)[0][:8]:  # the numpy closed form of the problem, independent of the expression layer
        assert abs(r["f"][i] - problems_ref.dual_arm_cost(r["x"][i], P[i])) < 1e-10
        assert np.abs(problems_ref.dual_arm_constraints(r["x"][i], P[i])).max() < 1e-8


def test_c5_from_the_zero_seed(torch_cuda):
    """dual_arm.py never calls reset_initial_seed: the reference solves from x0 = 0 (optas/solver.py:76)."""
    _horizon_config("dual_arm", 512, {}, zero_seed=True, min_converged=0.99)


def test_c2_parity_protocol_and_same_seed_basin(torch_cuda):
    """BASELINE config 2 at its batch (65536): KKT at 1e-8 on 512, SLSQP polish on 64, and protocol step (iii) on 512."""
    import ipm_ref
    import kkt_check
    import optas_b200
    import slsqp_driver
    from optas_b200 import problems

    prob = problems.lwr_ik()
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt")
    assert solver.tier_info()["tier"] == "team"
    lo = solver._lowered
    B = 65536
    P, X0 = prob.sample(B, seed=0)
    r = _solve(solver, P, X0)
    conv, acc = r["status"] == 0, r["status"] == 1
    print(f"\n[lwr_ik] B={B}: converged (1e-8) {int(conv.sum())}, acceptable (1e-6) {int(acc.sum())}, failed {int((r['status'] >= 2).sum())}")
    assert conv.mean() >= 0.999
    idx = np.where(conv)[0][:512]
    res = kkt_check.kkt_residual(prob, r["x"][idx], P[idx], r["lam"][idx][:, :lo.n_eq], r["lam"][idx][:, lo.n_eq:], scaled=True)
    assert res.max() <= KKT_TOL, res.max()
    op = slsqp_driver.OracleProblem(prob.opt)
    worst = 0.0
    for i in idx[:64]:
        pol = slsqp_driver.solve_slsqp(op, P[i], r["x"][i], form="split", options={"ftol": 1e-15, "maxiter": 200})
        worst = max(worst, np.abs(pol.x - r["x"][i]).max() / max(1.0, np.abs(pol.x).max()))
    assert worst <= X_RTOL, worst
    # (iii) same seed, same basin, against two independent CPU solvers started from the same x0
    ipm = ipm_ref.SparseIPM(lo)
    n = 512
    ok_s = same_s = ok_i = same_i = 0
    for i in np.arange(n):
        if not conv[i]:
            continue
        o = slsqp_driver.solve_slsqp(op, P[i], X0[i], form="split", options={"ftol": 1e-15, "maxiter": 500})
        if o.success:
            ok_s += 1
            same_s += bool(np.abs(o.x - r["x"][i]).max() < 1e-5)
        q = ipm.solve(P[i], X0[i], max_step=0.5)
        if q["status"] == 0:
            ok_i += 1
            same_i += bool(np.abs(q["x"] - r["x"][i]).max() < 1e-5)
    print(f"   same seed, same basin over the first {n} instances: vs scipy SLSQP {same_s}/{ok_s} = {same_s / max(1, ok_s):.3f} "
          f"(SLSQP success rate {ok_s / n:.3f}); vs the oracle interior point {same_i}/{ok_i} = {same_i / max(1, ok_i):.3f} "
          f"(its success rate {ok_i / n:.3f}); GPU success rate {conv[:n].mean():.3f}")
    # the same algorithm family from the same seed must land in the same basin almost always; an SQP method takes other
    # paths across the >= 4 local minima of this problem, so its agreement is reported and only bounded loosely
    assert same_i / max(1, ok_i) >= 0.95
    assert same_s / max(1, ok_s) >= 0.6
    assert conv[:n].mean() >= ok_i / n - 0.01  # the GPU solver fails no more often than the CPU restatement does
