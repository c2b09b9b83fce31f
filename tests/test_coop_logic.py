"""Cooperative tier (one instance per CTA; csrc/bo_coop.cpp + csrc/jit/bo_ipm_cta.cuh) on a GPU-less machine.

The generated kernel source compiles for the host (tests/hostsim.py: one "thread", the warp-wide lane
programs emulated lane by lane with the same reduction tree), so the pieces the tier adds -- tape
partitioning + per-class code generation, target-owned KKT assembly, the level-scheduled LDL' lane programs,
the CTA-parallel iteration -- are checked here against numpy and against the thread-per-instance tier.
A test harness, not a product path."""
import ctypes as C

import numpy as np
import pytest

from hostsim import HostSim

import optas_b200
from optas_b200 import problems


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _coop(prob, **kw):
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt", compile_only=True, coop=True, **kw)
    lo = solver._lowered
    tab, dtab = np.ascontiguousarray(solver.ldl_table()), np.ascontiguousarray(solver.dtable())
    sim = HostSim(solver.kernel_source(), lo.nx, lo.np_, lo.n_eq, lo.n_ineq, ldl_table=tab, dtable=dtab)
    return solver, sim, lo, tab, dtab


def _eval_tapes(sim, lo, tab, dtab, p, x, y, z):
    kkt = [np.zeros(max(n, 1)) for n in lo.kkt.out_sizes]
    sim.lib.hostsim_coop_kkt(_vp(tab), _vp(dtab), _vp(p), _vp(x), _vp(y), _vp(z), *[_vp(o) for o in kkt])
    fc = [np.zeros(max(n, 1)) for n in lo.fc.out_sizes]
    sim.lib.hostsim_coop_fc(_vp(tab), _vp(dtab), _vp(p), _vp(x), *[_vp(o) for o in fc])
    return [o[:n] for o, n in zip(kkt, lo.kkt.out_sizes)], [o[:n] for o, n in zip(fc, lo.fc.out_sizes)]


@pytest.mark.parametrize("generated", [True, False])
@pytest.mark.parametrize("name", ["point_mass_mpc", "dual_arm"])
def test_partitioned_tapes_equal_the_reference_interpreter(name, generated, monkeypatch):
    """Partitioning (independent sub-tapes, parameter-only prefix, partial sums reduced in fixed order) and the
    per-class code generation must not change what the tapes compute: compare with Tape.eval_numpy."""
    if not generated:
        monkeypatch.setenv("BO_NO_GEN_TAPES", "1")
    prob = getattr(problems, name)()
    solver, sim, lo, tab, dtab = _coop(prob)
    info = solver.tier_info()
    assert info["tier"] == "coop" and info["generated_tapes"] == int(generated)
    assert info["kkt_components"] >= 16
    rng = np.random.default_rng(5)
    P, X0 = prob.sample(1, seed=4)
    p = np.ascontiguousarray(P[0])
    x = X0[0] + 0.1 * rng.standard_normal(lo.nx)
    y, z = rng.standard_normal(max(lo.n_eq, 1))[:lo.n_eq], rng.random(max(lo.n_ineq, 1))[:lo.n_ineq] + 0.1
    y, z = np.ascontiguousarray(y), np.ascontiguousarray(z)
    kkt, fc = _eval_tapes(sim, lo, tab, dtab, p, x, y, z)
    for got, ref in zip(kkt, lo.kkt.eval_numpy([x, p, y, z])):
        if ref.size:
            assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    for got, ref in zip(fc, lo.fc.eval_numpy([x, p])):
        if ref.size:
            assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


def test_horizon_problems_compile_to_a_few_classes():
    """All stages of a horizon problem are isomorphic: thousands of components, a handful of functions."""
    s = optas_b200.B200Solver(problems.dual_arm().opt).setup("ipopt", compile_only=True)
    info = s.tier_info()
    assert info["tier"] == "coop" and info["generated_tapes"] == 1
    assert info["kkt_components"] > 1000 and info["kkt_classes"] <= 32
    assert info["kkt_code_rows"] < info["kkt_total_instr"] / 10
    assert info["smem_dynamic"] <= 227 * 1024
    # the IK problem (one kinematic chain, 10 x 10 KKT system) runs on the team tier (4 threads per instance, shared memory)
    assert optas_b200.B200Solver(problems.lwr_ik().opt).setup("ipopt", compile_only=True).tier_info()["tier"] == "team"


@pytest.mark.parametrize("segments", ["1", "2", "3"])
def test_lane_program_factorisation_and_solve(segments, monkeypatch):
    """Assembly + level-scheduled LDL' + both substitutions, as executed by the lane programs, against a dense
    numpy solve; the pivot signs must report the inertia (Debreu: positive definite (1,1) block <=> correct)."""
    monkeypatch.setenv("BO_SEGMENTS", segments)
    prob = problems.point_mass_mpc()
    solver, sim, lo, tab, dtab = _coop(prob)
    assert solver.tier_info()["segments"] == int(segments)
    rng = np.random.default_rng(0)
    P, X0 = prob.sample(1, seed=0)
    x = X0[0] + 0.1 * rng.standard_normal(lo.nx)
    y, z = rng.standard_normal(lo.n_eq), rng.random(lo.n_ineq) + 0.1
    _, _, _, _, JE, JI, H = lo.kkt.eval_numpy([x, P[0], y, z])
    sigma = rng.random(lo.n_ineq) + 0.5
    nx, me = lo.nx, lo.n_eq
    Hd = lo.hess.dense(H, (nx, nx), symmetric=True)
    JEd, JId = lo.jac_eq.dense(JE, (me, nx)), lo.jac_ineq.dense(JI, (lo.n_ineq, nx))
    sim.lib.hostsim_coop_linsolve.argtypes = [C.c_void_p] * 6 + [C.c_double] * 3 + [C.c_void_p, C.c_int]
    # merged: the right-hand side goes through the factor program (forward substitution folded into the factorisation)
    for (rho, dw, dcp, hscale), merged in [(c, m) for c in ((1e3, 0.5, 1e-3, 1.0), (1e6, 0.0, 0.0, 1.0), (10.0, 0.0, 0.0, -50.0)) for m in (0, 1)]:
        K = np.zeros((nx + me, nx + me))
        K[:nx, :nx] = hscale * Hd + JId.T @ np.diag(sigma) @ JId + rho * JEd.T @ JEd + dw * np.eye(nx)
        K[nx:, :nx], K[:nx, nx:] = JEd, JEd.T
        K[nx:, nx:] = -dcp * np.eye(me)
        rhs = rng.standard_normal(nx + me)
        sol = rhs.copy()
        Hs = np.ascontiguousarray(hscale * H)
        bad = sim.lib.hostsim_coop_linsolve(_vp(tab), _vp(dtab), _vp(Hs), _vp(np.ascontiguousarray(JE)),
                                            _vp(np.ascontiguousarray(JI)), _vp(sigma), rho, dw, dcp, _vp(sol), merged)
        pd = np.linalg.eigvalsh(K[:nx, :nx]).min() > 0
        if dcp == 0.0 and pd:
            # -dc = 0: the y pivots are -JE (..)^-1 JE' < 0 for full-rank JE
            pass
        assert (bad == 0) == bool(pd), (rho, dw, dcp, hscale, bad)
        if bad == 0:
            ref = np.linalg.solve(K, rhs)
            assert np.abs(sol - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())


def test_coop_tier_agrees_with_the_thread_per_instance_tier():
    """Same algorithm, same constants: on C3 both tiers must land on the same solutions (round-off apart)."""
    prob = problems.point_mass_mpc()
    P, X0 = prob.sample(24, seed=1)
    _, sim, lo, _, _ = _coop(prob)
    r = sim.solve(P, X0)
    ref_solver = optas_b200.B200Solver(prob.opt).setup("ipopt", compile_only=True, coop=False)
    assert ref_solver.tier_info()["tier"] == "sparse"
    ref = HostSim(ref_solver.kernel_source(), lo.nx, lo.np_, lo.n_eq, lo.n_ineq, ldl_table=ref_solver.ldl_table(),
                  dtable=ref_solver.dtable()).solve(P, X0)
    assert (r["status"] == ref["status"]).all() and (r["status"] == 0).all()
    assert (r["iters"] == ref["iters"]).all()
    assert np.abs(r["x"] - ref["x"]).max() < 1e-9
    assert np.abs(r["f"] - ref["f"]).max() < 1e-12


def test_coop_tier_table_indices_are_in_bounds():
    """Every operand of the lane programs addresses the factor / right-hand side (or their zero padding cell)."""
    prob = problems.point_mass_mpc()
    solver, _, lo, tab, _ = _coop(prob)
    info = solver.tier_info()
    nvals, nk = info["factor_vals"], lo.nx + lo.n_eq
    for slot, solve in ((5, False), (6, True), (7, True)):  # CT_PROG_FAC / FWD / BWD: packets of 1 + pk words
        h = tab[tab[slot]:]
        W, pk, stream = int(h[0]), int(h[1]) >> 8, int(h[2])
        assert pk in (1, 2, 4, 8)
        for w in range(W):
            first, n = int(h[4 + 2 * w]), int(h[5 + 2 * w])
            assert n % (pk + 1) == 0
            words = tab[stream + 64 * first: stream + 64 * (first + n)].reshape(-1, pk + 1, 32, 2).astype(np.int64) & 0xFFFFFFFF
            flags = words[:, 0, :, 0] & 0xF  # bits 4..6: log2 of the lanes per target of the round
            assert (words[:, 0, :, 0] == words[:, 0, :1, 0]).all() and (words[:, 0, :, 0] < 128).all()  # warp-uniform
            opened = False
            for f in flags[:, 0]:  # rounds are FIRST ... LAST sequences; a level ends only where a round ends (or on an empty packet)
                if f & 1:
                    assert not opened
                    opened = True
                if f & 2:
                    assert opened
                    opened = False
                if f & 4:
                    assert not opened
            tg = words[:, 0, :, 1] & 0x7FFF
            # factor targets beyond the factor values are right-hand-side entries (bp sits right behind vals: nvals + 1 + j)
            assert ((tg == 0x7FFF) | (tg < (nk if solve else nvals + 1 + nk))).all()
            assert solve or ((tg == 0x7FFF) | (tg != nvals)).all()
            assert (tg[(flags & 2) == 0] == 0x7FFF).all()
            ops = words[:, 1:]
            a, b, c = ops[..., 0] & 0xFFFF, ops[..., 0] >> 16, ops[..., 1]
            assert (a <= (nvals if solve else nvals + 1 + nk)).all() and (b <= (nk if solve else nvals)).all() and (c <= nvals).all()


@pytest.mark.parametrize("with_eq,with_ineq", [(False, True), (True, False), (False, False)])
def test_coop_tier_without_equalities_or_inequalities(with_eq, with_ineq):
    """Edge cases of the problem shape (n_eq = 0 and / or n_ineq = 0) through the cooperative tier, on a synthetic
    chain problem straight through the C ABI (lower_nlp -> bo_problem_create), against the thread-per-instance tier."""
    from optas_b200 import _capi, sym as cs
    from optas_b200.lowering import lower_nlp

    n = 24
    x, p = cs.SX.sym("x", n), cs.SX.sym("p", n)
    f = cs.sumsqr(x - p) + 0.5 * cs.sumsqr(cs.sin(x[1:]) - x[:-1]) + 0.1 * cs.sumsqr(x[1:] * x[:-1])
    c_eq = cs.vertcat(*[x[3 * k] + x[3 * k + 1] * x[3 * k + 2] - 0.3 for k in range(n // 3)]) if with_eq else cs.SX(0, 1)
    c_in = cs.vertcat(*[1.2 - x[k] * x[k] - 0.5 * x[k + 1] for k in range(n - 1)]) if with_ineq else cs.SX(0, 1)
    lo = lower_nlp(x, p, f, c_eq, c_in)
    assert (lo.n_eq > 0) == with_eq and (lo.n_ineq > 0) == with_ineq
    rng = np.random.default_rng(11)
    P = 0.6 * rng.standard_normal((6, n))
    X0 = np.zeros((6, n))
    res = {}
    for name, flag in (("coop", _capi.BO_FLAG_COOP), ("thread", _capi.BO_FLAG_NO_COOP)):
        h = _capi.ProblemHandle(lo, flags=_capi.BO_FLAG_COMPILE_ONLY | flag)
        assert h.tier_info()["tier"] == ("coop" if name == "coop" else "sparse")
        sim = HostSim(h.source(), lo.nx, lo.np_, lo.n_eq, lo.n_ineq, ldl_table=h.ldl_table(), dtable=h.dtable())
        res[name] = sim.solve(P, X0)
        assert (res[name]["status"] <= 1).all(), res[name]["status"]
    assert np.abs(res["coop"]["x"] - res["thread"]["x"]).max() < 1e-7
    assert np.abs(res["coop"]["f"] - res["thread"]["f"]).max() < 1e-9
    if with_ineq:
        xs = res["coop"]["x"]
        assert (1.2 - xs[:, :-1] ** 2 - 0.5 * xs[:, 1:] > -1e-8).all()
