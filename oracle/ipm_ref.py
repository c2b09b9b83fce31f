"""ORACLE (test infrastructure, never imported by optas_b200): a CPU restatement of what the reference's
hot path runs for the horizon configs -- ``casadi.nlpsol("solver", "ipopt", {x, p, f, g})`` called once per instance
(optas/solver.py:382, 395; example/figure_eight_plan.py:110, example/dual_arm.py:125, example/point_mass_mpc.py:143).

IPOPT is an un-vendored, unpinned dependency of the reference (setup.py:20-30 lists `casadi`; IPOPT + MUMPS ship
inside that wheel) and is NOT installable in this image, so its published algorithm is restated here:
Waechter & Biegler, "On the implementation of an interior-point filter line-search algorithm for large-scale
nonlinear programming", Math. Program. 106 (2006), Algorithm A with the paper's default constants:

    barrier problem in slack form     min f(x) - mu sum log s   s.t. c_E(x) = 0, c_I(x) - s = 0
    error measure E_mu (eq. 5), mu update (eq. 7: kappa_eps 10, kappa_mu 0.2, theta_mu 1.5), tau = max(0.99, 1 - mu)
    primal-dual system (eq. 13) reduced to (dx, dy), factored sparse (scipy SuperLU stands in for MUMPS)
    inertia correction (Algorithm IC: delta_w 1e-4 first, x100 / x8, kappa_w^- 1/3, delta_c 1e-8 mu^(1/4)); SuperLU
      returns no inertia, so the test uses Debreu's lemma: the KKT matrix has inertia (n, m, 0) iff
      W + Sigma + delta_w I + rho JE'JE is positive definite for large rho, checked by an LU with diagonal pivots only
      (all pivots positive, no row exchange); a singular system (rank-deficient JE) shows as a failed / inaccurate solve
    fraction-to-the-boundary rule (eq. 15), filter line search (Section 2.3: gamma_theta gamma_phi 1e-5, eta_phi 1e-8,
      s_phi 2.3, s_theta 1.1, theta_max 1e4 max(1, theta_0), theta_min 1e-4 max(1, theta_0)), multiplier reset (eq. 16)
    NOT restated: second-order correction, the restoration phase, adaptive mu, scaling, bound relaxation.

It reads the same lowered tapes the C ABI takes (f, grad f, c, sparse Jacobians, sparse Hessian of the Lagrangian --
the quantities CasADi hands IPOPT through its callbacks) and evaluates them with the oracle's C tape interpreter.
Used as: (1) the CPU baseline of bench.py for C4 / C5, where the reference's only other runnable formulation (dense
SLSQP, oracle/slsqp_driver.py) is O(n^3) per iteration with n = 693 / 1386; (2) an independent solver for the parity
protocol of SURVEY.md 8c on those sizes: *polish* (seeded at the GPU result it must stay there) and *same seed, same
basin*.  Parity with IPOPT's own iterates is UNPINNED (no IPOPT binary exists here).
"""

from __future__ import annotations

import os
import sys
from typing import Optional

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import splu

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from tape_vm import CTape  # noqa: E402

CONVERGED, ACCEPTABLE, MAX_ITER, LINE_SEARCH, NUMERICAL = 0, 1, 2, 3, 4


class SparseIPM:
    """One lowered problem (``optas_b200.lowering.LoweredProblem``: tapes + coordinate sparsity) bound to the CPU."""

    def __init__(self, lowered):
        lo = self.lo = lowered
        self.nx, self.me, self.mi = lo.nx, lo.n_eq, lo.n_ineq
        self.kkt = CTape(lo.kkt)
        self.fc = CTape(lo.fc)
        n = self.nx
        # symmetric Hessian from its lower triangle: value index per (row, col) coordinate, off-diagonals mirrored
        hr, hc = np.asarray(lo.hess.row), np.asarray(lo.hess.col)
        off = hr != hc
        self._h_rows = np.concatenate([hr, hc[off]])
        self._h_cols = np.concatenate([hc, hr[off]])
        self._h_src = np.concatenate([np.arange(len(hr)), np.arange(len(hr))[off]])
        self._je = (np.asarray(lo.jac_eq.row), np.asarray(lo.jac_eq.col))
        self._ji = (np.asarray(lo.jac_ineq.row), np.asarray(lo.jac_ineq.col))
        self._eye_x = sp.identity(n, format="csc")
        self._eye_e = sp.identity(self.me, format="csc")

    # -- evaluation --------------------------------------------------------------------------------
    def eval_kkt(self, x, p, y, z):
        f, g, ce, ci, je, ji, h = self.kkt(x[None, :], p[None, :], y[None, :], z[None, :])
        n = self.nx
        H = sp.csc_matrix((h[0][self._h_src], (self._h_rows, self._h_cols)), shape=(n, n))
        JE = sp.csr_matrix((je[0], self._je), shape=(self.me, n))
        JI = sp.csr_matrix((ji[0], self._ji), shape=(self.mi, n))
        return float(f[0, 0]), g[0], ce[0], ci[0], JE, JI, H

    def eval_fc(self, x, p):
        f, ce, ci = self.fc(x[None, :], p[None, :])
        return float(f[0, 0]), ce[0], ci[0]

    # -- the iteration -----------------------------------------------------------------------------
    def solve(self, p, x0, tol: float = 1e-8, acceptable_tol: float = 1e-6, max_iter: int = 300, mu_init: float = 0.1,
              max_step: float = 0.0, y0: Optional[np.ndarray] = None, z0: Optional[np.ndarray] = None,
              mu0: Optional[float] = None, fixed_mu: bool = False, trace: bool = False):
        """Returns dict(x, y, z, f, status, iters, kkt).  ``y0 / z0 / mu0`` warm-start the duals (used by the polish
        check, which starts at a candidate primal-dual solution); ``max_step`` > 0 caps ||alpha dx||_inf like the GPU
        back-end's option of the same name (0: IPOPT behaviour, no cap).  ``fixed_mu``: solve the barrier sub-problem of
        ``mu0`` only -- mu is never reduced and convergence is measured by E_mu (Waechter & Biegler eq. 5 with the
        complementarity residual s z - mu); the polish check uses this to refine a candidate on ITS central-path point."""
        kappa_eps, kappa_mu, theta_mu, tau_min, s_max = 10.0, 0.2, 1.5, 0.99, 100.0
        gamma_theta, gamma_phi, eta_phi, s_phi, s_theta, kappa_sigma = 1e-5, 1e-5, 1e-8, 2.3, 1.1, 1e10
        n, me, mi = self.nx, self.me, self.mi
        p = np.asarray(p, dtype=float)
        x = np.array(x0, dtype=float)
        mu = float(mu0) if mu0 is not None else mu_init
        f, ce, ci = self.eval_fc(x, p)
        s = np.maximum(ci, 1e-2 * np.maximum(1.0, np.abs(ci))) if mi else np.zeros(0)
        z = np.array(z0, dtype=float) if z0 is not None else (mu / s if mi else np.zeros(0))
        if z0 is not None and mi:
            s = np.maximum(ci, 1e-12)
            z = np.maximum(z, 1e-300)
        y = np.array(y0, dtype=float) if y0 is not None else np.zeros(me)
        filt = []
        dw_last, n_acc, theta_max, theta_min = 0.0, 0, np.inf, 0.0
        mu_min = tol * 0.1
        status, it, err0 = MAX_ITER, 0, np.inf
        for it in range(max_iter + 1):
            f, g, ce, ci, JE, JI, H = self.eval_kkt(x, p, y, z)
            rd = g - JE.T @ y - JI.T @ z
            e_dual = np.abs(rd).max(initial=0.0)
            e_prim = max(np.abs(ce).max(initial=0.0), np.abs(ci - s).max(initial=0.0))
            sum_z = np.abs(z).sum()
            sum_mult = np.abs(y).sum() + sum_z
            s_d = max(s_max, sum_mult / max(1, me + mi)) / s_max if me + mi else 1.0
            s_c = max(s_max, sum_z / max(1, mi)) / s_max if mi else 1.0
            err0 = max(e_dual / s_d, e_prim, (np.abs(s * z - mu).max(initial=0.0) if fixed_mu else (s * z).max(initial=0.0)) / s_c)
            if trace:
                print(f"it {it:3d} f {f:.6e} err0 {err0:.3e} (dual {e_dual / s_d:.3e} prim {e_prim:.3e}) mu {mu:.2e}")
            if not np.isfinite(err0) or not np.isfinite(f):
                status = NUMERICAL
                break
            if err0 <= tol:
                status = CONVERGED
                break
            n_acc = n_acc + 1 if err0 <= acceptable_tol else 0
            if n_acc >= 15:
                status = ACCEPTABLE
                break
            if it >= max_iter:
                status = MAX_ITER
                break
            if mi and not fixed_mu:
                for _ in range(8):
                    err_mu = max(e_dual / s_d, e_prim, np.abs(s * z - mu).max() / s_c)
                    if err_mu <= kappa_eps * mu and mu > mu_min:
                        mu = max(mu_min, min(kappa_mu * mu, mu ** theta_mu))
                        filt = []
                    else:
                        break
            tau = max(tau_min, 1.0 - mu)
            sigma = z / s if mi else np.zeros(0)
            theta0 = np.abs(ce).sum() + np.abs(ci - s).sum()
            phi0 = f - mu * np.log(s).sum()
            if it == 0:
                theta_max, theta_min = 1e4 * max(1.0, theta0), 1e-4 * max(1.0, theta0)
            # ---- direction (eq. 13 reduced), Algorithm IC with the curvature test ----
            W = H + JI.T @ sp.diags(sigma) @ JI if mi else H
            rhs_x = -rd - (JI.T @ ((z * (ci - s) + (s * z - mu)) / s) if mi else 0.0)  # -(grad L) - JI'((z cI - mu)/s)
            rhs = np.concatenate([rhs_x, -ce])
            dw, dc, attempt, sol = 0.0, 0.0, 0, None
            JtJ = (JE.T @ JE) * 1e6 if me else None
            while True:
                K = sp.bmat([[W + dw * self._eye_x, JE.T], [JE, -dc * self._eye_e]], format="csc") if me else (W + dw * self._eye_x).tocsc()
                ok = True
                try:
                    with np.errstate(all="ignore"):
                        sol = splu(K).solve(rhs)
                except RuntimeError:
                    ok = False
                singular = not ok or not np.isfinite(sol).all()
                if not singular:
                    res = np.abs(K @ sol - rhs).max(initial=0.0)
                    singular = res > 1e-6 * max(1.0, np.abs(rhs).max(initial=0.0), np.abs(sol).max(initial=0.0))
                if not (singular and me and dc == 0.0):
                    if _positive_definite((W + dw * self._eye_x + JtJ) if me else (W + dw * self._eye_x)):
                        if not singular:
                            break
                attempt += 1
                if attempt > 60:
                    status = NUMERICAL
                    sol = None
                    break
                if singular and me and dc == 0.0:
                    dc = 1e-8 * mu ** 0.25
                elif dw == 0.0:
                    dw = 1e-4 if dw_last == 0.0 else max(1e-20, dw_last / 3.0)
                else:
                    dw *= 100.0 if dw_last == 0.0 else 8.0
            if sol is None:
                break
            if dw > 0.0:
                dw_last = dw
            dx, dy = sol[:n], -sol[n:]
            ds = JI @ dx + (ci - s) if mi else np.zeros(0)
            dz = -z + mu / s - sigma * ds if mi else np.zeros(0)
            neg = ds < 0.0
            a_max = min(1.0, (-tau * s[neg] / ds[neg]).min()) if neg.any() else 1.0
            if max_step > 0.0 and a_max * np.abs(dx).max(initial=0.0) > max_step:
                a_max = max_step / np.abs(dx).max()
            negz = dz < 0.0
            a_z = min(1.0, (-tau * z[negz] / dz[negz]).min()) if negz.any() else 1.0
            dphi = g @ dx - (mu * (ds / s).sum() if mi else 0.0)
            # ---- filter line search ----
            a, accepted = a_max, False
            for _ls in range(40):
                xt, st = x + a * dx, s + a * ds
                with np.errstate(all="ignore"):
                    ft, cet, cit = self.eval_fc(xt, p)
                    thetat = np.abs(cet).sum() + np.abs(cit - st).sum()
                    phit = ft - mu * np.log(st).sum() if mi else ft
                if np.isfinite(phit) and np.isfinite(thetat) and thetat <= theta_max and \
                        all(thetat <= (1 - gamma_theta) * th or phit <= ph - gamma_phi * th for th, ph in filt):
                    ftype = dphi < 0.0 and theta0 <= theta_min and \
                        (theta0 <= 0.0 or np.log(a) + s_phi * np.log(-dphi) > s_theta * np.log(theta0))
                    slack = 10.0 * 2.2e-16 * abs(phi0)
                    if ftype:
                        armijo = phit - phi0 - slack <= eta_phi * a * dphi
                        if armijo:
                            accepted = True
                            break
                    elif thetat <= (1 - gamma_theta) * theta0 or phit - slack <= phi0 - gamma_phi * theta0:
                        filt.append(((1 - gamma_theta) * theta0, phi0 - gamma_phi * theta0))
                        accepted = True
                        break
                a *= 0.5
            if not accepted:
                status = ACCEPTABLE if err0 <= acceptable_tol else LINE_SEARCH
                break
            x = xt
            if mi:
                s = np.maximum(st, cit)  # slack reset (keeps c_I - s <= 0 from accumulating round-off)
                z = z + a_z * dz
                z = np.maximum(np.minimum(z, kappa_sigma * mu / s), mu / (kappa_sigma * s))
            y = y + a * dy
        return {"x": x, "y": y, "z": z, "f": f, "status": status, "iters": it, "kkt": err0}


def _positive_definite(M) -> bool:
    """LU with pivots restricted to the diagonal (SuperLU symmetric mode, threshold 0): for a symmetric matrix the
    pivots are those of LDL', all positive iff M is positive definite (Sylvester).  A row exchange means a zero pivot."""
    M = sp.csc_matrix(M)
    try:
        with np.errstate(all="ignore"):
            lu = splu(M, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options={"SymmetricMode": True})
    except RuntimeError:
        return False
    if not np.array_equal(lu.perm_r, lu.perm_c):
        return False
    d = lu.U.diagonal()
    return bool(np.isfinite(d).all() and (d > 0.0).all())


_CACHE = {}


def _ipm_for(factory):
    key = getattr(factory, "__name__", repr(factory))
    if key not in _CACHE:
        from optas_b200.lowering import lower_problem

        prob = factory()
        _CACHE[key] = (prob, SparseIPM(lower_problem(prob.opt)))
    return _CACHE[key]


def _worker(args):
    factory, P, X0, kw = args
    from slsqp_driver import _one_blas_thread

    _one_blas_thread()
    _, ipm = _ipm_for(factory)
    X = np.empty_like(X0)
    ok = np.zeros(len(X0), dtype=bool)
    nit = np.zeros(len(X0), dtype=np.int64)
    for i in range(len(X0)):
        r = ipm.solve(P[i], X0[i], **kw)
        X[i], ok[i], nit[i] = r["x"], r["status"] == CONVERGED, r["iters"]
    return X, ok, nit


def solve_batch(factory, P: np.ndarray, X0: np.ndarray, workers: int = 1, **kw):
    """A batch the only way the reference can run one: a loop, one nlpsol call per instance per worker process.
    ``factory``: picklable zero-argument callable returning an object with ``.opt`` (``optas_b200.problems.*``)."""
    if workers <= 1 or len(X0) <= 1:
        return _worker((factory, P, X0, kw))
    import multiprocessing as mp

    chunks = [(factory, Pc, Xc, kw) for Pc, Xc in zip(np.array_split(P, workers), np.array_split(X0, workers)) if len(Xc)]
    with mp.get_context("fork").Pool(len(chunks)) as pool:
        parts = pool.map(_worker, chunks)
    return (np.concatenate([a for a, _, _ in parts]), np.concatenate([b for _, b, _ in parts]),
            np.concatenate([c for _, _, c in parts]))
