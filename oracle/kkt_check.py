"""ORACLE (test infrastructure): KKT residual of a candidate solution, computed on the CPU from the
problem's own IR functions through the oracle tape interpreter -- step (i) of the parity protocol
(SURVEY.md 8c).  For ``min f s.t. c_eq = 0, c_ineq >= 0`` with multipliers (y, z):

    stationarity   || grad f - J_eq' y - J_ineq' z ||_inf
    feasibility    || c_eq ||_inf ,  || min(c_ineq, 0) ||_inf
    dual sign      || min(z, 0) ||_inf
    complementarity|| z * c_ineq ||_inf

``kkt_residual`` returns the max of those per instance.  If multipliers are not supplied they are
estimated by non-negative least squares on the active set.  With ``scaled=True`` stationarity and
complementarity are divided by IPOPT's scaling factors s_d = max(100, (|y|_1 + |z|_1) / (m + n)) / 100 and
s_c = max(100, |z|_1 / n) / 100 (Waechter & Biegler 2006, eq. 6) -- the quantity IPOPT's ``tol`` (and the GPU
back-end's) bounds, which is what "converged to 1e-8" means on the reference path.
"""

from __future__ import annotations

import os
import sys
from typing import Optional

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from slsqp_driver import OracleProblem  # noqa: E402


def kkt_terms(oprob: OracleProblem, x, p, y=None, z=None, active_tol: float = 1e-6, scaled: bool = False):
    g = oprob.df(x, p)
    ce, Je = oprob.c_eq(x, p) if oprob.constrained else (np.zeros(0), np.zeros((0, oprob.nx)))
    ci, Ji = oprob.c_ineq(x, p) if oprob.constrained else (np.zeros(0), np.zeros((0, oprob.nx)))
    if y is None or z is None:
        from scipy.optimize import lsq_linear

        act = np.where(ci <= active_tol)[0]
        A = np.concatenate([Je.T, Ji[act].T], axis=1)
        if A.shape[1]:
            lb = np.concatenate([-np.inf * np.ones(len(ce)), np.zeros(len(act))])
            sol = lsq_linear(A, g, bounds=(lb, np.inf * np.ones(A.shape[1]))).x
        else:
            sol = np.zeros(0)
        y = sol[:len(ce)]
        z = np.zeros(len(ci))
        z[act] = sol[len(ce):]
    y = np.asarray(y, dtype=float).reshape(-1)
    z = np.asarray(z, dtype=float).reshape(-1)
    stat = g - Je.T @ y - Ji.T @ z
    s_d, s_c = ipopt_scaling(y, z) if scaled else (1.0, 1.0)
    return {
        "stationarity": float(np.abs(stat).max(initial=0.0)) / s_d,
        "eq": float(np.abs(ce).max(initial=0.0)),
        "ineq": float(np.abs(np.minimum(ci, 0.0)).max(initial=0.0)),
        "dual_sign": float(np.abs(np.minimum(z, 0.0)).max(initial=0.0)),
        "complementarity": float(np.abs(z * ci).max(initial=0.0)) / s_c,
    }


def ipopt_scaling(y, z, s_max: float = 100.0):
    """(s_d, s_c) of Waechter & Biegler 2006, eq. 6."""
    n_mult = len(y) + len(z)
    s_d = max(s_max, (np.abs(y).sum() + np.abs(z).sum()) / n_mult) / s_max if n_mult else 1.0
    s_c = max(s_max, np.abs(z).sum() / len(z)) / s_max if len(z) else 1.0
    return s_d, s_c


_CACHE = {}


def kkt_residual(prob, X, P, lam_eq: Optional[np.ndarray] = None, lam_ineq: Optional[np.ndarray] = None,
                 scaled: bool = False) -> np.ndarray:
    """Per-instance max KKT term.  ``prob``: an ``optas_b200.problems.Problem`` (only ``.opt`` is used)."""
    key = id(prob.opt)
    if key not in _CACHE:
        _CACHE[key] = OracleProblem(prob.opt)
    op = _CACHE[key]
    X = np.atleast_2d(X)
    P = np.atleast_2d(P)
    out = np.empty(X.shape[0])
    for i in range(X.shape[0]):
        y = None if lam_eq is None else lam_eq[i]
        z = None if lam_ineq is None else lam_ineq[i]
        out[i] = max(kkt_terms(op, X[i], P[i] if P.shape[0] > 1 else P[0], y, z, scaled=scaled).values())
    return out
