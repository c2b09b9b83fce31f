"""ORACLE (test infrastructure): KKT residual of a candidate solution, computed on the CPU from the
problem's own IR functions through the oracle tape interpreter -- step (i) of the parity protocol
(SURVEY.md 8c).  For ``min f s.t. c_eq = 0, c_ineq >= 0`` with multipliers (y, z):

    stationarity   || grad f - J_eq' y - J_ineq' z ||_inf
    feasibility    || c_eq ||_inf ,  || min(c_ineq, 0) ||_inf
    dual sign      || min(z, 0) ||_inf
    complementarity|| z * c_ineq ||_inf

``kkt_residual`` returns the max of those per instance.  If multipliers are not supplied they are
estimated by non-negative least squares on the active set.
"""

from __future__ import annotations

import os
import sys
from typing import Optional

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from slsqp_driver import OracleProblem  # noqa: E402


def kkt_terms(oprob: OracleProblem, x, p, y=None, z=None, active_tol: float = 1e-6):
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
    return {
        "stationarity": float(np.abs(stat).max(initial=0.0)),
        "eq": float(np.abs(ce).max(initial=0.0)),
        "ineq": float(np.abs(np.minimum(ci, 0.0)).max(initial=0.0)),
        "dual_sign": float(np.abs(np.minimum(z, 0.0)).max(initial=0.0)),
        "complementarity": float(np.abs(z * ci).max(initial=0.0)),
    }


_CACHE = {}


def kkt_residual(prob, X, P, lam_eq: Optional[np.ndarray] = None, lam_ineq: Optional[np.ndarray] = None) -> np.ndarray:
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
        out[i] = max(kkt_terms(op, X[i], P[i] if P.shape[0] > 1 else P[0], y, z).values())
    return out
