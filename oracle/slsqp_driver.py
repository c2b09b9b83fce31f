"""ORACLE (test infrastructure, never imported by optas_b200): the reference's own *runnable*
solver path, ``ScipyMinimizeSolver("SLSQP")`` (optas/solver.py:587-813, offered as the alternative
back-end at example/example.py:41), restated on scipy with the problem functions evaluated by the
oracle's C tape interpreter (stand-in for CasADi's VM).

Formulation, exactly as the reference passes it to scipy (solver.py:652-679, 716-734, 783-792):

    minimize(fun=f, jac=df, x0=x0, method="SLSQP",
             constraints=[{"type": "ineq", "fun": v, "jac": dv}])          v = [k; g; a; -a; h; -h] >= 0

``form="v"`` reproduces that; ``form="split"`` hands scipy the equalities as equalities and is used
for the tight-tolerance "polish" runs of the parity protocol.  The reference runs
SLSQP with scipy's defaults (ftol 1e-6), which only determines x* to ~1e-4; parity tests therefore
pass ``options={"ftol": 1e-15, "maxiter": 500}`` and say so.

Parity pinning: tests/test_oracle.py checks this driver against the Booth known answer the
reference pins for every back-end (tests/test_solver.py:45-54: x=1, y=3, success) and against the
survey's provisional C1 golden.  IPOPT itself is not installable here: parity with the CasADi+IPOPT
path is UNPINNED beyond those.
"""

from __future__ import annotations

import os
import sys
from typing import Dict, Optional

import numpy as np
from scipy.optimize import minimize

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from tape_vm import CTape  # noqa: E402


class OracleProblem:
    """Numeric callbacks of an ``Optimization`` (any object with the reference's IR attributes
    f, df, v, dv, a, h, k, g and their Jacobian functions), evaluated by the C tape VM."""

    def __init__(self, opt):
        from optas_b200.tape import Tape
        from optas_b200 import sym as cs

        self.opt = opt
        self.nx, self.np_ = opt.nx, opt.np
        x, p = opt.x, opt.p

        def T(expr_list):
            return CTape(Tape.from_function(cs.Function("o", [x, p], expr_list)))

        self._f = T([opt.f(x, p), cs.jacobian(opt.f(x, p), x)])
        self.constrained = opt.v is not None and opt.nv > 0
        if self.constrained:
            v = opt.v(x, p)
            self._v = T([v, cs.jacobian(v, x)])
            eq = [fn(x, p) for fn in (opt.a, opt.h) if fn is not None and fn.numel_out() > 0]
            ineq = [fn(x, p) for fn in (opt.k, opt.g) if fn is not None and fn.numel_out() > 0]
            ce = cs.SX(cs.vertcat(*[cs.vec(e) for e in eq])) if eq else cs.SX(0, 1)
            ci = cs.SX(cs.vertcat(*[cs.vec(e) for e in ineq])) if ineq else cs.SX(0, 1)
            self.n_eq, self.n_ineq = ce.numel(), ci.numel()
            self._ce = T([ce, cs.jacobian(ce, x)])
            self._ci = T([ci, cs.jacobian(ci, x)])
        else:
            self.n_eq = self.n_ineq = 0
        self.nv = opt.nv if self.constrained else 0

    # all return numpy; Jacobians are dense [m, nx] (tape outputs are column-major flattenings)
    def f(self, x, p):
        return float(self._f(x, p)[0][0, 0])

    def df(self, x, p):
        return self._f(x, p)[1][0].copy()

    def _val_jac(self, tape, m, x, p):
        val, jac = tape(x, p)
        return val[0].copy(), jac[0].reshape(self.nx, m).T.copy()

    def v(self, x, p):
        return self._v(x, p)[0][0].copy()

    def dv(self, x, p):
        return self._val_jac(self._v, self.nv, x, p)[1]

    def c_eq(self, x, p):
        return self._val_jac(self._ce, self.n_eq, x, p)

    def c_ineq(self, x, p):
        return self._val_jac(self._ci, self.n_ineq, x, p)


def solve_slsqp(prob: OracleProblem, p: np.ndarray, x0: np.ndarray, form: str = "v", tol: Optional[float] = None,
                options: Optional[Dict] = None):
    """One instance through scipy SLSQP.  Returns the scipy OptimizeResult (``.x .success .nit .fun``)."""
    p = np.asarray(p, dtype=float)
    kw = {"fun": lambda x: prob.f(x, p), "jac": lambda x: prob.df(x, p), "x0": np.asarray(x0, dtype=float),
          "method": "SLSQP"}
    if tol is not None:
        kw["tol"] = tol
    if options is not None:
        kw["options"] = options
    if prob.constrained:
        if form == "v":
            kw["constraints"] = [{"type": "ineq", "fun": lambda x: prob.v(x, p), "jac": lambda x: prob.dv(x, p)}]
        elif form == "split":
            cons = []
            if prob.n_eq:
                cons.append({"type": "eq", "fun": lambda x: prob.c_eq(x, p)[0], "jac": lambda x: prob.c_eq(x, p)[1]})
            if prob.n_ineq:
                cons.append({"type": "ineq", "fun": lambda x: prob.c_ineq(x, p)[0], "jac": lambda x: prob.c_ineq(x, p)[1]})
            kw["constraints"] = cons
        else:
            raise ValueError(form)
    return minimize(**kw)


def _one_blas_thread():
    """One BLAS / OpenMP thread per worker process: the batch is parallel over instances (one process per core), nested
    library threads would only fight over the same cores."""
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(1)
    except Exception:
        pass


def _worker(args):
    opt_factory, P, X0, form, tol, options = args
    _one_blas_thread()
    prob = OracleProblem(opt_factory().opt)
    out = np.empty_like(X0)
    ok = np.zeros(len(X0), dtype=bool)
    nit = np.zeros(len(X0), dtype=np.int64)
    for i in range(len(X0)):
        r = solve_slsqp(prob, P[i], X0[i], form=form, tol=tol, options=options)
        out[i], ok[i], nit[i] = r.x, r.success, r.nit
    return out, ok, nit


def solve_batch(opt_factory, P: np.ndarray, X0: np.ndarray, workers: int = 1, form: str = "v",
                tol: Optional[float] = None, options: Optional[Dict] = None):
    """Batch = a loop, one instance at a time per worker process -- how the reference would have to
    be driven (it has no batching).  ``opt_factory`` is a picklable zero-argument
    callable returning an object with ``.opt`` (e.g. ``optas_b200.problems.lwr_ik``)."""
    if workers <= 1:
        return _worker((opt_factory, P, X0, form, tol, options))
    import multiprocessing as mp

    chunks = [(opt_factory, Pc, Xc, form, tol, options)
              for Pc, Xc in zip(np.array_split(P, workers), np.array_split(X0, workers)) if len(Xc)]
    with mp.get_context("fork").Pool(len(chunks)) as pool:
        parts = pool.map(_worker, chunks)
    return (np.concatenate([a for a, _, _ in parts]), np.concatenate([b for _, b, _ in parts]),
            np.concatenate([c for _, _, c in parts]))
