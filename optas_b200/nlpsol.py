"""``nlpsol`` / ``qpsol``: the CasADi factory signature, backed by the B200 solver.

This is the exact crossing point of the reference into native code (optas/solver.py:366-382:
``sol = cs.qpsol | cs.nlpsol``; ``self._solver = sol("solver", solver_name, problem, solver_options)``,
and :386-398: ``self._solver(x0=, p=, lbg=, ubg=)`` then ``.stats()``).  An object returned by
``nlpsol`` here is call-compatible with what CasADi returns, so the reference's unmodified
``CasADiSolver`` can sit on top of it.

Row classification happens at call time, when the bounds are known:
  lbg == ubg              -> equality      g - lbg = 0
  lbg finite              -> inequality    g - lbg >= 0
  ubg finite (< 1e10)     -> inequality    ubg - g >= 0
and, because the reference encodes every equality e = 0 as the pair (e >= 0, -e >= 0)
(optimization.py:47-51 with lbv = 0, ubv = 1e10, :302-303), two rows i, j with bounds [0, inf) whose
expressions are structural negations of each other are merged back into ONE equality -- otherwise
the feasible set has no interior and LICQ fails by construction.  Hash-consed
expression nodes make that test exact: ``neg(node_i) is node_j``.

``qpsol`` is the same object: a convex QP is an NLP the interior-point kernel solves in a handful
of iterations.
"""

from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from . import sym as cs
from .lowering import lower_nlp

INF_THRESHOLD = 1e10  # the reference's "infinity" (optimization.py:58); >= this means unbounded


def _col(v, n: int, default: float) -> np.ndarray:
    if v is None:
        return np.full(n, default)
    a = v.toarray() if isinstance(v, cs.DM) else np.asarray(v, dtype=float)
    a = a.reshape(-1)
    if a.size == 1 and n != 1:
        a = np.full(n, float(a[0]))
    if a.size != n:
        raise ValueError(f"expected {n} values, got {a.size}")
    return a.astype(float)


class NlpSolver:
    def __init__(self, name: str, plugin: str, problem: Dict, options: Optional[Dict] = None):
        self.name, self.plugin = name, plugin
        self.options = dict(options or {})
        self.x = cs.SX(problem["x"])
        self.p = cs.SX(problem["p"]) if "p" in problem and problem["p"] is not None else cs.SX(0, 1)
        self.f = cs.SX(problem["f"]) if "f" in problem else cs.SX(0.0)
        self.g = cs.SX(cs.vec(problem["g"])) if "g" in problem and problem["g"] is not None else cs.SX(0, 1)
        self.nx, self.np_, self.ng = self.x.numel(), self.p.numel(), self.g.numel()
        self._g_fun = cs.Function("g", [self.x, self.p], [self.g])
        self._compiled: Dict[bytes, Tuple] = {}
        self._stats: Dict = {}
        self.compile_only = bool(self.options.pop("b200.compile_only", False))
        # structural +-pairs among the rows of g
        nodes = self.g.nodes()
        index = {nd.idx: i for i, nd in enumerate(nodes)}
        self._partner = [-1] * self.ng
        for i, nd in enumerate(nodes):
            j = index.get(cs.n_unary(cs.OP_NEG, nd).idx, -1)
            if j > i and self._partner[i] < 0 and self._partner[j] < 0:
                self._partner[i], self._partner[j] = j, i

    # -- classification + compilation (cached per bound pattern) ----------------------------------
    def _build(self, lbx, ubx, lbg, ubg):
        from . import _capi

        key = np.concatenate([lbx, ubx, lbg, ubg]).tobytes()
        if key in self._compiled:
            return self._compiled[key]
        g_nodes = self.g.nodes()
        eq: List = []
        ineq: List = []
        rows_eq: List[Tuple[int, float, int]] = []    # (row of g, sign, partner row or -1)
        rows_ineq: List[Tuple[int, float]] = []       # (row of g or -1-k for bound on x_k, sign)
        skip = set()
        for i in range(self.ng):
            if i in skip:
                continue
            lo, up = lbg[i], ubg[i]
            gi = cs.SX(g_nodes[i])
            j = self._partner[i]
            if lo == up:
                eq.append(gi - lo)
                rows_eq.append((i, 1.0, -1))
            elif j > i and lo == 0.0 and up >= INF_THRESHOLD and lbg[j] == 0.0 and ubg[j] >= INF_THRESHOLD:
                eq.append(gi)
                rows_eq.append((i, 1.0, j))
                skip.add(j)
            else:
                if lo > -INF_THRESHOLD:
                    ineq.append(gi - lo)
                    rows_ineq.append((i, 1.0))
                if up < INF_THRESHOLD:
                    ineq.append(up - gi)
                    rows_ineq.append((i, -1.0))
        x_nodes = self.x.nodes()
        for k in range(self.nx):
            if lbx[k] == ubx[k]:
                eq.append(cs.SX(x_nodes[k]) - lbx[k])
                rows_eq.append((-1 - k, 1.0, -1))
                continue
            if lbx[k] > -INF_THRESHOLD:
                ineq.append(cs.SX(x_nodes[k]) - lbx[k])
                rows_ineq.append((-1 - k, 1.0))
            if ubx[k] < INF_THRESHOLD:
                ineq.append(ubx[k] - cs.SX(x_nodes[k]))
                rows_ineq.append((-1 - k, -1.0))
        c_eq = cs.SX(cs.vertcat(*eq)) if eq else cs.SX(0, 1)
        c_ineq = cs.SX(cs.vertcat(*ineq)) if ineq else cs.SX(0, 1)
        lowered = lower_nlp(self.x, self.p, self.f, c_eq, c_ineq)
        opts = {}
        for k, v in self.options.items():
            leaf = k.split(".")[-1]
            if leaf in ("max_iter", "max_trips", "tol", "acceptable_tol", "mu_init", "max_step"):
                opts[leaf] = int(v) if leaf in ("max_iter", "max_trips") else float(v)
        flags = _capi.BO_FLAG_COMPILE_ONLY if self.compile_only else 0
        handle = _capi.ProblemHandle(lowered, flags=flags, **opts)
        self._compiled[key] = (handle, lowered, rows_eq, rows_ineq)
        return self._compiled[key]

    # -- the CasADi call convention ----------------------------------------------------------------
    def __call__(self, x0=None, p=None, lbg=None, ubg=None, lbx=None, ubx=None, lam_g0=None, lam_x0=None) -> Dict:
        lbx_, ubx_ = _col(lbx, self.nx, -np.inf), _col(ubx, self.nx, np.inf)
        lbg_, ubg_ = _col(lbg, self.ng, -np.inf), _col(ubg, self.ng, np.inf)
        handle, lowered, rows_eq, rows_ineq = self._build(lbx_, ubx_, lbg_, ubg_)
        X0 = _col(x0, self.nx, 0.0)[None, :].copy()
        P = _col(p, self.np_, 0.0)[None, :].copy()
        X = np.empty((1, self.nx))
        lam = np.empty((1, lowered.n_eq + lowered.n_ineq))
        f = np.empty(1)
        status = np.empty(1, dtype=np.int32)
        iters = np.empty(1, dtype=np.int32)
        kkt = np.empty(1)
        handle.solve(1, P if self.np_ else None, X0, X, lam, f, status, iters, kkt)
        # multipliers in CasADi's convention: L = f + lam_g' g + lam_x' x, negative at active lower bounds
        lam_g, lam_x = np.zeros(self.ng), np.zeros(self.nx)
        y, z = lam[0, :lowered.n_eq], lam[0, lowered.n_eq:]
        for (row, sign, partner), yi in zip(rows_eq, y):
            if row < 0:
                lam_x[-1 - row] = -yi
            elif partner < 0:
                lam_g[row] = -yi
            else:
                lam_g[row], lam_g[partner] = -max(yi, 0.0), -max(-yi, 0.0)
        for (row, sign), zi in zip(rows_ineq, z):
            if row < 0:
                lam_x[-1 - row] += -sign * zi
            else:
                lam_g[row] += -sign * zi
        ok = int(status[0]) <= 1
        self._stats = {"success": ok, "iter_count": int(iters[0]),
                       "return_status": {0: "Solve_Succeeded", 1: "Solved_To_Acceptable_Level", 2: "Maximum_Iterations_Exceeded",
                                         3: "Restoration_Failed", 4: "Error_In_Step_Computation"}[int(status[0])],
                       "kkt_error": float(kkt[0])}
        return {"x": cs.DM(X[0]), "f": cs.DM(float(f[0])), "g": self._g_fun(X[0], P[0]) if self.ng else cs.DM.zeros(0, 1),
                "lam_g": cs.DM(lam_g), "lam_x": cs.DM(lam_x), "lam_p": cs.DM.zeros(self.np_, 1)}

    def stats(self) -> Dict:
        return self._stats


def nlpsol(name: str, plugin: str, problem: Dict, options: Optional[Dict] = None) -> NlpSolver:
    return NlpSolver(name, plugin, problem, options)


def qpsol(name: str, plugin: str, problem: Dict, options: Optional[Dict] = None) -> NlpSolver:
    return NlpSolver(name, plugin, problem, options)
