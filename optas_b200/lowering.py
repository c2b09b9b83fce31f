"""Lower an ``Optimization`` to the tapes the C ABI takes (``bo_problem_desc`` in include/b200optas.h).

The reference hands CasADi ``{x, p, f, g=v(x,p)}`` and lets ``nlpsol`` derive grad f, the
constraint Jacobian and the Hessian of the Lagrangian by AD inside the wheel
(optas/solver.py:346-382).  This module does that derivation on the
casadi-free graph layer, with two deliberate differences:

* equalities and inequalities are kept apart -- ``c_eq = [a; h]`` and ``c_ineq = [k; g]``
  (optimization.py:225-290) -- instead of the ``v = [k; g; a; -a; h; -h] >= 0`` stacking
  (optimization.py:27-51) whose +- pairs violate LICQ by construction;
* first and second derivatives are emitted *sparse*, as coordinate lists, and f / grad / c / J / H
  share one tape so that common sub-expressions of the kinematics are evaluated once.

Lagrangian sign convention: ``L = f - y'c_eq - z'c_ineq`` with ``z >= 0``.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

from . import sym as cs
from .tape import Tape


@dataclass
class Sparsity:
    row: np.ndarray  # int32 [nnz]
    col: np.ndarray  # int32 [nnz]

    @property
    def nnz(self) -> int:
        return int(self.row.shape[0])

    def dense(self, values: np.ndarray, shape: Tuple[int, int], symmetric: bool = False) -> np.ndarray:
        """Scatter ``values[..., nnz]`` into dense ``[..., rows, cols]``."""
        values = np.asarray(values, dtype=float)
        out = np.zeros(values.shape[:-1] + tuple(shape))
        out[..., self.row, self.col] = values
        if symmetric:
            off = self.row != self.col
            out[..., self.col[off], self.row[off]] = values[..., off]
        return out


@dataclass
class LoweredProblem:
    nx: int
    np_: int
    n_eq: int
    n_ineq: int
    fc: Tape        # (x, p)       -> f, c_eq, c_ineq
    kkt: Tape       # (x, p, y, z) -> f, grad, c_eq, c_ineq, Jeq nz, Jineq nz, H nz
    jac_eq: Sparsity
    jac_ineq: Sparsity
    hess: Sparsity  # lower triangle of the Lagrangian Hessian


def _coo(entries: dict, lower_only: bool = False):
    keys = sorted(k for k in entries if not lower_only or k[0] >= k[1])
    row = np.array([k[0] for k in keys], dtype=np.int32)
    col = np.array([k[1] for k in keys], dtype=np.int32)
    return Sparsity(row, col), [entries[k] for k in keys]


def constraint_vectors(opt):
    """(c_eq, c_ineq) as SX column vectors in the order [a; h] and [k; g]."""
    x, p = opt.x, opt.p
    eq = [fun(x, p) for fun in (opt.a, opt.h) if fun is not None and fun.numel_out() > 0]
    ineq = [fun(x, p) for fun in (opt.k, opt.g) if fun is not None and fun.numel_out() > 0]
    c_eq = cs.vertcat(*[cs.vec(e) for e in eq]) if eq else cs.SX(0, 1)
    c_ineq = cs.vertcat(*[cs.vec(e) for e in ineq]) if ineq else cs.SX(0, 1)
    return cs.SX(c_eq), cs.SX(c_ineq)


def lower_problem(opt) -> LoweredProblem:
    """Lower an ``Optimization`` (reference IR, optas/optimization.py:54-309)."""
    c_eq, c_ineq = constraint_vectors(opt)
    return lower_nlp(opt.x, opt.p, cs.SX(opt.f(opt.x, opt.p)), c_eq, c_ineq)


def lower_nlp(x, p, f, c_eq, c_ineq) -> LoweredProblem:
    """Lower ``min f(x,p) s.t. c_eq(x,p) = 0, c_ineq(x,p) >= 0`` given as SX column vectors."""
    x, p, f, c_eq, c_ineq = cs.SX(x), cs.SX(p), cs.SX(f), cs.SX(c_eq), cs.SX(c_ineq)
    nx, np_ = x.numel(), p.numel()
    n_eq, n_ineq = c_eq.numel(), c_ineq.numel()

    y = cs.SX.sym("__y", n_eq)
    z = cs.SX.sym("__z", n_ineq)

    _, g_entries = cs.jacobian_sparse(f, x)
    grad = [g_entries.get((0, j), cs.ZERO) for j in range(nx)]
    _, je_entries = cs.jacobian_sparse(c_eq, x)
    _, ji_entries = cs.jacobian_sparse(c_ineq, x)
    sp_eq, je_nodes = _coo(je_entries)
    sp_ineq, ji_nodes = _coo(ji_entries)

    # gradient of the Lagrangian assembled from the sparse first derivatives, then differentiated
    # once more (forward-over-forward; every partial is a {column: node} dict so the cost follows
    # the true sparsity)
    lag_grad = list(grad)
    y_nodes, z_nodes = y.nodes(), z.nodes()
    for (i, j), nd in je_entries.items():
        lag_grad[j] = cs.n_binary(cs.OP_SUB, lag_grad[j], cs.n_binary(cs.OP_MUL, y_nodes[i], nd))
    for (i, j), nd in ji_entries.items():
        lag_grad[j] = cs.n_binary(cs.OP_SUB, lag_grad[j], cs.n_binary(cs.OP_MUL, z_nodes[i], nd))
    parts = cs.forward_partials(lag_grad, x.nodes())
    h_entries = {}
    for i, dd in enumerate(parts):
        for j, nd in dd.items():
            if i >= j:
                h_entries[(i, j)] = nd
            elif (j, i) not in h_entries:
                h_entries[(j, i)] = nd  # symmetric counterpart (structurally present on one side only)
    sp_h, h_nodes = _coo(h_entries, lower_only=True)

    xs, ps = x.nodes(), p.nodes()
    f_node = f.nodes()
    ce_nodes, ci_nodes = c_eq.nodes(), c_ineq.nodes()
    fc = Tape.lower([xs, ps], [f_node, ce_nodes, ci_nodes])
    kkt = Tape.lower([xs, ps, y_nodes, z_nodes], [f_node, grad, ce_nodes, ci_nodes, je_nodes, ji_nodes, h_nodes])
    return LoweredProblem(nx, np_, n_eq, n_ineq, fc, kkt, sp_eq, sp_ineq, sp_h)


def lower_function(fun: "cs.Function") -> Tape:
    """Tape of a ``Function`` (inputs / outputs flattened column-major, one segment each)."""
    return Tape.from_function(fun)
