"""ORACLE (test infrastructure): plain-numpy closed forms of the horizon workloads, written from the
reference scripts and independent of the optas_b200 expression layer.

  dual_arm_cost / dual_arm_constraints   example/dual_arm.py:17-128 (C5)
      x = [ql (7x50, column-major); dql (7x49); qr; dqr],  p = [qcl; qcr]
      f = 0.01 ||dQl||^2 + 0.01 ||dQr||^2 + sum_t ||p_l(q_t) - path_l,t||^2 + (same for r)
      c = [ qc - q_0 (fix_configuration, builder.py:525-539) ; -(q_t + dt dq_t - q_{t+1}) (builder.py:419-469) ] per arm
  sparse_kkt_residual                    KKT stationarity / feasibility from the lowered tapes evaluated
                                         by the oracle's tape interpreter (for problems whose dense Jacobian
                                         Functions are too large to form: nx = 693 / 1386)
"""

from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import fk_ref  # noqa: E402
import tape_vm  # noqa: E402

_ROBOTS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "optas_b200", "robots")
T_DUAL = 50
DT_DUAL = 10.0 / (T_DUAL - 1)


def _lwr_with_base(y_offset: float) -> fk_ref.Chain:
    chain = fk_ref.Chain(os.path.join(_ROBOTS, "kuka_lwr.urdf"), "end_effector_ball")
    # add_base_frame("global_world", xyz=[0, y, 0]) (models.py:552-588): a fixed joint in front of the chain
    chain.joints = [("fixed", np.array([0.0, y_offset, 0.0]), np.zeros(3), np.array([1.0, 0.0, 0.0]), -1)] + chain.joints
    return chain


def _path(p0: np.ndarray, side: float) -> np.ndarray:
    """[T, 3] end-effector path of dual_arm.py:84-113 for start position p0 (side = +1 left, -1 right)."""
    p1 = p0 + np.array([-0.1, 0.1 * side, -0.2])
    p2 = p1 + np.array([0.0, 0.0, 0.3])
    out = np.empty((T_DUAL, 3))
    for i in range(T_DUAL):
        a_ = i / (T_DUAL - 1)
        if a_ < 0.4:
            a = a_ / 0.4
            out[i] = a * p1 + (1 - a) * p0
        elif a_ < 0.5:
            out[i] = p1
        else:
            a = (a_ - 0.5) / 0.5
            out[i] = a * p2 + (1 - a) * p1
    return out


def _split(x: np.ndarray):
    n, m = 7 * T_DUAL, 7 * (T_DUAL - 1)
    ql = x[:n].reshape(T_DUAL, 7)
    dql = x[n:n + m].reshape(T_DUAL - 1, 7)
    qr = x[n + m:2 * n + m].reshape(T_DUAL, 7)
    dqr = x[2 * n + m:].reshape(T_DUAL - 1, 7)
    return ql, dql, qr, dqr


def dual_arm_cost(x: np.ndarray, p: np.ndarray) -> float:
    ql, dql, qr, dqr = _split(np.asarray(x, dtype=float))
    f = 0.01 * (dql ** 2).sum() + 0.01 * (dqr ** 2).sum()
    for q, qc, y_off, side in ((ql, p[:7], -0.25, 1.0), (qr, p[7:], 0.25, -1.0)):
        chain = _lwr_with_base(y_off)
        pos = chain.fk(q)[1]
        p0 = chain.fk(qc[None, :])[1][0]
        f += ((pos - _path(p0, side)) ** 2).sum()
    return float(f)


def dual_arm_constraints(x: np.ndarray, p: np.ndarray) -> np.ndarray:
    """Linear equalities in the order the builder files them: fix(l), fix(r), dynamics(l), dynamics(r)."""
    ql, dql, qr, dqr = _split(np.asarray(x, dtype=float))
    fix_l, fix_r = p[:7] - ql[0], p[7:] - qr[0]
    # add_equality_constraint stores rhs - lhs with rhs = 0 (builder.py:349-352): the integrator residual enters negated
    dyn_l = -(ql[:-1] + DT_DUAL * dql - ql[1:]).reshape(-1)
    dyn_r = -(qr[:-1] + DT_DUAL * dqr - qr[1:]).reshape(-1)
    return np.concatenate([fix_l, fix_r, dyn_l, dyn_r])


def sparse_kkt_residual(lowered, x: np.ndarray, p: np.ndarray, y: np.ndarray, z: np.ndarray) -> dict:
    """Stationarity / feasibility / complementarity from the lowered kkt tape (oracle interpreter)."""
    f, g, ce, ci, je, ji, _h = tape_vm.CTape(lowered.kkt)(x[None, :], p[None, :], y[None, :], z[None, :])
    stat = g[0].copy()
    np.subtract.at(stat, lowered.jac_eq.col, je[0] * y[lowered.jac_eq.row])
    np.subtract.at(stat, lowered.jac_ineq.col, ji[0] * z[lowered.jac_ineq.row])
    return {"stationarity": float(np.abs(stat).max(initial=0.0)), "eq": float(np.abs(ce[0]).max(initial=0.0)),
            "ineq": float(np.abs(np.minimum(ci[0], 0.0)).max(initial=0.0)),
            "dual_sign": float(np.abs(np.minimum(z, 0.0)).max(initial=0.0)),
            "complementarity": float(np.abs(z * ci[0]).max(initial=0.0)), "f": float(f[0, 0])}
