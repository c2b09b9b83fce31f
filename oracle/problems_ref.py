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


def sparse_kkt_residual(lowered, x: np.ndarray, p: np.ndarray, y: np.ndarray, z: np.ndarray, scaled: bool = False) -> dict:
    """Stationarity / feasibility / complementarity from the lowered kkt tape (oracle interpreter); ``scaled``: as in
    kkt_check.kkt_terms (IPOPT's s_d / s_c)."""
    from kkt_check import ipopt_scaling

    s_d, s_c = ipopt_scaling(y, z) if scaled else (1.0, 1.0)
    f, g, ce, ci, je, ji, _h = tape_vm.CTape(lowered.kkt)(x[None, :], p[None, :], y[None, :], z[None, :])
    stat = g[0].copy()
    np.subtract.at(stat, lowered.jac_eq.col, je[0] * y[lowered.jac_eq.row])
    np.subtract.at(stat, lowered.jac_ineq.col, ji[0] * z[lowered.jac_ineq.row])
    return {"stationarity": float(np.abs(stat).max(initial=0.0)) / s_d, "eq": float(np.abs(ce[0]).max(initial=0.0)),
            "ineq": float(np.abs(np.minimum(ci[0], 0.0)).max(initial=0.0)),
            "dual_sign": float(np.abs(np.minimum(z, 0.0)).max(initial=0.0)),
            "complementarity": float(np.abs(z * ci[0]).max(initial=0.0)) / s_c, "f": float(f[0, 0])}


# ---------------------------------------------------------------------------------------------------------------
# C3 and C4 in closed form (plain numpy, written from the reference scripts; independent of optas_b200.sym / .tape), and a
# derivative check of the lowered tapes that uses them: central differences of these closed forms pin grad f and the
# constraint Jacobians, central differences of those pin the Hessian of the Lagrangian.  Nothing here touches the
# expression layer's own AD (optas_b200.sym.jacobian), which the tapes were derived with.
# ---------------------------------------------------------------------------------------------------------------
T_MPC, DT_MPC = 20, 0.05


def point_mass_fc(x: np.ndarray, p: np.ndarray):
    """example/point_mass_mpc.py:88-154.  x = [Y (2 x T, column-major); dY (2 x T)], p = [curr; dcurr; goal (2 x T); obs (2 x T)].
    Returns f, c_eq = [a] (42: dynamics 38, fix curr 2, fix dcurr 2), c_ineq = [k; g] (160 bounds, 20 obstacle rows)."""
    T, dt = T_MPC, DT_MPC
    x, p = np.asarray(x, dtype=float), np.asarray(p, dtype=float)
    Y, dY = x[:2 * T].reshape(T, 2).T, x[2 * T:].reshape(T, 2).T  # 2 x T
    curr, dcurr = p[0:2], p[2:4]
    goal, obs = p[4:4 + 2 * T].reshape(T, 2).T, p[4 + 2 * T:].reshape(T, 2).T
    w = 0.0025 / float(T)
    ddY = (dY[:, 1:] - dY[:, :-1]) / dt
    f = ((goal - Y) ** 2).sum() + w * (ddY ** 2).sum()
    vec = lambda A: A.T.reshape(-1)  # column-major flattening
    # builder.py:419-469 integrate_model_states: x_t + dt xdot_t - x_{t+1}, filed as an equality with rhs 0 (:349-352: rhs - lhs)
    dyn = -(Y[:, :-1] + dt * dY[:, :-1] - Y[:, 1:])
    c_eq = np.concatenate([vec(dyn), curr - Y[:, 0], dcurr - dY[:, 0]])
    # builder.py:471-523 enforce_model_limits -> bound constraint: lower (x - lo >= 0) then upper (up - x >= 0), per derivative
    k = np.concatenate([vec(Y + 1.5), vec(1.5 - Y), vec(dY + 1.0), vec(1.0 - dY)])
    g = ((obs - Y) ** 2).sum(axis=0) - (0.2 + 0.1) ** 2  # point_mass_mpc.py:127-131
    return float(f), c_eq, np.concatenate([k, g])


T_FIG8 = 50
DT_FIG8 = 10.0 / (T_FIG8 - 1)
_MED7 = None


def _med7() -> "fk_ref.Chain":
    global _MED7
    if _MED7 is None:
        _MED7 = fk_ref.Chain(os.path.join(_ROBOTS, "med7.urdf"), "lbr_link_ee")
    return _MED7


def figure_eight_fc(x: np.ndarray, p: np.ndarray, joint_limits: bool = True):
    """example/figure_eight_plan.py:54-107 (+ enforce_model_limits on q, which BASELINE.json's config 4 adds).
    x = [Q (7 x T, column-major); dQ (7 x (T-1))], p = qc.  c_eq = [a; h]: fix q_0 = qc (7), fix dq_0 = 0 (7), dynamics (7 (T-1)),
    then quat(q_t) = quat(qc) (4 T).  c_ineq = [Q - lo; up - Q]."""
    T, dt = T_FIG8, DT_FIG8
    x, qc = np.asarray(x, dtype=float), np.asarray(p, dtype=float)
    chain = _med7()
    Q = x[:7 * T].reshape(T, 7)          # rows = time steps
    dQ = x[7 * T:].reshape(T - 1, 7)
    Rc, pc = chain.fk(qc[None, :])
    Rc, pc = Rc[0], pc[0]
    quatc = chain.quaternion(qc[None, :])[0]
    t = np.linspace(0.0, 10.0, T)
    local = np.stack([0.2 * np.sin(t * np.pi * 0.5), 0.1 * np.sin(t * np.pi), np.zeros(T)], axis=1)  # [T, 3]
    path = pc[None, :] + local @ Rc.T
    pos = chain.fk(Q)[1]
    f = 1000.0 * ((path - pos) ** 2).sum() + 0.01 * (dQ ** 2).sum()
    dyn = -(Q[:-1] + dt * dQ - Q[1:])
    a = np.concatenate([qc - Q[0], 0.0 - dQ[0], dyn.reshape(-1)])
    h = (quatc[None, :] - chain.quaternion(Q)).reshape(-1)  # add_equality_constraint(lhs, rhs) files rhs - lhs
    c_ineq = np.concatenate([(Q - chain.lower[None, :]).reshape(-1), (chain.upper[None, :] - Q).reshape(-1)]) if joint_limits else np.zeros(0)
    return float(f), np.concatenate([a, h]), c_ineq


def check_tapes_against_closed_form(lowered, fc_closed, x, p, y, z, h: float = 1e-6) -> dict:
    """Max abs error of every output of the lowered kkt tape against the closed form: values directly, first derivatives by
    central differences of the closed form, the Lagrangian Hessian by central differences of the (now pinned) tape gradient
    of the Lagrangian.  All dense, so meant for a handful of points."""
    nx = lowered.nx
    kkt = tape_vm.CTape(lowered.kkt)
    f, g, ce, ci, je, ji, hh = [o[0] for o in kkt(x[None, :], p[None, :], y[None, :], z[None, :])]
    f0, ce0, ci0 = fc_closed(x, p)
    out = {"f": abs(f[0] - f0), "c_eq": float(np.abs(ce - ce0).max(initial=0.0)), "c_ineq": float(np.abs(ci - ci0).max(initial=0.0))}
    G = np.zeros(nx)
    JE = np.zeros((lowered.n_eq, nx))
    JI = np.zeros((lowered.n_ineq, nx))
    for j in range(nx):
        e = np.zeros(nx)
        e[j] = h
        fp, cep, cip = fc_closed(x + e, p)
        fm, cem, cim = fc_closed(x - e, p)
        G[j] = (fp - fm) / (2 * h)
        JE[:, j] = (cep - cem) / (2 * h)
        JI[:, j] = (cip - cim) / (2 * h)
    scale = max(1.0, np.abs(G).max())
    out["grad"] = float(np.abs(g - G).max()) / scale
    JE_t = np.zeros_like(JE)
    JE_t[lowered.jac_eq.row, lowered.jac_eq.col] = je
    JI_t = np.zeros_like(JI)
    JI_t[lowered.jac_ineq.row, lowered.jac_ineq.col] = ji
    out["jac_eq"] = float(np.abs(JE_t - JE).max(initial=0.0)) / max(1.0, np.abs(JE).max(initial=0.0))
    out["jac_ineq"] = float(np.abs(JI_t - JI).max(initial=0.0)) / max(1.0, np.abs(JI).max(initial=0.0))

    def lag_grad(xx):
        _, g_, _, _, je_, ji_, _ = [o[0] for o in kkt(xx[None, :], p[None, :], y[None, :], z[None, :])]
        r = g_.copy()
        np.subtract.at(r, lowered.jac_eq.col, je_ * y[lowered.jac_eq.row])
        np.subtract.at(r, lowered.jac_ineq.col, ji_ * z[lowered.jac_ineq.row])
        return r

    H = np.zeros((nx, nx))
    for j in range(nx):
        e = np.zeros(nx)
        e[j] = h
        H[:, j] = (lag_grad(x + e) - lag_grad(x - e)) / (2 * h)
    H_t = np.zeros((nx, nx))
    H_t[lowered.hess.row, lowered.hess.col] = hh
    H_t = H_t + np.tril(H_t, -1).T
    out["hess"] = float(np.abs(H_t - 0.5 * (H + H.T)).max()) / max(1.0, np.abs(H).max())
    return out
