"""CPU oracle (TEST INFRASTRUCTURE, not a product path): inverse dynamics in plain numpy.

Restates the reference's ``RobotModel.rnea`` (optas/models.py:1731-1884: Craig's recursive Newton-Euler sweep,
serial chain, first joint fixed = base, last joint fixed = tool, gravity -9.81 z as an upward base acceleration,
inertial origin rpy not applied) independently of ``optas_b200`` -- it parses the URDF itself.

Pinned by physics that does not go through the recursion (tests/test_oracle.py):
  * the gravity torques rnea(q, 0, 0) equal the gradient of the potential energy  sum_b m_b g z_com,b(q)
    taken by central differences through oracle/fk_ref.py's forward kinematics,
  * the joint-space inertia matrix  M[:, j] = rnea(q, 0, e_j) - rnea(q, 0, 0)  is symmetric positive definite,
  * the velocity terms are quadratic in qd:  c(q, 2 qd) = 4 c(q, qd).
The reference pins its own implementation only against pybullet (absent here) at atol 8e-2
(tests/test_models.py:1040-1052).
"""

from __future__ import annotations

import xml.etree.ElementTree as ET
from typing import List

import numpy as np

from fk_ref import Chain, angvec2r, rpy2r

GRAVITY = 9.81


def _bodies(urdf_path: str):
    root = ET.parse(urdf_path).getroot()
    f3 = lambda el, key: np.array([float(t) for t in el.get(key, "0 0 0").split()]) if el is not None else np.zeros(3)
    inert = []
    for link in root.findall("link"):
        el = link.find("inertial")
        if el is None:
            continue
        i = el.find("inertia")
        g = lambda k: float(i.get(k, 0.0))
        inert.append(dict(name=link.get("name"), m=float(el.find("mass").get("value")), c=f3(el.find("origin"), "xyz"),
                          I=np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])))
    joints = [j for j in root.findall("joint") if j.get("type") is not None]
    assert joints[0].get("type") == "fixed", "first joint must be fixed"
    out = []
    for j, body in zip(joints[1:], inert[1:]):
        assert j.get("type") in ("revolute", "continuous", "fixed")
        assert j.find("child").get("link") == body["name"], "serial chain in document order expected"
        ax = j.find("axis")
        axis = f3(ax, "xyz") if ax is not None else np.array([1.0, 0.0, 0.0])
        out.append(dict(body, r=f3(j.find("origin"), "xyz"), R0=rpy2r(f3(j.find("origin"), "rpy")), axis=axis / np.linalg.norm(axis)))
    return out


def rnea(urdf_path: str, q, qd, qdd) -> np.ndarray:
    """tau [ndof] for one (q, qd, qdd)."""
    bodies = _bodies(urdf_path)
    n = len(bodies)
    q, qd, qdd = (np.asarray(v, dtype=float).flatten() for v in (q, qd, qdd))
    R = [b["R0"] @ angvec2r(np.array([q[k]]), b["axis"])[0] if k < n - 1 else b["R0"] for k, b in enumerate(bodies)]
    om, omd, vd = [np.zeros(3)], [np.zeros(3)], [np.array([0.0, 0.0, GRAVITY])]
    F: List[np.ndarray] = []
    N: List[np.ndarray] = []
    for k, b in enumerate(bodies):
        E = R[k].T
        if k < n - 1:
            z = E @ b["axis"]
            o = E @ om[k] + z * qd[k]
            od = E @ omd[k] + np.cross(E @ om[k], z * qd[k]) + z * qdd[k]
        else:
            o, od = E @ om[k], E @ omd[k]
        a = E @ (vd[k] + np.cross(omd[k], b["r"]) + np.cross(om[k], np.cross(om[k], b["r"])))
        F.append(b["m"] * (a + np.cross(od, b["c"]) + np.cross(o, np.cross(o, b["c"]))))
        N.append(b["I"] @ od + np.cross(o, b["I"] @ o))
        om.append(o), omd.append(od), vd.append(a)
    f, m = F[-1], N[-1] + np.cross(bodies[-1]["c"], F[-1])
    tau = np.zeros(n - 1)
    for k in range(n - 2, -1, -1):
        m = N[k] + R[k + 1] @ m + np.cross(bodies[k]["c"], F[k]) + np.cross(bodies[k + 1]["r"], R[k + 1] @ f)
        f = R[k + 1] @ f + F[k]
        tau[k] = m @ (R[k].T @ bodies[k]["axis"])
    return tau


def potential_energy(urdf_path: str, q) -> float:
    """sum_b m_b g z_com,b(q) through the independent forward kinematics of fk_ref.Chain."""
    V = 0.0
    for b in _bodies(urdf_path):
        Rw, pw = Chain(urdf_path, b["name"]).fk(np.asarray(q, dtype=float)[None, :])
        V += b["m"] * GRAVITY * (pw[0] + Rw[0] @ b["c"])[2]
    return V
