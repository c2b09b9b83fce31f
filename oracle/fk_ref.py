"""ORACLE (test infrastructure, never imported by optas_b200): plain-numpy restatement of the
reference's forward kinematics, independent of the optas_b200 expression layer.

Follows, function by function:
  rotx/roty/rotz      optas/spatialmath.py:115-157
  rpy2r ("zyx")       optas/spatialmath.py:160-185   R = rotz(yaw) @ roty(pitch) @ rotx(roll)
  angvec2r            optas/spatialmath.py:89-99     Rodrigues: I + sin(t) K + (1 - cos(t)) K K, K = skew(unit(v))
  rt2tr / r2t         optas/spatialmath.py:102-112, 188-199
  Quaternion          optas/spatialmath.py:277-375   (xyzw; q0 * q1 composes like Rot(q1) Rot(q0))
  FK chain            optas/models.py:826-868        T = prod_j  rt2tr(rpy2r(rpy_j), xyz_j) [ @ r2t(angvec2r(q_j, axis_j)) ]
  quaternion chain    optas/models.py:1049-1088
  geometric Jacobian  optas/models.py:1199-1264      column j = [ z_j x (e - p_j) ; z_j ]

Parity pinning: the rotation primitives are checked against scipy.spatial.transform.Rotation -- the
oracle the reference's own tests/test_spatialmath.py uses (:123-129, :217-223, :415-425, :484-489) --
and the chain against the closed form of tests/tester_robot.urdf (tests/test_fk_oracle.py here).
All functions are vectorised over a leading batch axis.
"""

from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from typing import List, Tuple

import numpy as np

_ROBOTS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "optas_b200", "robots")


def rotx(t):
    t = np.asarray(t, dtype=float)
    c, s, o, z = np.cos(t), np.sin(t), np.ones_like(t), np.zeros_like(t)
    return np.stack([np.stack([o, z, z], -1), np.stack([z, c, -s], -1), np.stack([z, s, c], -1)], -2)


def roty(t):
    t = np.asarray(t, dtype=float)
    c, s, o, z = np.cos(t), np.sin(t), np.ones_like(t), np.zeros_like(t)
    return np.stack([np.stack([c, z, s], -1), np.stack([z, o, z], -1), np.stack([-s, z, c], -1)], -2)


def rotz(t):
    t = np.asarray(t, dtype=float)
    c, s, o, z = np.cos(t), np.sin(t), np.ones_like(t), np.zeros_like(t)
    return np.stack([np.stack([c, -s, z], -1), np.stack([s, c, z], -1), np.stack([z, z, o], -1)], -2)


def rpy2r(rpy, opt: str = "zyx"):
    r, p, y = (np.asarray(v, dtype=float) for v in rpy)
    if opt in ("xyz", "arm"):
        return rotx(y) @ roty(p) @ rotz(r)
    if opt in ("zyx", "vehicle"):
        return rotz(y) @ roty(p) @ rotx(r)
    if opt in ("yxz", "camera"):
        return roty(y) @ rotx(p) @ rotz(r)
    raise ValueError(opt)


def skew(v):
    x, y, z = v
    return np.array([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]])


def angvec2r(theta, v):
    """theta: [...]; v: 3-vector.  Returns [..., 3, 3]."""
    theta = np.asarray(theta, dtype=float)
    v = np.asarray(v, dtype=float)
    K = skew(v / np.linalg.norm(v))
    return np.eye(3) + np.sin(theta)[..., None, None] * K + (1.0 - np.cos(theta))[..., None, None] * (K @ K)


def quat_mul(q0, q1):
    """Reference ``Quaternion.__mul__`` (xyzw): self = q0, argument = q1."""
    x0, y0, z0, w0 = np.moveaxis(np.asarray(q0, dtype=float), -1, 0)
    x1, y1, z1, w1 = np.moveaxis(np.asarray(q1, dtype=float), -1, 0)
    return np.stack([
        x1 * w0 + y1 * z0 - z1 * y0 + w1 * x0,
        -x1 * z0 + y1 * w0 + z1 * x0 + w1 * y0,
        x1 * y0 - y1 * x0 + z1 * w0 + w1 * z0,
        -x1 * x0 - y1 * y0 - z1 * z0 + w1 * w0,
    ], -1)


def quat_fromrpy(rpy):
    r, p, y = (np.asarray(v, dtype=float) for v in rpy)
    cr, sr, cp, sp, cy, sy = np.cos(.5 * r), np.sin(.5 * r), np.cos(.5 * p), np.sin(.5 * p), np.cos(.5 * y), np.sin(.5 * y)
    q = np.stack([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                  cr * cp * cy + sr * sp * sy], -1)
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def quat_fromangvec(theta, v):
    theta = np.asarray(theta, dtype=float)
    v = np.asarray(v, dtype=float)
    u = v / np.linalg.norm(v)
    return np.concatenate([np.sin(.5 * theta)[..., None] * u, np.cos(.5 * theta)[..., None]], -1)


# ----------------------------------------------------------------------------------------------
# minimal URDF chain reader (independent of optas_b200.urdf)
# ----------------------------------------------------------------------------------------------


class Chain:
    """Root-to-tip list of (type, xyz, rpy, axis, actuated index or -1) plus joint limits."""

    def __init__(self, urdf_path: str, tip: str):
        root = ET.parse(urdf_path).getroot()
        joints = {}
        parent_of = {}
        order = []
        for j in root.findall("joint"):
            if j.get("type") is None:
                continue
            name = j.get("name")
            o = j.find("origin")
            ax = j.find("axis")
            lim = j.find("limit")
            f3 = lambda el, key: [float(t) for t in el.get(key, "0 0 0").split()] if el is not None else [0.0, 0.0, 0.0]
            joints[name] = dict(type=j.get("type"), parent=j.find("parent").get("link"), child=j.find("child").get("link"),
                                xyz=f3(o, "xyz"), rpy=f3(o, "rpy"),
                                axis=[float(t) for t in ax.get("xyz").split()] if ax is not None else [1.0, 0.0, 0.0],
                                lower=float(lim.get("lower", 0.0)) if lim is not None else -1e9,
                                upper=float(lim.get("upper", 0.0)) if lim is not None else 1e9)
            parent_of[joints[name]["child"]] = name
            order.append(name)
        actuated = [n for n in order if joints[n]["type"] != "fixed"]  # index = position among non-fixed joints
        self.ndof = len(actuated)
        self.lower = np.array([joints[n]["lower"] for n in actuated])
        self.upper = np.array([joints[n]["upper"] for n in actuated])
        chain: List[str] = []
        link = tip
        while link in parent_of:
            chain.append(parent_of[link])
            link = joints[parent_of[link]]["parent"]
        chain.reverse()
        self.joints = [(joints[n]["type"], np.array(joints[n]["xyz"]), np.array(joints[n]["rpy"]), np.array(joints[n]["axis"]),
                        actuated.index(n) if n in actuated else -1) for n in chain]

    def fk(self, q) -> Tuple[np.ndarray, np.ndarray]:
        """q [B, ndof] -> (R [B,3,3], p [B,3]) of the tip in the root frame (models.py:826-868)."""
        q = np.atleast_2d(np.asarray(q, dtype=float))
        B = q.shape[0]
        R = np.tile(np.eye(3), (B, 1, 1))
        p = np.zeros((B, 3))
        for typ, xyz, rpy, axis, idx in self.joints:
            p = p + R @ xyz
            R = R @ rpy2r(rpy)
            if typ in ("revolute", "continuous"):
                R = R @ angvec2r(q[:, idx], axis)
            elif typ == "prismatic":
                p = p + (R @ axis) * q[:, idx:idx + 1]
            elif typ != "fixed":
                raise NotImplementedError(typ)
        return R, p

    def position_and_linear_jacobian(self, q) -> Tuple[np.ndarray, np.ndarray]:
        """(p [B,3], J [B,3,ndof]) with J[:, :, j] = z_j x (e - p_j) (revolute) / z_j (prismatic),
        the top half of the geometric Jacobian of models.py:1199-1264."""
        q = np.atleast_2d(np.asarray(q, dtype=float))
        B = q.shape[0]
        R = np.tile(np.eye(3), (B, 1, 1))
        p = np.zeros((B, 3))
        frames = []
        for typ, xyz, rpy, axis, idx in self.joints:
            p = p + R @ xyz
            R = R @ rpy2r(rpy)
            if typ in ("revolute", "continuous"):
                frames.append((idx, "r", R @ axis, p.copy()))  # R(q) axis = axis, so z needs no joint rotation
                R = R @ angvec2r(q[:, idx], axis)
            elif typ == "prismatic":
                frames.append((idx, "p", R @ axis, p.copy()))
                p = p + (R @ axis) * q[:, idx:idx + 1]
        J = np.zeros((B, 3, self.ndof))
        for idx, kind, z, pj in frames:
            J[:, :, idx] = np.cross(z, p - pj) if kind == "r" else z
        return p, J

    def quaternion(self, q) -> np.ndarray:
        """Tip orientation as xyzw quaternion via the reference's quaternion chain (models.py:1049-1088)."""
        q = np.atleast_2d(np.asarray(q, dtype=float))
        quat = np.tile(np.array([0.0, 0.0, 0.0, 1.0]), (q.shape[0], 1))
        for typ, xyz, rpy, axis, idx in self.joints:
            quat = quat_mul(quat_fromrpy(rpy), quat)
            if typ in ("revolute", "continuous"):
                quat = quat_mul(quat_fromangvec(q[:, idx], axis), quat)
        return quat


_LWR = None


def lwr_chain() -> Chain:
    global _LWR
    if _LWR is None:
        _LWR = Chain(os.path.join(_ROBOTS, "kuka_lwr.urdf"), "end_effector_ball")
    return _LWR


def lwr_position_and_jacobian(q):
    """(p [B,3], J flattened column-major [B,21]) -- the layout of the `fk_jac` Function outputs."""
    p, J = lwr_chain().position_and_linear_jacobian(q)
    return p, J.transpose(0, 2, 1).reshape(J.shape[0], -1)
