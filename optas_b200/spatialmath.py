"""Spatial maths helpers over `optas_b200.sym` arrays (SX = symbolic, DM = numeric).

Mirrors the public surface of the reference's optas/spatialmath.py (rotations :89-199,
homogeneous transforms :102-264, ``unit`` :267-274, ``Quaternion`` :277-437) so that model and
example code reads the same.  Conventions that matter for parity (all pinned by the reference's
tests/test_spatialmath.py, replayed in tests/test_spatialmath.py here):

* ``rpy2r`` default order ``"zyx"`` means ``Rz(yaw) @ Ry(pitch) @ Rx(roll)`` (ref :160-185).
* quaternions are stored ``xyzw``; ``q0 * q1`` composes like ``Rot(q1) @ Rot(q0)`` (ref :298-312).
* list / tuple / 1-D ndarray arguments become COLUMN vectors (ref ``arrayify_args`` :21-70).
"""

import functools
import inspect
from typing import Callable, List, Tuple, Union

import numpy as np

from . import sym as cs
from .sym import DM, SX, cos, sin, vec

ArrayType = Union[DM, SX, List[float], Tuple[float], np.ndarray, float, int]
CasADiArrayType = Union[DM, SX]

pi = np.pi
eps = np.finfo(float).eps

_ARRAYLIKE = (DM, SX, list, tuple, np.ndarray, float, int, np.floating, np.integer)


def _arr(a):
    return a if isinstance(a, (DM, SX)) else DM(a)


def arrayify_args(fun: Callable) -> Callable:
    """Decorator: every array-like positional / keyword argument (and array-like default) is
    turned into an SX/DM before ``fun`` runs."""
    sig = inspect.signature(fun)
    defaults = {
        k: v.default
        for k, v in sig.parameters.items()
        if v.default is not inspect.Parameter.empty and isinstance(v.default, _ARRAYLIKE) and not isinstance(v.default, bool)
    }

    @functools.wraps(fun)
    def wrap(*args, **kwargs):
        args_use = [_arr(a) if isinstance(a, _ARRAYLIKE) and not isinstance(a, bool) else a for a in args]
        kwargs_use = {
            k: (_arr(v) if isinstance(v, _ARRAYLIKE) and not isinstance(v, bool) else v) for k, v in kwargs.items()
        }
        bound = sig.bind_partial(*args_use, **kwargs_use).arguments
        for k, v in defaults.items():
            if k not in bound:
                kwargs_use[k] = _arr(v)
        return fun(*args_use, **kwargs_use)

    return wrap


def I3() -> DM:
    return DM.eye(3)


def I4() -> DM:
    return DM.eye(4)


@arrayify_args
def skew(v: ArrayType) -> CasADiArrayType:
    """Skew-symmetric matrix of a scalar (2x2) or a 3-vector (3x3)."""
    if v.numel() == 1:
        return cs.vertcat(cs.horzcat(0.0, -v), cs.horzcat(v, 0.0))
    if v.numel() == 3:
        v = vec(v)
        x, y, z = v[0], v[1], v[2]
        return cs.vertcat(cs.horzcat(0.0, -z, y), cs.horzcat(z, 0.0, -x), cs.horzcat(-y, x, 0.0))
    raise ValueError("expecting a scalar or 3-vector")


@arrayify_args
def unit(v: ArrayType) -> CasADiArrayType:
    return v / cs.norm_fro(v)


@arrayify_args
def angvec2r(theta: ArrayType, v: ArrayType) -> CasADiArrayType:
    """Rodrigues' formula: rotation by ``theta`` about direction ``v``."""
    K = skew(unit(v))
    return I3() + sin(theta) * K + (1.0 - cos(theta)) * (K @ K)


@arrayify_args
def r2t(R: ArrayType) -> CasADiArrayType:
    return cs.vertcat(cs.horzcat(R, DM.zeros(3, 1)), DM([[0.0, 0.0, 0.0, 1.0]]))


@arrayify_args
def rotx(theta: ArrayType) -> CasADiArrayType:
    c, s = cos(theta), sin(theta)
    return cs.vertcat(DM([[1.0, 0.0, 0.0]]), cs.horzcat(0.0, c, -s), cs.horzcat(0.0, s, c))


@arrayify_args
def roty(theta: ArrayType) -> CasADiArrayType:
    c, s = cos(theta), sin(theta)
    return cs.vertcat(cs.horzcat(c, 0.0, s), DM([[0.0, 1.0, 0.0]]), cs.horzcat(-s, 0.0, c))


@arrayify_args
def rotz(theta: ArrayType) -> CasADiArrayType:
    c, s = cos(theta), sin(theta)
    return cs.vertcat(cs.horzcat(c, -s, 0.0), cs.horzcat(s, c, 0.0), DM([[0.0, 0.0, 1.0]]))


@arrayify_args
def rpy2r(rpy: ArrayType, opt: str = "zyx") -> CasADiArrayType:
    """Roll-pitch-yaw to SO(3).  ``opt``: 'zyx'/'vehicle', 'xyz'/'arm', 'yxz'/'camera'."""
    rpy = vec(rpy)
    r, p, y = rpy[0], rpy[1], rpy[2]
    if opt in ("xyz", "arm"):
        return rotx(y) @ roty(p) @ rotz(r)
    if opt in ("zyx", "vehicle"):
        return rotz(y) @ roty(p) @ rotx(r)
    if opt in ("yxz", "camera"):
        return roty(y) @ rotx(p) @ rotz(r)
    raise ValueError(
        f"didn't recognize given option {opt}, only allowed ['zyx', 'xyz', 'yxz', 'arm', 'vehicle', 'camera']")


@arrayify_args
def rt2tr(R: ArrayType, t: ArrayType) -> CasADiArrayType:
    return cs.vertcat(cs.horzcat(R, vec(t)), DM([[0.0, 0.0, 0.0, 1.0]]))


@arrayify_args
def t2r(T: ArrayType) -> CasADiArrayType:
    return T[:3, :3]


@arrayify_args
def transl(T: ArrayType) -> CasADiArrayType:
    return T[:3, 3]


@arrayify_args
def invt(T: ArrayType) -> CasADiArrayType:
    Rt = t2r(T).T
    return rt2tr(Rt, -(Rt @ transl(T)))


class Quaternion:
    """Quaternion in xyzw storage."""

    def __init__(self, x: ArrayType, y: ArrayType, z: ArrayType, w: ArrayType):
        self._q = cs.vertcat(x, y, z, w)

    def split(self):
        return cs.vertsplit(self._q)

    def __mul__(self, quat):
        assert isinstance(quat, Quaternion), "unsupported type"
        x0, y0, z0, w0 = self.split()
        x1, y1, z1, w1 = quat.split()
        return Quaternion(
            x1 * w0 + y1 * z0 - z1 * y0 + w1 * x0,
            -x1 * z0 + y1 * w0 + z1 * x0 + w1 * y0,
            x1 * y0 - y1 * x0 + z1 * w0 + w1 * z0,
            -x1 * x0 - y1 * y0 - z1 * z0 + w1 * w0,
        )

    def sumsqr(self):
        return cs.sumsqr(self._q)

    def inv(self):
        q = self._q
        qinv = cs.vertcat(-q[:3], q[3]) / self.sumsqr()
        return Quaternion(qinv[0], qinv[1], qinv[2], qinv[3])

    @staticmethod
    def fromrpy(rpy: ArrayType):
        rpy = vec(_arr(rpy))
        r, p, y = rpy[0], rpy[1], rpy[2]
        cr, sr = cos(0.5 * r), sin(0.5 * r)
        cp, sp = cos(0.5 * p), sin(0.5 * p)
        cy, sy = cos(0.5 * y), sin(0.5 * y)
        qx = sr * cp * cy - cr * sp * sy
        qy = cr * sp * cy + sr * cp * sy
        qz = cr * cp * sy - sr * sp * cy
        qw = cr * cp * cy + sr * sp * sy
        n = cs.sqrt(qx * qx + qy * qy + qz * qz + qw * qw)
        return Quaternion(qx / n, qy / n, qz / n, qw / n)

    @staticmethod
    def fromvec(q: ArrayType):
        q = _arr(q)
        return Quaternion(q[0], q[1], q[2], q[3])

    @staticmethod
    def fromangvec(theta: ArrayType, v: ArrayType):
        theta = _arr(theta)
        w = cos(0.5 * theta)
        xyz = sin(0.5 * theta) * unit(vec(_arr(v)))
        return Quaternion(xyz[0], xyz[1], xyz[2], w)

    def getquat(self):
        return self._q

    def getrpy(self):
        qx, qy, qz, qw = self.split()
        roll = cs.atan2(2.0 * (qw * qx + qy * qz), 1.0 - 2.0 * (qx * qx + qy * qy))
        sinp = 2.0 * (qw * qy - qz * qx)
        pitch = cs.if_else(cs.fabs(sinp) >= 1.0, pi / 2.0, cs.asin(sinp))
        yaw = cs.atan2(2.0 * (qw * qz + qx * qy), 1.0 - 2.0 * (qy * qy + qz * qz))
        return cs.vertcat(roll, pitch, yaw)

    def getrotm(self):
        """Reproduces the reference's formula verbatim in behaviour, including its known
        off-diagonal defects (ref spatialmath.py:426-437); only reached by the
        torque-control examples, which are out of scope."""
        x, y, z, w = self.split()
        return cs.vertcat(
            cs.horzcat(1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * x * y + 2 * w * y),
            cs.horzcat(2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * z),
            cs.horzcat(2 * x * z - 2 * w * y, 2 * y * z + w * w * x, 1 - 2 * x * x - 2 * y * y),
        )
