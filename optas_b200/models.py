"""Task and robot models: state naming, limits and URDF kinematics as expression graphs.

Public surface mirrors the reference's optas/models.py (Model :79-186, TaskModel :189-214,
RobotModel :233-1729) so problem-building scripts read the same:

* naming scheme ``{name}/{d*}{symbol}``, ``.../x`` (optimised), ``.../p`` (parameter)
  (ref :130-161),
* forward kinematics as the ordered product over ``urdf.get_chain(root, link)`` of the fixed
  joint-origin transform and, for actuated joints, the joint motion (ref :826-868),
* relative frames as ``T_link_world @ inv(T_base_world)`` (ref :884-898),
* quaternion chain ``fromrpy(rpy) * quat`` / ``fromangvec(q_i, axis) * quat`` (ref :1049-1088),
* geometric Jacobian columns ``[z x (e - p); z]`` for revolute and ``[z; 0]`` for prismatic joints
  (ref :1199-1264).

The whole ``get_[global_]link_<quantity>[_function]`` family is generated from one table instead
of being spelled out method by method.  Inverse dynamics: ``rnea`` (ref :1731-1884) as one
outward / inward sweep over per-body records.
"""

from __future__ import annotations

import functools
import os
import pathlib
from typing import Callable, Dict, List, Optional, Tuple, Union

import numpy as np

from . import sym as cs
from .spatialmath import (ArrayType, CasADiArrayType, I3, I4, Quaternion, angvec2r, arrayify_args,
                          invt, rpy2r, rt2tr, t2r, transl, unit, vec)
from .sym import DM, SX
from .urdf import URDF, Joint, Link, Pose


class Model:
    """Base model: a named state of dimension ``dim`` with optional time derivatives."""

    def __init__(self, name: str, dim: int, time_derivs: List[int], symbol: str,
                 dlim: Dict[int, Tuple[List[float]]], T: Union[None, int]):
        self.name = name
        self.dim = dim
        self.time_derivs = time_derivs
        self.symbol = symbol
        self.dlim = dlim
        self.T = T  # stored but (as in the reference, builder.py:89-99) not used by the builder

    def get_name(self) -> str:
        return self.name

    def _check_deriv(self, time_deriv: int) -> None:
        assert time_deriv in self.time_derivs, (
            f"Given time derivative time_deriv={time_deriv} is not recognized, only allowed {self.time_derivs}")

    def state_name(self, time_deriv: int) -> str:
        self._check_deriv(time_deriv)
        return f"{self.name}/{'d' * time_deriv}{self.symbol}"

    def state_parameter_name(self, time_deriv: int) -> str:
        return self.state_name(time_deriv) + "/p"

    def state_optimized_name(self, time_deriv: int) -> str:
        return self.state_name(time_deriv) + "/x"

    def get_limits(self, time_deriv: int):
        self._check_deriv(time_deriv)
        assert time_deriv in self.dlim.keys(), f"Limit for time derivative time_deriv={time_deriv} has not been given"
        return self.dlim[time_deriv]

    def in_limit(self, x: ArrayType, time_deriv: int) -> DM:
        lo, up = self.get_limits(time_deriv)
        x = x if isinstance(x, (DM, SX)) else DM(x)
        return cs.logic_all(cs.logic_and(DM(lo) <= x, x <= DM(up)))


class TaskModel(Model):
    def __init__(self, name: str, dim: int, time_derivs: List[int] = [0], symbol: str = "y",
                 dlim: Dict[int, Tuple[List[float]]] = {}, T: Union[None, int] = None,
                 is_discrete: bool = False):
        super().__init__(name, dim, time_derivs, symbol, dlim, T)
        self.is_discrete = is_discrete


class JointTypeNotSupported(NotImplementedError):
    def __init__(self, joint_type: str):
        super().__init__(
            f"{joint_type} joints are currently not supported\n"
            "if you require this joint type please raise an issue at https://github.com/cmower/optas/issues")


def _over_columns(fun: Callable) -> Callable:
    """If the joint state argument is a trajectory (n columns > 1) evaluate per column: vector
    results are stacked side by side, matrix results are returned as a list (ref :19-52)."""

    @functools.wraps(fun)
    def wrapper(self, link, q, *args, **kwargs):
        if q.shape[1] > 1:
            outs = [fun(self, link, q[:, i], *args, **kwargs) for i in range(q.shape[1])]
            return cs.horzcat(*outs) if outs[0].shape[1] == 1 else outs
        return fun(self, link, q, *args, **kwargs)

    return wrapper


class _ListFunction:
    """Function over a trajectory whose per-column result is a matrix: returns a list."""

    def __init__(self, F: cs.Function, n: int):
        self.F, self.n = F, n

    def __call__(self, Q):
        Q = Q if isinstance(Q, (DM, SX)) else DM(Q)
        assert Q.shape[1] == self.n, f"expected input with {self.n} columns, got {Q.shape[0]}-by-{Q.shape[1]}"
        return [self.F(q) for q in cs.horzsplit(Q)]

    def __getattr__(self, item):  # size_in/size_out/numel_in/... forward to the column function
        return getattr(self.F, item)


class _NumpyOutput:
    def __init__(self, F):
        self.F = F

    @staticmethod
    def _np(out):
        a = out.toarray()
        return a.flatten() if a.shape[1] == 1 else a

    def __call__(self, q):
        assert not isinstance(q, SX), "numpy_output=True was specified, you can not pass symbolic variables"
        out = self.F(q)
        return [self._np(o) for o in out] if isinstance(out, list) else self._np(out)


class RobotModel(Model):
    """Robot model loaded from a URDF (file or string) or a xacro file."""

    def __init__(self, urdf_filename: Union[None, str] = None, urdf_string: Union[None, str] = None,
                 xacro_filename: Union[None, str] = None, name: Union[None, str] = None,
                 time_derivs: List[int] = [0], qddlim: Union[None, ArrayType] = None,
                 T: Union[None, int] = None, param_joints: List[str] = []):
        self.xacro_filename = xacro_filename
        if xacro_filename is not None:
            urdf_string = self._expand_xacro(xacro_filename)

        self.urdf = None
        self.urdf_filename = None
        self.urdf_string = None
        if urdf_filename is not None:
            self.urdf_filename = urdf_filename
            self.urdf = URDF.from_xml_file(urdf_filename)
        if urdf_string is not None:
            self.urdf_string = urdf_string
            self.urdf = URDF.from_xml_string(urdf_string)
        assert self.urdf is not None, "You need to supply a urdf, either through filename or as a string"

        self.param_joints = param_joints
        dlim = {
            0: (self.lower_optimized_joint_limits, self.upper_optimized_joint_limits),
            1: (-self.velocity_optimized_joint_limits, self.velocity_optimized_joint_limits),
        }
        if qddlim:
            qddlim = vec(qddlim if isinstance(qddlim, (DM, SX)) else DM(qddlim))
            if qddlim.shape[0] == 1:
                qddlim = qddlim * DM.ones(self.ndof)
            assert qddlim.shape[0] == self.ndof, f"expected ddlim to have {self.ndof} elements"
            dlim[2] = -qddlim, qddlim

        if name is None:
            name = self.urdf.name
        super().__init__(name, self.ndof, time_derivs, "q", dlim, T)

    @staticmethod
    def _expand_xacro(xacro_filename: str) -> str:
        try:
            import xacro  # noqa: F401  (not available in the build image)
        except ImportError:
            # A pre-expanded twin "<file>.urdf" next to "<file>.urdf.xacro" is accepted instead.
            twin = xacro_filename[: -len(".xacro")] if xacro_filename.endswith(".xacro") else None
            if twin is not None and os.path.exists(twin):
                with open(twin, "r") as f:
                    return f.read()
            raise ImportError(
                "the 'xacro' package is not installed and no pre-expanded URDF was found next to "
                f"'{xacro_filename}'")
        try:
            return xacro.process(xacro_filename)
        except AttributeError:
            from io import StringIO

            buf = StringIO()
            xacro.process_file(xacro_filename).writexml(buf)
            return buf.getvalue()

    # -- bookkeeping ---------------------------------------------------------------------------
    def get_urdf(self):
        return self.urdf

    def get_urdf_dirname(self):
        fn = self.urdf_filename or self.xacro_filename
        return pathlib.Path(os.path.dirname(fn)) if fn is not None else None

    @property
    def joint_names(self) -> List[str]:
        return [j.name for j in self.urdf.joints]

    @property
    def link_names(self) -> List[str]:
        return [l.name for l in self.urdf.links]

    @property
    def actuated_joint_names(self) -> List[str]:
        return [j.name for j in self.urdf.joints if j.type != "fixed"]

    @property
    def parameter_joint_names(self) -> List[str]:
        return [j for j in self.actuated_joint_names if j in self.param_joints]

    @property
    def optimized_joint_names(self) -> List[str]:
        params = set(self.parameter_joint_names)
        return [j for j in self.actuated_joint_names if j not in params]

    @property
    def optimized_joint_indexes(self) -> List[int]:
        return [self.get_actuated_joint_index(j) for j in self.optimized_joint_names]

    @property
    def parameter_joint_indexes(self) -> List[int]:
        return [self.get_actuated_joint_index(j) for j in self.parameter_joint_names]

    def extract_parameter_dimensions(self, values):
        return values[self.parameter_joint_indexes, :]

    def extract_optimized_dimensions(self, values):
        return values[self.optimized_joint_indexes, :]

    @property
    def ndof(self) -> int:
        return len(self.actuated_joint_names)

    @property
    def num_opt_joints(self) -> int:
        return len(self.optimized_joint_names)

    @property
    def num_param_joints(self) -> int:
        return len(self.parameter_joint_names)

    # limits: a joint without <limit> is unbounded, encoded as -/+1e9 (ref :444-466)
    @staticmethod
    def get_joint_lower_limit(joint: Joint) -> float:
        return -1e9 if joint.limit is None else joint.limit.lower

    @staticmethod
    def get_joint_upper_limit(joint: Joint) -> float:
        return 1e9 if joint.limit is None else joint.limit.upper

    @staticmethod
    def get_velocity_joint_limit(joint: Joint) -> float:
        return 1e9 if joint.limit is None else joint.limit.velocity

    def _limits(self, getter, names) -> DM:
        names = set(names)
        vals = [getter(j) for j in self.urdf.joints if j.name in names]
        return DM(vals) if vals else DM.zeros(0, 1)

    @property
    def lower_actuated_joint_limits(self) -> DM:
        return self._limits(self.get_joint_lower_limit, self.actuated_joint_names)

    @property
    def upper_actuated_joint_limits(self) -> DM:
        return self._limits(self.get_joint_upper_limit, self.actuated_joint_names)

    @property
    def velocity_actuated_joint_limits(self) -> DM:
        return self._limits(self.get_velocity_joint_limit, self.actuated_joint_names)

    @property
    def lower_optimized_joint_limits(self) -> DM:
        return self._limits(self.get_joint_lower_limit, self.optimized_joint_names)

    @property
    def upper_optimized_joint_limits(self) -> DM:
        return self._limits(self.get_joint_upper_limit, self.optimized_joint_names)

    @property
    def velocity_optimized_joint_limits(self) -> DM:
        return self._limits(self.get_velocity_joint_limit, self.optimized_joint_names)

    def add_base_frame(self, base_link: str, xyz=None, rpy=None, joint_name: str = None) -> None:
        """Insert a new root link connected to the current root by a fixed joint (ref :552-588)."""
        child = self.urdf.get_root()
        if not isinstance(joint_name, str):
            joint_name = f"{base_link}_and_{child}_joint"
        self.urdf.add_link(Link(name=base_link))
        self.urdf.add_joint(Joint(name=joint_name, parent=base_link, child=child, joint_type="fixed",
                                  origin=Pose(xyz=xyz or [0.0] * 3, rpy=rpy or [0.0] * 3)))

    def get_root_link(self) -> str:
        return self.urdf.get_root()

    @staticmethod
    def get_link_visual_origin(link: Link):
        if link.visual is not None and link.visual.origin is not None:
            return DM(link.visual.origin.xyz), DM(link.visual.origin.rpy)
        return DM.zeros(3), DM.zeros(3)

    @staticmethod
    def get_joint_origin(joint: Joint):
        if joint.origin is not None:
            return DM(joint.origin.xyz), DM(joint.origin.rpy)
        return DM.zeros(3), DM.zeros(3)

    @staticmethod
    def get_joint_axis(joint: Joint) -> DM:
        return unit(DM(joint.axis) if joint.axis is not None else DM([1.0, 0.0, 0.0]))

    def get_actuated_joint_index(self, joint_name: str) -> int:
        return self.actuated_joint_names.index(joint_name)

    def get_random_joint_positions(self, n: int = 1, xlim=None, ylim=None, zlim=None, base_link=None) -> DM:
        lo = self.lower_actuated_joint_limits.toarray()
        hi = self.upper_actuated_joint_limits.toarray()
        boxes = [b for b in (xlim, ylim, zlim)]
        pos = None
        if isinstance(base_link, str) and any(b is not None for b in boxes):
            pos = [self.get_link_position_function(link, base_link) for link in self.link_names]

        def ok(q):
            if pos is None:
                return True
            for p in pos:
                pp = p(q).toarray().flatten()
                for k, b in enumerate(boxes):
                    if b is not None and not (b[0] <= pp[k] <= b[1]):
                        return False
            return True

        cols = []
        while len(cols) < n:
            q = DM(np.random.uniform(lo, hi))
            if ok(q):
                cols.append(q)
        return cs.horzcat(*cols)

    def get_random_pose_in_global_link(self, link_name: str) -> DM:
        return self.get_global_link_transform(link_name, self.get_random_joint_positions())

    # -- kinematics ------------------------------------------------------------------------------
    def _chain(self, link: str) -> List[Joint]:
        assert link in self.urdf.link_map, f"given link '{link}' does not appear in URDF"
        root = self.urdf.get_root()
        return [self.urdf.joint_map[j] for j in self.urdf.get_chain(root, link, links=False)]

    def _joint_value(self, joint: Joint, q):
        return q[self.get_actuated_joint_index(joint.name)]

    @arrayify_args
    @_over_columns
    def get_global_link_transform(self, link: str, q: ArrayType) -> CasADiArrayType:
        T = I4()
        for joint in self._chain(link):
            xyz, rpy = self.get_joint_origin(joint)
            T = T @ rt2tr(rpy2r(rpy), xyz)
            if joint.type == "fixed":
                continue
            qi = self._joint_value(joint, q)
            if joint.type in ("revolute", "continuous"):
                T = T @ cs.vertcat(cs.horzcat(angvec2r(qi, self.get_joint_axis(joint)), DM.zeros(3, 1)),
                                   DM([[0.0, 0.0, 0.0, 1.0]]))
            elif joint.type == "prismatic":
                T = T @ rt2tr(I3(), qi * self.get_joint_axis(joint))
            else:
                raise JointTypeNotSupported(joint.type)
        return T

    @arrayify_args
    @_over_columns
    def get_link_transform(self, link: str, q: ArrayType, base_link: str) -> CasADiArrayType:
        return self.get_global_link_transform(link, q) @ invt(self.get_global_link_transform(base_link, q))

    @arrayify_args
    @_over_columns
    def get_global_link_position(self, link: str, q: ArrayType) -> CasADiArrayType:
        return transl(self.get_global_link_transform(link, q))

    @arrayify_args
    @_over_columns
    def get_link_position(self, link: str, q: ArrayType, base_link: str) -> CasADiArrayType:
        return transl(self.get_link_transform(link, q, base_link))

    @arrayify_args
    @_over_columns
    def get_global_link_rotation(self, link: str, q: ArrayType) -> CasADiArrayType:
        return t2r(self.get_global_link_transform(link, q))

    @arrayify_args
    @_over_columns
    def get_link_rotation(self, link: str, q: ArrayType, base_link: str) -> CasADiArrayType:
        return t2r(self.get_link_transform(link, q, base_link))

    @arrayify_args
    @_over_columns
    def get_global_link_quaternion(self, link: str, q: ArrayType) -> CasADiArrayType:
        quat = Quaternion(0.0, 0.0, 0.0, 1.0)
        for joint in self._chain(link):
            _, rpy = self.get_joint_origin(joint)
            quat = Quaternion.fromrpy(rpy) * quat
            if joint.type == "fixed":
                continue
            if joint.type in ("revolute", "continuous"):
                quat = Quaternion.fromangvec(self._joint_value(joint, q), self.get_joint_axis(joint)) * quat
            elif joint.type != "prismatic":
                raise JointTypeNotSupported(joint.type)
        return quat.getquat()

    @arrayify_args
    @_over_columns
    def get_link_quaternion(self, link: str, q: ArrayType, base_link: str) -> CasADiArrayType:
        ql = Quaternion.fromvec(self.get_global_link_quaternion(link, q))
        qb = Quaternion.fromvec(self.get_global_link_quaternion(base_link, q))
        return (ql * qb.inv()).getquat()

    @arrayify_args
    @_over_columns
    def get_global_link_rpy(self, link: str, q: ArrayType) -> CasADiArrayType:
        return Quaternion.fromvec(self.get_global_link_quaternion(link, q)).getrpy()

    @arrayify_args
    @_over_columns
    def get_link_rpy(self, link: str, q: ArrayType, base_link: str) -> CasADiArrayType:
        return Quaternion.fromvec(self.get_link_quaternion(link, q, base_link)).getrpy()

    @arrayify_args
    @_over_columns
    def get_global_link_geometric_jacobian(self, link: str, q: ArrayType) -> CasADiArrayType:
        e = self.get_global_link_position(link, q)
        in_chain = {j.name for j in self._chain(link)}
        cols: Dict[int, CasADiArrayType] = {}
        for joint in self.urdf.joints:
            if joint.type == "fixed":
                continue
            k = self.get_actuated_joint_index(joint.name)
            if joint.name not in in_chain:
                cols[k] = DM.zeros(6, 1)
                continue
            axis = self.get_joint_axis(joint)
            # orientation / origin of the joint's child frame; rotating about the axis leaves the
            # axis itself unchanged so R(child) @ axis is the world-frame joint axis
            z = self.get_global_link_rotation(joint.child, q) @ axis
            if joint.type in ("revolute", "continuous"):
                p = self.get_global_link_position(joint.child, q)
                cols[k] = cs.vertcat(cs.cross(z, e - p), z)
            elif joint.type == "prismatic":
                cols[k] = cs.vertcat(z, DM.zeros(3, 1))
            else:
                raise JointTypeNotSupported(joint.type)
        return cs.horzcat(*[cols[k] for k in range(self.ndof)])

    @arrayify_args
    @_over_columns
    def get_link_geometric_jacobian(self, link: str, q: ArrayType, base_link: str) -> CasADiArrayType:
        J = self.get_global_link_geometric_jacobian(link, q)
        Rt = self.get_global_link_rotation(base_link, q).T
        return cs.vertcat(Rt @ J[:3, :], Rt @ J[3:, :])

    @arrayify_args
    @_over_columns
    def get_global_link_linear_jacobian(self, link: str, q: ArrayType) -> CasADiArrayType:
        return self.get_global_link_geometric_jacobian(link, q)[:3, :]

    @arrayify_args
    @_over_columns
    def get_link_linear_jacobian(self, link: str, q: ArrayType, base_link: str) -> CasADiArrayType:
        return self.get_link_geometric_jacobian(link, q, base_link)[:3, :]

    @arrayify_args
    @_over_columns
    def get_global_link_angular_geometric_jacobian(self, link: str, q: ArrayType) -> CasADiArrayType:
        return self.get_global_link_geometric_jacobian(link, q)[3:, :]

    @arrayify_args
    @_over_columns
    def get_link_angular_geometric_jacobian(self, link: str, q: ArrayType, base_link: str) -> CasADiArrayType:
        return self.get_link_geometric_jacobian(link, q, base_link)[3:, :]

    @arrayify_args
    @_over_columns
    def get_link_angular_analytical_jacobian(self, link: str, q: ArrayType, base_link: str) -> CasADiArrayType:
        qs = SX.sym("q_sym", self.ndof)
        Ja = cs.Function("Ja", [qs], [cs.jacobian(self.get_link_rpy(link, qs, base_link), qs)])
        return Ja(q)

    @arrayify_args
    @_over_columns
    def get_global_link_angular_analytical_jacobian(self, link: str, q: ArrayType) -> CasADiArrayType:
        return self.get_link_angular_analytical_jacobian(link, q, self.get_root_link())

    @arrayify_args
    @_over_columns
    def get_global_link_analytical_jacobian(self, link: str, q: ArrayType) -> CasADiArrayType:
        return cs.vertcat(self.get_global_link_linear_jacobian(link, q),
                          self.get_global_link_angular_analytical_jacobian(link, q))

    @arrayify_args
    @_over_columns
    def get_link_analytical_jacobian(self, link: str, q: ArrayType, base_link: str) -> CasADiArrayType:
        return cs.vertcat(self.get_link_linear_jacobian(link, q, base_link),
                          self.get_link_angular_analytical_jacobian(link, q, base_link))

    @_over_columns
    def _link_axis(self, link, q, axis, base_link):
        Tf = self.get_link_transform(link, q, base_link)
        if isinstance(axis, str):
            assert axis in ("x", "y", "z"), "axis must be either 'x', 'y', 'z' or a 3-array"
            return Tf[:3, "xyz".index(axis)]
        if isinstance(axis, (DM, SX)):
            a = unit(vec(axis))
            return a[0] * Tf[:3, 0] + a[1] * Tf[:3, 1] + a[2] * Tf[:3, 2]
        raise ValueError(f"did not recognize input for axis: {axis}")

    @arrayify_args
    def get_link_axis(self, link: str, q: ArrayType, axis: Union[str, ArrayType], base_link: str):
        return self._link_axis(link, q, axis, base_link)

    @arrayify_args
    def get_global_link_axis(self, link: str, q: ArrayType, axis: Union[str, ArrayType]):
        return self._link_axis(link, q, axis, self.get_root_link())

    # -- function factories ------------------------------------------------------------------------
    def make_function(self, label: str, link: str, method: Callable, n: int = 1,
                      base_link: Union[None, str] = None, axis=None, numpy_output: bool = False):
        """Wrap one of the kinematic methods as a Function of q (``n`` > 1: over a trajectory;
        vector outputs are mapped column-wise, matrix outputs give a list) (ref :729-824)."""
        q = SX.sym("q", self.ndof)
        args = [link, q]
        if axis is not None:
            args.append(axis)
        kwargs = {} if base_link is None else {"base_link": base_link}
        out = method(*args, **kwargs)
        F = cs.Function(label, [q], [out])
        if n > 1:
            F = F.map(n) if out.shape[1] == 1 else _ListFunction(F, n)
        return _NumpyOutput(F) if numpy_output else F

    def get_link_axis_function(self, link, axis, base_link, n=1, numpy_output=False):
        return self.make_function("a", link, self.get_link_axis, n=n, base_link=base_link, axis=axis,
                                  numpy_output=numpy_output)

    def get_global_link_axis_function(self, link, axis, n=1, numpy_output=False):
        return self.make_function("a", link, functools.partial(self.get_global_link_axis, axis=axis), n=n,
                                  numpy_output=numpy_output)

    # -- inverse dynamics ----------------------------------------------------------------------------
    def _rnea_bodies(self):
        """Per-body records for ``rnea``: (xyz, rpy, axis of the joint the body hangs on, mass, centre of mass,
        inertia tensor), base to tip.  As in the reference (:1743-1791): serial chain root -> last link, first joint
        fixed (its body is the base and carries no dynamics), every link with an <inertial> element after the first
        is a body, the inertial origin's rpy is not applied."""
        for joint in self.urdf.joint_map.values():
            if joint.type not in ("revolute", "continuous", "fixed"):
                raise JointTypeNotSupported(joint.type)
        if next(iter(self.urdf.joint_map.values())).type != "fixed":
            raise JointTypeNotSupported("First joint should be fixed")
        inert = [link.inertial for link in self.urdf.links if link.inertial is not None][1:]
        chain = self.urdf.get_chain(self.urdf.get_root(), self.link_names[-1], links=False)[1:]
        bodies = []
        for k, joint_name in enumerate(chain):
            joint = self.urdf.joint_map[joint_name]
            xyz, rpy = self.get_joint_origin(joint)
            bodies.append(dict(r=xyz, R0=rpy2r(rpy), axis=self.get_joint_axis(joint), m=float(inert[k].mass),
                               c=DM(inert[k].origin.xyz), I=DM(inert[k].inertia.to_matrix())))
        return bodies

    @arrayify_args
    def rnea(self, q: ArrayType, qd: ArrayType, qdd: ArrayType) -> CasADiArrayType:
        """Joint torques for (q, qd, qdd) by the recursive Newton-Euler algorithm (Craig, Introduction to
        Robotics, ch. 6; ref :1731-1884).  Revolute / continuous joints only; the chain's last joint is the fixed
        tool joint; gravity is -9.81 along the base z axis, entering as an upward base acceleration."""
        cross = lambda a, b: cs.cross(a, b)
        bodies = self._rnea_bodies()
        n = len(bodies)
        # parent -> child rotation of every joint (the tool joint does not move)
        R = [b["R0"] @ angvec2r(q[k], b["axis"]) if k < n - 1 else b["R0"] for k, b in enumerate(bodies)]
        # outward sweep: velocities / accelerations of each body frame, then the inertial force and moment
        w, dw, dv = DM.zeros(3), DM.zeros(3), DM([0.0, 0.0, 9.81])
        F, N = [], []
        for k, b in enumerate(bodies):
            E = R[k].T
            w_in = E @ w
            dv = E @ (dv + cross(dw, b["r"]) + cross(w, cross(w, b["r"])))
            if k < n - 1:
                z = E @ b["axis"]
                dw = E @ dw + cross(w_in, z * qd[k]) + z * qdd[k]
                w = w_in + z * qd[k]
            else:
                w, dw = w_in, E @ dw
            F.append(b["m"] * (dv + cross(dw, b["c"]) + cross(w, cross(w, b["c"]))))
            N.append(b["I"] @ dw + cross(w, b["I"] @ w))
        # inward sweep: force / moment each body passes to its parent, projected on the joint axis
        f = F[n - 1]
        m = N[n - 1] + cross(bodies[n - 1]["c"], F[n - 1])
        tau = [None] * (n - 1)
        for k in range(n - 2, -1, -1):
            Rf = R[k + 1] @ f
            m = N[k] + R[k + 1] @ m + cross(bodies[k]["c"], F[k]) + cross(bodies[k + 1]["r"], Rf)
            f = Rf + F[k]
            tau[k] = m.T @ (R[k].T @ bodies[k]["axis"])
        return cs.vertcat(*tau)


def _install_function_factories():
    """``get_global_link_X_function(link, n, numpy_output)`` / ``get_link_X_function(link, base_link,
    n, numpy_output)`` for every kinematic quantity X."""
    quantities = {
        "transform": "T", "position": "p", "rotation": "R", "quaternion": "quat", "rpy": "rpy",
        "geometric_jacobian": "J", "analytical_jacobian": "J_a", "linear_jacobian": "J_l",
        "angular_geometric_jacobian": "J_ag", "angular_analytical_jacobian": "J_aa",
    }
    for quantity, label in quantities.items():

        def global_factory(self, link, n=1, numpy_output=False, _q=quantity, _l=label):
            return self.make_function(_l, link, getattr(self, f"get_global_link_{_q}"), n=n,
                                      numpy_output=numpy_output)

        def local_factory(self, link, base_link, n=1, numpy_output=False, _q=quantity, _l=label):
            return self.make_function(_l, link, getattr(self, f"get_link_{_q}"), n=n, base_link=base_link,
                                      numpy_output=numpy_output)

        global_factory.__name__ = f"get_global_link_{quantity}_function"
        local_factory.__name__ = f"get_link_{quantity}_function"
        setattr(RobotModel, global_factory.__name__, global_factory)
        setattr(RobotModel, local_factory.__name__, local_factory)


_install_function_factories()
