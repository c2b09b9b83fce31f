"""Minimal URDF object model on the standard library's XML parser.

The reference delegates URDF handling to the third-party ``urdf_parser_py`` package
(reference: optas/models.py:15, used at :288-290, :338-354, :568-588, :838-847); that package
is absent from this image, so this module provides the small surface RobotModel needs (F15):
``URDF.from_xml_file/from_xml_string``, ``.name/.joints/.links/.joint_map/.link_map``,
``get_root()``, ``get_chain(root, tip, joints=True, links=True, fixed=True)``,
``add_link/add_joint`` and the ``Joint/Link/Pose/JointLimit`` records.
Kinematic information plus each link's ``<inertial>`` element (mass, centre-of-mass origin, inertia
tensor -- what ``RobotModel.rnea`` reads, ref :1753-1768) and a link's visual origin are kept;
collision geometry, meshes and transmissions are ignored.
"""

from __future__ import annotations

import xml.etree.ElementTree as ET
from typing import Dict, List, Optional


def _floats(text: Optional[str], n: int, default: float = 0.0) -> List[float]:
    if text is None:
        return [default] * n
    vals = [float(t) for t in text.split()]
    if len(vals) != n:
        raise ValueError(f"expected {n} numbers, got '{text}'")
    return vals


class Pose:
    def __init__(self, xyz=None, rpy=None):
        self.xyz = list(xyz) if xyz is not None else [0.0, 0.0, 0.0]
        self.rpy = list(rpy) if rpy is not None else [0.0, 0.0, 0.0]

    @staticmethod
    def from_xml(el: Optional[ET.Element]) -> Optional["Pose"]:
        if el is None:
            return None
        return Pose(_floats(el.get("xyz"), 3), _floats(el.get("rpy"), 3))


class JointLimit:
    def __init__(self, lower=0.0, upper=0.0, velocity=0.0, effort=0.0):
        self.lower, self.upper, self.velocity, self.effort = lower, upper, velocity, effort

    @staticmethod
    def from_xml(el: Optional[ET.Element]) -> Optional["JointLimit"]:
        if el is None:
            return None
        g = lambda k: float(el.get(k, 0.0))
        return JointLimit(g("lower"), g("upper"), g("velocity"), g("effort"))


class Visual:
    def __init__(self, origin: Optional[Pose] = None):
        self.origin = origin


class Inertia:
    """The six independent entries of a link's inertia tensor about its centre of mass."""

    def __init__(self, ixx=0.0, ixy=0.0, ixz=0.0, iyy=0.0, iyz=0.0, izz=0.0):
        self.ixx, self.ixy, self.ixz, self.iyy, self.iyz, self.izz = ixx, ixy, ixz, iyy, iyz, izz

    def to_matrix(self) -> List[List[float]]:
        return [[self.ixx, self.ixy, self.ixz], [self.ixy, self.iyy, self.iyz], [self.ixz, self.iyz, self.izz]]

    @staticmethod
    def from_xml(el: Optional[ET.Element]) -> "Inertia":
        if el is None:
            return Inertia()
        return Inertia(*(float(el.get(k, 0.0)) for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz")))


class Inertial:
    def __init__(self, mass: float = 0.0, inertia: Optional[Inertia] = None, origin: Optional[Pose] = None):
        self.mass = mass
        self.inertia = inertia if inertia is not None else Inertia()
        self.origin = origin if origin is not None else Pose()

    @staticmethod
    def from_xml(el: Optional[ET.Element]) -> Optional["Inertial"]:
        if el is None:
            return None
        mass = el.find("mass")
        return Inertial(float(mass.get("value", 0.0)) if mass is not None else 0.0, Inertia.from_xml(el.find("inertia")),
                        Pose.from_xml(el.find("origin")))


class Link:
    def __init__(self, name: str, visual: Optional[Visual] = None, inertial: Optional[Inertial] = None):
        self.name = name
        self.visual = visual
        self.inertial = inertial


class Joint:
    def __init__(self, name=None, parent=None, child=None, joint_type=None, axis=None, origin=None, limit=None):
        self.name = name
        self.parent = parent
        self.child = child
        self.type = joint_type
        self.axis = axis
        self.origin = origin
        self.limit = limit

    @property
    def joint_type(self):
        return self.type


class URDF:
    def __init__(self, name: str = "robot"):
        self.name = name
        self.joints: List[Joint] = []
        self.links: List[Link] = []
        self.joint_map: Dict[str, Joint] = {}
        self.link_map: Dict[str, Link] = {}
        self.parent_map: Dict[str, tuple] = {}  # child link -> (joint name, parent link)
        self.child_map: Dict[str, list] = {}  # parent link -> [(joint name, child link)]

    # -- construction -----------------------------------------------------------------------
    def add_link(self, link: Link) -> None:
        self.links.append(link)
        self.link_map[link.name] = link

    def add_joint(self, joint: Joint) -> None:
        self.joints.append(joint)
        self.joint_map[joint.name] = joint
        self.parent_map[joint.child] = (joint.name, joint.parent)
        self.child_map.setdefault(joint.parent, []).append((joint.name, joint.child))

    @staticmethod
    def from_xml_string(xml: str) -> "URDF":
        root = ET.fromstring(xml)
        if root.tag != "robot":
            raise ValueError("URDF: top-level element must be <robot>")
        out = URDF(root.get("name", "robot"))
        for el in root.findall("link"):
            vis = el.find("visual")
            visual = Visual(Pose.from_xml(vis.find("origin"))) if vis is not None else None
            out.add_link(Link(el.get("name"), visual, Inertial.from_xml(el.find("inertial"))))
        for el in root.findall("joint"):
            if el.get("type") is None:
                continue  # <joint name=.../> references inside <transmission> etc.
            ax = el.find("axis")
            jtype = el.get("type")
            axis = _floats(ax.get("xyz"), 3) if ax is not None else None
            if axis is None and jtype != "fixed":
                axis = [1.0, 0.0, 0.0]  # URDF default
            out.add_joint(
                Joint(
                    name=el.get("name"),
                    parent=el.find("parent").get("link"),
                    child=el.find("child").get("link"),
                    joint_type=jtype,
                    axis=axis,
                    origin=Pose.from_xml(el.find("origin")),
                    limit=JointLimit.from_xml(el.find("limit")),
                )
            )
        return out

    @staticmethod
    def from_xml_file(filename: str) -> "URDF":
        with open(filename, "r") as f:
            return URDF.from_xml_string(f.read())

    # -- queries ------------------------------------------------------------------------------
    def get_root(self) -> str:
        roots = [l.name for l in self.links if l.name not in self.parent_map]
        if len(roots) != 1:
            raise ValueError(f"URDF must have exactly one root link, found {roots}")
        return roots[0]

    def get_chain(self, root: str, tip: str, joints: bool = True, links: bool = True, fixed: bool = True) -> List[str]:
        chain = []
        if links:
            chain.append(tip)
        link = tip
        while link != root:
            if link not in self.parent_map:
                raise KeyError(f"link '{tip}' is not a descendant of '{root}'")
            joint, parent = self.parent_map[link]
            if joints and (fixed or self.joint_map[joint].type != "fixed"):
                chain.append(joint)
            if links:
                chain.append(parent)
            link = parent
        chain.reverse()
        return chain
