"""Derive kinematics-only robot descriptions from the reference's URDF files.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_robot_assets.py

For each robot the config scripts use (KUKA LWR for C1/C2/C5, LBR Med7 for
C4) it parses the reference URDF with this repo's reader and re-emits ONLY what forward kinematics
needs -- joint name / type / parent / child / origin / axis / limits -- plus each link's inertial
element (mass, centre of mass, inertia tensor: read by RobotModel.rnea) as a minimal URDF under
optas_b200/robots/.  Meshes, visuals, collision geometry, transmissions and gazebo tags
are dropped.  The numbers are printed with repr() so they round-trip exactly; tests/test_models.py
checks that FK on the derived file is bit-identical to FK on the original when that is present.
"""

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from optas_b200.urdf import URDF  # noqa: E402

REF = "/root/reference/example/robots"
SOURCES = {
    "kuka_lwr.urdf": os.path.join(REF, "kuka_lwr", "kuka_lwr.urdf"),
    "med7.urdf": os.path.join(REF, "kuka_lbr", "med7.urdf"),
    "planar_3dof.urdf": os.path.join(REF, "planar_3dof.urdf"),
}


def fmt(vals):
    return " ".join(repr(float(v)) for v in vals)


def emit(urdf: URDF) -> str:
    out = ['<?xml version="1.0"?>',
           "<!-- kinematics + inertials, derived by tests/golden/make_robot_assets.py -->",
           f'<robot name="{urdf.name}">']
    for link in urdf.links:
        if link.inertial is None:
            out.append(f'  <link name="{link.name}"/>')
            continue
        i, n = link.inertial, link.inertial.inertia
        out.append(f'  <link name="{link.name}"><inertial><mass value="{i.mass!r}"/>'
                   f'<origin rpy="{fmt(i.origin.rpy)}" xyz="{fmt(i.origin.xyz)}"/>')
        out.append(f'    <inertia ixx="{n.ixx!r}" iyy="{n.iyy!r}" izz="{n.izz!r}" ixy="{n.ixy!r}" ixz="{n.ixz!r}" iyz="{n.iyz!r}"/>'
                   f'</inertial></link>')
    for j in urdf.joints:  # one line per joint: kinematic tree first, then geometry, then limits
        parts = [f'<joint type="{j.type}" name="{j.name}"><child link="{j.child}"/><parent link="{j.parent}"/>']
        if j.origin is not None:
            parts.append(f'<origin rpy="{fmt(j.origin.rpy)}" xyz="{fmt(j.origin.xyz)}"/>')
        if j.axis is not None:
            parts.append(f'<axis xyz="{fmt(j.axis)}"/>')
        if j.limit is not None:
            parts.append(f'<limit upper="{j.limit.upper!r}" lower="{j.limit.lower!r}" '
                         f'effort="{j.limit.effort!r}" velocity="{j.limit.velocity!r}"/>')
        out.append("  " + "".join(parts) + "</joint>")
    out.append("</robot>")
    return "\n".join(out) + "\n"


# the reference's RNEA test robot (tests/test_models.py:1012-1052): a fixture, not a product asset
FIXTURES = {"tester_robot_revolute.urdf": "/root/reference/tests/tester_robot_revolute.urdf",
            "tester_robot.urdf": "/root/reference/tests/tester_robot.urdf"}

if __name__ == "__main__":
    for name, src in list(SOURCES.items()) + list(FIXTURES.items()):
        urdf = URDF.from_xml_file(src)
        dst = os.path.join(ROOT, "tests", "golden", name) if name in FIXTURES else os.path.join(ROOT, "optas_b200", "robots", name)
        with open(dst, "w") as fh:
            fh.write(emit(urdf))
        print(f"{src} -> {dst}: {len(urdf.links)} links, {len(urdf.joints)} joints")
