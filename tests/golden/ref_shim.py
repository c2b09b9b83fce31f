"""Import the UNMODIFIED reference package (`/root/reference/optas`) in this container with the
third-party modules it needs but which are not installable here replaced by stand-ins:

    casadi            -> optas_b200.sym            (this repo's casadi-free expression layer)
    urdf_parser_py    -> optas_b200.urdf           (stdlib-xml URDF reader)
    xacro, osqp, cvxopt, vtk, vtkmodules.*         -> inert stubs (never called on the solver path)

Purpose: (1) run the reference's own unit tests (tests/test_sx_container.py, test_optimization.py,
test_builder.py, test_spatialmath.py, test_optas_utils.py) against the expression layer, and
(2) let the reference's own models.py / builder.py / optimization.py generate golden vectors
(tests/golden/make_golden.py).  Only usable where /root/reference exists (not on the GPU box).
What this does NOT give: CasADi's or IPOPT's arithmetic -- those wheels are absent.
"""

import importlib
import os
import sys
import types

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {"__init__": lambda self, *a, **k: None})


def install():
    if not os.path.isdir(os.path.join(REF, "optas")):
        raise ImportError("the reference checkout is not available on this machine")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import optas_b200.sym as sym
    import optas_b200.urdf as urdf

    sym.casadi = sym  # the reference spells some names `cs.casadi.SX` (sx_container.py:12)
    sys.modules.setdefault("casadi", sym)
    pkg = types.ModuleType("urdf_parser_py")
    shim_urdf = types.ModuleType("urdf_parser_py.urdf")
    shim_urdf.__dict__.update({k: v for k, v in vars(urdf).items() if not k.startswith("__")})
    for geometry in ("Mesh", "Cylinder", "Sphere", "Box"):  # only the (out-of-scope) visualiser names these
        setattr(shim_urdf, geometry, type(geometry, (), {}))
    urdf = shim_urdf
    pkg.urdf = urdf
    sys.modules.setdefault("urdf_parser_py", pkg)
    sys.modules.setdefault("urdf_parser_py.urdf", urdf)
    for name in ("xacro", "osqp", "cvxopt", "vtk", "vtkmodules", "vtkmodules.vtkFiltersSources",
                 "pybullet", "pybullet_data"):
        sys.modules.setdefault(name, _Anything(name))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    return importlib.import_module("optas")
