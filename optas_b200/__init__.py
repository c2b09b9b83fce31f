"""optas_b200 -- B200-native batched NLP/QP solver back-end behind the OpTaS solver interface.

Namespace mirrors ``optas`` (reference: optas/__init__.py:1-41): the casadi-free expression layer
(`sym`) is re-exported wholesale the way the reference re-exports casadi, followed by spatial maths,
models, the builder and the solvers.  ``B200Solver`` is the product; everything else is the host
side that task-specification scripts need in order to reach it.
"""

import sys as _sys

from . import sym as cs
from .sym import *  # noqa: F401,F403  (SX, DM, Function, vertcat, sumsqr, jacobian, sin, cos, pi, ...)

# the star import brings in the scalar-symbol helper `sym()`, which would shadow the submodule
# attribute `optas_b200.sym` that `from . import sym` resolves to; put the module back
sym = _sys.modules[__name__ + ".sym"]
from .sym import DM, SX, Function
from .spatialmath import *  # noqa: F401,F403
from .spatialmath import ArrayType, arrayify_args
from .models import RobotModel, TaskModel, Model
from .builder import OptimizationBuilder
from .optimization import Optimization
from .solver import B200Solver, CasADiSolver, CVXOPTSolver, OSQPSolver, ScipyMinimizeSolver, Solver
from .nlpsol import nlpsol, qpsol

__version__ = "0.1.0"


@arrayify_args
def deg2rad(x: ArrayType):
    """Degrees -> radians (reference: optas/__init__.py:10-17)."""
    return (cs.pi / 180.0) * x


@arrayify_args
def rad2deg(x: ArrayType):
    """Radians -> degrees (reference: optas/__init__.py:20-27)."""
    return (180.0 / cs.pi) * x


@arrayify_args
def clip(x: ArrayType, lo, hi):
    """Element-wise clamp to [lo, hi] (reference: optas/__init__.py:30-41)."""
    return cs.fmax(cs.fmin(x, hi), lo)
