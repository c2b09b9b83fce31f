"""Problem IR: the functions and dimensions a solver back-end reads from a built problem.

Mirrors the attribute surface of the reference's optas/optimization.py (Optimization :54-309
and its seven problem classes :312-568) because that surface *is* the contract between
``OptimizationBuilder.build()`` and every ``Solver``:

    f, df, ddf                      cost, its Jacobian (1 x nx) and Hessian          (ref :192-204)
    P(p), q(p)                      f = x'Px + q'x (+ const) for quadratic costs     (ref :219-223)
    k, M(p), c(p)   nk              linear inequalities   k = Mx + c >= 0            (ref :225-247)
    a, A(p), b(p)   na              linear equalities     a = Ax + b == 0            (ref :249-260)
    g, dg, ddg      ng              nonlinear inequalities g >= 0                    (ref :262-281)
    h, dh, ddh      nh              nonlinear equalities   h == 0                    (ref :283-290)
    v, dv, ddv      nv, lbv, ubv    v = [k; g; a; -a; h; -h], 0 <= v <= 1e10         (ref :27-51, :292-306)

Unlike the reference, derivative functions (``df, ddf, dg, ddg, dh, ddh, dv, ddv, P, q``) are
derived lazily on first access: the stacked second derivatives ``jacobian(jacobian(.))``
(ref :20-21) are (n*nx) x nx objects that only the scipy trust-region methods ever read, and
the GPU back-end derives its own sparse Lagrangian Hessian instead (optas_b200/lowering.py).
"""

from __future__ import annotations

from typing import Callable, Dict, List, Optional

from . import sym as cs
from .sx_container import SXContainer

INF = 1.0e10  # the reference's finite stand-in for infinity (ref :58)


def derive_jacobian_and_hessian_functions(name: str, fun: cs.Function, x, p):
    """(d<name>, dd<name>): Jacobian of fun wrt x and the stacked Jacobian of that (ref :8-24)."""
    first = cs.jacobian(fun(x, p), x)
    second = cs.jacobian(first, x)
    return cs.Function("d" + name, [x, p], [first]), cs.Function("dd" + name, [x, p], [second])


def vertcon(x, p, ineq: Optional[List[cs.Function]] = None, eq: Optional[List[cs.Function]] = None) -> cs.Function:
    """Stack constraints as one ">= 0" vector: inequalities first, then each equality followed
    by its negation (ref :27-51; ordering pinned by the reference's test_optimization.py:44-70)."""
    rows = [fun(x, p) for fun in (ineq or [])]
    for fun in eq or []:
        e = fun(x, p)
        rows += [e, -e]
    return cs.Function("v", [x, p], [cs.vertcat(*rows)])


class _Lazy:
    """Descriptor: compute an attribute on first read by calling ``self._derive_<name>()``."""

    def __init__(self, name: str):
        self.name = name

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        cache = obj.__dict__.setdefault("_lazy_cache", {})
        if self.name not in cache:
            cache[self.name] = obj._derive(self.name)
        return cache[self.name]

    def __set__(self, obj, value):
        obj.__dict__.setdefault("_lazy_cache", {})[self.name] = value


class Optimization:
    inf = INF

    df = _Lazy("df")
    ddf = _Lazy("ddf")
    dg = _Lazy("dg")
    ddg = _Lazy("ddg")
    dh = _Lazy("dh")
    ddh = _Lazy("ddh")
    dv = _Lazy("dv")
    ddv = _Lazy("ddv")
    P = _Lazy("P")
    q = _Lazy("q")

    def __init__(self, decision_variables: SXContainer, parameters: SXContainer, cost_terms: SXContainer):
        self.models = None
        self.decision_variables = decision_variables
        self.parameters = parameters
        self.cost_terms = cost_terms
        self.lin_eq_constraints: Dict = {}
        self.lin_ineq_constraints: Dict = {}
        self.eq_constraints: Dict = {}
        self.ineq_constraints: Dict = {}
        self._quadratic_cost = False
        for fam in ("k", "a", "g", "h", "v"):
            setattr(self, fam, None)
            setattr(self, "n" + fam, 0)
            setattr(self, "lb" + fam, None)
            setattr(self, "ub" + fam, None)
        self.M = self.c = self.A = self.b = None
        self.x = decision_variables.vec()
        self.p = parameters.vec()
        self.nx = decision_variables.numel()
        self.np = parameters.numel()
        self.f = cs.Function("f", [self.x, self.p], [cs.sum1(cost_terms.vec())])

    # -- lazy derivative functions ---------------------------------------------------------
    def _derive(self, name: str):
        pair_of = {"df": "f", "ddf": "f", "dg": "g", "ddg": "g", "dh": "h", "ddh": "h", "dv": "v", "ddv": "v"}
        if name in pair_of:
            base = pair_of[name]
            fun = getattr(self, base)
            if fun is None:
                return None
            first, second = derive_jacobian_and_hessian_functions(base, fun, self.x, self.p)
            cache = self.__dict__.setdefault("_lazy_cache", {})
            cache["d" + base], cache["dd" + base] = first, second
            return cache[name]
        if name in ("P", "q"):
            if not self._quadratic_cost:
                return None
            if name == "P":
                return cs.Function("P", [self.p], [0.5 * self.ddf(self.x, self.p)])
            return cs.Function("q", [self.p], [cs.vec(self.df(cs.DM.zeros(self.nx), self.p))])
        raise AttributeError(name)

    # -- specification steps (same names as the reference so subclasses read alike) --------
    def set_models(self, models) -> None:
        self.models = models

    def specify_quadratic_cost(self) -> None:
        self._quadratic_cost = True

    def _linear_family(self, tag: str, Tag: str, off: str, container: SXContainer, equality: bool) -> None:
        x, p = self.x, self.p
        fun = cs.Function(tag, [x, p], [container.vec()])
        n = container.numel()
        setattr(self, tag, fun)
        setattr(self, "n" + tag, n)
        setattr(self, "lb" + tag, cs.DM.zeros(n))
        setattr(self, "ub" + tag, cs.DM.zeros(n) if equality else INF * cs.DM.ones(n))
        setattr(self, Tag, cs.Function(Tag, [p], [cs.jacobian(fun(x, p), x)]))
        setattr(self, off, cs.Function(off, [p], [fun(cs.DM.zeros(self.nx), p)]))

    def specify_linear_constraints(self, lin_ineq_constraints: SXContainer, lin_eq_constraints: SXContainer) -> None:
        self.lin_ineq_constraints = lin_ineq_constraints
        self.lin_eq_constraints = lin_eq_constraints
        self._linear_family("k", "M", "c", lin_ineq_constraints, equality=False)
        self._linear_family("a", "A", "b", lin_eq_constraints, equality=True)

    def _nonlinear_family(self, tag: str, container: SXContainer, equality: bool) -> None:
        n = container.numel()
        # the reference names both functions "g" (ref :275,:284); names are cosmetic
        setattr(self, tag, cs.Function("g", [self.x, self.p], [container.vec()]))
        setattr(self, "n" + tag, n)
        setattr(self, "lb" + tag, cs.DM.zeros(n))
        setattr(self, "ub" + tag, cs.DM.zeros(n) if equality else INF * cs.DM.ones(n))

    def specify_nonlinear_constraints(self, ineq_constraints: SXContainer, eq_constraints: SXContainer) -> None:
        self.ineq_constraints = ineq_constraints
        self.eq_constraints = eq_constraints
        self._nonlinear_family("g", ineq_constraints, equality=False)
        self._nonlinear_family("h", eq_constraints, equality=True)

    def specify_v(self, ineq: Optional[List[cs.Function]] = None, eq: Optional[List[cs.Function]] = None) -> None:
        self.v = vertcon(self.x, self.p, ineq=ineq, eq=eq)
        self.nv = self.v.numel_out()
        self.lbv = cs.DM.zeros(self.nv)
        self.ubv = INF * cs.DM.ones(self.nv)

    def has_discrete_variables(self) -> bool:
        return self.decision_variables.has_discrete_variables()


class QuadraticCostUnconstrained(Optimization):
    """min x'P x + q'x (ref :312-330)."""

    def __init__(self, decision_variables, parameters, cost_terms):
        super().__init__(decision_variables, parameters, cost_terms)
        self.specify_quadratic_cost()


class QuadraticCostLinearConstraints(Optimization):
    """Quadratic cost, k >= 0, a == 0 (ref :333-358)."""

    def __init__(self, decision_variables, parameters, cost_terms, lin_eq_constraints, lin_ineq_constraints):
        super().__init__(decision_variables, parameters, cost_terms)
        self.specify_quadratic_cost()
        self.specify_linear_constraints(lin_ineq_constraints, lin_eq_constraints)
        self.specify_v(ineq=[self.k], eq=[self.a])


class QuadraticCostNonlinearConstraints(Optimization):
    """Quadratic cost, linear and nonlinear constraints (ref :361-418)."""

    def __init__(self, decision_variables, parameters, cost_terms, lin_eq_constraints, lin_ineq_constraints,
                 eq_constraints, ineq_constraints):
        super().__init__(decision_variables, parameters, cost_terms)
        self.specify_quadratic_cost()
        self.specify_linear_constraints(lin_ineq_constraints, lin_eq_constraints)
        self.specify_nonlinear_constraints(ineq_constraints, eq_constraints)
        self.specify_v(ineq=[self.k, self.g], eq=[self.a, self.h])


class NonlinearCostUnconstrained(Optimization):
    """min f(x) (ref :421-435)."""


class NonlinearCostLinearConstraints(Optimization):
    """Nonlinear cost, k >= 0, a == 0 (ref :438-470)."""

    def __init__(self, decision_variables, parameters, cost_terms, lin_eq_constraints, lin_ineq_constraints):
        super().__init__(decision_variables, parameters, cost_terms)
        self.specify_linear_constraints(lin_ineq_constraints, lin_eq_constraints)
        self.specify_v(ineq=[self.k], eq=[self.a])


class NonlinearCostNonlinearConstraints(Optimization):
    """General NLP (ref :473-520)."""

    def __init__(self, decision_variables, parameters, cost_terms, lin_eq_constraints, lin_ineq_constraints,
                 eq_constraints, ineq_constraints):
        super().__init__(decision_variables, parameters, cost_terms)
        self.specify_linear_constraints(lin_ineq_constraints, lin_eq_constraints)
        self.specify_nonlinear_constraints(ineq_constraints, eq_constraints)
        self.specify_v(ineq=[self.k, self.g], eq=[self.a, self.h])


class MixedIntegerNonlinearCostNonlinearConstrained(NonlinearCostNonlinearConstraints):
    """Same IR with discrete variables flagged (ref :523-568).  The GPU back-end refuses these
    (branch-and-bound is out of scope)."""


QP_COST = {QuadraticCostUnconstrained, QuadraticCostLinearConstraints, QuadraticCostNonlinearConstraints}
UNCONSTRAINED_OPT = {QuadraticCostUnconstrained, NonlinearCostUnconstrained}
CONSTRAINED_OPT = {
    QuadraticCostLinearConstraints,
    QuadraticCostNonlinearConstraints,
    NonlinearCostLinearConstraints,
    NonlinearCostNonlinearConstraints,
    MixedIntegerNonlinearCostNonlinearConstrained,
}
