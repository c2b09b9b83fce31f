"""Problem builder: collects variables, parameters, cost terms and constraints, sorts the
constraints into linear / nonlinear families and emits one of the `optimization` classes.

API-compatible with the reference's ``optas.OptimizationBuilder`` (optas/builder.py:11-635) so
task-specification scripts (example/*.py) read the same.  Behaviour that parity depends on:

* per model and time derivative ``d`` the builder owns ``{name}/{d*}sym/x`` of shape
  ``dim x (T - d)`` (``dim x T`` when ``derivs_align``) and, for robots, the parameter block
  ``.../p`` of shape ``num_param_joints x same`` -- possibly 0 rows (ref :89-99);
* a constraint written ``lhs <= rhs`` is stored as ``rhs - lhs`` (">= 0" form, ref :287-317), an
  equality as ``rhs - lhs`` ("== 0", ref :337-360); each goes to the *linear* family iff it is
  affine in the decision variables (ref :220-226);
* ``integrate_model_states`` adds ``x_t + dt * xd_t - x_{t+1} == 0`` column by column (ref :419-469);
* ``build`` picks the class from (cost quadratic?, any nonlinear constraint?, any linear
  constraint?, any discrete variable?) (ref :545-635).
"""

from __future__ import annotations

from typing import List, Optional, Union

from . import sym as cs
from .models import Model, RobotModel, TaskModel
from .optimization import (MixedIntegerNonlinearCostNonlinearConstrained, NonlinearCostLinearConstraints,
                           NonlinearCostNonlinearConstraints, NonlinearCostUnconstrained, Optimization,
                           QuadraticCostLinearConstraints, QuadraticCostNonlinearConstraints,
                           QuadraticCostUnconstrained)
from .spatialmath import ArrayType, CasADiArrayType, arrayify_args
from .sx_container import SXContainer


class OptimizationBuilder:
    def __init__(self, T: int, robots: Union[RobotModel, List[RobotModel]] = [],
                 tasks: Union[TaskModel, List[TaskModel]] = [], derivs_align: bool = False):
        assert T > 0, "T must be strictly positive"
        robots = robots if isinstance(robots, list) else [robots]
        tasks = tasks if isinstance(tasks, list) else [tasks]
        self.T = T
        self.derivs_align = derivs_align
        self._models: List[Model] = list(robots) + list(tasks)
        if self._models and not derivs_align:
            need = 1 + max(d for m in self._models for d in m.time_derivs)
            assert T >= need, f"T={T} is too low, it should be at least {need}"
        names = [m.get_name() for m in self._models]
        assert len(set(names)) == len(names), "each model should have a unique name"

        self._decision_variables = SXContainer()
        self._parameters = SXContainer()
        self._cost_terms = SXContainer()
        self._lin_eq_constraints = SXContainer()
        self._lin_ineq_constraints = SXContainer()
        self._ineq_constraints = SXContainer()
        self._eq_constraints = SXContainer()

        for m in self._models:
            for d in m.time_derivs:
                cols = T if derivs_align else T - d
                if isinstance(m, RobotModel):
                    self.add_decision_variables(m.state_optimized_name(d), m.num_opt_joints, cols)
                    self.add_parameter(m.state_parameter_name(d), m.num_param_joints, cols)
                else:
                    self.add_decision_variables(m.state_optimized_name(d), m.dim, cols, m.is_discrete)

    # -- model look-up ----------------------------------------------------------------------
    def get_model_names(self) -> List[str]:
        return [m.name for m in self._models]

    def get_model_index(self, name: str) -> int:
        return self.get_model_names().index(name)

    def get_model(self, name: str) -> Model:
        return self._models[self.get_model_index(name)]

    def _model_block(self, container: SXContainer, name: str, time_deriv: int, optimized: bool):
        m = self.get_model(name)
        assert time_deriv in m.time_derivs, \
            f"model '{name}', was not specified with time derivative to order {time_deriv}"
        label = m.state_optimized_name(time_deriv) if optimized else m.state_parameter_name(time_deriv)
        return container[label]

    def get_model_states(self, name: str, time_deriv: int = 0) -> CasADiArrayType:
        return self._model_block(self._decision_variables, name, time_deriv, True)

    def get_model_state(self, name: str, t: int, time_deriv: int = 0) -> CasADiArrayType:
        return self.get_model_states(name, time_deriv)[:, t]

    def get_model_parameters(self, name: str, time_deriv: int = 0) -> CasADiArrayType:
        return self._model_block(self._parameters, name, time_deriv, False)

    def get_model_parameter(self, name: str, t: int, time_deriv: int = 0) -> CasADiArrayType:
        return self.get_model_parameters(name, time_deriv)[:, t]

    def get_robot_states_and_parameters(self, name: str, time_deriv: int = 0) -> CasADiArrayType:
        """Full ``ndof x T`` joint array with optimised rows from x and the rest from p (ref :171-205)."""
        m = self.get_model(name)
        assert isinstance(m, RobotModel), "this method only applies to robot models"
        states = self.get_model_states(name, time_deriv)
        params = self.get_model_parameters(name, time_deriv)
        full = cs.SX.zeros(m.dim, max(1, self.T - time_deriv))
        for row, j in enumerate(m.parameter_joint_indexes):
            full[j, :] = params[row, :]
        for row, j in enumerate(m.optimized_joint_indexes):
            full[j, :] = states[row, :]
        return full

    # -- classification helpers -------------------------------------------------------------
    def _x(self):
        return self._decision_variables.vec()

    def _p(self):
        return self._parameters.vec()

    def _is_linear_in_x(self, y) -> bool:
        return cs.is_linear(y, self._x())

    def _cost(self):
        return cs.sum1(self._cost_terms.vec())

    def is_cost_quadratic(self) -> bool:
        return cs.is_quadratic(self._cost(), self._x())

    # -- variables / parameters / cost ------------------------------------------------------
    def add_decision_variables(self, name: str, m: int = 1, n: int = 1, is_discrete: bool = False) -> cs.SX:
        x = cs.SX.sym(name, m, n)
        self._decision_variables[name] = x
        if is_discrete:
            self._decision_variables.variable_is_discrete(name)
        return x

    def add_parameter(self, name: str, m: int = 1, n: int = 1) -> cs.SX:
        p = cs.SX.sym(name, m, n)
        self._parameters[name] = p
        return p

    @arrayify_args
    def add_cost_term(self, name: str, cost_term) -> None:
        cost_term = cs.vec(cost_term)
        assert cost_term.shape == (1, 1), "cost term must be scalar"
        self._cost_terms[name] = cost_term

    # -- constraints ------------------------------------------------------------------------
    def _file(self, name: str, residual, linear_box: SXContainer, nonlinear_box: SXContainer) -> None:
        (linear_box if self._is_linear_in_x(residual) else nonlinear_box)[name] = residual

    @arrayify_args
    def add_leq_inequality_constraint(self, name: str, lhs, rhs=None) -> None:
        """lhs <= rhs (rhs defaults to zeros)."""
        if rhs is None:
            rhs = cs.DM.zeros(*lhs.shape)
        self._file(name, rhs - lhs, self._lin_ineq_constraints, self._ineq_constraints)

    @arrayify_args
    def add_geq_inequality_constraint(self, name: str, lhs, rhs=None) -> None:
        """lhs >= rhs (rhs defaults to zeros)."""
        if rhs is None:
            rhs = cs.DM.zeros(*lhs.shape)
        self.add_leq_inequality_constraint(name, rhs, lhs)

    @arrayify_args
    def add_bound_inequality_constraint(self, name: str, lhs, mid, rhs) -> None:
        """lhs <= mid <= rhs, stored as two families ``name_l`` and ``name_r`` (ref :319-335)."""
        self.add_leq_inequality_constraint(name + "_l", lhs, mid)
        self.add_leq_inequality_constraint(name + "_r", mid, rhs)

    @arrayify_args
    def add_equality_constraint(self, name: str, lhs, rhs=None, reduce_constraint: bool = False) -> None:
        """lhs == rhs; ``reduce_constraint`` collapses it to the scalar ``||rhs - lhs||^2 == 0``."""
        if rhs is None:
            rhs = cs.DM.zeros(*lhs.shape)
        residual = rhs - lhs
        if reduce_constraint:
            residual = cs.sumsqr(residual)
        self._file(name, residual, self._lin_eq_constraints, self._eq_constraints)

    def sphere_collision_avoidance_constraints(self, name, obstacle_names, link_names=None, base_link=None) -> None:
        """Per time step, link and obstacle: ``(r_link + r_obs)^2 <= ||p_link(q_t) - p_obs||^2``
        with the radii / obstacle positions as new parameters (ref :366-417)."""
        model = self.get_model(name)
        assert isinstance(model, RobotModel), "this method only applies to robot models"
        base_link = model.get_root_link() if base_link is None else base_link
        Q = self.get_model_states(name)
        link_names = model.link_names if link_names is None else link_names
        assert len(link_names), "at least one link should be named"
        assert len(obstacle_names), "at least one obstacle should be named"
        links = [(ln, model.get_link_position_function(ln, base_link), self.add_parameter(ln + "_radii"))
                 for ln in link_names]
        obstacles = [(on, self.add_parameter(on + "_position", 3), self.add_parameter(on + "_radii"))
                     for on in obstacle_names]
        for t in range(Q.shape[1]):
            q = Q[:, t]
            for ln, pos, link_radius in links:
                centre = pos(q)
                for on, obs_pos, obs_radius in obstacles:
                    self.add_leq_inequality_constraint(
                        f"sphere_col_avoid_{t}_{ln}_{on}", (link_radius + obs_radius) ** 2, cs.sumsqr(centre - obs_pos))

    def integrate_model_states(self, name: str, time_deriv: int, dt) -> None:
        """Explicit-Euler consistency between derivative orders ``time_deriv - 1`` and ``time_deriv``."""
        n = self.T - (1 if self.derivs_align else time_deriv)
        if isinstance(dt, (float, int)):
            dt = dt * cs.DM.ones(n)
        dt = cs.vec(dt)
        if dt.shape[0] == 1:
            dt = dt * cs.DM.ones(n)
        dt = dt.T
        assert dt.shape[1] == n, f"The array for dt has an incorrect length, expected {n}, got {dt.shape[1]}"
        xd = self.get_model_states(name, time_deriv)
        x = self.get_model_states(name, time_deriv - 1)
        if self.derivs_align:
            xd = xd[:, :-1]
        # residual columns x_t + dt_t * xd_t - x_{t+1}; the broadcast of the 1 x n row of dt over
        # the dim x n block is the column-wise map the reference builds with Function.map (ref :432-438)
        residual = x[:, :-1] + cs.repmat(dt, x.shape[0], 1) * xd - x[:, 1:]
        self.add_equality_constraint(f"__integrate_model_states_{name}_{time_deriv}__", residual)

    def enforce_model_limits(self, name: str, time_deriv: int = 0, lo=None, up=None, safe_frac: float = 1.0) -> None:
        assert 0.0 < safe_frac <= 1.0, f"Given safe_frac '{safe_frac}' must be in range (0, 1]."
        x = self.get_model_states(name, time_deriv)
        if lo is None or up is None:
            mlo, mup = self.get_model(name).get_limits(time_deriv)
            lo = mlo if lo is None else lo
            up = mup if up is None else up
        if safe_frac < 1.0:
            centre, half = 0.5 * (lo + up), 0.5 * safe_frac * (up - lo)
            lo, up = centre - half, centre + half
        self.add_bound_inequality_constraint(f"__{name}_model_limit_{time_deriv}__", lo, x, up)

    def initial_configuration(self, name: str, init=None, time_deriv: int = 0) -> None:
        x0 = self.get_model_state(name, 0, time_deriv=time_deriv)
        self.add_equality_constraint(f"__{name}_initial_configuration_{time_deriv}__", lhs=x0, rhs=init)

    def fix_configuration(self, name: str, config=None, time_deriv: int = 0, t: int = 0) -> None:
        xt = self.get_model_state(name, t, time_deriv=time_deriv)
        self.add_equality_constraint(f"__{name}_fix_configuration_{time_deriv}_{t}__", lhs=xt, rhs=config)

    # -- build ------------------------------------------------------------------------------
    def build(self) -> Optimization:
        base = (self._decision_variables, self._parameters, self._cost_terms)
        linear = (self._lin_eq_constraints, self._lin_ineq_constraints)
        nonlinear = (self._eq_constraints, self._ineq_constraints)
        n_linear = sum(c.numel() for c in linear)
        n_nonlinear = sum(c.numel() for c in nonlinear)
        quadratic = self.is_cost_quadratic()
        if self._decision_variables.has_discrete_variables():
            opt = MixedIntegerNonlinearCostNonlinearConstrained(*base, *linear, *nonlinear)
        elif n_nonlinear > 0:
            cls = QuadraticCostNonlinearConstraints if quadratic else NonlinearCostNonlinearConstraints
            opt = cls(*base, *linear, *nonlinear)
        elif n_linear > 0:
            cls = QuadraticCostLinearConstraints if quadratic else NonlinearCostLinearConstraints
            opt = cls(*base, *linear)
        else:
            opt = (QuadraticCostUnconstrained if quadratic else NonlinearCostUnconstrained)(*base)
        opt.set_models(self._models)
        return opt
