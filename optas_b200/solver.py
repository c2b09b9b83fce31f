"""Solver front: the plug-in boundary of OpTaS, with the B200 back-end behind it.

``Solver`` restates the reference's abstract base (optas/solver.py:61-315): same members, same
argument meaning, same error behaviour.  ``B200Solver`` is the drop-in for the concrete
``CasADiSolver`` (ref :321-419) and ``ScipyMinimizeSolver`` (ref :587-813): construct it with a
built ``Optimization``, call ``setup(...)`` with either spelling, then ``reset_parameters`` /
``reset_initial_seed`` / ``solve`` exactly as task scripts do (example/example.py:40-60).

Batch extension (the reason this back-end exists): any value handed to ``reset_parameters`` /
``reset_initial_seed`` may carry a leading batch axis -- ``[B, m, n]`` or ``[B, m*n]`` for an
``m x n`` label -- and then ``solve()`` returns ``[B, m, n]`` arrays and ``stats()`` per-instance
status / iteration / KKT-error arrays.  Without a batch axis everything behaves (and is typed)
like the reference: ``solve()`` returns a dict of ``DM``.

All numerical work goes through libb200optas.so (ctypes, GIL released); there is no CPU solve
path in this module -- without the library or without a GPU ``setup`` / ``solve`` raise.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import sym as cs
from .lowering import LoweredProblem, lower_problem
from .models import RobotModel
from .optimization import CONSTRAINED_OPT, MixedIntegerNonlinearCostNonlinearConstrained, Optimization
from .spatialmath import ArrayType, CasADiArrayType


class Solver(ABC):
    """Abstract solver (reference: optas/solver.py:61-315)."""

    def __init__(self, optimization: Optimization, error_on_fail: bool = False):
        self.opt = optimization
        self.x0 = cs.DM.zeros(optimization.nx)  # ref :76 -- an unset seed is zeros
        self.p = cs.DM.zeros(optimization.np)
        self._p_dict: Dict = {}
        self._error_on_fail = error_on_fail
        self._solution = None

    @property
    def opt_type(self) -> type:
        return type(self.opt)

    @abstractmethod
    def setup(self, *args, **kwargs):
        """Must return ``self`` (ref :98-101)."""

    def reset_initial_seed(self, x0: Dict[str, ArrayType]) -> None:
        """Missing labels become zeros, unknown labels are ignored (ref :103-108)."""
        self.x0 = self.opt.decision_variables.dict2vec(x0)

    def reset_parameters(self, p: Dict[str, ArrayType]) -> None:
        self.p = self.opt.parameters.dict2vec(p)
        self._p_dict = self.opt.parameters.vec2dict(self.p)

    @abstractmethod
    def _solve(self) -> CasADiArrayType:
        pass

    def _with_full_states(self, solution: Dict, p_dict: Dict, zeros, take_rows) -> Dict:
        """Add ``name/q`` entries rebuilt from ``name/q/x`` and ``name/q/p`` (ref :137-155)."""
        for model in self.opt.models or []:
            for d in model.time_derivs:
                full, opt_name = model.state_name(d), model.state_optimized_name(d)
                if isinstance(model, RobotModel) and model.num_param_joints > 0:
                    par_name = model.state_parameter_name(d)
                    merged = zeros(model.dim, solution[opt_name])
                    take_rows(merged, model.optimized_joint_indexes, solution[opt_name])
                    take_rows(merged, model.parameter_joint_indexes, p_dict[par_name])
                    solution[full] = merged
                else:
                    solution[full] = solution[opt_name]
        return solution

    def solve(self) -> Dict[str, CasADiArrayType]:
        solution = self.opt.decision_variables.vec2dict(self._solve())
        if self._error_on_fail and not self.did_solve():
            raise RuntimeError("Solver failed!")

        def zeros(dim, like):
            return cs.DM.zeros(dim, like.shape[1])

        def take_rows(dst, rows, src):
            dst[rows, :] = src

        return self._with_full_states(solution, self._p_dict, zeros, take_rows)

    @abstractmethod
    def stats(self):
        pass

    @abstractmethod
    def did_solve(self) -> bool:
        pass

    @abstractmethod
    def number_of_iterations(self) -> int:
        pass

    # -- diagnostics (ref :167-314) -----------------------------------------------------------
    def violated_constraints(self, x: Dict[str, ArrayType], p: Dict[str, ArrayType]) -> Tuple:
        xv = self.opt.decision_variables.dict2vec(x)
        pv = self.opt.parameters.dict2vec(p)

        @dataclass
        class ViolatedConstraint:
            label: str
            ctype: str
            diff: cs.DM
            pattern: cs.DM

            def __str__(self):
                return f"\n{self.label} [{self.ctype}]:\n{self.pattern}\n"

            def __repr__(self):
                info = str(self)
                width = max(len(line) for line in info.split("\n"))
                return "=" * width + info + "-" * width + "\n"

            @property
            def verbose_info(self):
                return str(self) + f"{self.diff}\n"

        def family(container, ctype) -> List:
            out = []
            for label, expr in container.items():
                diff = cs.Function("fun", [self.opt.x, self.opt.p], [expr])(xv, pv)
                out.append(ViolatedConstraint(label, ctype, diff, diff >= 0.0))
            return out

        return (family(self.opt.lin_eq_constraints, "lin_eq"), family(self.opt.eq_constraints, "eq"),
                family(self.opt.lin_ineq_constraints, "lin_ineq"), family(self.opt.ineq_constraints, "ineq"))

    @staticmethod
    def interpolate(traj: cs.DM, T: float, **interp_args):
        from scipy.interpolate import interp1d

        assert isinstance(traj, cs.DM), f"traj is incorrect type, got '{type(traj)}', expected casadi.DM'"
        return interp1d(np.linspace(0, T, traj.shape[1]), traj.toarray(), **interp_args)

    def evaluate_cost(self, x: Dict[str, ArrayType], p: Dict[str, ArrayType]) -> CasADiArrayType:
        return self.opt.f(self.opt.decision_variables.dict2vec(x), self.opt.parameters.dict2vec(p))

    def evaluate_cost_terms(self, x: Dict[str, ArrayType], p: Dict[str, ArrayType]) -> List:
        xv = self.opt.decision_variables.dict2vec(x)
        pv = self.opt.parameters.dict2vec(p)
        return [cs.Function("fun", [self.opt.x, self.opt.p], [expr])(xv, pv) for expr in self.opt.cost_terms.values()]


# ----------------------------------------------------------------------------------------------
# batched marshalling: dict of (possibly batched) arrays <-> [B, n] row-major matrix
# ----------------------------------------------------------------------------------------------


def _as_numpy(v) -> np.ndarray:
    if isinstance(v, (cs.DM,)):
        return v.toarray()
    if isinstance(v, cs.SX):
        return cs.DM(v).toarray()
    if hasattr(v, "detach") and hasattr(v, "cpu"):  # torch tensor
        return v.detach().cpu().numpy()
    return np.asarray(v, dtype=float)


def _batch_of(value: np.ndarray, m: int, n: int) -> Optional[int]:
    """Batch size carried by ``value`` for an m x n label, or None when it is one instance."""
    if value.ndim == 3:
        return int(value.shape[0])
    if value.ndim == 2 and value.shape != (m, n) and value.size != m * n and value.shape[1] == m * n:
        return int(value.shape[0])
    return None


_PINNED = None  # None: not probed yet; False: unavailable


def host_array(shape, dtype=np.float64, zero: bool = False) -> np.ndarray:
    """Host array for data that crosses PCIe: page-locked (torch's caching host allocator) when a CUDA device is
    present, so that the library's cuMemcpy*Async runs at DMA speed; plain numpy otherwise.  The numpy view keeps
    the pinned block alive, a later solve never overwrites an array handed to the caller."""
    global _PINNED
    n = int(np.prod(shape))
    if _PINNED is None:
        try:
            import torch

            _PINNED = torch if torch.cuda.is_available() else False
        except Exception:  # torch is plumbing, not a requirement
            _PINNED = False
    if _PINNED and n * np.dtype(dtype).itemsize >= (1 << 16):
        tdt = {np.dtype(np.float64): _PINNED.float64, np.dtype(np.int32): _PINNED.int32}[np.dtype(dtype)]
        t = _PINNED.zeros(shape, dtype=tdt, pin_memory=True) if zero else _PINNED.empty(shape, dtype=tdt, pin_memory=True)
        return t.numpy()
    return np.zeros(shape, dtype=dtype) if zero else np.empty(shape, dtype=dtype)


def pack_batch(container, d: Dict[str, ArrayType]) -> Tuple[np.ndarray, Optional[int]]:
    """Vectorised ``SXContainer.dict2vec`` (ref sx_container.py:113-123) over a batch.

    Returns ``(M, B)``: ``M`` is ``[B or 1, numel]`` float64 C-contiguous, every row the
    column-major flattening ``vec()`` layout of one instance; ``B`` is None when no value was batched.
    """
    layout = container.offsets()
    total = container.numel()
    values, B = {}, None
    for label, (off, m, n) in layout.items():
        if label not in d or d[label] is None:
            continue
        v = _as_numpy(d[label])
        b = _batch_of(v, m, n)
        if b is not None:
            if B is not None and b != B:
                raise ValueError(f"'{label}': batch size {b} does not match {B}")
            B = b
        values[label] = (v, b)
    out = host_array((B or 1, total))
    if total == 0:
        return out, B
    # one threaded pass of the library over all labels (bo_pack_rows): missing labels become zeros like the reference's
    # dict2vec, [m, n] values are broadcast over the batch, every instance's matrix lands column-major (vec() layout)
    from . import _capi

    segments, keep = [], []
    for label, (off, m, n) in layout.items():
        if m * n == 0:
            continue
        if label not in values:
            segments.append((None, (0, 0, 0), m, n, off))
            continue
        v, b = values[label]
        if v.dtype != np.float64:
            v = v.astype(np.float64)
        keep.append(v)
        e = v.itemsize
        if b is None:
            if v.size != m * n:
                raise ValueError(f"'{label}': expected {m * n} elements, got {v.size}")
            if v.ndim == 2 and v.shape == (m, n):
                st = (0, v.strides[0] // e, v.strides[1] // e)
            else:  # a flat vector is the column-major flattening (reference: cs.vec)
                v = np.ascontiguousarray(v).reshape(-1)
                keep.append(v)
                st = (0, 1, m)
        elif v.ndim == 3:
            if v.shape[1:] != (m, n):
                raise ValueError(f"'{label}': expected [B, {m}, {n}], got {list(v.shape)}")
            st = (v.strides[0] // e, v.strides[1] // e, v.strides[2] // e)
        else:  # [B, m*n]: every row already the column-major flattening of one instance
            st = (v.strides[0] // e, v.strides[1] // e, m * (v.strides[1] // e))
        segments.append((v, st, m, n, off))
    _capi.pack_rows(out, segments)
    return out, B


def unpack_batch(container, M: np.ndarray) -> Dict[str, np.ndarray]:
    """Vectorised ``vec2dict``: ``[B, numel]`` -> label -> ``[B, m, n]``."""
    out = {}
    B = M.shape[0]
    for label, (off, m, n) in container.offsets().items():
        out[label] = M[:, off:off + m * n].reshape(B, n, m).transpose(0, 2, 1)
    return out


# ----------------------------------------------------------------------------------------------
# the B200 back-end
# ----------------------------------------------------------------------------------------------

_NLP_NAMES = {"ipopt", "knitro", "snopt", "worhp", "scpgen", "sqpmethod", "b200"}
_QP_NAMES = {"cplex", "gurobi", "ooqp", "qpoases", "sqic", "nlp", "osqp", "cvxopt"}
_MI_NAMES = {"bonmin", "knitro"}
_SCIPY_CONSTRAINED = {"COBYLA", "SLSQP", "trust-constr"}
_SCIPY_METHODS = {"Nelder-Mead", "Powell", "CG", "BFGS", "Newton-CG", "L-BFGS-B", "TNC", "COBYLA", "SLSQP",
                  "trust-constr", "dogleg", "trust-ncg", "trust-exact", "trust-krylov"}


class B200Solver(Solver):
    """Batched interior-point / Newton-KKT solver on B200 (libb200optas.so) behind the OpTaS API.

    ``setup`` accepts both reference spellings:
      * ``setup("ipopt", {"ipopt.max_iter": 100, ...})``  (CasADiSolver.setup, ref :333)
      * ``setup(method="SLSQP", tol=1e-6, options={"maxiter": 100})``  (ScipyMinimizeSolver.setup, ref :619)
    The solver name / method selects nothing but option parsing: every problem class runs the
    same fused GPU kernel.  Returns ``self``.
    """

    def __init__(self, optimization: Optimization, error_on_fail: bool = False):
        super().__init__(optimization, error_on_fail)
        self._handle = None
        self._lowered: Optional[LoweredProblem] = None
        self._stats: Optional[Dict] = None
        self._batch: Optional[int] = None
        self._X0 = np.zeros((1, optimization.nx))
        self._P = np.zeros((1, optimization.np))
        self._x0_batched = self._p_batched = None

    # -- setup --------------------------------------------------------------------------------
    def setup(self, solver_name: Optional[str] = None, solver_options: Optional[Dict] = None, *,
              method: Optional[str] = None, tol: Optional[float] = None, options: Optional[Dict] = None,
              compile_only: bool = False, timing: bool = False, threads_per_block: int = 0, max_trips: int = 0,
              blocks_per_sm: int = 0, pivoted_ldl: bool = False, coop: Optional[bool] = None, team: Optional[bool] = None,
              qp: Optional[bool] = None, schedule: Optional[str] = None):
        """``schedule="seed_infeasibility"`` (device-resident ``solve_raw`` only): the batch is processed in the order of
        decreasing constraint violation of the seed -- see ``_solve_scheduled``."""
        from . import _capi

        if schedule not in (None, "natural", "seed_infeasibility"):
            raise ValueError(f"unknown schedule '{schedule}'")
        self._schedule = schedule if schedule != "natural" else None
        self._sched = None

        name = solver_name if solver_name is not None else (method if method is not None else "ipopt")
        solver_options = dict(solver_options or {})  # never mutate the caller's dict
        if name in _SCIPY_METHODS:
            # ScipyMinimizeSolver.setup semantics (ref :619-639)
            if self.opt_type in CONSTRAINED_OPT and name not in _SCIPY_CONSTRAINED:
                raise TypeError(f"optimization problem has constraints, the method '{name}' is not suitable")
        elif name not in _NLP_NAMES | _QP_NAMES | _MI_NAMES:
            raise ValueError(f"solver '{name}' does not support this problem type")
        if self.opt.has_discrete_variables() or isinstance(self.opt, MixedIntegerNonlinearCostNonlinearConstrained):
            raise NotImplementedError("mixed-integer problems are out of scope for the B200 back-end")

        max_iter, acc_tol, mu_init, max_step = 0, 0.0, 0.0, 0.0
        tol_use = float(tol) if tol is not None else 0.0
        for key, val in solver_options.items():
            leaf = key.split(".")[-1]
            if leaf == "max_iter":
                max_iter = int(val)
            elif leaf == "tol":
                tol_use = float(val)
            elif leaf == "acceptable_tol":
                acc_tol = float(val)
            elif leaf == "mu_init":
                mu_init = float(val)
            elif leaf == "max_step":
                max_step = float(val)
            elif leaf == "max_trips":
                max_trips = int(val)
        for key, val in (options or {}).items():
            if key == "maxiter":
                max_iter = int(val)
            elif key in ("ftol", "gtol", "xtol") and tol is None:
                pass  # scipy's per-method tolerances have no KKT-error equivalent; `tol` is honoured

        self._lowered = lower_problem(self.opt)
        flags = ((_capi.BO_FLAG_COMPILE_ONLY if compile_only else 0) | (_capi.BO_FLAG_TIMING if timing else 0)
                 | (_capi.BO_FLAG_PIVOTED_LDL if pivoted_ldl else 0)
                 | (_capi.BO_FLAG_COOP if coop else 0) | (_capi.BO_FLAG_NO_COOP if coop is False else 0)
                 | (_capi.BO_FLAG_NO_TEAM if team is False else 0) | (_capi.BO_FLAG_TEAM if team else 0)
                 | (_capi.BO_FLAG_NO_QP if qp is False else 0))
        self._handle = _capi.ProblemHandle(self._lowered, flags=flags, max_iter=max_iter, tol=tol_use,
                                           acceptable_tol=acc_tol, mu_init=mu_init, max_step=max_step,
                                           threads_per_block=threads_per_block, max_trips=max_trips,
                                           blocks_per_sm=blocks_per_sm)
        self._stats = None
        return self

    # -- inputs -------------------------------------------------------------------------------
    def reset_initial_seed(self, x0: Dict[str, ArrayType]) -> None:
        self._X0, self._x0_batched = pack_batch(self.opt.decision_variables, x0)
        self.x0 = cs.DM(self._X0[0])

    def reset_parameters(self, p: Dict[str, ArrayType]) -> None:
        self._P, self._p_batched = pack_batch(self.opt.parameters, p)
        self.p = cs.DM(self._P[0])
        self._p_dict = self.opt.parameters.vec2dict(self.p)

    # -- raw batched call (numpy or torch tensors, host or device) ---------------------------
    def solve_raw(self, P, X0, X, lam=None, f=None, status=None, iters=None, kkt=None, stream: int = 0) -> None:
        """Direct ``bo_solve``: ``P [B, np]``, ``X0 [B, nx]`` (or None), ``X [B, nx]`` out, all float64
        C-contiguous; numpy arrays (host) or torch tensors (host or cuda).  No copies are made here."""
        if self._handle is None:
            raise RuntimeError("call setup() first")
        B = int(X.shape[0])
        if getattr(self, "_schedule", None) == "seed_infeasibility" and self._can_schedule(P, X0, X):
            self._solve_scheduled(B, P, X0, X, lam, f, status, iters, kkt, stream)
            return
        self._handle.solve(B, P, X0, X, lam, f, status, iters, kkt, stream)

    # -- longest-first scheduling of a device-resident batch -------------------------------------
    def seed_infeasibility_function(self, compile_only: bool = False):
        """theta(x, p) = sum max(0, -v(x, p))^2 as a streaming-kernel function (inputs ``[B, nx]``, ``[B, np]``, output ``[B, 1]``)."""
        from .function import B200Function

        x, p = self.opt.x, self.opt.p
        theta = cs.sumsqr(cs.fmax(0.0, -self.opt.v(x, p)))
        return B200Function(cs.Function("seed_infeasibility", [x, p], [theta]), compile_only=compile_only)

    def _can_schedule(self, P, X0, X) -> bool:
        lo = self._lowered
        if P is None or X0 is None or lo.np_ == 0 or lo.n_eq + lo.n_ineq == 0:
            return False
        return all(type(t).__module__.startswith("torch") and t.is_cuda for t in (P, X0, X))

    def _solve_scheduled(self, B, P, X0, X, lam, f, status, iters, kkt, stream) -> None:
        """A launch of the persistent solver kernel ends with its slowest instance, and an instance that starts in the last
        wave and then needs ten times the median number of iterations keeps the whole GPU waiting (C2, 65536 instances:
        5.98 ms in the caller's order, 4.3 ms per 65536 at a batch of a million).  Which instances are slow is not known in
        advance, but it correlates with how far the seed is from feasibility, so the batch is handed to the kernel in the
        order of decreasing seed infeasibility theta(x0, p) = sum max(0, -v(x0, p))^2 (v = the reference's stacked constraint
        vector, optas/optimization.py:27-51): longest-processing-time-first with theta as the predictor.  Every instance
        is solved exactly as before (bitwise: instances are independent); only the order in which lanes pick them up
        changes.  Measured on B200 (tools/order_probe.py, kernel only): caller's order 6.05 ms, random order 6.03 ms,
        decreasing 2-norm 5.03 ms, max-norm 5.01 ms, 1-norm 5.58 ms, 16 quantile buckets of the 2-norm 4.98 ms, increasing
        2-norm (the worst case) 7.10 ms -- a handful of instances decide the tail, so the predictor matters.
        theta is one evaluation of the constraint tape per instance through the streaming kernel (bo_function_eval); the
        sort, the gather of the input rows and the scatter of the results are torch device ops on the caller's stream."""
        import torch

        lo = self._lowered
        sc = self._sched
        if sc is None or sc["B"] != B or sc["device"] != X.device:
            if sc is None:
                from ._capi import BoError

                try:
                    fun = self.seed_infeasibility_function()
                except BoError as err:  # e.g. more doubles per instance than the streaming kernel's shared-memory pipeline takes
                    import warnings

                    warnings.warn(f"schedule='seed_infeasibility' is not available for this problem ({err}); using the caller's order")
                    self._schedule = None
                    self._handle.solve(B, P, X0, X, lam, f, status, iters, kkt, stream)
                    return
            else:
                fun = sc["fun"]
            kw = dict(device=X.device)
            sc = self._sched = {
                "B": B, "device": X.device, "fun": fun, "theta": torch.empty((B, 1), dtype=torch.float64, **kw),
                "P": torch.empty((B, lo.np_), dtype=torch.float64, **kw), "X0": torch.empty((B, lo.nx), dtype=torch.float64, **kw),
                "X": torch.empty((B, lo.nx), dtype=torch.float64, **kw),
                "lam": torch.empty((B, lo.n_eq + lo.n_ineq), dtype=torch.float64, **kw), "f": torch.empty(B, dtype=torch.float64, **kw),
                "status": torch.empty(B, dtype=torch.int32, **kw), "iters": torch.empty(B, dtype=torch.int32, **kw),
                "kkt": torch.empty(B, dtype=torch.float64, **kw)}
        ts = torch.cuda.ExternalStream(stream) if stream else torch.cuda.current_stream()
        with torch.cuda.stream(ts):  # the torch ops below and the two kernels of the library share one stream
            sc["fun"].eval_raw(B, [X0, P], [sc["theta"]], ts.cuda_stream)
            order = torch.argsort(sc["theta"].view(-1), descending=True)
            torch.index_select(P, 0, order, out=sc["P"])
            torch.index_select(X0, 0, order, out=sc["X0"])
            outs = {"lam": lam, "f": f, "status": status, "iters": iters, "kkt": kkt}
            self._handle.solve(B, sc["P"], sc["X0"], sc["X"], *[None if outs[k] is None else sc[k] for k in ("lam", "f", "status", "iters", "kkt")],
                               ts.cuda_stream)
            X.index_copy_(0, order, sc["X"])
            for k, dst in outs.items():
                if dst is not None:
                    dst.index_copy_(0, order, sc[k])

    def solve_arrays(self, P: np.ndarray, X0: Optional[np.ndarray]) -> Dict[str, np.ndarray]:
        """Batched solve on host arrays in ``vec()`` layout: ``P [B, np]``, ``X0 [B, nx]`` (None = zeros).
        Returns ``x [B, nx]``, ``lam [B, n_eq + n_ineq]``, ``f``, ``status``, ``iters``, ``kkt`` ([B] each)."""
        if self._handle is None:
            raise RuntimeError("call setup() first")
        lo = self._lowered
        n = int(P.shape[0]) if lo.np_ else int(X0.shape[0])
        P = np.ascontiguousarray(P, dtype=np.float64)
        X0 = None if X0 is None else np.ascontiguousarray(X0, dtype=np.float64)
        out = self._result_buffers(n)
        self._handle.solve(n, P if lo.np_ else None, X0, out["x"], out["lam"], out["f"], out["status"], out["iters"],
                           out["kkt"])
        return out

    def _result_buffers(self, n: int) -> Dict[str, np.ndarray]:
        """Page-locked result arrays for a batch of ``n``.  Allocating them costs 1.2 ms per call at n = 65536 (measured,
        tools/e2e_breakdown2.py), so sets are recycled -- but only sets nobody else still holds: a set is reused when the
        reference count of every array in it shows this pool as the only owner (views handed out by ``solve()`` keep
        their base array alive, so a caller who still holds a solution keeps its buffers).  A result handed to the caller
        is therefore never overwritten by a later solve."""
        import sys

        lo = self._lowered
        pool = self.__dict__.setdefault("_pool", [])
        for st in pool:
            if st["n"] == n and all(sys.getrefcount(a) <= 3 for a in st["arrays"].values()):  # pool dict + loop variable + argument
                return dict(st["arrays"])
        arrays = {"x": host_array((n, lo.nx)), "lam": host_array((n, lo.n_eq + lo.n_ineq)), "f": host_array((n,)),
                  "status": host_array((n,), np.int32), "iters": host_array((n,), np.int32), "kkt": host_array((n,))}
        pool.append({"n": n, "arrays": arrays})
        if len(pool) > 4:
            pool.pop(0)
        return dict(arrays)

    # -- solve --------------------------------------------------------------------------------
    def _run(self) -> np.ndarray:
        if self._handle is None:
            raise RuntimeError("call setup() first")
        B = self._x0_batched or self._p_batched
        if self._x0_batched and self._p_batched and self._x0_batched != self._p_batched:
            raise ValueError(f"seed batch {self._x0_batched} != parameter batch {self._p_batched}")
        self._batch = B
        n = B or 1
        def rows(M, width):
            if M.shape == (n, width) and M.flags.c_contiguous:
                return M
            out = host_array((n, width))
            out[...] = M
            return out

        X0, P = rows(self._X0, self.opt.nx), rows(self._P, self.opt.np)
        lo = self._lowered
        r = self.solve_arrays(P, X0)
        X, lam, f, status, iters, kkt = r["x"], r["lam"], r["f"], r["status"], r["iters"], r["kkt"]
        ok = status <= 1
        self._stats = {
            "success": bool(ok.all()),
            "iter_count": int(iters.max()),
            "return_status": "Solve_Succeeded" if ok.all() else "Solve_Failed",
            "n_instances": n,
            "n_converged": int(ok.sum()),
            "status": status,
            "iterations": iters,
            "kkt_error": kkt,
            "solution": {"x": X if B else cs.DM(X[0]), "f": f if B else cs.DM(f[0]),
                         "lam_eq": lam[:, :lo.n_eq], "lam_ineq": lam[:, lo.n_eq:]},
        }
        self._solution = self._stats["solution"]
        return X

    def _solve(self):
        X = self._run()
        return X if self._batch else X[0]

    def solve(self) -> Dict:
        X = self._run()
        if self._batch is None:
            solution = self.opt.decision_variables.vec2dict(X[0])
            if self._error_on_fail and not self.did_solve():
                raise RuntimeError("Solver failed!")

            def zeros(dim, like):
                return cs.DM.zeros(dim, like.shape[1])

            def take_rows(dst, rows, src):
                dst[rows, :] = src

            return self._with_full_states(solution, self._p_dict, zeros, take_rows)

        if self._error_on_fail and not self.did_solve():
            raise RuntimeError("Solver failed!")
        solution = unpack_batch(self.opt.decision_variables, X)
        p_dict = unpack_batch(self.opt.parameters, np.broadcast_to(self._P, (X.shape[0], self.opt.np)))

        def zeros(dim, like):
            return np.zeros((like.shape[0], dim, like.shape[2]))

        def take_rows(dst, rows, src):
            dst[:, rows, :] = src

        return self._with_full_states(solution, p_dict, zeros, take_rows)

    # -- outcome ------------------------------------------------------------------------------
    def stats(self) -> Dict:
        return self._stats

    def did_solve(self) -> bool:
        return bool(self._stats["success"])

    def number_of_iterations(self) -> int:
        return int(self._stats["iter_count"])

    # -- batched diagnostics on the GPU (ref :167-314 evaluates them one instance at a time) ----------
    def _diagnostic_function(self, key: str, exprs: List):
        """Streaming kernel (K1) evaluating ``exprs(x, p)``; built and compiled on first use."""
        from .function import B200Function

        cache = self.__dict__.setdefault("_diag_functions", {})
        if key not in cache:
            # one output segment holding every expression (column-major flattening each), split again on the host
            sizes = [cs.SX(e).numel() for e in exprs]
            fun = B200Function(cs.Function(key, [self.opt.x, self.opt.p], [cs.vertcat(*[cs.vec(cs.SX(e)) for e in exprs])]))
            cache[key] = (fun, np.cumsum([0] + sizes))

        fun, offs = cache[key]

        def call(X, P):
            out = fun(X, P)
            return [out[:, offs[k]:offs[k + 1]] for k in range(len(offs) - 1)]

        return call

    def _diag_inputs(self, x: Dict[str, ArrayType], p: Dict[str, ArrayType]):
        X, bx = pack_batch(self.opt.decision_variables, x)
        P, bp = pack_batch(self.opt.parameters, p)
        B = bx or bp
        if B is None:
            return None
        return np.ascontiguousarray(np.broadcast_to(X, (B, X.shape[1]))), np.ascontiguousarray(np.broadcast_to(P, (B, P.shape[1]))), B

    def evaluate_cost_terms(self, x: Dict[str, ArrayType], p: Dict[str, ArrayType]) -> List:
        """Reference semantics for one instance (ref :303-314); with a batch axis on any value every cost term is
        evaluated for the whole batch by one launch of the streaming kernel and comes back as a ``[B]`` array."""
        packed = self._diag_inputs(x, p)
        if packed is None:
            return super().evaluate_cost_terms(x, p)
        X, P, B = packed
        terms = list(self.opt.cost_terms.values())
        if not terms:
            return []
        out = self._diagnostic_function("cost_terms", terms)(X, P)
        return [o[:, 0] if o.shape[1] == 1 else o for o in out]

    def evaluate_cost(self, x: Dict[str, ArrayType], p: Dict[str, ArrayType]):
        packed = self._diag_inputs(x, p)
        if packed is None:
            return super().evaluate_cost(x, p)
        X, P, B = packed
        return self._diagnostic_function("cost", [self.opt.f(self.opt.x, self.opt.p)])(X, P)[0][:, 0]

    def violated_constraints(self, x: Dict[str, ArrayType], p: Dict[str, ArrayType]) -> Tuple:
        """Reference semantics for one instance (ref :167-238).  Batched: every constraint family is evaluated on the
        GPU; ``diff`` is ``[B, m, n]`` and ``pattern`` the boolean ``diff >= 0`` (the reference's convention)."""
        packed = self._diag_inputs(x, p)
        if packed is None:
            return super().violated_constraints(x, p)
        X, P, B = packed

        @dataclass
        class ViolatedConstraintBatch:
            label: str
            ctype: str
            diff: np.ndarray
            pattern: np.ndarray

            @property
            def n_violated(self) -> np.ndarray:  # per instance
                bad = self.diff < 0.0 if self.ctype.endswith("ineq") else self.diff != 0.0
                return bad.reshape(bad.shape[0], -1).sum(axis=1)

        def family(container, ctype) -> List:
            items = list(container.items())
            if not items:
                return []
            out = self._diagnostic_function("constraints_" + ctype, [e for _, e in items])(X, P)
            res = []
            for (label, expr), o in zip(items, out):
                m, n = cs.SX(expr).shape
                diff = o.reshape(B, n, m).transpose(0, 2, 1)  # column-major flattening per instance
                res.append(ViolatedConstraintBatch(label, ctype, diff, diff >= 0.0))
            return res

        return (family(self.opt.lin_eq_constraints, "lin_eq"), family(self.opt.eq_constraints, "eq"),
                family(self.opt.lin_ineq_constraints, "lin_ineq"), family(self.opt.ineq_constraints, "ineq"))

    # -- closed-loop residency ---------------------------------------------------
    def capture_tick(self, P, X0, X, status=None, iters=None, lam=None, f=None, kkt=None):
        """Capture one solve on device-resident tensors into a CUDA graph and return it (``graph.replay()`` runs a
        tick: counter reset + the solver kernel, no host work but the replay call).  The tensors are the graph's
        fixed buffers: write the new parameters into ``P`` (and the seed into ``X0``; pass ``X0 = X`` to warm-start
        every tick from the previous solution, example/point_mass_mpc.py:156-175) before each replay."""
        import torch

        self.solve_raw(P, X0, X, lam, f, status, iters, kkt, stream=torch.cuda.current_stream().cuda_stream)  # warm-up: allocations
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(graph, stream=side):
            self.solve_raw(P, X0, X, lam, f, status, iters, kkt, stream=side.cuda_stream)
        return graph

    def kernel_info(self) -> Dict:
        return self._handle.kernel_info()

    def tier_info(self) -> Dict:
        """Which kernel tier the problem was lowered to, and the sizes of its plan."""
        return self._handle.tier_info()

    def kernel_source(self) -> str:
        return self._handle.source()

    def ldl_table(self) -> np.ndarray:
        """Tables of the sparse KKT factorisation (empty for the dense tiers)."""
        return self._handle.ldl_table()

    def dtable(self) -> np.ndarray:
        return self._handle.dtable()


class CasADiSolver(B200Solver):
    """Name-compatible drop-in for ``optas.CasADiSolver`` (ref :321-419): same constructor, same
    ``setup(solver_name, solver_options)``; the work runs on the B200 back-end."""


class ScipyMinimizeSolver(B200Solver):
    """Name-compatible drop-in for ``optas.ScipyMinimizeSolver`` (ref :587-813): same constructor,
    same ``setup(method, tol, options)``; the work runs on the B200 back-end.  (The CPU restatement
    of the reference's scipy path lives in oracle/slsqp_driver.py and is test infrastructure.)"""

    def setup(self, method: str = "SLSQP", tol: Optional[float] = None, options: Optional[Dict] = None, **kw):
        return super().setup(method=method, tol=tol, options=options, **kw)


class OSQPSolver(B200Solver):
    """Name-compatible drop-in for ``optas.OSQPSolver`` (ref :426-507): QP-only, ``setup(use_warm_start,
    settings={})``; the QP is solved by the same GPU interior-point kernel (a convex QP takes a handful
    of Newton iterations).  ``use_warm_start=False`` discards the seed like the reference does."""

    def setup(self, use_warm_start: bool = False, settings: Optional[Dict] = None, **kw):
        from .optimization import QP_COST

        assert self.opt_type in QP_COST, "OSQP cannot solve this type of problem"
        self.use_warm_start = use_warm_start
        return super().setup("osqp", dict(settings or {}), **kw)

    def reset_initial_seed(self, x0: Dict[str, ArrayType]) -> None:
        super().reset_initial_seed(x0 if self.use_warm_start else {})


class CVXOPTSolver(B200Solver):
    """Name-compatible drop-in for ``optas.CVXOPTSolver`` (ref :514-580): QP-only, ``setup(solver_settings={})``."""

    def setup(self, solver_settings: Optional[Dict] = None, **kw):
        from .optimization import QP_COST

        assert self.opt_type in QP_COST, "CVXOPT cannot solve this problem"
        return super().setup("cvxopt", dict(solver_settings or {}), **kw)
