"""TEST SHIM (not part of the optas_b200 package; SURVEY.md section 2 marks optas/templates.py out of scope): exists only so
that tests/test_scripts_unchanged.py can run the reference's example scripts, which subclass it, without editing them.

Loop glue between a task script and a solver: the ``Manager`` base class the reference's example scripts derive
their planners / controllers from (reference: optas/templates.py:10-105).  Host orchestration only -- it owns a
``Solver`` (here: the B200 back-end behind ``CasADiSolver`` & co.), calls ``solve()`` and keeps the last solution.

The ROS managers of the reference (templates.py:108-319: publishers / subscribers around the same four hooks) are not
mirrored: they need a ROS runtime and add nothing on the solver path.
"""

from __future__ import annotations

import abc
import time
from typing import Callable, Dict, Union

import yaml


class Manager(abc.ABC):
    """Derive, implement ``setup_solver`` (build the problem, return the solver), ``is_ready``, ``reset`` (feed
    parameters / seed) and ``get_target`` (pick the interesting part of ``self.solution``); then call ``solve()``."""

    def __init__(self, config_filename: Union[None, str] = None, record_solver_perf: bool = False):
        self.reset_manager()
        self.config_filename = config_filename
        self.record_solver_perf = record_solver_perf
        self.config = self._load_configuration(config_filename)
        self.solver = self.setup_solver()
        self.solve: Callable[[], None] = self._solve_and_time if record_solver_perf else self._solve

    def reset_manager(self) -> None:
        self.num_solves = 0
        self.solver_duration = None
        self.solution = None

    @staticmethod
    def _load_configuration(filename) -> Dict:
        if not filename:
            return {}
        with open(filename, "rb") as fh:
            return yaml.load(fh, Loader=yaml.FullLoader)

    def _solve(self) -> None:
        self.solution = self.solver.solve()
        self.num_solves += 1

    def _solve_and_time(self) -> None:
        t0 = time.perf_counter()
        self._solve()
        self.solver_duration = time.perf_counter() - t0

    def get_solver_duration(self) -> float:
        return self.solver_duration

    def is_first_solve(self) -> bool:
        return self.num_solves == 0

    @abc.abstractmethod
    def setup_solver(self):
        ...

    @abc.abstractmethod
    def is_ready(self) -> bool:
        ...

    @abc.abstractmethod
    def reset(self) -> None:
        ...

    @abc.abstractmethod
    def get_target(self):
        ...
