"""Ordered name -> SX container; defines the memory layout of decision variables and parameters.

Mirror of the reference's optas/sx_container.py:18-130.  Layout contract (F2): ``vec()`` is the
concatenation of the column-major flattening of every entry in insertion order;
``dict2vec`` substitutes zeros for missing labels and ignores unknown labels (ref :113-123,
pinned by tests/test_sx_container.py:41-51); ``vec2dict`` slices and reshapes column-major.
"""

import collections
from typing import Dict, List, Union

import numpy as np

from . import sym as cs
from .sym import DM, SX


class SXContainer(collections.OrderedDict):
    ## class-level on purpose: the reference shares this dict between all containers
    ## (ref sx_container.py:22)
    is_discrete = {}

    def __add__(self, other):
        assert isinstance(other, SXContainer), f"cannot add SXContainer with a variable of type {type(other)}"
        out = SXContainer()
        for label, value in self.items():
            out[label] = value
        for label, value in other.items():
            out[label] = value
        out.is_discrete = {**self.is_discrete, **other.is_discrete}
        return out

    def __setitem__(self, label: str, value) -> None:
        assert isinstance(value, (SX, float)), f"value must be of type SX/float, not {type(value)}"
        if label in self:
            raise KeyError(f"'{label}' already exists")
        super().__setitem__(label, SX(value))
        self.is_discrete[label] = False

    def variable_is_discrete(self, label: str) -> None:
        assert label in self, f"'{label}' was not found"
        self.is_discrete[label] = True

    def has_discrete_variables(self) -> bool:
        return any(self.is_discrete.values())

    def discrete(self) -> List[bool]:
        out = []
        for label, value in self.items():
            out += [self.is_discrete[label]] * value.numel()
        return out

    def vec(self) -> SX:
        return SX(cs.vertcat(*[cs.vec(v) for v in self.values()])) if len(self) else SX(0, 1)

    def numel(self) -> int:
        return sum(v.numel() for v in self.values())

    def offsets(self) -> Dict[str, tuple]:
        """label -> (offset, rows, cols) inside ``vec()`` (used by the batched marshalling)."""
        out, off = {}, 0
        for label, value in self.items():
            m, n = value.shape
            out[label] = (off, m, n)
            off += m * n
        return out

    def vec2dict(self, vec_) -> dict:
        v = DM(vec_)
        flat = v._a.flatten(order="F")
        out, off = {}, 0
        for label, value in self.items():
            m, n = value.shape
            out[label] = DM(flat[off:off + m * n].reshape((m, n), order="F"))
            off += m * n
        return out

    def dict2vec(self, d: Dict) -> Union[DM, SX]:
        parts = []
        for label, value in self.items():
            v = d.get(label)
            if v is None:
                v = DM.zeros(*value.shape)
            v = v if isinstance(v, (DM, SX)) else DM(v)
            if v.numel() != value.numel():
                raise ValueError(f"'{label}': expected {value.numel()} elements, got {v.numel()}")
            parts.append(cs.vec(v))
        if not parts:
            return DM.zeros(0, 1)
        out = cs.vertcat(*parts)
        if out.shape == (0, 0):
            out = DM.zeros(0, 1)
        return out

    def zero(self) -> dict:
        return {label: DM.zeros(*value.shape) for label, value in self.items()}
