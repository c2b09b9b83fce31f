"""ORACLE (test infrastructure): tape interpreters -- the CPU stand-in for CasADi's SX virtual
machine (what evaluates ``opt.f / df / v / dv`` inside the reference's solver callbacks,
optas/solver.py:716-734).  Two independent implementations of the tape semantics documented in
include/b200optas.h:

* ``eval_numpy``  -- vectorised numpy, written against the opcode table only;
* ``CTape``       -- ctypes wrapper of oracle/tape_vm.c (scalar C, gcc -O3), used for the timed
                     CPU baseline.

Neither is imported by ``optas_b200``.
"""

from __future__ import annotations

import ctypes as C
import os
import re
from typing import List, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libtape_vm.so")


def _opcodes() -> dict:
    text = open(os.path.join(_HERE, "..", "include", "bo_opcodes.h")).read()
    return {m.group(1): int(m.group(2)) for m in re.finditer(r"#define BO_OP_(\w+) (\d+)", text)}


OP = _opcodes()
_UN = {
    "NEG": np.negative, "SQ": np.square, "SQRT": np.sqrt, "SIN": np.sin, "COS": np.cos, "TAN": np.tan,
    "ASIN": np.arcsin, "ACOS": np.arccos, "ATAN": np.arctan, "FABS": np.fabs, "EXP": np.exp, "LOG": np.log,
    "NOT": lambda a: (a == 0.0).astype(float), "SIGN": np.sign, "FLOOR": np.floor, "CEIL": np.ceil,
    "TANH": np.tanh, "SINH": np.sinh, "COSH": np.cosh,
}
_BI = {
    "ADD": np.add, "SUB": np.subtract, "MUL": np.multiply, "DIV": np.divide, "ATAN2": np.arctan2,
    "FMIN": np.fmin, "FMAX": np.fmax, "POW": np.power,
    "LT": lambda a, b: (a < b).astype(float), "LE": lambda a, b: (a <= b).astype(float),
    "EQ": lambda a, b: (a == b).astype(float), "NE": lambda a, b: (a != b).astype(float),
    "AND": lambda a, b: ((a != 0.0) & (b != 0.0)).astype(float), "OR": lambda a, b: ((a != 0.0) | (b != 0.0)).astype(float),
}
_UN_BY_CODE = {OP[k]: v for k, v in _UN.items()}
_BI_BY_CODE = {OP[k]: v for k, v in _BI.items()}


def eval_numpy(tape, inputs: Sequence[np.ndarray]) -> List[np.ndarray]:
    """inputs[k]: [B, in_sizes[k]] -> outputs[k]: [B, out_sizes[k]]."""
    ins = [np.atleast_2d(np.asarray(a, dtype=float)) for a in inputs]
    B = max([a.shape[0] for a in ins] + [1])
    outs = [np.zeros((B, n)) for n in tape.out_sizes]
    work = [None] * int(tape.n_work)
    with np.errstate(all="ignore"):
        for op, dst, a, b in np.asarray(tape.instr).tolist():
            c, op = op >> 8, op & 0xFF
            if op == OP["INPUT"]:
                work[dst] = ins[b][:, a]
            elif op == OP["CONST"]:
                work[dst] = np.full(B, tape.consts[a])
            elif op == OP["OUTPUT"]:
                outs[b][:, a] = work[dst]
            elif op == OP["IF_ELSE"]:
                work[dst] = np.where(work[c] != 0.0, work[a], work[b])
            elif op in _UN_BY_CODE:
                work[dst] = _UN_BY_CODE[op](work[a])
            else:
                work[dst] = _BI_BY_CODE[op](work[a], work[b])
    return outs


class CTape:
    """One tape bound to the C interpreter."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(_SO):
                raise ImportError(f"{_SO} missing: run `make -C oracle`")
            cls._lib = C.CDLL(_SO)
            vp = C.c_void_p
            cls._lib.tape_eval_batch.argtypes = [vp, C.c_int64, vp, vp, C.c_int64, C.c_int32, vp, vp, C.c_int32, vp, vp]
        return cls._lib

    def __init__(self, tape):
        self.instr = np.ascontiguousarray(tape.instr, dtype=np.int32)
        self.consts = np.ascontiguousarray(tape.consts, dtype=np.float64)
        self.work = np.zeros(int(tape.n_work))
        self.in_sizes = np.ascontiguousarray(tape.in_sizes, dtype=np.int32)
        self.out_sizes = np.ascontiguousarray(tape.out_sizes, dtype=np.int32)
        assert len(self.in_sizes) <= 32 and len(self.out_sizes) <= 32
        self._in_ptrs = (C.c_void_p * len(self.in_sizes))()
        self._out_ptrs = (C.c_void_p * len(self.out_sizes))()

    def __call__(self, *inputs) -> List[np.ndarray]:
        ins = [np.ascontiguousarray(np.atleast_2d(a), dtype=np.float64) for a in inputs]
        B = max([a.shape[0] for a, n in zip(ins, self.in_sizes) if n > 0] + [1])
        outs = [np.zeros((B, int(n))) for n in self.out_sizes]
        for k, a in enumerate(ins):
            self._in_ptrs[k] = a.ctypes.data
        for k, a in enumerate(outs):
            self._out_ptrs[k] = a.ctypes.data
        self.lib().tape_eval_batch(self.instr.ctypes.data, self.instr.shape[0], self.consts.ctypes.data,
                                   self.work.ctypes.data, B, len(self.in_sizes), self.in_sizes.ctypes.data,
                                   self._in_ptrs, len(self.out_sizes), self.out_sizes.ctypes.data, self._out_ptrs)
        return outs
