"""Batched evaluation of an expression ``Function`` on the GPU (streaming kernel "K1").

Replaces ``cs.Function(...).map(n)`` evaluation of model functions on the reference path
(optas/models.py:786-787: ``make_function`` builds a CasADi Function and maps it over ``n``
columns, which CasADi then walks serially on one core).  Here the function's tape is lowered to
straight-line sm_100a code and streamed over the batch by libb200optas (bo_function_eval).

Layout: every input / output segment is ``[B, numel]`` float64, each row the column-major
flattening of that argument for one instance (so a 3 x 7 Jacobian output is ``[B, 21]`` with
``out.reshape(B, 7, 3).transpose(0, 2, 1)`` giving ``[B, 3, 7]``).  There is no CPU path.
"""

from __future__ import annotations

from typing import List, Sequence

import numpy as np

from . import _capi
from .tape import Tape


class B200Function:
    def __init__(self, fun, compile_only: bool = False, timing: bool = False, threads_per_block: int = 0):
        self.fun = fun
        self.tape = fun if isinstance(fun, Tape) else Tape.from_function(fun)
        flags = (_capi.BO_FLAG_COMPILE_ONLY if compile_only else 0) | (_capi.BO_FLAG_TIMING if timing else 0)
        self._handle = _capi.FunctionHandle(self.tape, flags=flags, threads_per_block=threads_per_block)
        self.in_sizes: List[int] = list(self.tape.in_sizes)
        self.out_sizes: List[int] = list(self.tape.out_sizes)

    def kernel_info(self) -> dict:
        return self._handle.kernel_info()

    def kernel_source(self) -> str:
        return self._handle.source()

    def kernel_time(self):
        return self._handle.kernel_time()

    def eval_raw(self, B: int, ins: Sequence, outs: Sequence, stream: int = 0) -> None:
        """Direct ``bo_function_eval`` on preallocated float64 C-contiguous buffers (numpy arrays or
        torch tensors, host or cuda).  Asynchronous on ``stream`` when every buffer is on the device."""
        self._handle.eval(B, ins, outs, stream)

    def __call__(self, *args):
        """Host convenience: numpy in, numpy out (copies staged by the library)."""
        if len(args) != len(self.in_sizes):
            raise TypeError(f"expected {len(self.in_sizes)} arguments, got {len(args)}")
        ins = []
        B = None
        for a, n in zip(args, self.in_sizes):
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.ndim == 1:
                a = a.reshape(1, -1)
            if a.shape[1] != n:
                raise ValueError(f"argument has {a.shape[1]} elements per instance, expected {n}")
            B = a.shape[0] if B is None else B
            if a.shape[0] != B:
                raise ValueError("all arguments must have the same batch size")
            ins.append(a)
        outs = [np.empty((B, n)) for n in self.out_sizes]
        self._handle.eval(B, ins, outs)
        return outs[0] if len(outs) == 1 else tuple(outs)
