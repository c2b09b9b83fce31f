"""ctypes binding of libb200optas.so (include/b200optas.h).  Thin on purpose: every structure
here is a field-for-field image of the C header, and every call releases the GIL.

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C optas_b200/csrc``; if it
is missing, importing this module raises -- there is no Python or CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from .tape import Tape

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200optas.so")

BO_ABI_VERSION = 1
BO_OK, BO_ERR_INVALID, BO_ERR_COMPILE, BO_ERR_CUDA, BO_ERR_NO_DEVICE, BO_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
BO_FLAG_COMPILE_ONLY, BO_FLAG_VERBOSE, BO_FLAG_NO_CACHE, BO_FLAG_TIMING, BO_FLAG_PIVOTED_LDL = 1, 2, 4, 8, 16
BO_FLAG_COOP, BO_FLAG_NO_COOP, BO_FLAG_NO_TEAM, BO_FLAG_PIPELINE, BO_FLAG_TEAM, BO_FLAG_NO_QP = 32, 64, 128, 256, 512, 1024
TIER_NAMES = {0: "dense", 1: "sparse", 2: "large", 3: "coop", 4: "team", 5: "qp", 6: "qp_sparse"}
STATUS_NAMES = {0: "converged", 1: "acceptable", 2: "max_iter", 3: "line_search", 4: "numerical"}

EXPORTS = [
    "bo_abi_version", "bo_last_error", "bo_device_count",
    "bo_problem_create", "bo_problem_destroy", "bo_problem_source", "bo_problem_ldl_table", "bo_problem_dtable", "bo_problem_kernel_info",
    "bo_problem_tier_info", "bo_problem_options",
    "bo_solve", "bo_problem_kernel_time", "bo_pack_rows",
    "bo_function_create", "bo_function_destroy", "bo_function_eval", "bo_function_source",
    "bo_function_kernel_info", "bo_function_kernel_time",
]


class BoError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libb200optas error {code}: {message}")
        self.code = code


class bo_tape(C.Structure):
    _fields_ = [
        ("instr", C.POINTER(C.c_int32)), ("n_instr", C.c_int64),
        ("consts", C.POINTER(C.c_double)), ("n_consts", C.c_int32),
        ("n_work", C.c_int32),
        ("n_in", C.c_int32), ("in_sizes", C.POINTER(C.c_int32)),
        ("n_out", C.c_int32), ("out_sizes", C.POINTER(C.c_int32)),
    ]


class bo_sparsity(C.Structure):
    _fields_ = [("nnz", C.c_int32), ("row", C.POINTER(C.c_int32)), ("col", C.POINTER(C.c_int32))]


class bo_problem_desc(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("np", C.c_int32), ("n_eq", C.c_int32), ("n_ineq", C.c_int32),
        ("fc", bo_tape), ("kkt", bo_tape),
        ("jac_eq", bo_sparsity), ("jac_ineq", bo_sparsity), ("hess", bo_sparsity),
    ]


class bo_options(C.Structure):
    _fields_ = [
        ("flags", C.c_uint32), ("max_iter", C.c_int32),
        ("tol", C.c_double), ("acceptable_tol", C.c_double), ("mu_init", C.c_double), ("max_step", C.c_double),
        ("cache_dir", C.c_char_p), ("include_dir", C.c_char_p),
        ("threads_per_block", C.c_int32), ("max_trips", C.c_int32), ("blocks_per_sm", C.c_int32), ("device", C.c_int32),
        ("reserved", C.c_int32 * 4),
    ]


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C optas_b200/csrc` (there is no fallback path)")
    lib = C.CDLL(LIB_PATH)
    vp, i32p, f64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)
    lib.bo_abi_version.restype = C.c_int
    lib.bo_last_error.restype = C.c_char_p
    lib.bo_device_count.restype = C.c_int
    lib.bo_problem_create.argtypes = [C.POINTER(bo_problem_desc), C.POINTER(bo_options), C.POINTER(vp)]
    lib.bo_problem_destroy.argtypes = [vp]
    lib.bo_problem_source.argtypes = [vp, C.c_char_p, C.c_int64]
    lib.bo_problem_source.restype = C.c_int64
    lib.bo_problem_ldl_table.argtypes = [vp, i32p, C.c_int64]
    lib.bo_problem_ldl_table.restype = C.c_int64
    lib.bo_problem_dtable.argtypes = [vp, f64p, C.c_int64]
    lib.bo_problem_dtable.restype = C.c_int64
    lib.bo_problem_kernel_info.argtypes = [vp, i32p, i32p, i32p]
    lib.bo_problem_tier_info.argtypes = [vp, C.POINTER(C.c_int64), C.c_int32]
    lib.bo_problem_options.argtypes = [vp, C.POINTER(bo_options)]
    lib.bo_solve.argtypes = [vp, C.c_int64] + [vp] * 8 + [vp]
    lib.bo_problem_kernel_time.argtypes = [vp, f64p, C.POINTER(C.c_int64)]
    lib.bo_pack_rows.argtypes = [vp, C.c_int64, C.c_int64, C.c_int32, C.POINTER(vp), C.POINTER(C.c_int64), i32p, i32p, C.POINTER(C.c_int64), C.c_int32]
    lib.bo_function_create.argtypes = [C.POINTER(bo_tape), C.POINTER(bo_options), C.POINTER(vp)]
    lib.bo_function_destroy.argtypes = [vp]
    lib.bo_function_eval.argtypes = [vp, C.c_int64, C.POINTER(vp), C.POINTER(vp), vp]
    lib.bo_function_source.argtypes = [vp, C.c_char_p, C.c_int64]
    lib.bo_function_source.restype = C.c_int64
    lib.bo_function_kernel_info.argtypes = [vp, i32p, i32p, i32p]
    lib.bo_function_kernel_time.argtypes = [vp, f64p, C.POINTER(C.c_int64)]
    if lib.bo_abi_version() != BO_ABI_VERSION:
        raise ImportError(f"libb200optas ABI {lib.bo_abi_version()} != binding ABI {BO_ABI_VERSION}")
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != BO_OK:
        raise BoError(code, load().bo_last_error().decode(errors="replace"))


def device_count() -> int:
    return int(load().bo_device_count())


class _Keep:
    """Keeps the numpy arrays behind a C structure alive."""

    def __init__(self):
        self.refs: List = []

    def i32(self, a) -> C.POINTER(C.c_int32):
        arr = np.ascontiguousarray(a, dtype=np.int32)
        self.refs.append(arr)
        return arr.ctypes.data_as(C.POINTER(C.c_int32))

    def f64(self, a) -> C.POINTER(C.c_double):
        arr = np.ascontiguousarray(a, dtype=np.float64)
        self.refs.append(arr)
        return arr.ctypes.data_as(C.POINTER(C.c_double))


def tape_struct(t: Tape, keep: _Keep) -> bo_tape:
    return bo_tape(
        instr=keep.i32(t.instr.reshape(-1)), n_instr=t.n_instr,
        consts=keep.f64(t.consts), n_consts=int(t.consts.shape[0]),
        n_work=int(t.n_work),
        n_in=len(t.in_sizes), in_sizes=keep.i32(t.in_sizes),
        n_out=len(t.out_sizes), out_sizes=keep.i32(t.out_sizes),
    )


def options_struct(flags: int = 0, max_iter: int = 0, tol: float = 0.0, acceptable_tol: float = 0.0,
                   mu_init: float = 0.0, max_step: float = 0.0, cache_dir: Optional[str] = None, include_dir: Optional[str] = None,
                   threads_per_block: int = 0, max_trips: int = 0, blocks_per_sm: int = 0, device: Optional[int] = None) -> bo_options:
    return bo_options(
        flags=flags, max_iter=max_iter, tol=tol, acceptable_tol=acceptable_tol, mu_init=mu_init, max_step=max_step,
        cache_dir=cache_dir.encode() if cache_dir else None,
        include_dir=include_dir.encode() if include_dir else None,
        threads_per_block=threads_per_block, max_trips=max_trips, blocks_per_sm=blocks_per_sm,
        device=_current_device_plus_one() if device is None else int(device) + 1,
    )


def _current_device_plus_one() -> int:
    """Device ordinal + 1 the caller has selected through torch (``torch.cuda.set_device`` is lazy: no context may be
    current yet, and the library would otherwise bind the handle to device 0); 0 = let the library look at the thread's
    current context."""
    import sys

    torch = sys.modules.get("torch")
    try:
        if torch is not None and torch.cuda.is_available():
            return int(torch.cuda.current_device()) + 1
    except Exception:
        pass
    return 0


def pack_rows(dst: np.ndarray, segments) -> None:
    """``bo_pack_rows``: ``dst`` is ``[B, total]`` float64 C-contiguous; ``segments`` is a list of
    ``(array or None, (batch, row, column) strides in elements, m, n, offset)``; arrays are float64 views of any layout."""
    B, total = dst.shape
    k = len(segments)
    ptrs = (C.c_void_p * max(k, 1))(*[None if a is None else a.ctypes.data for a, _, _, _, _ in segments])
    strides = (C.c_int64 * max(3 * k, 1))(*[int(v) for _, st, _, _, _ in segments for v in st])
    ms = (C.c_int32 * max(k, 1))(*[int(m) for _, _, m, _, _ in segments])
    ns = (C.c_int32 * max(k, 1))(*[int(n) for _, _, _, n, _ in segments])
    offs = (C.c_int64 * max(k, 1))(*[int(o) for _, _, _, _, o in segments])
    check(load().bo_pack_rows(dst.ctypes.data, B, total, k, ptrs, strides, ms, ns, offs, 0))


def _source(getter, handle) -> str:
    n = getter(handle, None, 0)
    buf = C.create_string_buffer(int(n) + 1)
    getter(handle, buf, int(n) + 1)
    return buf.value.decode()


def _kernel_info(getter, handle) -> dict:
    r, l, s = C.c_int32(-1), C.c_int32(-1), C.c_int32(-1)
    check(getter(handle, C.byref(r), C.byref(l), C.byref(s)))
    return {"registers": r.value, "local_bytes": l.value, "smem_bytes": s.value}


def _ptr(a) -> Optional[int]:
    """Raw address of a numpy array / torch tensor / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return int(a.data_ptr())
    raise TypeError(f"cannot take the address of {type(a)}")


def _check_buffer(name: str, a, rows: int, width: Optional[int], kind: str) -> None:
    """The library takes raw addresses: refuse anything that is not a C-contiguous ``[rows, width]`` (or ``[rows]``)
    array of the expected element type -- a float32 or transposed tensor would be read / written out of bounds."""
    if a is None:
        return
    want = {"f64": "float64", "i32": "int32"}[kind]
    dt = str(a.dtype).replace("torch.", "")
    if dt != want:
        raise ValueError(f"{name}: expected {want}, got {dt}")
    contiguous = a.flags.c_contiguous if isinstance(a, np.ndarray) else bool(a.is_contiguous())
    if not contiguous:
        raise ValueError(f"{name}: must be C-contiguous")
    shape = tuple(int(d) for d in a.shape)
    numel = int(np.prod(shape)) if shape else 1
    need = rows * (width if width is not None else 1)
    if numel != need or (len(shape) >= 1 and shape[0] != rows and need > 0):
        raise ValueError(f"{name}: expected {rows} rows of {width if width is not None else 1} elements, got shape {list(shape)}")


def _check_devices(arrays) -> None:
    devs = {str(a.device) for a in arrays if a is not None and hasattr(a, "data_ptr") and getattr(a, "is_cuda", False)}
    if len(devs) > 1:
        raise ValueError(f"buffers live on different devices: {sorted(devs)}")


class ProblemHandle:
    """Owns a ``bo_problem*``."""

    def __init__(self, lowered, **opts):
        lib = load()
        keep = _Keep()
        desc = bo_problem_desc(
            nx=lowered.nx, np=lowered.np_, n_eq=lowered.n_eq, n_ineq=lowered.n_ineq,
            fc=tape_struct(lowered.fc, keep), kkt=tape_struct(lowered.kkt, keep),
            jac_eq=bo_sparsity(lowered.jac_eq.nnz, keep.i32(lowered.jac_eq.row), keep.i32(lowered.jac_eq.col)),
            jac_ineq=bo_sparsity(lowered.jac_ineq.nnz, keep.i32(lowered.jac_ineq.row), keep.i32(lowered.jac_ineq.col)),
            hess=bo_sparsity(lowered.hess.nnz, keep.i32(lowered.hess.row), keep.i32(lowered.hess.col)),
        )
        o = options_struct(**opts)
        h = C.c_void_p()
        check(lib.bo_problem_create(C.byref(desc), C.byref(o), C.byref(h)))
        self._h = h
        self.lowered = lowered

    def source(self) -> str:
        return _source(load().bo_problem_source, self._h)

    def ldl_table(self) -> np.ndarray:
        n = load().bo_problem_ldl_table(self._h, None, 0)
        out = np.zeros(max(int(n), 1), dtype=np.int32)
        if n > 0:
            load().bo_problem_ldl_table(self._h, out.ctypes.data_as(C.POINTER(C.c_int32)), int(n))
        return out[:int(n)]

    def dtable(self) -> np.ndarray:
        n = load().bo_problem_dtable(self._h, None, 0)
        out = np.zeros(max(int(n), 1))
        if n > 0:
            load().bo_problem_dtable(self._h, out.ctypes.data_as(C.POINTER(C.c_double)), int(n))
        return out[:int(n)]

    def kernel_info(self) -> dict:
        return _kernel_info(load().bo_problem_kernel_info, self._h)

    def tier_info(self) -> dict:
        v = (C.c_int64 * 32)()
        check(load().bo_problem_tier_info(self._h, v, 32))
        keys = ["tier", "threads_per_block", "smem_dynamic", "levels", "nsub_fc", "nsub_kkt", "n_pe_kkt", "n_partial",
                "kkt_max_len", "kkt_total_instr", "ldl_g", "solve_g", "factor_madds", "n_work", "kkt_components",
                "fc_max_len", "fc_total_instr", "kkt_pre_len", "scratch_doubles", "factor_vals", "blocks_per_sm", "n_sm",
                "n_work_fc", "ldl_warps", "factor_steps", "solve_steps", "kkt_wstride", "fc_wstride", "segments",
                "generated_tapes", "kkt_classes", "kkt_code_rows"]
        d = {k: int(v[i]) for i, k in enumerate(keys)}
        d["tier"] = TIER_NAMES[d["tier"]]
        return d

    def options(self) -> dict:
        """The options in effect after the library resolved its defaults."""
        o = bo_options()
        check(load().bo_problem_options(self._h, C.byref(o)))
        return {k: getattr(o, k) for k in ("flags", "max_iter", "tol", "acceptable_tol", "mu_init", "max_step",
                                           "threads_per_block", "max_trips", "blocks_per_sm", "device")}

    def solve(self, B: int, p, x0, x, lam=None, f=None, status=None, iters=None, kkt=None, stream: int = 0) -> None:
        lo = self.lowered
        if lo.np_ > 0 and p is None:
            raise ValueError("p: required (np > 0)")
        _check_buffer("p", p if lo.np_ > 0 else None, B, lo.np_, "f64")
        _check_buffer("x0", x0, B, lo.nx, "f64")
        _check_buffer("x", x, B, lo.nx, "f64")
        _check_buffer("lam", lam, B, lo.n_eq + lo.n_ineq, "f64")
        _check_buffer("f", f, B, None, "f64")
        _check_buffer("status", status, B, None, "i32")
        _check_buffer("iters", iters, B, None, "i32")
        _check_buffer("kkt", kkt, B, None, "f64")
        _check_devices([p, x0, x, lam, f, status, iters, kkt])
        check(load().bo_solve(self._h, B, _ptr(p), _ptr(x0), _ptr(x), _ptr(lam), _ptr(f), _ptr(status), _ptr(iters),
                              _ptr(kkt), stream or None))

    def kernel_time(self):
        ms, n = C.c_double(0.0), C.c_int64(0)
        check(load().bo_problem_kernel_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def close(self) -> None:
        if getattr(self, "_h", None):
            load().bo_problem_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FunctionHandle:
    """Owns a ``bo_function*`` (streaming evaluation of one tape)."""

    def __init__(self, tape: Tape, **opts):
        lib = load()
        keep = _Keep()
        ts = tape_struct(tape, keep)
        o = options_struct(**opts)
        h = C.c_void_p()
        check(lib.bo_function_create(C.byref(ts), C.byref(o), C.byref(h)))
        self._h = h
        self.tape = tape

    def source(self) -> str:
        return _source(load().bo_function_source, self._h)

    def kernel_info(self) -> dict:
        return _kernel_info(load().bo_function_kernel_info, self._h)

    def eval(self, B: int, ins: Sequence, outs: Sequence, stream: int = 0) -> None:
        n_in, n_out = len(self.tape.in_sizes), len(self.tape.out_sizes)
        if len(ins) != n_in or len(outs) != n_out:
            raise ValueError(f"expected {n_in} inputs and {n_out} outputs")
        for k, a in enumerate(ins):
            _check_buffer(f"in[{k}]", a, B, int(self.tape.in_sizes[k]), "f64")
        for k, a in enumerate(outs):
            _check_buffer(f"out[{k}]", a, B, int(self.tape.out_sizes[k]), "f64")
        _check_devices(list(ins) + list(outs))
        in_arr = (C.c_void_p * n_in)(*[_ptr(a) for a in ins])
        out_arr = (C.c_void_p * n_out)(*[_ptr(a) for a in outs])
        check(load().bo_function_eval(self._h, B, in_arr, out_arr, stream or None))

    def kernel_time(self):
        ms, n = C.c_double(0.0), C.c_int64(0)
        check(load().bo_function_kernel_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def close(self) -> None:
        if getattr(self, "_h", None):
            load().bo_function_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
