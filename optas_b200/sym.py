"""Casadi-free symbolic layer: scalar expression DAG + 2-D array wrappers.

The reference builds every cost / constraint / kinematics expression as a CasADi ``SX``
graph and hands numeric data round as ``DM`` (reference: optas/__init__.py:2 re-exports the
whole casadi namespace).  CasADi is not installable in this image, so this module provides the
slice of that surface which ``optas/{spatialmath,models,builder,optimization,sx_container,
solver}.py`` and the example scripts rely on (census in SURVEY.md section 9, items F1-F14):

* ``SX`` - always-2-D array of scalar expression nodes (hash-consed DAG => automatic CSE,
  constant folding, structural zeros).
* ``DM`` - always-2-D float64 array (thin wrapper over numpy).
* free functions ``vertcat/horzcat/vec/reshape/sumsqr/jacobian/is_linear/...`` and ``Function``
  (callable symbolically = substitution, numerically = tape evaluation, ``.map(n)``).

Design choices that differ from CasADi on purpose (this is the front half of a GPU backend,
not a CasADi clone): derivatives are taken by one sparse forward sweep over the DAG
(partials kept as ``{input index: node}`` dicts), and ``Function`` lowers itself to the
same SSA tape format (`optas_b200.tape`) that the CUDA virtual machine executes.
"""

from __future__ import annotations

import math
import numbers
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

# --------------------------------------------------------------------------------------
# opcodes (shared with optas_b200/csrc/bo_opcodes.h and oracle/tape_vm.c; a test checks
# that the three agree)
# --------------------------------------------------------------------------------------

OP_CONST = 1
OP_INPUT = 2
OP_OUTPUT = 3
# binary
OP_ADD = 10
OP_SUB = 11
OP_MUL = 12
OP_DIV = 13
OP_ATAN2 = 14
OP_FMIN = 15
OP_FMAX = 16
OP_POW = 17
OP_LT = 18
OP_LE = 19
OP_EQ = 20
OP_NE = 21
OP_AND = 22
OP_OR = 23
# unary
OP_NEG = 40
OP_SQ = 41
OP_SQRT = 42
OP_SIN = 43
OP_COS = 44
OP_TAN = 45
OP_ASIN = 46
OP_ACOS = 47
OP_ATAN = 48
OP_FABS = 49
OP_EXP = 50
OP_LOG = 51
OP_NOT = 52
OP_SIGN = 53
OP_FLOOR = 54
OP_CEIL = 55
OP_TANH = 56
OP_SINH = 57
OP_COSH = 58
# ternary
OP_IF_ELSE = 70
# leaf (symbolic only; never reaches a tape)
OP_SYM = 0

BINARY_OPS = {OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_ATAN2, OP_FMIN, OP_FMAX, OP_POW, OP_LT,
              OP_LE, OP_EQ, OP_NE, OP_AND, OP_OR}
UNARY_OPS = {OP_NEG, OP_SQ, OP_SQRT, OP_SIN, OP_COS, OP_TAN, OP_ASIN, OP_ACOS, OP_ATAN,
             OP_FABS, OP_EXP, OP_LOG, OP_NOT, OP_SIGN, OP_FLOOR, OP_CEIL, OP_TANH, OP_SINH,
             OP_COSH}

OP_NAMES = {v: k[3:] for k, v in list(globals().items()) if k.startswith("OP_")}


def _sign(x):
    return (x > 0) - (x < 0)


_NUM_UNARY = {
    OP_NEG: lambda a: -a,
    OP_SQ: lambda a: a * a,
    OP_SQRT: np.sqrt,
    OP_SIN: np.sin,
    OP_COS: np.cos,
    OP_TAN: np.tan,
    OP_ASIN: np.arcsin,
    OP_ACOS: np.arccos,
    OP_ATAN: np.arctan,
    OP_FABS: np.fabs,
    OP_EXP: np.exp,
    OP_LOG: np.log,
    OP_NOT: lambda a: np.asarray(a == 0.0, dtype=float),
    OP_SIGN: np.sign,
    OP_FLOOR: np.floor,
    OP_CEIL: np.ceil,
    OP_TANH: np.tanh,
    OP_SINH: np.sinh,
    OP_COSH: np.cosh,
}

_NUM_BINARY = {
    OP_ADD: lambda a, b: a + b,
    OP_SUB: lambda a, b: a - b,
    OP_MUL: lambda a, b: a * b,
    OP_DIV: lambda a, b: a / b,
    OP_ATAN2: np.arctan2,
    OP_FMIN: np.fmin,
    OP_FMAX: np.fmax,
    OP_POW: np.power,
    OP_LT: lambda a, b: np.asarray(a < b, dtype=float),
    OP_LE: lambda a, b: np.asarray(a <= b, dtype=float),
    OP_EQ: lambda a, b: np.asarray(a == b, dtype=float),
    OP_NE: lambda a, b: np.asarray(a != b, dtype=float),
    OP_AND: lambda a, b: np.asarray((a != 0.0) & (b != 0.0), dtype=float),
    OP_OR: lambda a, b: np.asarray((a != 0.0) | (b != 0.0), dtype=float),
}


# --------------------------------------------------------------------------------------
# scalar nodes
# --------------------------------------------------------------------------------------


class Node:
    """One scalar expression.  ``idx`` increases with creation, so sorting by ``idx`` is a
    topological order (children are always created before their parents)."""

    __slots__ = ("op", "a", "b", "c", "val", "name", "idx")
    _count = 0

    def __init__(self, op, a=None, b=None, c=None, val=None, name=None):
        self.op = op
        self.a = a
        self.b = b
        self.c = c
        self.val = val
        self.name = name
        self.idx = Node._count
        Node._count += 1

    @property
    def is_const(self):
        return self.op == OP_CONST

    def __repr__(self):
        if self.op == OP_CONST:
            return repr(self.val)
        if self.op == OP_SYM:
            return self.name
        if self.op in UNARY_OPS:
            return f"{OP_NAMES[self.op].lower()}({self.a!r})"
        if self.op == OP_IF_ELSE:
            return f"if_else({self.a!r},{self.b!r},{self.c!r})"
        return f"{OP_NAMES[self.op].lower()}({self.a!r},{self.b!r})"


_const_table: Dict[float, Node] = {}
_expr_table: Dict[tuple, Node] = {}


def const(v) -> Node:
    v = float(v)
    key = v if v == v else "nan"
    if v == 0.0:
        key = 0.0 if math.copysign(1.0, v) > 0 else "-0"
    n = _const_table.get(key)
    if n is None:
        n = Node(OP_CONST, val=v)
        _const_table[key] = n
    return n


ZERO = const(0.0)
ONE = const(1.0)
TWO = const(2.0)
MINUS_ONE = const(-1.0)


def sym(name: str) -> Node:
    return Node(OP_SYM, name=name)


def _intern(op, a, b=None, c=None) -> Node:
    key = (op, a.idx, -1 if b is None else b.idx, -1 if c is None else c.idx)
    n = _expr_table.get(key)
    if n is None:
        n = Node(op, a, b, c)
        _expr_table[key] = n
    return n


def _is_zero(n: Node) -> bool:
    return n.op == OP_CONST and n.val == 0.0


def _is_one(n: Node) -> bool:
    return n.op == OP_CONST and n.val == 1.0


def _is_mone(n: Node) -> bool:
    return n.op == OP_CONST and n.val == -1.0


def n_unary(op: int, a: Node) -> Node:
    if a.op == OP_CONST:
        with np.errstate(all="ignore"):
            return const(float(_NUM_UNARY[op](a.val)))
    if op == OP_NEG:
        if a.op == OP_NEG:
            return a.a
        if a.op == OP_SUB:
            return n_binary(OP_SUB, a.b, a.a)
    if op == OP_SQ and a.op == OP_NEG:
        return _intern(OP_SQ, a.a)
    if op == OP_FABS and a.op in (OP_FABS, OP_SQ):
        return a
    if op == OP_COS and a.op == OP_NEG:
        return _intern(OP_COS, a.a)
    return _intern(op, a)


def n_binary(op: int, a: Node, b: Node) -> Node:
    if a.op == OP_CONST and b.op == OP_CONST:
        with np.errstate(all="ignore"):
            return const(float(_NUM_BINARY[op](a.val, b.val)))
    if op == OP_ADD:
        if _is_zero(a):
            return b
        if _is_zero(b):
            return a
        if b.op == OP_NEG:
            return n_binary(OP_SUB, a, b.a)
        if a.op == OP_NEG:
            return n_binary(OP_SUB, b, a.a)
        if a.idx > b.idx:  # canonical order for commutative ops (better CSE)
            a, b = b, a
    elif op == OP_SUB:
        if _is_zero(b):
            return a
        if _is_zero(a):
            return n_unary(OP_NEG, b)
        if a is b:
            return ZERO
        if b.op == OP_NEG:
            return n_binary(OP_ADD, a, b.a)
    elif op == OP_MUL:
        if _is_zero(a) or _is_zero(b):
            return ZERO
        if _is_one(a):
            return b
        if _is_one(b):
            return a
        if _is_mone(a):
            return n_unary(OP_NEG, b)
        if _is_mone(b):
            return n_unary(OP_NEG, a)
        if a is b:
            return n_unary(OP_SQ, a)
        if a.op == OP_NEG and b.op == OP_NEG:
            return n_binary(OP_MUL, a.a, b.a)
        if a.op == OP_NEG:
            return n_unary(OP_NEG, n_binary(OP_MUL, a.a, b))
        if b.op == OP_NEG:
            return n_unary(OP_NEG, n_binary(OP_MUL, a, b.a))
        if a.idx > b.idx:
            a, b = b, a
    elif op == OP_DIV:
        if _is_zero(a):
            return ZERO
        if _is_one(b):
            return a
        if _is_mone(b):
            return n_unary(OP_NEG, a)
        if a is b:
            return ONE
    elif op == OP_POW:
        if b.op == OP_CONST:
            if b.val == 1.0:
                return a
            if b.val == 2.0:
                return n_unary(OP_SQ, a)
            if b.val == 0.0:
                return ONE
            if b.val == 0.5:
                return n_unary(OP_SQRT, a)
            if b.val == -1.0:
                return n_binary(OP_DIV, ONE, a)
    return _intern(op, a, b)


def n_if_else(c: Node, a: Node, b: Node) -> Node:
    if c.op == OP_CONST:
        return a if c.val != 0.0 else b
    if a is b:
        return a
    return _intern(OP_IF_ELSE, c, a, b)


# --------------------------------------------------------------------------------------
# array wrappers
# --------------------------------------------------------------------------------------

ArrayLike = Union["SX", "DM", numbers.Number, Sequence, np.ndarray]


def _as_2d_float(x) -> np.ndarray:
    """numbers / lists / ndarrays -> 2-D float64 (1-D becomes a COLUMN, as casadi.DM(list))."""
    if isinstance(x, DM):
        return x._a
    a = np.array(x, dtype=float)
    if a.ndim == 0:
        return a.reshape(1, 1)
    if a.ndim == 1:
        return a.reshape(-1, 1)
    if a.ndim == 2:
        return a
    raise ValueError(f"cannot make a 2-D array from shape {a.shape}")


def _normalize_index(key, m, n):
    """Return (rows, cols, mode): mode 'lin' for single (column-major linear) indexing."""

    def ax(k, size):
        if isinstance(k, slice):
            return list(range(*k.indices(size))), False
        if isinstance(k, (list, tuple, np.ndarray)):
            out = [int(i) + size if int(i) < 0 else int(i) for i in k]
            return out, False
        if isinstance(k, (DM,)):
            out = [int(i) for i in k._a.flatten(order="F")]
            return out, False
        i = int(k)
        if i < 0:
            i += size
        if not (0 <= i < size):
            raise IndexError(f"index {k} out of range for size {size}")
        return [i], True

    if isinstance(key, tuple):
        if len(key) != 2:
            raise IndexError("only 1 or 2 indices are supported")
        r, _ = ax(key[0], m)
        c, _ = ax(key[1], n)
        return r, c, "rc"
    lin, _ = ax(key, m * n)
    return lin, None, "lin"


class _Mat:
    """Shared behaviour of SX and DM (always 2-D, column-major semantics as in CasADi)."""

    __array_priority__ = 1000.0
    _a: np.ndarray

    # ---- shape -------------------------------------------------------------------------
    @property
    def shape(self) -> Tuple[int, int]:
        return self._a.shape

    def size(self, axis=None):
        if axis is None:
            return self._a.shape
        return self._a.shape[axis - 1]

    def size1(self):
        return self._a.shape[0]

    def size2(self):
        return self._a.shape[1]

    def numel(self):
        return self._a.shape[0] * self._a.shape[1]

    def is_empty(self):
        return self.numel() == 0

    def is_scalar(self):
        return self._a.shape == (1, 1)

    def is_vector(self):
        return 1 in self._a.shape

    def is_column(self):
        return self._a.shape[1] == 1

    @property
    def T(self):
        return type(self)._wrap(self._a.T.copy())

    def __len__(self):
        raise TypeError("len() of a 2-D casadi-style array is ambiguous; use .shape")

    # ---- indexing ----------------------------------------------------------------------
    def __getitem__(self, key):
        m, n = self._a.shape
        r, c, mode = _normalize_index(key, m, n)
        if mode == "lin":
            flat = self._a.flatten(order="F")
            out = np.empty((len(r), 1), dtype=self._a.dtype)
            for i, k in enumerate(r):
                out[i, 0] = flat[k]
            return type(self)._wrap(out)
        out = self._a[np.ix_(r, c)] if (r and c) else np.empty((len(r), len(c)), dtype=self._a.dtype)
        return type(self)._wrap(np.array(out, copy=True))

    def __setitem__(self, key, value):
        m, n = self._a.shape
        r, c, mode = _normalize_index(key, m, n)
        val = self._coerce(value)._a
        if mode == "lin":
            cnt = len(r)
            src = _bcast_to(val, (cnt, 1)) if val.shape != (1, cnt) else val.T
            for i, k in enumerate(r):
                self._a[k % m, k // m] = src[i, 0]
            return
        tgt = (len(r), len(c))
        if val.shape != tgt and val.shape == (tgt[1], tgt[0]) and 1 in tgt:
            val = val.T  # casadi tolerates row/column vector mismatch on assignment
        src = _bcast_to(val, tgt)
        for i, ri in enumerate(r):
            for j, cj in enumerate(c):
                self._a[ri, cj] = src[i, j]

    def __iter__(self):
        raise TypeError("casadi-style arrays are not iterable; use vertsplit/horzsplit")

    # ---- arithmetic dispatch -------------------------------------------------------------
    def _binary(self, other, op, swap=False):
        o = _to_mat(other)
        if o is None:
            return NotImplemented
        a, b = (o, self) if swap else (self, o)
        return _elementwise_binary(op, a, b)

    def __add__(self, o):
        return self._binary(o, OP_ADD)

    def __radd__(self, o):
        return self._binary(o, OP_ADD, True)

    def __sub__(self, o):
        return self._binary(o, OP_SUB)

    def __rsub__(self, o):
        return self._binary(o, OP_SUB, True)

    def __mul__(self, o):
        return self._binary(o, OP_MUL)

    def __rmul__(self, o):
        return self._binary(o, OP_MUL, True)

    def __truediv__(self, o):
        return self._binary(o, OP_DIV)

    def __rtruediv__(self, o):
        return self._binary(o, OP_DIV, True)

    def __pow__(self, o):
        return self._binary(o, OP_POW)

    def __rpow__(self, o):
        return self._binary(o, OP_POW, True)

    def __lt__(self, o):
        return self._binary(o, OP_LT)

    def __le__(self, o):
        return self._binary(o, OP_LE)

    def __gt__(self, o):
        return self._binary(o, OP_LT, True)

    def __ge__(self, o):
        return self._binary(o, OP_LE, True)

    def __eq__(self, o):  # elementwise, as in casadi
        return self._binary(o, OP_EQ)

    def __ne__(self, o):
        return self._binary(o, OP_NE)

    __hash__ = None

    def __neg__(self):
        return _elementwise_unary(OP_NEG, self)

    def __pos__(self):
        return self

    def __abs__(self):
        return _elementwise_unary(OP_FABS, self)

    def __matmul__(self, o):
        o = _to_mat(o)
        if o is None:
            return NotImplemented
        return mtimes(self, o)

    def __rmatmul__(self, o):
        o = _to_mat(o)
        if o is None:
            return NotImplemented
        return mtimes(o, self)


def _bcast_to(a: np.ndarray, shape) -> np.ndarray:
    if a.shape == tuple(shape):
        return a
    if a.shape == (1, 1):
        out = np.empty(shape, dtype=a.dtype)
        out[...] = a[0, 0]
        return out
    if a.shape[0] == shape[0] and a.shape[1] == 1:
        return np.repeat(a, shape[1], axis=1)
    if a.shape[1] == shape[1] and a.shape[0] == 1:
        return np.repeat(a, shape[0], axis=0)
    raise ValueError(f"dimension mismatch: cannot use shape {a.shape} where {tuple(shape)} is expected")


def _result_shape(sa, sb):
    if sa == sb:
        return sa
    if sa == (1, 1):
        return sb
    if sb == (1, 1):
        return sa
    # repmat-style broadcast of a column against a matrix / a row against a matrix (F4)
    if sa[0] == sb[0] and 1 in (sa[1], sb[1]):
        return (sa[0], max(sa[1], sb[1]))
    if sa[1] == sb[1] and 1 in (sa[0], sb[0]):
        return (max(sa[0], sb[0]), sa[1])
    raise ValueError(f"dimension mismatch for elementwise operation: {sa} vs {sb}")


class DM(_Mat):
    """Dense numeric 2-D array (float64)."""

    def __init__(self, *args):
        if len(args) == 0:
            self._a = np.zeros((0, 0))
        elif len(args) == 1:
            x = args[0]
            if isinstance(x, SX):
                self._a = x._numeric_values()
            else:
                self._a = np.array(_as_2d_float(x), dtype=float, copy=True)
        elif len(args) == 2:
            self._a = np.zeros((int(args[0]), int(args[1])))
        else:
            raise TypeError("DM(): unsupported arguments")

    @classmethod
    def _wrap(cls, a: np.ndarray) -> "DM":
        out = cls.__new__(cls)
        out._a = a
        return out

    def _coerce(self, value) -> "DM":
        if isinstance(value, SX):
            return DM(value)
        return value if isinstance(value, DM) else DM(value)

    @staticmethod
    def zeros(m=1, n=1):
        if isinstance(m, tuple):
            m, n = m
        return DM._wrap(np.zeros((int(m), int(n))))

    @staticmethod
    def ones(m=1, n=1):
        if isinstance(m, tuple):
            m, n = m
        return DM._wrap(np.ones((int(m), int(n))))

    @staticmethod
    def eye(n):
        return DM._wrap(np.eye(int(n)))

    @staticmethod
    def inf(m=1, n=1):
        return DM._wrap(np.full((int(m), int(n)), np.inf))

    def toarray(self, simplify=False) -> np.ndarray:
        a = np.array(self._a, copy=True)
        if simplify:
            if a.shape == (1, 1):
                return float(a[0, 0])
            if 1 in a.shape:
                return a.flatten()
        return a

    full = toarray

    def nonzeros(self):
        return list(self._a.flatten(order="F"))

    def __array__(self, dtype=None, copy=None):
        a = self._a if dtype is None else self._a.astype(dtype)
        return np.array(a, copy=True)

    def __float__(self):
        if self._a.size != 1:
            raise TypeError("only 1-by-1 DM can be converted to float")
        return float(self._a.flat[0])

    def __int__(self):
        return int(float(self))

    def __bool__(self):
        if self._a.size != 1:
            raise TypeError("only 1-by-1 DM can be converted to bool")
        return bool(self._a.flat[0] != 0.0)

    def __repr__(self):
        return f"DM({np.array2string(self._a, separator=', ')})"

    __str__ = __repr__


class SX(_Mat):
    """2-D array of scalar expression nodes."""

    def __init__(self, *args):
        if len(args) == 0:
            self._a = np.empty((0, 0), dtype=object)
        elif len(args) == 1:
            x = args[0]
            if isinstance(x, SX):
                self._a = x._a.copy()
            elif isinstance(x, Node):
                self._a = np.empty((1, 1), dtype=object)
                self._a[0, 0] = x
            else:
                self._a = _nodes_from_float(_as_2d_float(x))
        elif len(args) == 2:
            self._a = _nodes_from_float(np.zeros((int(args[0]), int(args[1]))))
        else:
            raise TypeError("SX(): unsupported arguments")

    @classmethod
    def _wrap(cls, a: np.ndarray) -> "SX":
        out = cls.__new__(cls)
        out._a = a
        return out

    def _coerce(self, value) -> "SX":
        return value if isinstance(value, SX) else SX(value)

    @staticmethod
    def sym(name: str, m: int = 1, n: int = 1) -> "SX":
        m, n = int(m), int(n)
        a = np.empty((m, n), dtype=object)
        scalar = m == 1 and n == 1
        k = 0
        for j in range(n):  # column-major numbering, as casadi
            for i in range(m):
                a[i, j] = sym(name if scalar else f"{name}_{k}")
                k += 1
        return SX._wrap(a)

    @staticmethod
    def zeros(m=1, n=1):
        if isinstance(m, tuple):
            m, n = m
        return SX(int(m), int(n))

    @staticmethod
    def ones(m=1, n=1):
        if isinstance(m, tuple):
            m, n = m
        return SX(np.ones((int(m), int(n))))

    @staticmethod
    def eye(n):
        return SX(np.eye(int(n)))

    def nodes(self) -> List[Node]:
        """Column-major list of the scalar nodes."""
        return list(self._a.flatten(order="F"))

    def is_constant(self) -> bool:
        return all(nd.op == OP_CONST for nd in self._a.flat)

    def is_symbolic(self) -> bool:
        return all(nd.op == OP_SYM for nd in self._a.flat)

    def _numeric_values(self) -> np.ndarray:
        out = np.empty(self._a.shape)
        for i in range(out.shape[0]):
            for j in range(out.shape[1]):
                nd = self._a[i, j]
                if nd.op != OP_CONST:
                    raise TypeError("SX expression is not constant; cannot convert to a number")
                out[i, j] = nd.val
        return out

    def __float__(self):
        if self._a.size != 1:
            raise TypeError("only 1-by-1 SX can be converted to float")
        return float(self._numeric_values().flat[0])

    def __bool__(self):
        if self._a.size != 1:
            raise TypeError("only 1-by-1 SX can be converted to bool")
        nd = self._a.flat[0]
        if nd.op != OP_CONST:
            raise TypeError("truth value of a symbolic expression is undefined")
        return nd.val != 0.0

    def __repr__(self):
        m, n = self._a.shape
        if m * n > 16:
            return f"SX({m}x{n})"
        return "SX(" + repr(self._a.tolist()) + ")"

    __str__ = __repr__


def _nodes_from_float(a: np.ndarray) -> np.ndarray:
    out = np.empty(a.shape, dtype=object)
    for i in range(a.shape[0]):
        for j in range(a.shape[1]):
            out[i, j] = const(a[i, j])
    return out


def _to_mat(x) -> Optional[_Mat]:
    if isinstance(x, _Mat):
        return x
    if isinstance(x, Node):
        return SX(x)
    if isinstance(x, (numbers.Number, list, tuple, np.ndarray, np.generic)):
        try:
            return DM(x)
        except (ValueError, TypeError):
            return None
    return None


def _mat(x) -> _Mat:
    m = _to_mat(x)
    if m is None:
        raise TypeError(f"cannot interpret {type(x)} as a casadi-style array")
    return m


def _elementwise_unary(op, a):
    a = _mat(a)
    if isinstance(a, DM):
        with np.errstate(all="ignore"):
            return DM._wrap(np.asarray(_NUM_UNARY[op](a._a), dtype=float).reshape(a.shape))
    out = np.empty(a.shape, dtype=object)
    for i in range(a.shape[0]):
        for j in range(a.shape[1]):
            out[i, j] = n_unary(op, a._a[i, j])
    return SX._wrap(out)


def _elementwise_binary(op, a, b):
    a, b = _mat(a), _mat(b)
    shape = _result_shape(a.shape, b.shape)
    if isinstance(a, DM) and isinstance(b, DM):
        with np.errstate(all="ignore"):
            res = _NUM_BINARY[op](_bcast_to(a._a, shape), _bcast_to(b._a, shape))
        return DM._wrap(np.asarray(res, dtype=float).reshape(shape))
    aa = _bcast_to(a._a if isinstance(a, SX) else _nodes_from_float(a._a), shape)
    bb = _bcast_to(b._a if isinstance(b, SX) else _nodes_from_float(b._a), shape)
    out = np.empty(shape, dtype=object)
    for i in range(shape[0]):
        for j in range(shape[1]):
            out[i, j] = n_binary(op, aa[i, j], bb[i, j])
    return SX._wrap(out)


# --------------------------------------------------------------------------------------
# casadi-style free functions
# --------------------------------------------------------------------------------------


def _any_sx(items) -> bool:
    return any(isinstance(i, SX) for i in items)


def _node_array(m: _Mat) -> np.ndarray:
    return m._a if isinstance(m, SX) else _nodes_from_float(m._a)


def _cat(items, axis):
    mats = [_mat(i) for i in items]
    # casadi lets empty blocks (0 along the concatenation axis, or 0-by-0) vanish
    keep = [m for m in mats if not (m.shape[axis] == 0 or m.shape == (0, 0))]
    if not keep:
        other = max((m.shape[1 - axis] for m in mats), default=0)
        return DM._wrap(np.zeros((0, other) if axis == 0 else (other, 0)))
    other = keep[0].shape[1 - axis]
    for m in keep:
        if m.shape[1 - axis] != other:
            raise ValueError(f"concatenation dimension mismatch: {[k.shape for k in keep]}")
    if _any_sx(keep):
        return SX._wrap(np.concatenate([_node_array(m) for m in keep], axis=axis))
    return DM._wrap(np.concatenate([m._a for m in keep], axis=axis))


def vertcat(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)) and not _is_numeric_seq(args[0]):
        args = tuple(args[0])
    return _cat(args, 0)


def horzcat(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)) and not _is_numeric_seq(args[0]):
        args = tuple(args[0])
    return _cat(args, 1)


def _is_numeric_seq(x) -> bool:
    return all(isinstance(i, numbers.Number) for i in x)


def veccat(*args):
    return vertcat(*[vec(a) for a in args])


def vec(x):
    x = _mat(x)
    return type(x)._wrap(x._a.reshape((-1, 1), order="F").copy())


def reshape(x, *shape):
    x = _mat(x)
    if len(shape) == 1:
        shape = tuple(shape[0])
    m, n = int(shape[0]), int(shape[1])
    if m == -1:
        m = x.numel() // n
    if n == -1:
        n = x.numel() // m
    return type(x)._wrap(x._a.reshape((m, n), order="F").copy())


def transpose(x):
    return _mat(x).T


def vertsplit(x, incr=1):
    x = _mat(x)
    m = x.shape[0]
    if isinstance(incr, (list, tuple)):
        offs = list(incr)
    else:
        offs = list(range(0, m, int(incr))) + [m]
    return [x[offs[i]:offs[i + 1], :] for i in range(len(offs) - 1)]


def horzsplit(x, incr=1):
    x = _mat(x)
    n = x.shape[1]
    if isinstance(incr, (list, tuple)):
        offs = list(incr)
    else:
        offs = list(range(0, n, int(incr))) + [n]
    return [x[:, offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]


def repmat(x, m, n=1):
    x = _mat(x)
    return type(x)._wrap(np.tile(x._a, (int(m), int(n))))


def _tree_sum(nodes: List[Node]) -> Node:
    """Balanced summation tree (keeps derivative dicts small and rounding well-behaved)."""
    if not nodes:
        return ZERO
    while len(nodes) > 1:
        nxt = [n_binary(OP_ADD, nodes[i], nodes[i + 1]) for i in range(0, len(nodes) - 1, 2)]
        if len(nodes) % 2:
            nxt.append(nodes[-1])
        nodes = nxt
    return nodes[0]


def sum1(x):
    """Sum over rows -> 1-by-n."""
    x = _mat(x)
    if isinstance(x, DM):
        return DM._wrap(x._a.sum(axis=0, keepdims=True))
    out = np.empty((1, x.shape[1]), dtype=object)
    for j in range(x.shape[1]):
        out[0, j] = _tree_sum(list(x._a[:, j]))
    return SX._wrap(out)


def sum2(x):
    """Sum over columns -> m-by-1."""
    x = _mat(x)
    if isinstance(x, DM):
        return DM._wrap(x._a.sum(axis=1, keepdims=True))
    out = np.empty((x.shape[0], 1), dtype=object)
    for i in range(x.shape[0]):
        out[i, 0] = _tree_sum(list(x._a[i, :]))
    return SX._wrap(out)


def sumsqr(x):
    x = _mat(x)
    if isinstance(x, DM):
        return DM._wrap(np.array([[float((x._a ** 2).sum())]]))
    return SX(_tree_sum([n_unary(OP_SQ, nd) for nd in x.nodes()]))


def dot(a, b):
    a, b = _mat(a), _mat(b)
    if a.shape != b.shape:
        raise ValueError("dot: dimension mismatch")
    if isinstance(a, DM) and isinstance(b, DM):
        return DM._wrap(np.array([[float((a._a * b._a).sum())]]))
    aa, bb = _node_array(a), _node_array(b)
    return SX(_tree_sum([n_binary(OP_MUL, x, y) for x, y in zip(aa.flatten(order="F"), bb.flatten(order="F"))]))


def norm_fro(x):
    return sqrt(sumsqr(x))


norm_2 = norm_fro  # only ever applied to vectors on this path


def norm_1(x):
    return sum1(vec(fabs(x)))


def norm_inf(x):
    return mmax(fabs(x))


def mmax(x):
    x = _mat(x)
    if isinstance(x, DM):
        return DM(float(x._a.max()))
    nodes = x.nodes()
    out = nodes[0]
    for nd in nodes[1:]:
        out = n_binary(OP_FMAX, out, nd)
    return SX(out)


def mmin(x):
    x = _mat(x)
    if isinstance(x, DM):
        return DM(float(x._a.min()))
    nodes = x.nodes()
    out = nodes[0]
    for nd in nodes[1:]:
        out = n_binary(OP_FMIN, out, nd)
    return SX(out)


def _mtimes2(a, b):
    a, b = _mat(a), _mat(b)
    if a.shape == (1, 1) or b.shape == (1, 1):
        return _elementwise_binary(OP_MUL, a, b)
    if a.shape[1] != b.shape[0]:
        raise ValueError(f"matrix product dimension mismatch: {a.shape} @ {b.shape}")
    if isinstance(a, DM) and isinstance(b, DM):
        return DM._wrap(a._a @ b._a)
    aa, bb = _node_array(a), _node_array(b)
    m, k, n = a.shape[0], a.shape[1], b.shape[1]
    out = np.empty((m, n), dtype=object)
    for i in range(m):
        for j in range(n):
            terms = []
            for l in range(k):
                t = n_binary(OP_MUL, aa[i, l], bb[l, j])
                if not _is_zero(t):
                    terms.append(t)
            # left-to-right accumulation like a plain triple loop
            acc = ZERO
            for t in terms:
                acc = n_binary(OP_ADD, acc, t)
            out[i, j] = acc
    return SX._wrap(out)


def mtimes(a, b=None, *more):
    if b is None:
        if isinstance(a, (list, tuple)):
            return mtimes(*a)
        raise TypeError("mtimes needs at least two operands")
    out = _mtimes2(a, b)
    for c in more:
        out = _mtimes2(out, c)
    return out


def cross(a, b, dim=-1):
    a, b = _mat(a), _mat(b)
    a, b = vec(a), vec(b)
    if a.shape != (3, 1) or b.shape != (3, 1):
        raise ValueError("cross: expecting two 3-vectors")
    a0, a1, a2 = a[0], a[1], a[2]
    b0, b1, b2 = b[0], b[1], b[2]
    return vertcat(a1 * b2 - a2 * b1, a2 * b0 - a0 * b2, a0 * b1 - a1 * b0)


def diag(x):
    x = _mat(x)
    m, n = x.shape
    if 1 in (m, n):
        k = m * n
        flat = x._a.flatten(order="F")
        if isinstance(x, DM):
            return DM._wrap(np.diag(flat))
        out = _nodes_from_float(np.zeros((k, k)))
        for i in range(k):
            out[i, i] = flat[i]
        return SX._wrap(out)
    k = min(m, n)
    return type(x)._wrap(np.array([[x._a[i, i]] for i in range(k)], dtype=x._a.dtype).reshape(k, 1))


def trace(x):
    x = _mat(x)
    return sum1(diag(x))


def linspace(a, b, n):
    a, b = _mat(a), _mat(b)
    n = int(n)
    if isinstance(a, DM) and isinstance(b, DM) and a.shape == (1, 1) and b.shape == (1, 1):
        return DM._wrap(np.linspace(float(a), float(b), n).reshape(-1, 1))
    rows = []
    for i in range(n):
        w = i / float(n - 1)
        rows.append((a + (b - a) * w).T)
    return vertcat(*rows)


def _unary_fn(op):
    def fn(x):
        return _elementwise_unary(op, x)

    fn.__name__ = OP_NAMES[op].lower()
    return fn


def _binary_fn(op):
    def fn(a, b):
        return _elementwise_binary(op, a, b)

    fn.__name__ = OP_NAMES[op].lower()
    return fn


sin = _unary_fn(OP_SIN)
cos = _unary_fn(OP_COS)
tan = _unary_fn(OP_TAN)
asin = arcsin = _unary_fn(OP_ASIN)
acos = arccos = _unary_fn(OP_ACOS)
atan = arctan = _unary_fn(OP_ATAN)
sqrt = _unary_fn(OP_SQRT)
fabs = _unary_fn(OP_FABS)
exp = _unary_fn(OP_EXP)
log = _unary_fn(OP_LOG)
sign = _unary_fn(OP_SIGN)
floor = _unary_fn(OP_FLOOR)
ceil = _unary_fn(OP_CEIL)
tanh = _unary_fn(OP_TANH)
sinh = _unary_fn(OP_SINH)
cosh = _unary_fn(OP_COSH)
sq = _unary_fn(OP_SQ)
logic_not = _unary_fn(OP_NOT)
atan2 = arctan2 = _binary_fn(OP_ATAN2)
fmin = _binary_fn(OP_FMIN)
fmax = _binary_fn(OP_FMAX)
power = _binary_fn(OP_POW)
logic_and = _binary_fn(OP_AND)
logic_or = _binary_fn(OP_OR)
plus = _binary_fn(OP_ADD)
minus = _binary_fn(OP_SUB)
times = _binary_fn(OP_MUL)
rdivide = _binary_fn(OP_DIV)


def logic_all(x):
    x = _mat(x)
    if isinstance(x, DM):
        return DM(float(bool((x._a != 0.0).all())))
    nodes = x.nodes()
    out = ONE
    for nd in nodes:
        out = n_binary(OP_AND, out, nd)
    return SX(out)


def logic_any(x):
    x = _mat(x)
    if isinstance(x, DM):
        return DM(float(bool((x._a != 0.0).any())))
    nodes = x.nodes()
    out = ZERO
    for nd in nodes:
        out = n_binary(OP_OR, out, nd)
    return SX(out)


def if_else(c, a, b, short_circuit=False):
    c, a, b = _mat(c), _mat(a), _mat(b)
    shape = _result_shape(_result_shape(c.shape, a.shape), b.shape)
    if isinstance(c, DM) and isinstance(a, DM) and isinstance(b, DM):
        return DM._wrap(np.where(_bcast_to(c._a, shape) != 0.0, _bcast_to(a._a, shape), _bcast_to(b._a, shape)))
    cc, aa, bb = (_bcast_to(_node_array(m), shape) for m in (c, a, b))
    out = np.empty(shape, dtype=object)
    for i in range(shape[0]):
        for j in range(shape[1]):
            out[i, j] = n_if_else(cc[i, j], aa[i, j], bb[i, j])
    return SX._wrap(out)


def inv(x):
    x = _mat(x)
    if isinstance(x, DM):
        return DM._wrap(np.linalg.inv(x._a))
    raise NotImplementedError("symbolic matrix inverse is not used on the OpTaS solver path")


def solve(a, b):
    a, b = _mat(a), _mat(b)
    if isinstance(a, DM) and isinstance(b, DM):
        return DM._wrap(np.linalg.solve(a._a, b._a))
    raise NotImplementedError("symbolic linear solve is not used on the OpTaS solver path")


def det(x):
    x = _mat(x)
    if isinstance(x, DM):
        return DM(float(np.linalg.det(x._a)))
    raise NotImplementedError("symbolic determinant is not used on the OpTaS solver path")


# --------------------------------------------------------------------------------------
# graph utilities: topological order, dependency, substitution, differentiation
# --------------------------------------------------------------------------------------


def topo_sort(outputs: Iterable[Node]) -> List[Node]:
    """All nodes reachable from ``outputs`` in dependency order (children first)."""
    seen = {}
    stack = [o for o in outputs]
    while stack:
        nd = stack.pop()
        if nd.idx in seen:
            continue
        seen[nd.idx] = nd
        if nd.a is not None:
            stack.append(nd.a)
        if nd.b is not None:
            stack.append(nd.b)
        if nd.c is not None:
            stack.append(nd.c)
    return [seen[k] for k in sorted(seen)]


def symvar(x) -> List[Node]:
    x = _mat(x)
    if isinstance(x, DM):
        return []
    return [nd for nd in topo_sort(x.nodes()) if nd.op == OP_SYM]


def depends_on(expr, x) -> bool:
    expr, x = _mat(expr), _mat(x)
    if isinstance(expr, DM) or isinstance(x, DM):
        return False
    xs = {nd.idx for nd in x.nodes()}
    return any(nd.idx in xs for nd in topo_sort(expr.nodes()))


def _rebuild(nd: Node, a, b, c) -> Node:
    if nd.op in UNARY_OPS:
        return n_unary(nd.op, a)
    if nd.op == OP_IF_ELSE:
        return n_if_else(a, b, c)
    return n_binary(nd.op, a, b)


def substitute_nodes(outputs: List[Node], mapping: Dict[int, Node]) -> List[Node]:
    """Replace leaf symbols (by node idx) with other nodes; everything else is rebuilt
    through the simplifying constructors."""
    memo: Dict[int, Node] = dict(mapping)
    for nd in topo_sort(outputs):
        if nd.idx in memo:
            continue
        if nd.op in (OP_CONST, OP_SYM):
            memo[nd.idx] = nd
            continue
        a = memo[nd.a.idx]
        b = memo[nd.b.idx] if nd.b is not None else None
        c = memo[nd.c.idx] if nd.c is not None else None
        if a is nd.a and b is nd.b and c is nd.c:
            memo[nd.idx] = nd
        else:
            memo[nd.idx] = _rebuild(nd, a, b, c)
    return [memo[o.idx] for o in outputs]


def substitute(expr, old, new):
    expr, old, new = _mat(expr), _mat(old), _mat(new)
    if isinstance(expr, DM):
        return expr
    if old.numel() != new.numel():
        raise ValueError("substitute: old and new must have the same number of elements")
    mapping = {o.idx: n for o, n in zip(old.nodes(), _node_array(new).flatten(order="F"))}
    res = substitute_nodes(expr.nodes(), mapping)
    out = np.empty(expr.numel(), dtype=object)
    out[:] = res
    return SX._wrap(out.reshape(expr.shape, order="F"))


def _partials(nd: Node, a: Node, b: Optional[Node], c: Optional[Node]):
    """Local partial derivatives (as nodes) of nd wrt its operands."""
    op = nd.op
    if op == OP_ADD:
        return ONE, ONE, None
    if op == OP_SUB:
        return ONE, MINUS_ONE, None
    if op == OP_MUL:
        return b, a, None
    if op == OP_DIV:
        # d(a/b) = 1/b da - (a/b)/b db
        return n_binary(OP_DIV, ONE, b), n_unary(OP_NEG, n_binary(OP_DIV, nd, b)), None
    if op == OP_NEG:
        return MINUS_ONE, None, None
    if op == OP_SQ:
        return n_binary(OP_MUL, TWO, a), None, None
    if op == OP_SQRT:
        return n_binary(OP_DIV, const(0.5), nd), None, None
    if op == OP_SIN:
        return n_unary(OP_COS, a), None, None
    if op == OP_COS:
        return n_unary(OP_NEG, n_unary(OP_SIN, a)), None, None
    if op == OP_TAN:
        return n_binary(OP_ADD, ONE, n_unary(OP_SQ, nd)), None, None
    if op == OP_ASIN:
        return n_binary(OP_DIV, ONE, n_unary(OP_SQRT, n_binary(OP_SUB, ONE, n_unary(OP_SQ, a)))), None, None
    if op == OP_ACOS:
        return n_unary(OP_NEG, n_binary(OP_DIV, ONE, n_unary(OP_SQRT, n_binary(OP_SUB, ONE, n_unary(OP_SQ, a))))), None, None
    if op == OP_ATAN:
        return n_binary(OP_DIV, ONE, n_binary(OP_ADD, ONE, n_unary(OP_SQ, a))), None, None
    if op == OP_ATAN2:
        den = n_binary(OP_ADD, n_unary(OP_SQ, a), n_unary(OP_SQ, b))
        return n_binary(OP_DIV, b, den), n_unary(OP_NEG, n_binary(OP_DIV, a, den)), None
    if op == OP_FABS:
        return n_unary(OP_SIGN, a), None, None
    if op == OP_EXP:
        return nd, None, None
    if op == OP_LOG:
        return n_binary(OP_DIV, ONE, a), None, None
    if op == OP_TANH:
        return n_binary(OP_SUB, ONE, n_unary(OP_SQ, nd)), None, None
    if op == OP_SINH:
        return n_unary(OP_COSH, a), None, None
    if op == OP_COSH:
        return n_unary(OP_SINH, a), None, None
    if op == OP_POW:
        # d(a^b) = b a^(b-1) da + a^b log(a) db
        da = n_binary(OP_MUL, b, n_binary(OP_POW, a, n_binary(OP_SUB, b, ONE)))
        db = ZERO if b.op == OP_CONST else n_binary(OP_MUL, nd, n_unary(OP_LOG, a))
        return da, db, None
    if op == OP_FMIN:
        sel = n_binary(OP_LE, a, b)
        return sel, n_unary(OP_NOT, sel), None
    if op == OP_FMAX:
        sel = n_binary(OP_LE, b, a)
        return sel, n_unary(OP_NOT, sel), None
    if op == OP_IF_ELSE:
        # operands are (cond, then, else) stored as (a, b, c)
        return ZERO, n_if_else(a, ONE, ZERO), n_if_else(a, ZERO, ONE)
    if op in (OP_LT, OP_LE, OP_EQ, OP_NE, OP_AND, OP_OR, OP_NOT, OP_SIGN, OP_FLOOR, OP_CEIL):
        return ZERO, ZERO, None
    raise NotImplementedError(f"derivative of op {OP_NAMES.get(op, op)}")


def forward_partials(outputs: List[Node], inputs: List[Node]) -> List[Dict[int, Node]]:
    """One sparse forward sweep: for every output a dict {input position: d out / d in}."""
    pos = {nd.idx: k for k, nd in enumerate(inputs)}
    d: Dict[int, Dict[int, Node]] = {}
    empty: Dict[int, Node] = {}
    for nd in topo_sort(outputs):
        if nd.op == OP_CONST:
            continue
        if nd.op == OP_SYM:
            k = pos.get(nd.idx)
            if k is not None:
                d[nd.idx] = {k: ONE}
            continue
        da = d.get(nd.a.idx, empty)
        db = d.get(nd.b.idx, empty) if nd.b is not None else empty
        dc = d.get(nd.c.idx, empty) if nd.c is not None else empty
        if not da and not db and not dc:
            continue
        pa, pb, pc = _partials(nd, nd.a, nd.b, nd.c)
        acc: Dict[int, Node] = {}
        for part, dd in ((pa, da), (pb, db), (pc, dc)):
            if not dd or part is None or _is_zero(part):
                continue
            for k, v in dd.items():
                t = n_binary(OP_MUL, part, v)
                prev = acc.get(k)
                acc[k] = t if prev is None else n_binary(OP_ADD, prev, t)
        acc = {k: v for k, v in acc.items() if not _is_zero(v)}
        if acc:
            d[nd.idx] = acc
    return [d.get(o.idx, empty) for o in outputs]


def jacobian_sparse(expr, x) -> Tuple[Tuple[int, int], Dict[Tuple[int, int], Node]]:
    """Structurally sparse Jacobian: ((rows, cols), {(i, j): node})."""
    expr, x = _mat(expr), _mat(x)
    ne, nx = expr.numel(), x.numel()
    if isinstance(expr, DM) or ne == 0 or nx == 0:
        return (ne, nx), {}
    if not isinstance(x, SX) or not x.is_symbolic():
        raise ValueError("jacobian: second argument must be purely symbolic")
    parts = forward_partials(expr.nodes(), x.nodes())
    entries = {}
    for i, dd in enumerate(parts):
        for j, nd in dd.items():
            entries[(i, j)] = nd
    return (ne, nx), entries


def jacobian(expr, x):
    (ne, nx), entries = jacobian_sparse(expr, x)
    out = np.empty((ne, nx), dtype=object)
    out[...] = ZERO
    for (i, j), nd in entries.items():
        out[i, j] = nd
    return SX._wrap(out)


def gradient(expr, x):
    expr = _mat(expr)
    if expr.numel() != 1:
        raise ValueError("gradient: expression must be scalar")
    return jacobian(expr, x).T


def hessian(expr, x):
    g = gradient(expr, x)
    return jacobian(g, x), g


def jtimes(expr, x, v, tr=False):
    J = jacobian(expr, x)
    return (J.T if tr else J) @ _mat(v)


_DEG_INF = 1 << 20


def _degrees(outputs: List[Node], xs: List[Node]) -> List[int]:
    """Polynomial degree of each output in xs (``_DEG_INF`` if not polynomial)."""
    xset = {nd.idx for nd in xs}
    deg: Dict[int, int] = {}
    for nd in topo_sort(outputs):
        op = nd.op
        if op == OP_CONST:
            deg[nd.idx] = 0
        elif op == OP_SYM:
            deg[nd.idx] = 1 if nd.idx in xset else 0
        elif op in (OP_ADD, OP_SUB):
            deg[nd.idx] = max(deg[nd.a.idx], deg[nd.b.idx])
        elif op == OP_MUL:
            deg[nd.idx] = min(_DEG_INF, deg[nd.a.idx] + deg[nd.b.idx])
        elif op == OP_DIV:
            deg[nd.idx] = deg[nd.a.idx] if deg[nd.b.idx] == 0 else _DEG_INF
        elif op == OP_NEG:
            deg[nd.idx] = deg[nd.a.idx]
        elif op == OP_SQ:
            deg[nd.idx] = min(_DEG_INF, 2 * deg[nd.a.idx])
        elif op == OP_POW and nd.b.op == OP_CONST and float(nd.b.val).is_integer() and nd.b.val >= 0:
            deg[nd.idx] = min(_DEG_INF, int(nd.b.val) * deg[nd.a.idx])
        else:
            kids = [k for k in (nd.a, nd.b, nd.c) if k is not None]
            deg[nd.idx] = 0 if all(deg[k.idx] == 0 for k in kids) else _DEG_INF
    return [deg[o.idx] for o in outputs]


def is_linear(expr, x) -> bool:
    """True when expr is affine in x (F11)."""
    expr, x = _mat(expr), _mat(x)
    if isinstance(expr, DM) or expr.numel() == 0:
        return True
    return max(_degrees(expr.nodes(), x.nodes()), default=0) <= 1


def is_quadratic(expr, x) -> bool:
    """True when expr has polynomial degree <= 2 in x (F12)."""
    expr, x = _mat(expr), _mat(x)
    if isinstance(expr, DM) or expr.numel() == 0:
        return True
    return max(_degrees(expr.nodes(), x.nodes()), default=0) <= 2


# --------------------------------------------------------------------------------------
# Function
# --------------------------------------------------------------------------------------


class Function:
    """``Function(name, [inputs...], [outputs...])`` (F8).  Called with any SX argument it
    substitutes (inlines); called with numbers only it evaluates and returns DM."""

    def __init__(self, name: str, inputs: Sequence, outputs: Sequence, *unused, **unused_kw):
        self._name = name
        self._in = [SX(_mat(i)) if not isinstance(i, SX) else i for i in inputs]
        for i in self._in:
            if not i.is_symbolic():
                raise ValueError(f"Function '{name}': inputs must be purely symbolic")
        self._out = [_mat(o) for o in outputs]
        self._tape = None

    # -- introspection ---------------------------------------------------------------------
    def name(self):
        return self._name

    def n_in(self):
        return len(self._in)

    def n_out(self):
        return len(self._out)

    def size_in(self, i):
        return self._in[i].shape

    def size_out(self, i):
        return self._out[i].shape

    def size1_in(self, i):
        return self._in[i].shape[0]

    def size2_in(self, i):
        return self._in[i].shape[1]

    def size1_out(self, i):
        return self._out[i].shape[0]

    def size2_out(self, i):
        return self._out[i].shape[1]

    def numel_in(self, i=None):
        if i is None:
            return sum(x.numel() for x in self._in)
        return self._in[i].numel()

    def numel_out(self, i=None):
        if i is None:
            return sum(x.numel() for x in self._out)
        return self._out[i].numel()

    def sx_in(self, i=None):
        return self._in if i is None else self._in[i]

    def sx_out(self, i=None):
        return self._out if i is None else self._out[i]

    # -- calling ---------------------------------------------------------------------------
    def _check_args(self, args):
        if len(args) != len(self._in):
            raise TypeError(f"Function '{self._name}' takes {len(self._in)} arguments ({len(args)} given)")
        mats = []
        for a, i in zip(args, self._in):
            a = _mat(a)
            if a.shape != i.shape:
                if a.numel() == i.numel() and (a.is_vector() or a.numel() == 0) and (i.is_vector() or i.numel() == 0):
                    a = reshape(a, i.shape)  # casadi accepts a transposed vector
                elif a.shape == (1, 1):
                    a = repmat(a, *i.shape)
                else:
                    raise ValueError(
                        f"Function '{self._name}': argument has shape {a.shape}, expected {i.shape}")
            mats.append(a)
        return mats

    def __call__(self, *args, **kwargs):
        if kwargs:
            raise TypeError("keyword calling convention is not supported")
        mats = self._check_args(args)
        if any(isinstance(m, SX) for m in mats):
            res = self._call_symbolic(mats)
        else:
            res = self._call_numeric(mats)
        return res[0] if len(res) == 1 else tuple(res)

    def call(self, args):
        mats = self._check_args(list(args))
        if any(isinstance(m, SX) for m in mats):
            return self._call_symbolic(mats)
        return self._call_numeric(mats)

    def _call_symbolic(self, mats):
        mapping = {}
        for m, i in zip(mats, self._in):
            for old, new in zip(i.nodes(), _node_array(m).flatten(order="F")):
                mapping[old.idx] = new
        outs = []
        for o in self._out:
            if isinstance(o, DM):
                outs.append(DM(o))
                continue
            res = substitute_nodes(o.nodes(), mapping)
            arr = np.empty(o.numel(), dtype=object)
            arr[:] = res
            outs.append(SX._wrap(arr.reshape(o.shape, order="F")))
        return outs

    def _call_numeric(self, mats):
        from .tape import Tape  # local import: tape depends on this module

        if self._tape is None:
            self._tape = Tape.from_function(self)
        flat_in = [m._a.flatten(order="F") for m in mats]
        flat_out = self._tape.eval_numpy(flat_in)
        return [DM._wrap(np.array(v, dtype=float).reshape(o.shape, order="F")) for v, o in zip(flat_out, self._out)]

    # -- map ---------------------------------------------------------------------------------
    def map(self, n: int, *unused):
        """Map over columns: every m-by-1 input becomes m-by-n, outputs k-by-1 -> k-by-n (F9)."""
        return _MappedFunction(self, int(n))

    def __repr__(self):
        ins = ",".join(f"i{k}{list(i.shape)}" for k, i in enumerate(self._in))
        outs = ",".join(f"o{k}{list(o.shape)}" for k, o in enumerate(self._out))
        return f"Function({self._name}:({ins})->({outs}))"


class _MappedFunction:
    def __init__(self, f: Function, n: int):
        self.f = f
        self.n = n

    def name(self):
        return f"map{self.n}_{self.f.name()}"

    def n_in(self):
        return self.f.n_in()

    def n_out(self):
        return self.f.n_out()

    def size_in(self, i):
        m, k = self.f.size_in(i)
        return (m, k * self.n)

    def size_out(self, i):
        m, k = self.f.size_out(i)
        return (m, k * self.n)

    def size1_in(self, i):
        return self.size_in(i)[0]

    def size2_in(self, i):
        return self.size_in(i)[1]

    def size1_out(self, i):
        return self.size_out(i)[0]

    def size2_out(self, i):
        return self.size_out(i)[1]

    def numel_in(self, i=None):
        return self.f.numel_in(i) * self.n

    def numel_out(self, i=None):
        return self.f.numel_out(i) * self.n

    def __call__(self, *args):
        if len(args) != self.f.n_in():
            raise TypeError(f"mapped function takes {self.f.n_in()} arguments ({len(args)} given)")
        mats = [_mat(a) for a in args]
        cols = []
        for k in range(self.n):
            call_args = []
            for a, i in zip(mats, range(self.f.n_in())):
                w = self.f.size_in(i)[1]
                if a.shape[1] == w * self.n:
                    call_args.append(a[:, k * w:(k + 1) * w])
                elif a.shape[1] == w:  # non-repeated argument is broadcast to every column
                    call_args.append(a)
                else:
                    raise ValueError("mapped function: argument has wrong number of columns")
            res = self.f.call(call_args)
            cols.append(res)
        outs = [horzcat(*[c[o] for c in cols]) for o in range(self.f.n_out())]
        return outs[0] if len(outs) == 1 else tuple(outs)


# constants that ``from casadi import *`` would provide (reference: optas/__init__.py:2)
pi = math.pi
inf = math.inf
