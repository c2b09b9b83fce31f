"""SSA tapes: the lowered form of expression graphs that the CUDA virtual machine executes.

CasADi evaluates ``Function`` objects by interpreting an instruction tape over a scalar work
vector (the "SX virtual machine" that runs inside ``nlpsol`` on the reference path,
reference: optas/solver.py:395 -> casadi).  This module lowers `optas_b200.sym` graphs to the
equivalent structure for this backend:

* ``instr``  int32 [n, 4] rows ``(op | c << 8, dst, a, b)``; opcodes are `sym.OP_*`.
  - ``OP_INPUT``:  ``work[dst] = inputs[segment b][element a]``
  - ``OP_CONST``:  ``work[dst] = consts[a]``
  - ``OP_OUTPUT``: ``outputs[segment b][element a] = work[dst]``
  - unary/binary:  ``work[dst] = op(work[a], work[b])``
  - ``OP_IF_ELSE``: ``work[dst] = work[c] != 0 ? work[a] : work[b]`` (``c`` in the high bits)
* ``consts`` float64 [n_const]
* ``n_work`` number of work slots after liveness-based slot reuse.

A tape may be split into *groups* (independent sub-tapes that each write a disjoint part of
the outputs) so that the GPU can run one thread per (instance, group).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import sym as S


def _schedule(out_nodes: Sequence["S.Node"], in_pos: Dict[int, tuple]) -> List["S.Node"]:
    """Evaluation order of the DAG under ``out_nodes`` that keeps few values alive at a time.

    Plain creation order (what ``topo_sort`` gives) interleaves the sub-graphs of all time steps of a
    horizon problem, so tens of thousands of values are live at once (C5: 15 459 work slots = 124 KB per
    instance, far beyond any cache).  Here the outputs are visited grouped by the LAST decision variable
    (element of the first input segment) they depend on -- all outputs of one stage together, outputs that depend on everything (the cost)
    last -- and each is evaluated depth-first, so a stage's kinematics are computed, consumed by that
    stage's gradient / Jacobian / Hessian entries and freed before the next stage starts."""
    reach: Dict[int, int] = {}
    base = S.topo_sort(out_nodes)
    for nd in base:
        if nd.op == S.OP_SYM:
            s, e = in_pos.get(nd.idx, (0, 0))
            reach[nd.idx] = e if s == 0 else -1  # position in the FIRST input segment (the decision variables)
        elif nd.op == S.OP_CONST:
            reach[nd.idx] = -1
        else:
            r = reach[nd.a.idx]
            if nd.b is not None:
                r = max(r, reach[nd.b.idx])
            if nd.c is not None:
                r = max(r, reach[nd.c.idx])
            reach[nd.idx] = r
    visit_order = sorted(range(len(out_nodes)), key=lambda k: (reach[out_nodes[k].idx], k))
    done = set()
    order: List["S.Node"] = []
    for k in visit_order:
        root = out_nodes[k]
        if root.idx in done:
            continue
        stack = [(root, 0)]
        while stack:
            nd, state = stack.pop()
            if nd.idx in done:
                continue
            kids = [ch for ch in (nd.a, nd.b, nd.c) if ch is not None]
            if state < len(kids):
                stack.append((nd, state + 1))
                ch = kids[state]
                if ch.idx not in done:
                    stack.append((ch, 0))
            else:
                done.add(nd.idx)
                order.append(nd)
    return order


@dataclass
class Tape:
    instr: np.ndarray  # int32 [n, 4]
    consts: np.ndarray  # float64 [n_const]
    n_work: int
    in_sizes: List[int]
    out_sizes: List[int]
    # instruction ranges of the groups: group g is instr[group_ptr[g]:group_ptr[g+1]]
    group_ptr: np.ndarray = field(default_factory=lambda: np.zeros(2, dtype=np.int32))

    @property
    def n_instr(self) -> int:
        return int(self.instr.shape[0])

    @property
    def n_groups(self) -> int:
        return int(self.group_ptr.shape[0] - 1)

    # ---------------------------------------------------------------------------------------
    @staticmethod
    def from_function(f: "S.Function") -> "Tape":
        ins = [i.nodes() for i in f._in]
        outs = []
        for o in f._out:
            outs.append(o.nodes() if isinstance(o, S.SX) else [S.const(v) for v in o._a.flatten(order="F")])
        # single-expression model functions (K1): creation order schedules best (measured: 84.7 % vs 83.3 %
        # of HBM peak for fk_jac); the stage-grouped order is for the long solver tapes
        return Tape.lower(ins, outs, order="creation")

    @staticmethod
    def lower(inputs: Sequence[Sequence[S.Node]], outputs: Sequence[Sequence[S.Node]],
              groups: Optional[Sequence[Sequence[tuple]]] = None, order: str = "staged") -> "Tape":
        """Lower to a tape.

        inputs:  per input segment, the list of symbol nodes (position = element index).
        outputs: per output segment, the list of nodes to store.
        groups:  optional partition of the outputs: a list of groups, each a list of
                 ``(segment, element)`` pairs.  Default: one group holding everything.
        """
        in_pos: Dict[int, tuple] = {}
        for s, seg in enumerate(inputs):
            for e, nd in enumerate(seg):
                if nd.op != S.OP_SYM:
                    raise ValueError("tape inputs must be symbols")
                in_pos[nd.idx] = (s, e)
        if groups is None:
            groups = [[(s, e) for s, seg in enumerate(outputs) for e in range(len(seg))]]

        const_index: Dict[int, int] = {}
        consts: List[float] = []
        rows: List[tuple] = []
        group_ptr = [0]
        n_work = 0

        order_mode = order
        REMAT_GAP = 48  # a leaf (input / constant) whose next use is further away than this is re-loaded there

        for grp in groups:
            out_nodes = [outputs[s][e] for (s, e) in grp]
            order = S.topo_sort(out_nodes) if order_mode == "creation" else _schedule(out_nodes, in_pos)
            # positions (in `order`) at which every node is used as an operand
            uses: Dict[int, List[int]] = {}
            for k, nd in enumerate(order):
                for ch in (nd.a, nd.b, nd.c):
                    if ch is not None:
                        lst = uses.setdefault(ch.idx, [])
                        if not lst or lst[-1] != k:
                            lst.append(k)
            use_ptr: Dict[int, int] = {}
            stores: Dict[int, List[tuple]] = {}
            for (s, e), nd in zip(grp, out_nodes):
                stores.setdefault(nd.idx, []).append((s, e))
            slot_of: Dict[int, int] = {}
            free: List[int] = []
            high = 0

            def take_slot() -> int:
                nonlocal high
                if free:
                    return free.pop()
                high += 1
                return high - 1

            def emit_leaf(nd) -> int:
                dst = take_slot()
                if nd.op == S.OP_CONST:
                    ci = const_index.get(nd.idx)
                    if ci is None:
                        ci = len(consts)
                        consts.append(nd.val)
                        const_index[nd.idx] = ci
                    rows.append((S.OP_CONST, dst, ci, 0))
                else:
                    if nd.idx not in in_pos:
                        raise ValueError(f"free symbol '{nd.name}' is not an input of the tape")
                    s, e = in_pos[nd.idx]
                    rows.append((S.OP_INPUT, dst, e, s))
                slot_of[nd.idx] = dst
                return dst

            for k, nd in enumerate(order):
                leaf = nd.op in (S.OP_CONST, S.OP_SYM)
                if leaf:
                    # leaves are materialised on demand (below); a leaf that is itself an output is stored here
                    if nd.idx in stores:
                        dst = slot_of[nd.idx] if nd.idx in slot_of else emit_leaf(nd)
                        for (s, e) in stores[nd.idx]:
                            rows.append((S.OP_OUTPUT, dst, e, s))
                        if not uses.get(nd.idx):
                            free.append(slot_of.pop(nd.idx))
                    continue
                kids = [ch for ch in (nd.a, nd.b, nd.c) if ch is not None]
                for ch in kids:
                    if ch.idx not in slot_of:  # a leaf not (or no longer) resident
                        emit_leaf(ch)
                ops = [slot_of[ch.idx] if ch is not None else 0 for ch in (nd.a, nd.b, nd.c)]
                # release operand slots: last use, or a leaf whose next use is far away
                for ch in {c.idx: c for c in kids}.values():
                    lst = uses[ch.idx]
                    ptr = use_ptr.get(ch.idx, 0)
                    while ptr < len(lst) and lst[ptr] <= k:
                        ptr += 1
                    use_ptr[ch.idx] = ptr
                    is_leaf = ch.op in (S.OP_CONST, S.OP_SYM)
                    if ptr >= len(lst) or (is_leaf and lst[ptr] - k > REMAT_GAP):
                        free.append(slot_of.pop(ch.idx))
                dst = take_slot()
                slot_of[nd.idx] = dst
                if nd.op == S.OP_IF_ELSE:
                    rows.append((S.OP_IF_ELSE | (ops[0] << 8), dst, ops[1], ops[2]))
                else:
                    rows.append((nd.op, dst, ops[0], ops[1]))
                for (s, e) in stores.get(nd.idx, ()):
                    rows.append((S.OP_OUTPUT, dst, e, s))
                if not uses.get(nd.idx):
                    free.append(slot_of.pop(nd.idx))  # value is never read again (pure output)
            n_work = max(n_work, high)
            group_ptr.append(len(rows))

        instr = np.array(rows, dtype=np.int64).reshape(-1, 4)
        if instr.size and (instr[:, 0] >> 8).max() >= (1 << 23):
            raise ValueError("tape too large for the if_else operand encoding")
        return Tape(
            instr=instr.astype(np.int32),
            consts=np.array(consts, dtype=np.float64),
            n_work=max(1, n_work),
            in_sizes=[len(s) for s in inputs],
            out_sizes=[len(s) for s in outputs],
            group_ptr=np.array(group_ptr, dtype=np.int32),
        )

    # ---------------------------------------------------------------------------------------
    def eval_numpy(self, inputs: Sequence[np.ndarray]) -> List[np.ndarray]:
        """Reference interpreter.  Each input is [n] (one instance) or [n, B] (batched)."""
        ins = [np.asarray(v, dtype=float) for v in inputs]
        batched = any(v.ndim == 2 for v in ins)
        B = max((v.shape[1] for v in ins if v.ndim == 2), default=1)
        shape = (B,) if batched else ()
        outs = [np.zeros((n,) + shape) for n in self.out_sizes]
        work: List = [None] * self.n_work
        UN, BI = S._NUM_UNARY, S._NUM_BINARY
        with np.errstate(all="ignore"):
            for op, dst, a, b in self.instr.tolist():
                c = op >> 8
                op &= 0xFF
                if op == S.OP_INPUT:
                    work[dst] = ins[b][a]
                elif op == S.OP_CONST:
                    work[dst] = self.consts[a]
                elif op == S.OP_OUTPUT:
                    outs[b][a] = work[dst]
                elif op == S.OP_IF_ELSE:
                    work[dst] = np.where(np.asarray(work[c]) != 0.0, work[a], work[b])
                elif op in UN:
                    work[dst] = UN[op](work[a])
                else:
                    work[dst] = BI[op](work[a], work[b])
        return outs

    def op_histogram(self) -> Dict[str, int]:
        ops, counts = np.unique(self.instr[:, 0] & 0xFF, return_counts=True)
        return {S.OP_NAMES.get(int(o), str(o)): int(c) for o, c in zip(ops, counts)}
