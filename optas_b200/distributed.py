"""Multi-GPU: shard the batch axis, one process per GPU, no data-path collective.

Problem instances are independent, so a batch of B instances is cut into
contiguous slices, rank r solving rows ``shard_slice(B, r, world)`` on its own GPU with its own
``bo_problem`` handle.  The only communication is the result gather afterwards (and the counters
the benchmark reduces); per-instance results are bit-identical for every world size because no
arithmetic crosses instances (tests/test_gpu.py::test_shard_equivalence..., tests/test_distributed.py).
"""

from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import numpy as np


def shard_slice(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [lo, hi) of rank ``rank``: contiguous, sizes differ by at most one, larger shards first."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def solve_sharded(solve_fn: Callable[[np.ndarray, np.ndarray], Dict[str, np.ndarray]], P: np.ndarray, X0: np.ndarray,
                  group=None, gather: bool = True) -> Dict[str, np.ndarray]:
    """Solve the rows of (P, X0) owned by this rank with ``solve_fn`` and (optionally) all-gather
    the per-instance results so that every rank returns the full batch in the original order.

    ``solve_fn(P_shard, X0_shard)`` returns a dict of arrays whose first axis is the shard's batch
    axis (for the product path: ``lambda P, X0: solver.solve_arrays(P, X0)``).
    Works with any initialised ``torch.distributed`` backend (NCCL on GPUs, gloo in CPU tests).
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = X0.shape[0]
    lo, hi = shard_slice(B, rank, world)
    local = solve_fn(np.ascontiguousarray(P[lo:hi]), np.ascontiguousarray(X0[lo:hi]))
    if world == 1 or not gather:
        return local
    use_cuda = dist.get_backend(group) == "nccl"
    out: Dict[str, np.ndarray] = {}
    biggest = shard_slice(B, 0, world)[1]
    for key, arr in local.items():
        arr = np.ascontiguousarray(arr)
        pad = np.zeros((biggest,) + arr.shape[1:], dtype=arr.dtype)
        pad[: arr.shape[0]] = arr
        t = torch.from_numpy(pad)
        if use_cuda:
            t = t.cuda()
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t, group=group)
        rows = []
        for r, part in enumerate(parts):
            a, b = shard_slice(B, r, world)
            rows.append(part[: b - a].cpu().numpy())
        out[key] = np.concatenate(rows, axis=0)
    return out


def reduce_counts(values: np.ndarray, op: str = "sum", group=None) -> np.ndarray:
    """All-reduce a small float64 vector of counters / timings (sum or max)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return np.asarray(values, dtype=np.float64)
    t = torch.tensor(np.asarray(values, dtype=np.float64))
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX, group=group)
    return t.cpu().numpy()
