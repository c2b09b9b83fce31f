"""Host-side sharding logic under torch.distributed (gloo, world_size 2, CPU).  The compute inside
each rank is the host-compiled solver harness (tests/hostsim.py) standing in for the GPU call; what
is under test is the slicing / gather / counter-reduction plumbing of optas_b200.distributed."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_slices_partition_the_batch():
    from optas_b200.distributed import shard_slice

    for B in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_slice(B, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_slice(10, 2, 2)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    import optas_b200
    from hostsim import HostSim
    from optas_b200 import problems
    from optas_b200.distributed import reduce_counts, solve_sharded

    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = problems.lwr_ik()
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt", compile_only=True)
    lo = solver._lowered
    sim = HostSim(solver.kernel_source(), lo.nx, lo.np_, lo.n_eq, lo.n_ineq, ldl_table=solver.ldl_table())
    P, X0 = prob.sample(101, seed=9)  # odd size: ragged shards
    res = solve_sharded(lambda p, x0: sim.solve(p, x0), P, X0)
    n_ok = reduce_counts(np.array([float((res["status"] <= 1).sum())]), "max")
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), n_ok=n_ok, **res)
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import optas_b200
    from hostsim import HostSim
    from optas_b200 import problems

    prob = problems.lwr_ik()
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt", compile_only=True)
    lo = solver._lowered
    sim = HostSim(solver.kernel_source(), lo.nx, lo.np_, lo.n_eq, lo.n_ineq, ldl_table=solver.ldl_table())
    P, X0 = prob.sample(101, seed=9)
    whole = sim.solve(P, X0)
    for key in ("x", "lam", "f", "status", "iters", "kkt"):
        assert np.array_equal(r0[key], r1[key]), key          # every rank holds the full gathered batch
        assert np.array_equal(r0[key], whole[key]), key       # and it is bit-identical to the unsharded solve
    assert r0["n_ok"][0] == (whole["status"] <= 1).sum()
