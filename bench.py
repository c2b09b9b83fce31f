#!/usr/bin/env python
"""bench.py -- headline benchmark: IK problem-instances solved per second on B200.

Workload (BASELINE.json configs[1], "C2"): KUKA LWR 7-DoF position IK (example/example.py's
problem), batch = 65536 instances per GPU with random reachable p_goal, seed q_nominal.
One "step" = one pass of the hot path (bo_solve: the fused interior-point kernel) over one batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

* default arm: `value` = converged instances / s with inputs resident in HBM (CUDA events round each
  step, L2 flushed between steps, max over ranks); `e2e` = the same through the public solver API
  (reset_parameters / reset_initial_seed / solve with HOST arrays: H2D + kernel + D2H inside the
  timed region); `roofline` = the FK/Jacobian streaming kernel (HBM-bound, 248 B / evaluation),
  timed live here on inputs larger than L2; `solve_kernel` explains the solver kernel itself
  (FP64-issue / latency bound, not HBM bound -- see DESIGN.md); `cpu_baseline` = the CPU oracle
  (reference's scipy-SLSQP formulation) on a bounded sample.
* `--impl reference`: the reference's own CPU solver path for this workload -- its
  ScipyMinimizeSolver("SLSQP") formulation restated in oracle/ (CasADi/IPOPT cannot be installed in
  this image) -- on all host cores, one bounded sample per step.
Multi-GPU: launched by torchrun, one rank per GPU; instances are independent, so the batch axis is
sharded with no data-path collective ("weak" scaling: 65536 instances per rank).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# Exactly ONE line goes to stdout (the JSON result of rank 0): everything else that libraries write to file descriptor 1
# (NCCL prints its version banner there) is sent to stderr.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(result: dict) -> None:
    _RESULT_OUT.write(json.dumps(result) + "\n")
    _RESULT_OUT.flush()

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH_PER_GPU = 65536
FK_BATCH = 1 << 22  # 4 Mi evaluations: 1.04 GB of algorithmic traffic per launch (> 126 MB L2)
FK_BYTES_PER_EVAL = 248  # 7 q in + 3 p out + 21 J out, float64
CPU_SAMPLE = 16384  # ~10 s of SLSQP on 16 host cores


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._th = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([t.strip() for t in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self) -> dict:
        import statistics

        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def measured_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def fp64_peak_tflops(dev, stream) -> dict:
    """Measured FP64 FMA throughput of this GPU: a tape of 16 independent Horner chains (128 multiply-adds
    each, 4096 flop per evaluation, 256 B of I/O) pushed through the same streaming kernel as K1.
    MEASURED_PEAKS.json has no FP64 figure; this is the denominator for the solver kernel's FP64 fraction."""
    import torch
    import optas_b200.sym as cs
    from optas_b200.function import B200Function

    x = cs.SX.sym("x", 16)
    outs = []
    for k in range(16):
        v = x[k]
        for j in range(128):
            v = v * x[(k + 1) % 16] + (0.5 + 0.001 * j)
        outs.append(v)
    fn = B200Function(cs.Function("fma", [x], [cs.vertcat(*outs)]), timing=True)
    Bf = 1 << 21
    xin = torch.rand((Bf, 16), dtype=torch.float64, device=dev) * 0.5
    out = torch.empty((Bf, 16), dtype=torch.float64, device=dev)
    for _ in range(3):
        fn.eval_raw(Bf, [xin], [out], stream=stream)
    torch.cuda.synchronize()
    fn.kernel_time()
    for _ in range(5):
        fn.eval_raw(Bf, [xin], [out], stream=stream)
    torch.cuda.synchronize()
    ms, n = fn.kernel_time()
    flops = Bf * 16 * 128 * 2
    return {"tflops": flops / (ms / n * 1e-3) / 1e12, "how": "16x128 FMA Horner chains per evaluation, 2 Mi evaluations, bo_eval_kernel",
            "registers": fn.kernel_info()["registers"]}


def tape_flops(tape) -> int:
    """Arithmetic operations of one evaluation of a tape (each +,-,*,/,sqrt,sq,neg = 1; sin/cos = 1 each)."""
    hist = tape.op_histogram()
    return int(sum(v for k, v in hist.items() if k not in ("INPUT", "OUTPUT", "CONST")))


def cpu_baseline(sample: int, workers: int) -> dict:
    """The CPU oracle (reference's SLSQP formulation, scipy defaults as the reference runs it) on
    `sample` instances of the same workload, one instance at a time per worker process."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import slsqp_driver
    from optas_b200 import problems

    prob = problems.lwr_ik()
    P, X0 = prob.sample(sample, seed=12345)
    slsqp_driver.solve_batch(problems.lwr_ik, P[:workers * 2], X0[:workers * 2], workers=workers)  # warm the pool path
    t0 = time.perf_counter()
    X, ok, nit = slsqp_driver.solve_batch(problems.lwr_ik, P, X0, workers=workers)
    dt = time.perf_counter() - t0
    return {"value": float(ok.sum() / dt), "unit": "instances/s", "cores": workers, "kind": "port",
            "sample": f"{sample} instances of the C2 workload, scipy SLSQP (reference ScipyMinimizeSolver formulation, "
                      f"default tolerances) on the oracle's C tape VM, {workers} worker processes; "
                      f"{int(ok.sum())}/{sample} reported success, mean {float(nit.mean()):.1f} iterations",
            "seconds": dt}


def workload_config(batch: int, world: int) -> dict:
    """`config` of the JSON line -- identical for the B200 arm and the reference arm."""
    return {"workload": "C2: KUKA LWR 7-DoF IK (example/example.py), random reachable p_goal, seed q_nominal",
            "instances_per_gpu": batch, "global_instances": batch * world, "parallelism": f"batch-sharded x{world}",
            "l2_flush_between_steps": True, "solver_tolerance": 1e-8,
            "counted": "instances reported converged only"}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = os.cpu_count() or 1
    sample = CPU_SAMPLE // 4  # per step; the whole --steps K --warmup W run stays within minutes
    times, solved = [], []
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import slsqp_driver
    from optas_b200 import problems

    prob = problems.lwr_ik()
    for step in range(args.warmup + args.steps):
        P, X0 = prob.sample(sample, seed=1000 + step)
        t0 = time.perf_counter()
        X, ok, nit = slsqp_driver.solve_batch(problems.lwr_ik, P, X0, workers=workers)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
            solved.append(int(ok.sum()))
    total_t = sum(times)
    value = sum(solved) / total_t
    line = {
        "impl": "reference", "metric": "IK problem-instances solved/sec", "value": value, "unit": "instances/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.batch, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "instances/s", "cores": workers, "kind": "port",
                         "sample": f"{sample} instances of the workload per step (bounded sample of the {args.batch}-instance "
                                   f"batch), {workers} worker processes; reference's ScipyMinimizeSolver('SLSQP') formulation "
                                   "restated in oracle/ (CasADi+IPOPT is not installable in this image)"},
        "e2e": {"value": value, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


OTHER_CONFIGS = {
    # name: (problem factory, total batch of BASELINE.json, scaling, solver options, label)
    "c3": ("point_mass_mpc", 16384, "weak", {}, "C3: point_mass_mpc.py Controller tick, T=20 (nx 80, 42 eq, 180 ineq)"),
    "c4": ("figure_eight", 4096, "weak", {"max_iter": 400, "max_trips": 2500},
           "C4: figure_eight_plan.py T=50 + joint-limit bounds (nx 693, 557 eq, 700 ineq)"),
    "c5": ("dual_arm", 32768, "strong", {}, "C5: dual_arm.py T=50 (nx 1386, 700 eq), 32768 instances sharded over the GPUs"),
    # SURVEY.md 8f-3 rows (not BASELINE configs): the rest of the RobotModel surface through the same solver
    "jsp": ("joint_space_planner", 8192, "weak", {},
            "8f-3: simple_joint_space_planner.py T=20 (nx 280, 154 eq incl. pose goal, 40 link-height ineq)"),
    "qp": ("planar_idk", 65536, "weak", {},
           "8f-1: planar_idk.py differential-IK QP (QuadraticCostLinearConstraints: nx 3, 2 eq, 8 ineq) through the general kernel"),
    "aik": ("lwr_axis_ik", 65536, "weak", {},
            "8f-3: sphere_collision_avoidance.py first stage, position + tool-axis IK (nx 21, 20 eq, 14 bounds)"),
}


def run_other_config(args) -> None:
    """Extra bench lines for C3 / C4 / C5 (not read by the driver): device-resident `value` and `e2e`."""
    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    import optas_b200
    from optas_b200 import problems
    from optas_b200.distributed import shard_slice

    factory, total, scaling, opts, label = OTHER_CONFIGS[args.config]
    prob = getattr(problems, factory)()
    if scaling == "strong":
        lo_, hi_ = shard_slice(total, rank, world)
        P, X0 = prob.sample(total)
        P, X0 = np.ascontiguousarray(P[lo_:hi_]), np.ascontiguousarray(X0[lo_:hi_])
    else:
        P, X0 = prob.sample(total, seed=rank + 1)
    B = X0.shape[0]
    # the e2e leg copies its inputs from page-locked host memory (as the bench contract says)
    from optas_b200.solver import host_array

    def pinned(a):
        out = host_array(a.shape)
        out[...] = a
        return out

    P, X0 = pinned(P), pinned(X0)
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt", opts, timing=True)
    Pd, X0d = torch.from_numpy(P).to(dev), torch.from_numpy(X0).to(dev)
    Xd = torch.empty_like(X0d)
    std = torch.empty(B, dtype=torch.int32, device=dev)
    itd = torch.empty(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    steps, warmup = max(1, min(args.steps, 3)), 1
    for _ in range(warmup):
        solver.solve_raw(Pd, X0d, Xd, None, None, std, itd, None, stream=stream)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        solver.solve_raw(Pd, X0d, Xd, None, None, std, itd, None, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1)
    n_conv = int((std <= 1).sum().item())
    for _ in range(3):  # untimed, and holding the result like the timed loop does: two sets of page-locked result buffers
        r = solver.solve_arrays(P, X0)  # exist only from the third call on (tools/e2e_probe.py: 56 / 41 / 5.1 / 5.1 ms per call)
    t0 = time.perf_counter()
    for _ in range(steps):
        r = solver.solve_arrays(P, X0)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
    c = torch.tensor([float(n_conv), float(B)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    if rank == 0:
        emit({
            "metric": "problem-instances solved/sec", "value": float(c[0]) * steps / (float(t[0]) * 1e-3), "unit": "instances/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": float(t[0]) / steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": label, "global_instances": int(c[1]), "parallelism": f"batch-sharded x{world}"},
            "converged_fraction": float(c[0]) / float(c[1]), "mean_iterations": float(itd.float().mean().item()),
            "e2e": {"value": float(c[0]) * steps / float(t[1]), "unit": "instances/s",
                    "h2d_bytes_per_step": int(P.nbytes + X0.nbytes), "d2h_bytes_per_step": int(r["x"].nbytes + r["lam"].nbytes + 28 * B)},
            "gpu_launches": steps, "solve_kernel": solver.kernel_info(),
            "tier": {k: v for k, v in solver.tier_info().items() if k in (
                "tier", "threads_per_block", "smem_dynamic", "blocks_per_sm", "levels", "segments", "factor_vals",
                "factor_madds", "factor_steps", "solve_steps", "ldl_warps", "ldl_g", "solve_g", "generated_tapes",
                "kkt_components", "kkt_classes", "kkt_code_rows", "kkt_total_instr")}})
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5", "jsp", "aik", "qp"],
                    help="c2 (default) is the headline line the driver reads; c3/c4/c5 print an extra line for the other "
                         "BASELINE.json configs (device-resident value + e2e only)")
    args = ap.parse_args()
    if args.config != "c2":
        run_other_config(args)
        return
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the B200 path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    import optas_b200
    from optas_b200 import problems
    from optas_b200.function import B200Function

    B = args.batch
    prob = problems.lwr_ik()
    P, X0 = prob.sample(B, seed=rank)  # every rank gets its own shard of the (conceptually global) batch
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt", timing=True)
    nx, npar = prob.opt.nx, prob.opt.np
    nlam = solver._lowered.n_eq + solver._lowered.n_ineq

    # ---------------- device-resident arm (`value`) ----------------
    Pd, X0d = torch.from_numpy(P).to(dev), torch.from_numpy(X0).to(dev)
    Xd = torch.empty((B, nx), dtype=torch.float64, device=dev)
    lamd = torch.empty((B, nlam), dtype=torch.float64, device=dev)
    fd = torch.empty(B, dtype=torch.float64, device=dev)
    std = torch.empty(B, dtype=torch.int32, device=dev)
    itd = torch.empty(B, dtype=torch.int32, device=dev)
    kktd = torch.empty(B, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream

    def step_device():
        solver.solve_raw(Pd, X0d, Xd, lamd, fd, std, itd, kktd, stream=stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    solver._handle.kernel_time()  # reset the in-library launch timers
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:
        barrier()
        for a, b in ev:
            flush.zero_()  # L2 flush between timed iterations (outside the event pair)
            a.record()
            step_device()
            b.record()
        barrier()
        step_ms = [a.elapsed_time(b) for a, b in ev]
        dev_ms = float(sum(step_ms))
        kernel_ms, kernel_n = solver._handle.kernel_time()
        n_conv = int((std <= 1).sum().item())
        iters_total = int(itd.sum().item())

        # ---------------- end-to-end arm (`e2e`): public API, host arrays ----------------
        p_dict, x0_dict = prob.param_dict(P), prob.seed_dict(X0)
        for _ in range(3):  # untimed; holds the result like the timed loop (page-locked buffer sets, tools/e2e_probe.py)
            solver.reset_parameters(p_dict)
            solver.reset_initial_seed(x0_dict)
            sol = solver.solve()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            solver.reset_parameters(p_dict)
            solver.reset_initial_seed(x0_dict)
            sol = solver.solve()
            n_conv_e2e = solver.stats()["n_converged"]
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0

        # ---------------- roofline arm: FK/Jacobian streaming kernel ----------------
        fk = B200Function(prob.functions["fk_jac"], timing=True)
        lo = torch.from_numpy(prob.models["robot"].lower_actuated_joint_limits.toarray().flatten()).to(dev)
        up = torch.from_numpy(prob.models["robot"].upper_actuated_joint_limits.toarray().flatten()).to(dev)
        q = lo + (up - lo) * torch.rand((FK_BATCH, 7), dtype=torch.float64, device=dev)
        p_out = torch.empty((FK_BATCH, 3), dtype=torch.float64, device=dev)
        J_out = torch.empty((FK_BATCH, 21), dtype=torch.float64, device=dev)
        for _ in range(3):
            fk.eval_raw(FK_BATCH, [q], [p_out, J_out], stream=stream)
        torch.cuda.synchronize()
        fk.kernel_time()
        fk_launches = 10
        for _ in range(fk_launches):
            fk.eval_raw(FK_BATCH, [q], [p_out, J_out], stream=stream)
        torch.cuda.synchronize()
        fk_ms_total, fk_n = fk.kernel_time()
        fp64 = fp64_peak_tflops(dev, stream)
    fk_ms = fk_ms_total / fk_n

    # max over ranks of the device time, sum over ranks of the solved instances
    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
    c = torch.tensor([n_conv, n_conv_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_s_max = float(t[0]), float(t[1])
    conv_total, conv_e2e_total = float(c[0]), float(c[1])

    if rank == 0:
        peaks = measured_peaks()
        fk_gbs = FK_BATCH * FK_BYTES_PER_EVAL / (fk_ms * 1e-3) / 1e9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "fk_jac_traffic.json")
        if os.path.exists(prof):
            with open(prof) as fh:
                traffic = json.load(fh).get("dram_bytes_per_launch")
        line = {
            "metric": "IK problem-instances solved/sec", "value": conv_total * args.steps / (dev_ms_max * 1e-3),
            "unit": "instances/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(B, world),
            "converged_fraction": conv_total / (B * world),
            "mean_iterations": iters_total / B,
            "e2e": {"value": conv_e2e_total * args.steps / e2e_s_max, "unit": "instances/s",
                    "h2d_bytes_per_step": B * (npar + nx) * 8,
                    "d2h_bytes_per_step": B * (nx * 8 + nlam * 8 + 8 + 4 + 4 + 8),
                    "api": "B200Solver.reset_parameters/reset_initial_seed/solve with host numpy arrays"},
            "gpu_launches": int(kernel_n),
            "roofline": {"kernel": "bo_eval_kernel (FK position + linear Jacobian, LWR 7-DoF)", "bound": "hbm",
                         "achieved": fk_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": fk_gbs / peaks["hbm_gbs"],
                         "traffic": traffic, "peak_source": peaks["source"] + " (burst copy figure; kernel timed alone)",
                         "bytes_per_eval": FK_BYTES_PER_EVAL, "evals_per_launch": FK_BATCH, "ms_per_launch": fk_ms,
                         "launches_timed": int(fk_n)},
            "solve_kernel": {"kernel": "bo_solve_kernel", "bound": "fp64 issue / iteration latency (not HBM)",
                             "ms_per_launch": kernel_ms / max(1, kernel_n),
                             "algorithmic_io_bytes_per_launch": B * 200,
                             "io_gbs": B * 200 / (kernel_ms / max(1, kernel_n) * 1e-3) / 1e9,
                             "registers": solver.kernel_info()["registers"],
                             "local_bytes": solver.kernel_info()["local_bytes"],
                             "tape_flop_per_iteration": tape_flops(solver._lowered.kkt) + tape_flops(solver._lowered.fc),
                             "achieved_tflops_tapes_only": iters_total * (tape_flops(solver._lowered.kkt) + tape_flops(solver._lowered.fc))
                             / (kernel_ms / max(1, kernel_n) * 1e-3) / 1e12,
                             "fp64_peak_measured": fp64},
            "clocks": clocks.summary(),
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(CPU_SAMPLE, os.cpu_count() or 1)
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
