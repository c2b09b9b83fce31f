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


def fp64_microbench() -> dict:
    """tools/fp64_microbench.cu (built by __graft_entry__.build()): scalar DFMA with 8 independent chains per thread and
    mma.sync.m8n8k4.f64 (DMMA) with 8 accumulator tiles per warp, all SMs.  Returns {} when the binary is missing."""
    exe = os.path.join(ROOT, "tools", "_build", "fp64_microbench")
    if not os.path.exists(exe):
        return {}
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
        rows = [json.loads(line) for line in out.splitlines() if line.startswith("{")]
        return {r["kernel"]: r["tflops"] for r in rows}
    except Exception:
        return {}


def fp64_peak_tflops(dev, stream) -> dict:
    """Measured FP64 throughput of this GPU (MEASURED_PEAKS.json has no FP64 figure): the denominator for the solver
    kernels' FP64 fraction is the scalar-DFMA microbenchmark (tools/fp64_microbench.cu, 8 chains per thread); the DMMA
    (mma.sync.m8n8k4.f64) figure is reported beside it.  Also kept: the same measurement through the product's own
    streaming kernel K1 -- a tape of 16 independent Horner chains (128 multiply-adds each) -- which reaches less because
    every evaluation also moves 256 B."""
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
    k1 = flops / (ms / n * 1e-3) / 1e12
    mb = fp64_microbench()
    out = {"tflops": max(k1, mb.get("dfma", 0.0)), "how": "max of: scalar DFMA microbenchmark (tools/fp64_microbench.cu, 8 chains per thread, all SMs); "
           "16x128 FMA Horner chains per evaluation through bo_eval_kernel", "dfma_microbench_tflops": mb.get("dfma"),
           "dmma_m8n8k4_microbench_tflops": mb.get("dmma m8n8k4"), "horner_through_k1_tflops": k1,
           "registers": fn.kernel_info()["registers"]}
    return out


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
            "counted": "instances with status 0 (scaled KKT error <= 1e-8); status 1 (<= 1e-6, stalled) is reported "
                       "separately as acceptable_fraction and NOT counted"}


# The other BASELINE.json configs, reported inside the headline line under "configs" (and alone with --config NAME).
#   factory, BASELINE batch, scaling over --gpus N, solver options, seed ("own": the sampler's; "zeros": x0 = 0 as the
#   reference script's first call has it, optas/solver.py:76), timed steps, CPU-baseline solver and sample per core, label
CONFIGS = {
    "c3": dict(factory="point_mass_mpc", batch=16384, scaling="weak", opts={}, seed="own", steps=3, cpu=("slsqp", 64),
               label="C3: point_mass_mpc.py Controller tick, T=20 (nx 80, 42 eq, 180 ineq), seed = hold-position trajectory"),
    "c3_zero_seed": dict(factory="point_mass_mpc", batch=16384, scaling="weak", opts={}, seed="zeros", steps=3, cpu=None,
                         label="C3 as SURVEY 8d defines it: x0 = 0 (the script's cold first tick, point_mass_mpc.py:157-161)"),
    "c4": dict(factory="figure_eight", batch=4096, scaling="weak", opts={"max_iter": 400, "max_trips": 2500}, seed="own", steps=1,
               cpu=("ipm", 2),
               label="C4: figure_eight_plan.py T=50 + joint-limit bounds (nx 693, 557 eq, 700 ineq), seed q = qc tiled, dq = 0 "
                     "(figure_eight_plan.py:123-124); the reference script has no joint limits, BASELINE.json adds them"),
    "c5": dict(factory="dual_arm", batch=32768, scaling="strong", opts={}, seed="own", steps=1, cpu=("ipm", 4),
               label="C5: dual_arm.py T=50 (nx 1386, 700 eq), 32768 instances sharded over the GPUs, seed q = qc tiled, dq = 0"),
    "c5_zero_seed": dict(factory="dual_arm", batch=32768, scaling="strong", opts={}, seed="zeros", steps=1, cpu=None,
                         label="C5 with the reference script's own seed x0 = 0 (dual_arm.py never calls reset_initial_seed)"),
    # SURVEY.md 8f rows (not BASELINE configs; only with --config NAME)
    "jsp": dict(factory="joint_space_planner", batch=8192, scaling="weak", opts={}, seed="own", steps=3, cpu=None,
                label="8f-3: simple_joint_space_planner.py T=20 (nx 280, 154 eq incl. pose goal, 40 link-height ineq)"),
    "qp": dict(factory="planar_idk", batch=65536, scaling="weak", opts={}, seed="own", steps=3, cpu=None,
               label="8f-1: planar_idk.py differential-IK QP (QuadraticCostLinearConstraints: nx 3, 2 eq, 8 ineq), dedicated QP "
                     "iteration (Mehrotra predictor-corrector, one tape evaluation per instance; csrc/jit/bo_qp_reg.cuh)"),
    "qp_general": dict(factory="planar_idk", batch=65536, scaling="weak", opts={}, setup={"qp": False}, seed="own", steps=3, cpu=None,
                       label="8f-1 comparison: the same QP batch through the general interior-point kernel (round 1's path)"),
    "aik": dict(factory="lwr_axis_ik", batch=65536, scaling="weak", opts={}, seed="own", steps=3, cpu=None,
                label="8f-3: sphere_collision_avoidance.py first stage, position + tool-axis IK (nx 21, 20 eq, 14 bounds)"),
}
DEFAULT_CONFIGS = ["c3", "c3_zero_seed", "c4", "c5", "c5_zero_seed", "qp", "qp_general"]


def config_inputs(name: str, rank: int, world: int):
    """(problem, P, X0) of this rank for one config: weak = the BASELINE batch per rank, strong = a contiguous shard."""
    import numpy as np
    from optas_b200 import problems
    from optas_b200.distributed import shard_slice

    cfg = CONFIGS[name]
    prob = getattr(problems, cfg["factory"])()
    if cfg["scaling"] == "strong":
        lo_, hi_ = shard_slice(cfg["batch"], rank, world)
        P, X0 = prob.sample(cfg["batch"])
        P, X0 = np.ascontiguousarray(P[lo_:hi_]), np.ascontiguousarray(X0[lo_:hi_])
    else:
        P, X0 = prob.sample(cfg["batch"], seed=rank + 1)
    if cfg["seed"] == "zeros":
        X0 = np.zeros_like(X0)
    return prob, P, X0


def config_cpu_baseline(name: str, workers: int) -> dict:
    """CPU arm of one config on a bounded sample: C3 = the reference's runnable SLSQP formulation (oracle/slsqp_driver.py);
    C4 / C5 (nx 693 / 1386, where dense SLSQP is O(n^3) per iteration) = the oracle's restatement of the interior-point
    algorithm the reference calls there (oracle/ipm_ref.py).  One instance at a time per worker process."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    from optas_b200 import problems

    cfg = CONFIGS[name]
    kind, per_core = cfg["cpu"]
    factory = getattr(problems, cfg["factory"])
    n = per_core * workers
    P, X0 = factory().sample(n, seed=4242)
    if cfg["seed"] == "zeros":
        X0 = np.zeros_like(X0)
    if kind == "slsqp":
        import slsqp_driver as drv

        how = "scipy SLSQP, the reference's ScipyMinimizeSolver formulation with scipy defaults (oracle/slsqp_driver.py)"
        kw = {}
    else:
        import ipm_ref as drv

        how = ("sparse interior point restating the reference's nlpsol('ipopt') call (Waechter-Biegler Algorithm A, "
               "scipy SuperLU; oracle/ipm_ref.py), tol 1e-8, step cap 0.5 as on the GPU")
        kw = {"max_step": 0.5}
    drv.solve_batch(factory, P[:workers], X0[:workers], workers=workers, **kw)  # builds the tapes in every worker once
    t0 = time.perf_counter()
    X, ok, nit = drv.solve_batch(factory, P, X0, workers=workers, **kw)
    dt = time.perf_counter() - t0
    return {"value": float(ok.sum() / dt), "unit": "instances/s", "cores": workers, "kind": "port",
            "sample": f"{n} instances of the workload ({per_core} per core; FLAGGED: a sample, not the {cfg['batch']}-instance "
                      f"batch), {how}, {workers} worker processes; {int(ok.sum())}/{n} converged, mean {float(nit.mean()):.1f} iterations",
            "seconds": dt}


def run_config(name: str, rank: int, world: int, dev, dist, fp64_tflops, with_cpu: bool) -> dict:
    """One BASELINE config on this job's GPUs: device-resident `value`, `e2e` from page-locked host arrays through
    B200Solver.solve_arrays, an FP64 `roofline` (these kernels move ~KB per instance: FP64 issue / latency bound, not
    HBM) and, at N=1, the CPU arm on a bounded sample.  Collective calls inside: every rank must call it."""
    import numpy as np
    import torch
    import optas_b200
    from optas_b200.solver import host_array

    cfg = CONFIGS[name]
    prob, P, X0 = config_inputs(name, rank, world)
    B = X0.shape[0]

    def pinned(a):
        out = host_array(a.shape)
        out[...] = a
        return out

    P, X0 = pinned(P), pinned(X0)
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt", cfg["opts"], timing=True, **cfg.get("setup", {}))
    lo = solver._lowered
    Pd, X0d = torch.from_numpy(P).to(dev), torch.from_numpy(X0).to(dev)
    Xd = torch.empty_like(X0d)
    std = torch.empty(B, dtype=torch.int32, device=dev)
    itd = torch.empty(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    steps = cfg["steps"]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    solver.solve_raw(Pd, X0d, Xd, None, None, std, itd, None, stream=stream)  # warm-up (also sizes the scratch)
    barrier()
    solver._handle.kernel_time()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.zero_()
        a.record()
        solver.solve_raw(Pd, X0d, Xd, None, None, std, itd, None, stream=stream)
        b.record()
    barrier()
    dev_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    kernel_ms, kernel_n = solver._handle.kernel_time()
    n_conv = int((std == 0).sum().item())
    n_acc = int((std == 1).sum().item())
    iters_total = int(itd.sum().item())
    # end to end: page-locked host arrays in, page-locked results out (two warm calls: the second set of result buffers
    # of torch's caching host allocator exists from then on, tools/e2e_probe.py)
    for _ in range(2):
        r = solver.solve_arrays(P, X0)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        r = solver.solve_arrays(P, X0)
    e2e_s = time.perf_counter() - t0
    n_conv_e2e = int((r["status"] == 0).sum())
    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
    c = torch.tensor([n_conv, n_acc, iters_total, B, n_conv_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dev_ms, e2e_s = float(t[0]), float(t[1])
    conv, acc, its, total, conv_e2e = (float(v) for v in c)
    tier = solver.tier_info()
    # FP64 work per interior-point iteration that the plan makes explicit: both tapes once, one factorisation
    # (multiply-add = 2 flop); re-factorisations, substitutions and the extra trial points are NOT counted, so this is
    # a lower bound on the arithmetic actually done
    flop_iter = tape_flops(lo.kkt) + tape_flops(lo.fc) + 2 * int(tier.get("factor_madds", 0))
    flop_launch = its / world * flop_iter
    if tier["tier"].startswith("qp"):
        # QP iteration: the kkt tape twice per instance (seed, final check); per iteration one dense LDL' of the
        # (nx + n_eq)-row system (n^3 / 3 multiply-adds) -- substitutions and the assembly not counted
        nk = lo.nx + lo.n_eq
        flop_iter = 2 * nk ** 3 // 3
        flop_launch = its / world * flop_iter + 2 * tape_flops(lo.kkt) * B
    achieved = flop_launch / (kernel_ms / max(1, kernel_n) * 1e-3) / 1e12  # this rank's launch, TFLOP/s
    out = {
        "workload": cfg["label"], "scaling": cfg["scaling"], "global_instances": int(total), "n_gpus": world,
        "value": conv * steps / (dev_ms * 1e-3), "unit": "instances/s", "steps": steps, "ms_per_step": dev_ms / steps,
        "converged_fraction": conv / total, "acceptable_fraction": acc / total, "mean_iterations": its / total,
        "e2e": {"value": conv_e2e * steps / e2e_s, "unit": "instances/s", "h2d_bytes_per_step": int(P.nbytes + X0.nbytes),
                "d2h_bytes_per_step": int(r["x"].nbytes + r["lam"].nbytes + 28 * B), "api": "B200Solver.solve_arrays (page-locked host arrays, read / written in place by the kernel over PCIe)"},
        "gpu_launches": int(kernel_n),
        "roofline": {"kernel": f"bo_solve_kernel ({tier['tier']} tier)", "bound": "fp64 issue / dependent-instruction latency (not HBM, not tensor)",
                     "achieved": achieved, "peak": fp64_tflops, "unit": "TFLOP/s", "frac": achieved / fp64_tflops if fp64_tflops else None,
                     "traffic": None, "flop_per_iteration": flop_iter, "factor_madds": int(tier.get("factor_madds", 0)),
                     "ms_per_launch": kernel_ms / max(1, kernel_n),
                     "peak_source": "FP64 FMA microbenchmark of this run (fp64_peak_measured); MEASURED_PEAKS.json has no FP64 figure",
                     "algorithmic_io_bytes_per_instance": 8 * (lo.np_ + 2 * lo.nx) + 8},
        "tier": {k: v for k, v in tier.items() if k in ("tier", "threads_per_block", "smem_dynamic", "blocks_per_sm", "levels",
                                                          "factor_vals", "factor_madds", "kkt_total_instr")},
        "solve_kernel": solver.kernel_info(),
    }
    if with_cpu and cfg["cpu"] is not None and rank == 0:
        out["cpu_baseline"] = config_cpu_baseline(name, os.cpu_count() or 1)
    del solver
    return out


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
    cfg = workload_config(args.batch, max(1, args.gpus))
    cfg["reference_arm_sample"] = (f"each step solves a {sample}-instance sample of the {args.batch}-instance batch on {workers} host "
                                   "cores (ms_per_step is for that sample); rates compare, batch sizes do not")
    line = {
        "impl": "reference", "metric": "IK problem-instances solved/sec", "value": value, "unit": "instances/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "instances/s", "cores": workers, "kind": "port",
                         "sample": f"{sample} instances of the workload per step (bounded sample of the {args.batch}-instance "
                                   f"batch), {workers} worker processes; reference's ScipyMinimizeSolver('SLSQP') formulation "
                                   "restated in oracle/ (CasADi+IPOPT is not installable in this image)"},
        "e2e": {"value": value, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_configs:
        line["configs"] = {}
        for name in DEFAULT_CONFIGS:
            if CONFIGS[name]["cpu"] is None:
                continue
            cb = config_cpu_baseline(name, workers)
            line["configs"][name] = {"workload": CONFIGS[name]["label"], "value": cb["value"], "unit": "instances/s",
                                     "cpu_baseline": cb}
    emit(line)


def run_other_config(args) -> None:
    """`--config NAME`: one config alone, printed as a full bench line."""
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
    fp64 = fp64_peak_tflops(dev, torch.cuda.current_stream().cuda_stream)
    with ClockSampler(local_rank) as clocks:
        r = run_config(args.config, rank, world, dev, dist, fp64["tflops"], with_cpu=not args.no_cpu_baseline and world == 1)
    if rank == 0:
        line = {"metric": "problem-instances solved/sec", "value": r["value"], "unit": "instances/s", "n_gpus": world,
                "steps": r["steps"], "warmup": 1, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": r["scaling"],
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": r["workload"], "global_instances": r["global_instances"], "parallelism": f"batch-sharded x{world}",
                           "l2_flush_between_steps": True, "counted": "status 0 only"},
                "fp64_peak_measured": fp64, "clocks": clocks.summary()}
        line.update({k: v for k, v in r.items() if k not in ("value", "unit", "steps", "ms_per_step", "scaling", "workload", "n_gpus", "global_instances")})
        emit(line)
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
    ap.add_argument("--schedule", default="seed_infeasibility", choices=["seed_infeasibility", "natural"],
                    help="C2 device-resident arm: order in which the batch is handed to the solver kernel")
    ap.add_argument("--no-configs", action="store_true", help="headline C2 only: skip the `configs` block (C3 / C4 / C5)")
    ap.add_argument("--config", default="c2", choices=["c2"] + sorted(CONFIGS),
                    help="c2 (default) is the line the driver reads; it carries the other BASELINE.json configs under "
                         "`configs`.  Any other name prints that config alone as a full line.")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.config != "c2":
        run_other_config(args)
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
    # device-resident batches are handed to the kernel most-infeasible-seed first (B200Solver._solve_scheduled: the predictor
    # kernel, the sort, the gather and the scatter are inside the timed step); --schedule natural = the caller's order
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt", timing=True, schedule=args.schedule)
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
        n_conv = int((std == 0).sum().item())
        n_acc = int((std == 1).sum().item())
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
            n_conv_e2e = int((solver.stats()["status"] == 0).sum())
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
    del fk, q, p_out, J_out

    # max over ranks of the device time, sum over ranks of the solved instances
    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
    c = torch.tensor([n_conv, n_conv_e2e, n_acc, iters_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_s_max = float(t[0]), float(t[1])
    conv_total, conv_e2e_total, acc_total, iters_all = (float(v) for v in c)
    info = solver.kernel_info()
    tier = solver.tier_info()
    kkt_flops, fc_flops = tape_flops(solver._lowered.kkt), tape_flops(solver._lowered.fc)
    del solver, Pd, X0d, Xd, lamd

    # ---------------- the other BASELINE configs (every rank takes part) ----------------
    configs = {}
    if not args.no_configs:
        for name in DEFAULT_CONFIGS:
            configs[name] = run_config(name, rank, world, dev, dist, fp64["tflops"], with_cpu=not args.no_cpu_baseline and world == 1)

    if rank == 0:
        peaks = measured_peaks()
        fk_gbs = FK_BATCH * FK_BYTES_PER_EVAL / (fk_ms * 1e-3) / 1e9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "fk_jac_traffic.json")
        if os.path.exists(prof):
            with open(prof) as fh:
                traffic = json.load(fh).get("dram_bytes_per_launch")
        ms_launch = kernel_ms / max(1, kernel_n)
        line = {
            "metric": "IK problem-instances solved/sec", "value": conv_total * args.steps / (dev_ms_max * 1e-3),
            "unit": "instances/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(B, world),
            "converged_fraction": conv_total / (B * world),
            "acceptable_fraction": acc_total / (B * world),
            "mean_iterations": iters_all / (B * world),
            "e2e": {"value": conv_e2e_total * args.steps / e2e_s_max, "unit": "instances/s",
                    "h2d_bytes_per_step": B * (npar + nx) * 8,
                    "d2h_bytes_per_step": B * (nx * 8 + nlam * 8 + 8 + 4 + 4 + 8),
                    "api": "B200Solver.reset_parameters/reset_initial_seed/solve with host numpy arrays",
                    "transfers": "dict arrays are packed into page-locked rows (bo_pack_rows), which the kernel reads and "
                                 "whose page-locked result arrays it writes in place over PCIe (zero copy, inside the timed "
                                 "region: the bytes above cross the bus every step); B200OPTAS_ZERO_COPY=0 = staged copies"},
            "gpu_launches": int(kernel_n) + (args.steps if args.schedule == "seed_infeasibility" else 0),
            "schedule": {"order": args.schedule, "what": "seed_infeasibility: per step one bo_eval_kernel launch evaluates theta(x0, p) = "
                         "sum max(0, -v)^2 per instance, torch sorts / gathers on the same stream, bo_solve_kernel takes the batch "
                         "most-infeasible-seed first, results are scattered back (all inside the timed step; results bitwise "
                         "those of the natural order, tests/test_gpu.py)"},
            "roofline": {"kernel": "bo_eval_kernel (FK position + linear Jacobian, LWR 7-DoF)", "bound": "hbm",
                         "achieved": fk_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": fk_gbs / peaks["hbm_gbs"],
                         "traffic": traffic, "peak_source": peaks["source"] + " (burst copy figure; kernel timed alone)",
                         "bytes_per_eval": FK_BYTES_PER_EVAL, "evals_per_launch": FK_BATCH, "ms_per_launch": fk_ms,
                         "launches_timed": int(fk_n)},
            "solve_kernel": {"kernel": f"bo_solve_kernel ({tier['tier']} tier)", "bound": "fp64 issue / iteration latency (not HBM)",
                             "ms_per_launch": ms_launch,
                             "algorithmic_io_bytes_per_launch": B * 200,
                             "io_gbs": B * 200 / (ms_launch * 1e-3) / 1e9,
                             "registers": info["registers"], "local_bytes": info["local_bytes"],
                             "smem_bytes_per_cta": tier.get("smem_dynamic", 0) + max(0, info.get("smem_bytes", 0)),
                             "threads_per_block": tier.get("threads_per_block"), "blocks_per_sm": tier.get("blocks_per_sm"),
                             "tape_flop_per_iteration": kkt_flops + fc_flops,
                             "achieved_tflops_tapes_only": iters_all / world * (kkt_flops + fc_flops) / (ms_launch * 1e-3) / 1e12,
                             "fp64_peak_measured": fp64},
            "clocks": clocks.summary(),
        }
        line["solve_kernel"]["frac_of_fp64_peak_tapes_only"] = line["solve_kernel"]["achieved_tflops_tapes_only"] / fp64["tflops"]
        if configs:
            line["configs"] = configs
            line["gpu_launches_configs"] = int(sum(c.get("gpu_launches", 0) for c in configs.values()))
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(CPU_SAMPLE, os.cpu_count() or 1)
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
