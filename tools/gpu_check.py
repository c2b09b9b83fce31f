"""Quick end-to-end check on a GPU box: smoke(), then raw timings of the two kernels."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as g

g.smoke()
import optas_b200
from optas_b200 import problems
from optas_b200.function import B200Function

prob = problems.lwr_ik()
B = 65536
P, X0 = prob.sample(B, 0)
solver = optas_b200.B200Solver(prob.opt).setup("ipopt", timing=True)
print("solver kernel", solver.kernel_info())
dev = torch.device("cuda:0")
Pd, X0d = torch.from_numpy(P).to(dev), torch.from_numpy(X0).to(dev)
Xd = torch.empty_like(X0d)
st = torch.empty(B, dtype=torch.int32, device=dev); it = torch.empty(B, dtype=torch.int32, device=dev)
kkt = torch.empty(B, dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    solver.solve_raw(Pd, X0d, Xd, None, None, st, it, kkt, stream=stream)
torch.cuda.synchronize()
solver._handle.kernel_time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    solver.solve_raw(Pd, X0d, Xd, None, None, st, it, kkt, stream=stream)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
kms, kn = solver._handle.kernel_time()
print(f"C2 device-resident: {ms:.3f} ms/solve of {B} -> {B/ms*1e3:.3e} inst/s; kernel avg {kms/kn:.3f} ms over {kn}")
stc = st.cpu().numpy(); itc = it.cpu().numpy()
print("status hist", np.bincount(stc, minlength=5), "iters mean", itc.mean(), "max", itc.max())
# host path
t0 = time.perf_counter()
solver.reset_parameters(prob.param_dict(P)); solver.reset_initial_seed(prob.seed_dict(X0)); sol = solver.solve()
t1 = time.perf_counter()
print(f"C2 host dict path: {(t1-t0)*1e3:.2f} ms -> {B/(t1-t0):.3e} inst/s; converged {solver.stats()['n_converged']}")

fk = B200Function(prob.functions["fk_jac"], timing=True)
print("fk kernel", fk.kernel_info())
for Bf in (65536, 1 << 22):
    q = torch.rand(Bf, 7, dtype=torch.float64, device=dev) * 2 - 1
    p_out = torch.empty(Bf, 3, dtype=torch.float64, device=dev); J_out = torch.empty(Bf, 21, dtype=torch.float64, device=dev)
    for _ in range(3):
        fk.eval_raw(Bf, [q], [p_out, J_out], stream=stream)
    torch.cuda.synchronize(); fk.kernel_time()
    e0.record()
    for _ in range(10):
        fk.eval_raw(Bf, [q], [p_out, J_out], stream=stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    kms, kn = fk.kernel_time()
    gb = Bf * 248 / 1e9
    print(f"FK+J B={Bf}: {ms:.4f} ms (kernel avg {kms/kn:.4f}) -> {gb/(ms*1e-3):.1f} GB/s algorithmic, {Bf/ms*1e3:.3e} evals/s")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fk_ref
    pr, Jr = fk_ref.lwr_position_and_jacobian(q[:1000].cpu().numpy())
    print("   max err vs oracle", np.abs(p_out[:1000].cpu().numpy() - pr).max(), np.abs(J_out[:1000].cpu().numpy() - Jr).max())
