"""K1 (FK + Jacobian streaming kernel): threads per CTA sweep."""
import sys; sys.path.insert(0, ".")
import torch
from optas_b200 import problems
from optas_b200.function import B200Function
prob = problems.lwr_ik()
B = 1 << 22
q = torch.rand((B, 7), dtype=torch.float64, device="cuda") * 4 - 2
p = torch.empty((B, 3), dtype=torch.float64, device="cuda"); J = torch.empty((B, 21), dtype=torch.float64, device="cuda")
for tpb in (32, 64, 96, 128, 160, 192, 256):
    fk = B200Function(prob.functions["fk_jac"], timing=True, threads_per_block=tpb)
    for _ in range(3): fk.eval_raw(B, [q], [p, J])
    torch.cuda.synchronize(); fk.kernel_time()
    for _ in range(20): fk.eval_raw(B, [q], [p, J])
    torch.cuda.synchronize(); ms, n = fk.kernel_time()
    print(tpb, fk.kernel_info(), f"{ms/n:.4f} ms -> {B*248/(ms/n*1e-3)/1e9:.1f} GB/s ({B*248/(ms/n*1e-3)/1e9/6542.1:.3f} of peak)", flush=True)
