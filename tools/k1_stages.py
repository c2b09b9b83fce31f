"""K1 (FK + Jacobian streaming kernel): output buffering (B200OPTAS_K1_OUT_STAGES = 2 | 1) x threads per CTA."""
import os, sys; sys.path.insert(0, ".")
import torch
from optas_b200 import problems
from optas_b200.function import B200Function
prob = problems.lwr_ik()
B = 1 << 22
q = torch.rand((B, 7), dtype=torch.float64, device="cuda") * 4 - 2
ref = None
for stages, tpb, minb in [(2, 128, 1), (2, 64, 1), (1, 64, 1), (1, 128, 1), (1, 192, 1), (1, 256, 1), (1, 128, 5), (1, 128, 6), (1, 64, 10), (1, 64, 12), (2, 64, 7)]:
    os.environ["B200OPTAS_K1_OUT_STAGES"] = str(stages)
    os.environ["B200OPTAS_JIT_DEFINES"] = f"-DBO_MIN_BLOCKS={minb}" if minb > 1 else ""
    if True:
        p = torch.zeros((B, 3), dtype=torch.float64, device="cuda"); J = torch.zeros((B, 21), dtype=torch.float64, device="cuda")
        fk = B200Function(prob.functions["fk_jac"], timing=True, threads_per_block=tpb)
        for _ in range(3): fk.eval_raw(B, [q], [p, J])
        torch.cuda.synchronize(); fk.kernel_time()
        for _ in range(20): fk.eval_raw(B, [q], [p, J])
        torch.cuda.synchronize(); ms, n = fk.kernel_time()
        if ref is None: ref = (p.clone(), J.clone())
        same = bool(torch.equal(p, ref[0]) and torch.equal(J, ref[1]))
        print(f"out_stages {stages} tpb {tpb} min_blocks {minb}", fk.kernel_info(), f"{ms/n:.4f} ms -> {B*248/(ms/n*1e-3)/1e9:.1f} GB/s ({B*248/(ms/n*1e-3)/1e9/6542.1:.3f} of peak) bitwise_same={same}", flush=True)
