"""K1 (FK + Jacobian streaming kernel): creation-order vs stage-grouped DFS tape scheduling."""
import os, sys; sys.path.insert(0, ".")
import torch
for order in ("creation", "dfs", "creation", "dfs"):
    os.environ["B200OPTAS_TAPE_ORDER"] = order
    import importlib, optas_b200.tape, optas_b200.problems, optas_b200.function
    from optas_b200 import problems
    from optas_b200.function import B200Function
    prob = problems.lwr_ik()
    fk = B200Function(prob.functions["fk_jac"], timing=True)
    B = 1 << 22
    q = torch.rand((B, 7), dtype=torch.float64, device="cuda") * 4 - 2
    p = torch.empty((B, 3), dtype=torch.float64, device="cuda"); J = torch.empty((B, 21), dtype=torch.float64, device="cuda")
    for _ in range(3): fk.eval_raw(B, [q], [p, J])
    torch.cuda.synchronize(); fk.kernel_time()
    for _ in range(20): fk.eval_raw(B, [q], [p, J])
    torch.cuda.synchronize(); ms, n = fk.kernel_time()
    print(order, fk.kernel_info(), f"{ms/n:.4f} ms -> {B*248/(ms/n*1e-3)/1e9:.1f} GB/s", flush=True)
