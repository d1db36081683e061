"""Sweeps the TMA ring depth / CTAs per SM of the fused CG K1 kernel on the benchmark matrix and prints
event-timed per-kernel durations (KB200_STAGES / KB200_CTAS_PER_SM are read when the operator is planned)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "krylov.jl_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import krylov_b200 as kb  # noqa: E402
from krylov_b200.problems import div_grad_csr  # noqa: E402

N = int(os.environ.get("SWEEP_N", "215"))
dev = torch.device("cuda", 0)
rp, ci, va = div_grad_csr(N, xp=torch, device=dev)
n, nnz = N ** 3, int(va.numel())
b = torch.ones(n, dtype=torch.float64, device=dev)
B_cg = nnz * 12 + (n + 1) * 4 + 9 * n * 8
B_k1 = nnz * 12 + (n + 1) * 4 + 4 * n * 8
B_k2 = 6 * n * 8
configs = [(4, 2), (3, 3), (2, 4), (3, 2), (2, 3), (2, 2), (6, 1), (4, 1), (1, 4), (1, 6)]
if len(sys.argv) > 1:
    configs = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for cfg in configs:
    stages, cps = cfg[0], cfg[1]
    pf = cfg[2] if len(cfg) > 2 else 1
    os.environ["KB200_STAGES"], os.environ["KB200_CTAS_PER_SM"], os.environ["KB200_PREFETCH"] = str(stages), str(cps), str(pf)
    ws = kb.CgWorkspace(n, n, np.float64, device="cuda")
    ws.set_operator((rp.clone(), ci, va))          # new tuple id -> re-upload + re-plan with this config
    for _ in range(2):
        ws.solve(None, b, atol=0.0, rtol=0.0, itmax=100)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.ExternalStream(kb.lib().krylov_b200_stream(ws._h), device=dev)
    torch.cuda.synchronize()
    e0.record(st)
    for _ in range(3):
        ws.solve(None, b, atol=0.0, rtol=0.0, itmax=100)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 300
    ws.solve(None, b, atol=0.0, rtol=0.0, itmax=100, time_kernels=True)
    k1, k2, cnt = ws.kernel_times
    print(json.dumps(dict(stages=stages, ctas_per_sm=cps, prefetch=pf, us_per_iter=round(1e3 * ms, 1), it_per_s=round(1e3 / ms, 1),
                          frac_Bcg=round(B_cg / (ms * 1e-3) / 1e9 / 6574.8, 4), k1_us=round(1e3 * k1, 1), k2_us=round(1e3 * k2, 1),
                          k1_GBs=round(B_k1 / (k1 * 1e-3) / 1e9) if k1 else None, k2_GBs=round(B_k2 / (k2 * 1e-3) / 1e9) if k2 else None,
                          timed=cnt)), flush=True)
    ws.free()
