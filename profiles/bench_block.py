"""block_gmres! throughput on the BASELINE Poisson matrix (get_div_grad(215), n = 9.9e6) for p right-hand sides,
restarted with memory = 5, against the algorithmic bytes of the panel formulation:

    B_step(k) = matrix + n p v (4 k + 7)       k = 1..memory within a cycle
    (SpMM 2, first product 2, k fused update+product passes 4k - 1, panel QR 2 + 2), plus per restart cycle the
    residual (matrix + 10 n p v incl. the QR of R0) and the solution update (3 memory + 4) n p v

One JSON line per p.   python profiles/bench_block.py [p ...] [--small]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "krylov.jl_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import krylov_b200 as kb  # noqa: E402
from krylov_b200 import problems as P  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
small = "--small" in sys.argv
ps = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [4, 8, 16]
N = 48 if small else 215
dev = torch.device("cuda", 0)
rp, ci, va = P.div_grad_csr(N, xp=torch, device=dev)
n, nnz = N ** 3, int(va.shape[0])
mem, cycles = 5, 2
for p in ps:
    ws = kb.BlockGmresWorkspace(n, n, p, np.float64, memory=mem, device="cuda")
    ws.set_operator((rp, ci, va))
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    B = torch.randn((n, p), dtype=torch.float64, device=dev, generator=g)      # full-rank block (a rank-deficient one takes the 4p-pass Householder path)
    kw = dict(atol=0.0, rtol=0.0, itmax=mem * cycles, restart=True)
    ws.solve(None, B, **kw)
    st = torch.cuda.ExternalStream(kb.lib().krylov_b200_stream(ws._h), device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    l0 = ws.launches
    e0.record(st)
    reps = 2
    for _ in range(reps):
        ws.solve(None, B, **kw)
    e1.record(st)
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / reps
    steps = ws.stats.niter
    matrix = nnz * 12 + (n + 1) * 4
    panel = n * p * 8
    bytes_cycle = sum(matrix + panel * (4 * k + 7) for k in range(1, mem + 1)) + (matrix + 10 * panel) + panel * (3 * mem + 4)
    B_total = bytes_cycle * cycles
    print(json.dumps(dict(solver=f"block_gmres(memory={mem}, restart)", p=p, workload=f"get_div_grad({N}) f64, {steps} block iterations/solve",
                          block_iterations_per_s=round(steps / sec, 2), rhs_iterations_per_s=round(steps * p / sec, 1),
                          ms_per_block_iteration=round(sec / steps * 1e3, 3), launches_per_block_iteration=round((ws.launches - l0) / reps / steps, 1),
                          bytes_per_solve=B_total, achieved_GBs=round(B_total / sec / 1e9, 1), frac_of_measured_hbm=round(B_total / sec / 1e9 / PEAK, 4),
                          qr_fallbacks="n/a")), flush=True)
    ws.free()
    del B
