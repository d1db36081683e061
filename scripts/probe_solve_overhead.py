"""Where does the per-solve fixed cost go?  Times cg! solves of 25 / 50 / 100 / 200 / 400 iterations (device events on
the workspace stream and host wall clock) and fits time = a + b * iterations."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "krylov.jl_b200")]
import numpy as np, torch
import krylov_b200 as kb
from krylov_b200.problems import div_grad_csr
dev = torch.device("cuda", 0)
N = 215
rp, ci, va = div_grad_csr(N, xp=torch, device=dev)
n = N ** 3
b = torch.ones(n, dtype=torch.float64, device=dev)
ws = kb.CgWorkspace(n, n, np.float64, device="cuda")
ws.set_operator((rp, ci, va))
stream = torch.cuda.ExternalStream(kb.lib().krylov_b200_stream(ws._h), device=dev)
rows = []
for iters in (25, 50, 100, 200, 400):
    kw = dict(atol=0.0, rtol=0.0, itmax=iters)
    for _ in range(3):
        ws.solve(None, b, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record(stream)
    reps = 10
    for _ in range(reps):
        ws.solve(None, b, **kw)
    e1.record(stream); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    ev = e0.elapsed_time(e1) / reps
    rows.append((iters, ev, wall))
    print(f"iters {iters:4d}: {ev:8.3f} ms/solve (events)  {wall:8.3f} ms/solve (wall)  {1e3 * ev / iters:7.2f} us/iteration")
x = np.array([r[0] for r in rows], float); y = np.array([r[1] for r in rows])
bfit, afit = np.polyfit(x, y, 1)
print(f"fit: {afit * 1e3:.1f} us fixed per solve + {bfit * 1e3:.2f} us per iteration")
