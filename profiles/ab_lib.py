"""A/B of two builds of libkrylov_b200.so on the SAME GPU: each (library, stages x CTAs) pair runs in its own
process (KB200_LIB selects the .so), interleaved, and reports fused-CG microseconds per iteration on cfg2."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path[:0] = [%r, %r]
import numpy as np, torch
import krylov_b200 as kb
from krylov_b200.problems import div_grad_csr
N = 215
dev = torch.device("cuda", 0)
rp, ci, va = div_grad_csr(N, xp=torch, device=dev)
n = N ** 3
b = torch.ones(n, dtype=torch.float64, device=dev)
ws = kb.CgWorkspace(n, n, np.float64, device="cuda")
ws.set_operator((rp, ci, va))
for _ in range(3):
    ws.solve(None, b, atol=0.0, rtol=0.0, itmax=100)
st = torch.cuda.ExternalStream(kb.lib().krylov_b200_stream(ws._h), device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(st)
for _ in range(5):
    ws.solve(None, b, atol=0.0, rtol=0.0, itmax=100)
e1.record(st); torch.cuda.synchronize()
print(json.dumps(dict(lib=os.path.basename(os.environ.get("KB200_LIB", "current")), stages=os.environ.get("KB200_STAGES"),
                      ctas=os.environ.get("KB200_CTAS_PER_SM"), prefetch=os.environ.get("KB200_PREFETCH"),
                      us_per_iter=round(e0.elapsed_time(e1) * 1e3 / 500, 1))))
''' % (ROOT, os.path.join(ROOT, "krylov.jl_b200"))
old = os.path.join(ROOT, "krylov.jl_b200", "lib_ab", "libkrylov_b200_f96d5dc.so")
runs = [(old, "2", "3", "0"), (None, "2", "3", "0"), (None, "3", "2", "0"), (None, "4", "2", "0"), (old, "2", "3", "0"), (None, "2", "3", "0"), (None, "3", "3", "0")]
for lib, s, c, p in runs:
    env = dict(os.environ, KB200_STAGES=s, KB200_CTAS_PER_SM=c, KB200_PREFETCH=p)
    if lib:
        env["KB200_LIB"] = lib
    else:
        env.pop("KB200_LIB", None)
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(out.stdout.strip() or out.stderr[-400:], flush=True)
