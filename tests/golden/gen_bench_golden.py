"""Golden residual history for bench.py's cfg5 parity block (BASELINE config 5: cg! on get_div_grad(464,464,464),
b = ones, 50 fixed iterations).  n ~ 1e8 is too large for an oracle solve inside the bench run, so the history of
the CPU oracle (oracle/krylov_oracle.c: oracle_cg_timed_f64, the cg.jl:195-268 loop) is generated here once and
committed as tests/golden/bench_cg_poisson464.json.

    python tests/golden/gen_bench_golden.py [N] [iters] [threads]     (about 25 GB of host RAM for N = 464)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "krylov.jl_b200")]
import numpy as np  # noqa: E402

from krylov_b200.problems import div_grad_csr  # noqa: E402
from oracle import oracle as O  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 464
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
threads = int(sys.argv[3]) if len(sys.argv) > 3 else (os.cpu_count() or 1)
rp, ci, va = div_grad_csr(N)
b = np.ones(N ** 3)
t, x, rn, hist = O.cg_timed(rp, ci, va, b, iters, threads, history=True)
out = dict(problem=f"get_div_grad({N},{N},{N}), b = ones, Float64", n=N ** 3, nnz=int(len(va)), iters=iters,
           oracle="oracle_cg_timed_f64 (cg.jl:195-268 restated), OpenMP threads = %d" % threads,
           residuals=[float(h) for h in hist], x_norm=float(np.linalg.norm(x)), seconds=round(t, 1))
name = "bench_cg_poisson%d.json" % N
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), name), "w"), indent=1)
print(name, "written:", iters, "iterations in", round(t, 1), "s; final rNorm", hist[-1])
