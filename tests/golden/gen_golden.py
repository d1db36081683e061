"""Regenerates tests/golden/oracle_histories.json from the CPU oracle.

    python tests/golden/gen_golden.py

The reference (Julia) cannot run in this image, so these vectors are outputs of
the oracle -- the restatement pinned by tests/test_oracle_kat.py -- not of Krylov.jl
itself ("parity unpinned" for per-iteration histories, see oracle/README.md).
They freeze the oracle's behaviour so the GPU parity tests and later oracle edits
are checked against fixed numbers.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "krylov.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

import cases  # noqa: E402
from oracle import oracle as O  # noqa: E402

out = {}
for name in cases.NAMES:
    x, st = cases.run_oracle(O, name)
    out[name] = dict(niter=st["niter"], solved=st["solved"], inconsistent=st["inconsistent"], status=st["status"],
                     residuals=[float(v) for v in st["residuals"]], xnorm=float(np.linalg.norm(x.astype(np.float64))),
                     x_head=[float(v) for v in x[:8]])
json.dump(out, open(os.path.join(HERE, "oracle_histories.json"), "w"), indent=0)
print({k: (v["niter"], v["status"]) for k, v in out.items()})


# ---- block_gmres (oracle/krylov_oracle_block.h): n x p blocks, residual = Frobenius norm ----------------------
import scipy.sparse as sp  # noqa: E402
from krylov_b200 import problems as P  # noqa: E402


def block_case(p, **kw):
    rp, ci, va = P.kron_unsymmetric_csr(8)
    n = len(rp) - 1
    A = sp.csr_matrix((va, ci, rp), shape=(n, n))
    B = A @ np.cos(np.outer(np.arange(1, n + 1), np.arange(1, p + 1)))        # deterministic full-rank block
    X, st = O.block_gmres(A, B, **kw)
    return dict(p=p, kw=kw, niter=st["niter"], status=st["status"], residuals=[float(v) for v in st["residuals"]],
                xnorm=float(np.linalg.norm(X)), x_head=[float(v) for v in X[:4].ravel()])


blk = {f"block_gmres_kron8_p{p}_{'restart' if r else 'full'}": block_case(p, memory=6, restart=r) for p in (2, 3, 8) for r in (False, True)}
json.dump(blk, open(os.path.join(HERE, "oracle_block.json"), "w"), indent=0)
print({k: (v["niter"], v["status"]) for k, v in blk.items()})
