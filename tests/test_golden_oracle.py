"""The oracle reproduces the committed golden vectors (tests/golden/oracle_histories.json)."""
import json
import os

import numpy as np
import pytest

import cases

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_histories.json")))


@pytest.mark.parametrize("name", cases.NAMES)
def test_oracle_matches_golden(O, name):
    x, st = cases.run_oracle(O, name)
    g = GOLD[name]
    assert st["niter"] == g["niter"] and st["status"] == g["status"] and st["solved"] == g["solved"]
    assert np.allclose(st["residuals"], g["residuals"], rtol=1e-12, atol=0)
    assert np.allclose(x[:8], g["x_head"], rtol=1e-10)


BLOCK = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_block.json")))


@pytest.mark.parametrize("name", sorted(BLOCK))
def test_block_oracle_matches_golden(O, name):
    import scipy.sparse as sp
    from krylov_b200 import problems as P
    g = BLOCK[name]
    rp, ci, va = P.kron_unsymmetric_csr(8)
    n = len(rp) - 1
    A = sp.csr_matrix((va, ci, rp), shape=(n, n))
    B = A @ np.cos(np.outer(np.arange(1, n + 1), np.arange(1, g["p"] + 1)))
    X, st = O.block_gmres(A, B, **g["kw"])
    assert st["niter"] == g["niter"] and st["status"] == g["status"]
    assert np.allclose(st["residuals"], g["residuals"], rtol=1e-10, atol=0)
    assert np.allclose(X[:4].ravel(), g["x_head"], rtol=1e-9, atol=1e-12)
