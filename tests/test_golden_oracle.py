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
