"""CPU: the bench contract.  (1) `bench.py --impl reference` runs without a GPU and prints ONE JSON line with the keys
the driver reads (same `config.workload` string as our arm).  (2) The committed round-2 bench lines (profiles/) carry
the blocks the task statement asks for -- roofline, e2e, parity (green), cfg5 -- at every N."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_runs_on_cpu_and_matches_the_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "poisson32",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "it/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"]["workload"] == bench.workload_name(*bench.WORKLOADS["poisson32"])


@pytest.mark.parametrize("fn,n", [("r2_bench_n1_default_final.json", 1), ("r2_bench_n2_final.json", 2), ("r2_bench_n4_final.json", 4),
                                  ("r2_bench_n8_final.json", 8)])
def test_committed_bench_lines_carry_green_parity_and_cfg5(fn, n):
    path = os.path.join(ROOT, "profiles", fn)
    if not os.path.exists(path):
        pytest.skip(fn + " not committed")
    d = json.loads(open(path).read().strip().splitlines()[-1])
    assert d["n_gpus"] == n and d["metric"] == "CG iterations/s" and d["dtype"] == "f64"
    assert d["roofline"]["bound"] == "hbm" and 0 < d["roofline"]["frac"] <= 1.0
    assert d["parity"]["ok"] is True and d["parity"]["max_rel_dev"] <= 1e-6 and d["parity"]["iters_compared"] >= 100
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["gpu_launches"] > 0
    assert d["cfg5"]["parity"]["ok"] is True and d["cfg5"]["value"] > 0 and d["cfg5"]["n_gpus"] == n
    assert not d["clocks"]["reasons"] or set(d["clocks"]["reasons"]) <= {"sw_power_cap"}
    if n == 1:
        assert d["cpu_baseline"]["value"] > 0
        assert {e["solver"] for e in d["extra"]} == {"gmres(30)", "bicgstab"}
    else:
        assert d["parity"]["ranks_agree"] is True and d["cfg5"]["parity"]["ranks_agree"] is True
