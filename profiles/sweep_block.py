"""A/B of the block_gmres kernel choices on the same GPU: every configuration runs profiles/bench_block.py in its own
process.   python profiles/sweep_block.py [p ...]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ps = [a for a in sys.argv[1:]] or ["8", "16"]
configs = [("default", {}), ("prefetch", {"KB200_FAST_PREFETCH": "1"}), ("spmm=rows", {"KB200_SPMM": "rows"}),
           ("fast_tpr=alt", {"KB200_FAST_TPR": "alt"}), ("generic tiled kernels", {"KB200_BLOCK_GENERIC": "1"})]
if "--quick" in sys.argv:
    configs = configs[:2]
    ps = [a for a in ps if a != "--quick"]
for name, env in configs:
    e = dict(os.environ, **env)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "bench_block.py")] + ps, env=e, capture_output=True, text=True)
    for line in out.stdout.splitlines():
        if line.startswith("{"):
            d = json.loads(line)
            print(json.dumps(dict(config=name, p=d["p"], ms_per_block_iteration=d["ms_per_block_iteration"],
                                  frac_of_measured_hbm=d["frac_of_measured_hbm"])), flush=True)
    if out.returncode != 0:
        print("FAILED", name, out.stderr[-400:], flush=True)
