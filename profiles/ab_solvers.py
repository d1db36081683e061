"""Same-box A/B of library builds on the fused phases of MINRES / GMRES / BiCGSTAB: every (library, solver) pair
runs profiles/bench_solvers.py in its own process (KB200_LIB selects the .so), interleaved twice.

    python profiles/ab_solvers.py krylov.jl_b200/lib_ab/libkrylov_b200_<sha>.so [...]   # "current" is always included
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = [None] + [os.path.abspath(p) for p in sys.argv[1:]]
for rep in range(2):
    for solver in ("minres", "gmres"):
        for lib in libs:
            env = dict(os.environ)
            env.pop("KB200_LIB", None)
            if lib:
                env["KB200_LIB"] = lib
            out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "bench_solvers.py"), solver],
                                 env=env, capture_output=True, text=True)
            for line in out.stdout.splitlines():
                if line.startswith("{"):
                    d = json.loads(line)
                    print(json.dumps(dict(lib=os.path.basename(lib) if lib else "current", solver=d["solver"], fused=d["fused"],
                                          us_per_iteration=d["us_per_iteration"], frac=d["frac_of_measured_hbm"])), flush=True)
            if out.returncode != 0:
                print("FAILED", lib, solver, out.stderr[-300:], flush=True)
