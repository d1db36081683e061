"""GPU (>= 2 devices): row-partitioned fused CG over NVLink P2P vs the single-GPU path and the oracle."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import krylov_b200 as kb
    return kb.device_count()


@pytest.mark.parametrize("world", [2])
def test_distributed_cg_matches_single_gpu_and_oracle(tmp_path, O, world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    N = 24
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, json
        sys.path[:0] = [{ROOT!r}, {os.path.join(ROOT, 'krylov.jl_b200')!r}]
        import numpy as np, torch, torch.distributed as dist
        from krylov_b200 import dist as D, _lib
        rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local); dev = torch.device("cuda", local)
        _lib.lib().krylov_b200_set_device(local)
        dist.init_process_group("nccl", device_id=dev)
        csr, hr, ho, nloc = D.make_poisson_rank({N}, rank, world, torch, dev)
        ws = D.DistCgWorkspace(csr, hr, ho, rank, world)
        b = torch.ones(nloc, dtype=torch.float64, device=dev)
        for rep in range(2):                       # second solve re-uses the workspace (buffer swap bookkeeping)
            ws.solve(b, atol=0.0, rtol=1e-8, history=True)
        st = ws.stats
        xs = [None] * world
        dist.all_gather_object(xs, ws.x.cpu().numpy())
        if rank == 0:
            json.dump(dict(niter=st.niter, residuals=st.residuals, status=st.status, x=np.concatenate(xs).tolist()),
                      open({str(tmp_path / 'out.json')!r}, "w"))
        ws.free(); dist.destroy_process_group()
    """))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    res = json.load(open(tmp_path / "out.json"))
    A, b = O.sparse_laplacian(N)
    xo, so = O.cg(A, b, atol=0.0, rtol=1e-8)
    assert res["niter"] == so["niter"] and res["status"] == so["status"]
    assert np.allclose(res["residuals"], so["residuals"], rtol=1e-6)
    assert np.linalg.norm(np.array(res["x"]) - xo) <= 1e-6 * np.linalg.norm(xo)
