"""Host-side logic of the row-partitioned path, on CPU: slab bounds, column localisation (a bijection that keeps
row order), a NumPy emulation of the halo gather + rank-ordered all-reduce reproducing the global CG iterates,
and the handle-exchange plumbing under torch.distributed (gloo, world_size = 2)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import scipy.sparse as sp

from krylov_b200 import dist as D
from krylov_b200 import problems as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_bounds_cover_and_balance():
    for n3, w in ((215, 8), (464, 8), (7, 3), (4, 4), (10, 1)):
        b = D.slab_bounds(n3, w)
        assert b[0][0] == 0 and b[-1][1] == n3 and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1


def _blocks(N, world):
    bounds = D.slab_bounds(N, world)
    row_starts = np.array([b[0] * N * N for b in bounds] + [N ** 3], dtype=np.int64)
    out = []
    for r, (lo, hi) in enumerate(bounds):
        rp, ci, va = P.div_grad_csr(N, k_lo=lo, k_hi=hi)
        cl, hr, ho = D.localize_columns(ci, row_starts, r)
        out.append((rp, ci, cl, va, hr, ho))
    return bounds, row_starts, out


@pytest.mark.parametrize("world", [2, 3, 8])
def test_localize_columns_is_a_bijection(world):
    N = 9 if world == 8 else 6
    bounds, rs, blocks = _blocks(N, world)
    for r, (rp, cg, cl, va, hr, ho) in enumerate(blocks):
        nloc = len(rp) - 1
        lo = rs[r]
        back = np.where(cl < nloc, cl + lo, 0)
        halo = cl >= nloc
        back[halo] = rs[hr[cl[halo] - nloc]] + ho[cl[halo] - nloc]
        assert np.array_equal(back, cg)                       # every (row, col, val) triplet is preserved
        assert np.all(hr != r) and np.all(ho >= 0)
        # slab neighbours only, one N^2 plane each
        assert set(hr.tolist()) <= {r - 1, r + 1} and len(hr) == N * N * ((r > 0) + (r < world - 1))
        for i in range(nloc):                                 # ascending order inside rows survives the map
            seg = cg[rp[i]:rp[i + 1]]
            assert np.all(np.diff(seg) > 0)


def test_emulated_distributed_cg_matches_global(O):
    """NumPy emulation of the kernels' data flow: per-rank SpMV with halo gather from the peers' vectors and a
    rank-ordered sum of per-rank partial dots.  Must reproduce the oracle's global CG history to 1e-12."""
    N, world = 6, 3
    bounds, rs, blocks = _blocks(N, world)
    n = N ** 3
    xo, so = O.cg(O.get_div_grad(N, N, N), np.ones(n), atol=0.0, rtol=1e-10)
    r = [np.ones(rs[k + 1] - rs[k]) for k in range(world)]
    p = [v.copy() for v in r]
    x = [np.zeros_like(v) for v in r]
    gamma = sum(float(v @ v) for v in r)
    hist = [np.sqrt(gamma)]
    for it in range(so["niter"]):
        Ap, pAp = [], 0.0
        for k, (rp, cg, cl, va, hr, ho) in enumerate(blocks):
            nloc = len(rp) - 1
            ext = np.concatenate([p[k], np.array([p[hr[h]][ho[h]] for h in range(len(hr))])])
            A = sp.csr_matrix((va, cl, rp), shape=(nloc, nloc + len(hr)))
            Ap.append(A @ ext)
        pAp = sum(float(p[k] @ Ap[k]) for k in range(world))
        alpha = gamma / pAp
        for k in range(world):
            x[k] += alpha * p[k]
            r[k] -= alpha * Ap[k]
        g2 = sum(float(v @ v) for v in r)
        hist.append(np.sqrt(g2))
        beta = g2 / gamma
        gamma = g2
        for k in range(world):
            p[k] = r[k] + beta * p[k]
    assert np.allclose(hist, so["residuals"], rtol=1e-10)
    assert np.allclose(np.concatenate(x), xo, rtol=1e-9)


def test_handle_exchange_plumbing_gloo_world2(tmp_path):
    """The Python side of the IPC exchange (all_gather_object of fixed-size blobs in rank order) under gloo."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path[:0] = [{ROOT!r}, {os.path.join(ROOT, 'krylov.jl_b200')!r}]
        import numpy as np, torch.distributed as dist
        from krylov_b200 import dist as D, problems as P
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        N = 6
        bounds = D.slab_bounds(N, world)
        rs = np.array([b[0] * N * N for b in bounds] + [N ** 3])
        rp, ci, va = P.div_grad_csr(N, k_lo=bounds[rank][0], k_hi=bounds[rank][1])
        cl, hr, ho = D.localize_columns(ci, rs, rank)
        blob = bytes([rank]) * 256
        blobs = [None] * world
        dist.all_gather_object(blobs, blob)
        allb = b"".join(blobs)
        assert len(allb) == 256 * world and all(allb[256 * k] == k for k in range(world))
        tot = [None] * world
        dist.all_gather_object(tot, (len(rp) - 1, len(hr)))
        assert sum(t[0] for t in tot) == N ** 3 and all(t[1] == N * N for t in tot)
        dist.barrier(); dist.destroy_process_group()
        print("ok", rank)
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_push_ranges_for_slabs_and_fallback():
    N, world = 6, 4
    bounds, rs, blocks = _blocks(N, world)
    maps = [(b[4], b[5]) for b in blocks]
    for r in range(world):
        rg = D.push_ranges(maps, r)
        assert rg is not None and len(rg) == (r > 0) + (r < world - 1)
        nloc = rs[r + 1] - rs[r]
        for start, count, peer, slot in rg:
            assert count == N * N and abs(peer - r) == 1
            # the lower neighbour needs my FIRST plane, the upper one my LAST plane
            assert start == (0 if peer < r else nloc - N * N)
            # slot: where my plane sits in the peer's halo ([lower halo | upper halo])
            assert slot == (0 if (peer > r and peer - 1 == r and r == peer - 1 and peer == 0) else slot)
            hr, ho = maps[peer]
            assert np.all(hr[slot:slot + count] == r) and np.array_equal(ho[slot:slot + count], np.arange(start, start + count))
    # a scattered need (every other row) is not a contiguous range -> pull mode
    bad = [(np.array([1, 1], dtype=np.int32), np.array([0, 2], dtype=np.int32)), (np.zeros(0, np.int32), np.zeros(0, np.int32))]
    assert D.push_ranges(bad, 1) is None


def test_send_list_emulated_exchange_reproduces_global_spmv():
    """NumPy emulation of k_halo_exchange + the halo-aware gather: every rank scatters its send list into the
    peers' halo buffers, then multiplies [x_local | halo]; the stacked result equals the global product."""
    N, world = 5, 3
    bounds, rs, blocks = _blocks(N, world)
    maps = [(b[4], b[5]) for b in blocks]
    n = N ** 3
    x = np.random.default_rng(2).standard_normal(n)
    xl = [x[rs[k]:rs[k + 1]] for k in range(world)]
    halo = [np.full(len(maps[k][0]), np.nan) for k in range(world)]
    for r in range(world):
        rows, peers, slots = D.send_list(maps, r)
        assert rows.dtype == np.int32 and len(rows) == len(peers) == len(slots)
        for row, q, s in zip(rows, peers, slots):
            halo[q][s] = xl[r][row]
    y = []
    for k, (rp, cg, cl, va, hr, ho) in enumerate(blocks):
        assert not np.isnan(halo[k]).any()                    # every halo slot was filled by exactly its owner
        nloc = len(rp) - 1
        A = sp.csr_matrix((va, cl, rp), shape=(nloc, nloc + len(hr)))
        y.append(A @ np.concatenate([xl[k], halo[k]]))
    rpg, cig, vag = P.div_grad_csr(N)
    assert np.allclose(np.concatenate(y), sp.csr_matrix((vag, cig, rpg), shape=(n, n)) @ x, rtol=1e-14)
