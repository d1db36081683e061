"""GPU: input validation and lifetime rules added in round 2 (ADVICE): malformed CSR operators are errors instead
of out-of-bounds reads, a block handle may be freed through either free entry point, device inputs produced on
torch's stream are ordered before the library's private stream."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _lap(n):
    return sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1], format="csr")


def test_malformed_csr_is_rejected(kb):
    n = 40
    A = _lap(n)
    rp, ci, va = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()
    ws = kb.CgWorkspace(n, n, np.float64)
    bad_rp = rp.copy(); bad_rp[5], bad_rp[6] = rp[6], rp[5]                 # not monotone
    with pytest.raises(kb.B200Error, match="row pointers"):
        ws.set_operator((bad_rp, ci, va))
    bad_rp = rp.copy(); bad_rp[-1] -= 1                                      # does not end at nnz
    with pytest.raises(kb.B200Error, match="row pointers"):
        ws.set_operator((bad_rp, ci, va))
    with pytest.raises(kb.B200Error, match="row pointers"):
        ws.set_operator((rp + 1, ci, va, 0))                                 # 1-based pointers declared 0-based
    bad_ci = ci.copy(); bad_ci[3] = -1
    with pytest.raises(kb.B200Error, match="negative column"):
        ws.set_operator((rp, bad_ci, va))
    bad_ci = ci.copy(); bad_ci[-1] = n + 7                                   # column outside the operator
    ws.set_operator((rp, bad_ci, va))
    with pytest.raises(kb.B200Error, match="column index inconsistent|largest column"):
        ws.solve(None, np.ones(n))
    ws.set_operator((rp, ci, va))                                            # a good operator still works afterwards
    ws.solve(None, np.ones(n), atol=0.0, rtol=1e-10)
    assert ws.stats.solved and np.linalg.norm(A @ ws.x - 1) <= 1e-8 * np.sqrt(n)
    ws.free()


def test_block_handle_through_both_free_entry_points(kb):
    L = kb.lib()
    for free_first in ("krylov_workspace_free", "krylov_block_workspace_free"):
        ws = kb.BlockGmresWorkspace(64, 64, 4, np.float64, memory=3)
        h = C.c_void_p(ws._h.value)
        assert getattr(L, free_first)(h) == 0
        assert L.krylov_block_workspace_free(h) == 1 and L.krylov_workspace_free(h) == 1      # double free is safe
        ws._h = C.c_void_p()                                                                  # already freed
    out = (C.c_double * 4)()
    ws = kb.CgWorkspace(8, 8, np.float64)
    assert L.krylov_b200_get_history(ws._h, 0, None, 4) == -1                                 # NULL output buffer
    assert L.krylov_b200_get_history(ws._h, 0, out, -1) == -1
    ws.free()


def test_device_inputs_are_ordered_after_the_producer_stream(kb):
    """b is produced by a chain of torch kernels on torch's current stream immediately before the solve; the mirror
    orders the library's non-blocking stream behind it (krylov_b200_wait_stream), so the result equals the one from
    a fully synchronised b."""
    import torch
    from krylov_b200 import problems as P
    dev = torch.device("cuda", 0)
    N = 48
    rp, ci, va = P.div_grad_csr(N, xp=torch, device=dev)
    n = N ** 3
    ws = kb.CgWorkspace(n, n, np.float64, device="cuda")
    ws.set_operator((rp, ci, va))
    ref = None
    for rep in range(3):
        b = torch.zeros(n, dtype=torch.float64, device=dev)
        for k in range(40):                       # a few hundred microseconds of queued work that b depends on
            b = b + torch.sin(torch.arange(n, device=dev, dtype=torch.float64) * (k + 1) * 1e-3)
        ws.solve(None, b, atol=0.0, rtol=0.0, itmax=20, history=True)     # no torch.cuda.synchronize() in between
        hist = list(ws.stats.residuals)
        if ref is None:
            torch.cuda.synchronize()
            ws.solve(None, b, atol=0.0, rtol=0.0, itmax=20, history=True)
            ref = list(ws.stats.residuals)
        assert hist == ref
    ws.free()
