"""GPU: the reference's own C conformance programs, compiled from /root/reference into oracle/_ref/ and
linked to libkrylov_b200.so (oracle/build_ref.sh), plus a ctypes re-statement of the C-level checks."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from krylov_b200 import _lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def test_reference_basic_cg_example():
    exe = os.path.join(REF, "basic_cg")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/basic_cg was not built (reference tree absent at build time)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert re.search(r"Solved:\s*yes\s+niter:\s*3", out.stdout), out.stdout      # basic_cg.c:13-15
    assert re.search(r"1\.00\s+1\.00\s+1\.00\s+1\.00\s+1\.00", out.stdout)


def test_reference_test_api_program():
    """interfaces/test/C/test_api.c, unmodified: every check passes (its DQGMRES workspace-option check included,
    since dqgmres! joined the sibling solvers)."""
    exe = os.path.join(REF, "test_api")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/test_api was not built (reference tree absent at build time)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    fails = re.findall(r"FAIL\s+(.*?)\s+\(", out.stdout)
    assert not fails, out.stdout + out.stderr
    m = re.search(r"(\d+) checks passed, (\d+) failed", out.stdout)
    assert m and int(m.group(1)) >= 40 and int(m.group(2)) == 0, out.stdout


def _tridiag_cb(n, diag=2.0, off=-1.0, dt=np.float64):
    def mv(xp, yp, _ud):
        x = np.ctypeslib.as_array(C.cast(xp, C.POINTER(C.c_double if dt == np.float64 else C.c_float)), shape=(n,))
        y = np.ctypeslib.as_array(C.cast(yp, C.POINTER(C.c_double if dt == np.float64 else C.c_float)), shape=(n,))
        y[:] = diag * x
        y[1:] += off * x[:-1]
        y[:-1] += off * x[1:]
    return _lib.MATVEC(mv)


def test_c_level_solver_rows():
    """interfaces/test/C/test_all_solvers.c rows for cg / minres / gmres / bicgstab (:131,134,141,148):
    20x20 tridiagonal, x_true = ones, ||x-1||/sqrt(n) <= 1e-6."""
    L = _lib.lib()
    n = 20
    mv = _tridiag_cb(n)
    b = np.zeros(n)
    ones = np.ones(n)
    mv(ones.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), None)
    null = _lib.MATVEC()
    for solver in (_lib.KRYLOV_CG, _lib.KRYLOV_MINRES, _lib.KRYLOV_GMRES, _lib.KRYLOV_BICGSTAB):
        ws = C.c_void_p()
        assert L.krylov_workspace_create(solver, n, n, 1, 0, None, C.byref(ws)) == 0
        o = L.krylov_default_options()
        o.atol = o.rtol = 1e-10
        assert L.krylov_solve(ws, mv, null, null, null, b.ctypes.data_as(C.c_void_p), None, None, C.byref(o)) == 0
        x = np.empty(n)
        assert L.krylov_get_x(ws, x.ctypes.data_as(C.c_void_p), n) == 0
        assert L.krylov_is_solved(ws) == 1 and L.krylov_niter(ws) > 0 and L.krylov_elapsed_time(ws) > 0
        assert np.linalg.norm(x - 1) / np.sqrt(n) <= 1e-6
        assert L.krylov_get_y(ws, x.ctypes.data_as(C.c_void_p), n) == -2
        assert L.krylov_warm_start2(ws, None, None, n, n) == -2
        assert L.krylov_workspace_free(ws) == 0 and L.krylov_workspace_free(ws) == 1


def test_device_buffers_mode():
    """KRYLOV_CUDA: b / x are device pointers (torch tensors give the memory), CSR operator attached."""
    torch = pytest.importorskip("torch")
    import krylov_b200 as kb
    from krylov_b200 import problems as P
    N = 24
    rp, ci, va = P.div_grad_csr(N, xp=torch, device="cuda")
    n = N ** 3
    b = torch.ones(n, dtype=torch.float64, device="cuda")
    ws = kb.CgWorkspace(n, n, np.float64, device="cuda")
    ws.solve((rp, ci, va), b, history=True)
    x = ws.x
    assert x.is_cuda and ws.stats.solved
    rph, cih, vah = P.div_grad_csr(N)
    import scipy.sparse as sp
    A = sp.csr_matrix((vah, cih, rph), shape=(n, n))
    assert np.linalg.norm(np.ones(n) - A @ x.cpu().numpy()) <= 1e-6 * np.sqrt(n)
    ws.free()
