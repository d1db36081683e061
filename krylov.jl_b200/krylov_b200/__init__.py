"""krylov_b200 -- host-side mirror of Krylov.jl's workspace/solver API for the
B200 path, over the C ABI of libkrylov_b200.so.

Names and argument meanings follow the reference (src/interface.jl:67-246,
src/krylov_workspaces.jl, src/workspace_accessors.jl:140-204):

    ws = CgWorkspace(A, b)            # or krylov_workspace("cg", A, b)
    cg_(ws, A, b; atol, rtol, ...)    # Julia's cg!(ws, A, b; ...)
    x, stats = cg(A, b, ...)          # out-of-place
    solution(ws), statistics(ws), issolved(ws), iteration_count(ws), warm_start_(ws, x0)

`A` is a scipy.sparse matrix (or anything scipy can turn into CSR) that is
uploaded once into HBM as int32/0-based CSR, or a Python callable
`A(x_host) -> y_host` (matrix-free, staged through pinned memory like the
reference's C callback operator).  `b`, `x0`, `c` are NumPy arrays (host) or
torch CUDA tensors (device; zero-copy).  All arithmetic runs in the CUDA
library; this module contains no numerical code.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

from . import _lib
from ._lib import (KRYLOV_CPU, KRYLOV_CUDA, KRYLOV_FLOAT32, KRYLOV_FLOAT64, SOLVER_IDS, KrylovB200Options,
                   KrylovB200Stats, KrylovOptions, KrylovWorkspaceOptions, lib)

__all__ = ["CgWorkspace", "GmresWorkspace", "BicgstabWorkspace", "MinresWorkspace", "KrylovWorkspace", "SimpleStats",
           "cg", "cg_", "gmres", "gmres_", "bicgstab", "bicgstab_", "minres", "minres_", "krylov_workspace",
           "krylov_solve", "krylov_solve_", "solution", "statistics", "results", "issolved", "iteration_count",
           "elapsed_time", "Aprod_count", "warm_start_", "device_count", "B200Error",
           "FomWorkspace", "FgmresWorkspace", "CgsWorkspace", "CgLanczosWorkspace", "fom", "fom_", "fgmres", "fgmres_",
           "cgs", "cgs_", "cg_lanczos", "cg_lanczos_", "CrWorkspace", "DiomWorkspace", "DqgmresWorkspace", "cr", "cr_", "diom",
           "diom_", "dqgmres", "dqgmres_", "BlockGmresWorkspace", "block_gmres", "block_gmres_", "CsrOperator"]


class B200Error(RuntimeError):
    """Raised where the reference raises ErrorException (status -1 from the C ABI)."""


@dataclass
class SimpleStats:
    """src/krylov_stats.jl:24-36"""
    niter: int = 0
    solved: bool = False
    inconsistent: bool = False
    indefinite: bool = False
    npcCount: int = 0
    residuals: list = field(default_factory=list)
    Aresiduals: list = field(default_factory=list)
    Acond: list = field(default_factory=list)
    allocation_timer: float = 0.0
    timer: float = 0.0
    status: str = "unknown"
    Anorm: float = math.nan          # LanczosStats (src/krylov_stats.jl), cg_lanczos! only


def device_count() -> int:
    return lib().krylov_b200_device_count()


def _dtype_id(dtype) -> int:
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return KRYLOV_FLOAT64
    if dtype == np.float32:
        return KRYLOV_FLOAT32
    raise B200Error(f"unsupported element type {dtype} (Float32/Float64 only on this path)")


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _ptr(a):
    """(pointer, keepalive) of a NumPy array or torch tensor; None -> NULL."""
    if a is None:
        return None, None
    if _is_torch(a):
        a = a.contiguous()
        return C.c_void_p(a.data_ptr()), a
    return a.ctypes.data_as(C.c_void_p), a


class CsrOperator:
    """A CSR operator resident in HBM, independent of any workspace:

        A = CsrOperator.read_mtx("bcsstk01.mtx")     # Matrix Market ingestion (benchmark/benchmarks.jl:23-33)
        At = A.transpose()                           # A^T = A^H for the real types of this path
        kb.cg(A, b)                                  # solvers accept it like a SciPy matrix
    """

    def __init__(self, csr_handle, ctx, dtype, owns_ctx=True):
        self._csr, self._ctx, self.dtype, self._owns_ctx = csr_handle, ctx, np.dtype(dtype), owns_ctx
        n, nnz = C.c_int(), C.c_longlong()
        lib().kb200_csr_info(self._csr, C.byref(n), C.byref(nnz))
        self.shape, self.nnz = (n.value, n.value), nnz.value

    @classmethod
    def read_mtx(cls, path, dtype=np.float64, device: int = -1):
        ctx = lib().kb200_ctx_create(device)
        if not ctx:
            raise B200Error(_lib.last_error())
        h = lib().kb200_csr_read_mtx(ctx, os.fsencode(path), _dtype_id(dtype))
        if not h:
            lib().kb200_ctx_destroy(ctx)
            raise B200Error(_lib.last_error())
        return cls(h, ctx, dtype)

    @classmethod
    def from_scipy(cls, A, dtype=None, device: int = -1):
        import scipy.sparse as sp
        M = sp.csr_matrix(A)
        M.sort_indices()
        dtype = np.dtype(dtype or M.dtype)
        ctx = lib().kb200_ctx_create(device)
        rp, ci, va = np.ascontiguousarray(M.indptr), np.ascontiguousarray(M.indices), np.ascontiguousarray(M.data, dtype=dtype)
        h = lib().kb200_csr_create(ctx, _dtype_id(dtype), M.shape[0], int(M.nnz), rp.ctypes.data_as(C.c_void_p),
                                   ci.astype(rp.dtype).ctypes.data_as(C.c_void_p), va.ctypes.data_as(C.c_void_p), 0, rp.dtype.itemsize, 0)
        if not h:
            lib().kb200_ctx_destroy(ctx)
            raise B200Error(_lib.last_error())
        return cls(h, ctx, dtype)

    def transpose(self):
        h = lib().kb200_csr_transpose(self._ctx, self._csr)
        if not h:
            raise B200Error(_lib.last_error())
        out = CsrOperator(h, self._ctx, self.dtype, owns_ctx=False)
        out._parent = self                       # shares (and keeps alive) the context
        return out

    T = property(transpose)

    def to_scipy(self):
        import scipy.sparse as sp
        n = self.shape[0]
        rp, ci, va = np.empty(n + 1, np.int32), np.empty(self.nnz, np.int32), np.empty(self.nnz, self.dtype)
        if lib().kb200_csr_download(self._ctx, self._csr, rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p),
                                    va.ctypes.data_as(C.c_void_p)) != 0:
            raise B200Error(_lib.last_error())
        return sp.csr_matrix((va, ci, rp), shape=self.shape)

    def matvec(self, x):
        """y = A x on the GPU (host arrays in and out)."""
        x = np.ascontiguousarray(x, dtype=self.dtype)
        n = self.shape[0]
        L = lib()
        dx, dy = L.kb200_alloc(x.nbytes), L.kb200_alloc(x.nbytes)
        try:
            L.kb200_h2d(dx, x.ctypes.data_as(C.c_void_p), x.nbytes)
            if L.kb200_spmv_csr(self._ctx, self._csr, dx, dy, 0) != 0 or L.kb200_sync(self._ctx) != 0:
                raise B200Error(_lib.last_error())
            y = np.empty(n, self.dtype)
            L.kb200_d2h(y.ctypes.data_as(C.c_void_p), dy, y.nbytes)
            return y
        finally:
            L.kb200_free(dx)
            L.kb200_free(dy)

    def free(self):
        if getattr(self, "_csr", None):
            lib().kb200_csr_destroy(self._csr)
            self._csr = None
            if self._owns_ctx and self._ctx:
                lib().kb200_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class KrylovWorkspace:
    """One workspace = all device vectors of one solver (krylov_workspaces.jl)."""

    solver = ""
    nA = 1  # operator products per iteration (workspace_accessors.jl:101-139)

    def __init__(self, m_or_A, n_or_b=None, dtype=None, *, memory: int = 0, window: int = 0, device: str = "host",
                 solver: Optional[str] = None):
        if solver:
            self.solver = solver
        A = None
        if hasattr(m_or_A, "shape") and not isinstance(m_or_A, (int, np.integer)):   # (A, b) constructor
            A, b = m_or_A, n_or_b
            m, n = A.shape
            if dtype is None:
                dtype = (b.cpu().numpy().dtype if _is_torch(b) else np.asarray(b).dtype) if b is not None else A.dtype
            if b is not None and _is_torch(b):
                device = "cuda"
        else:
            m, n = int(m_or_A), int(n_or_b)
            dtype = dtype or np.float64
        self.m, self.n = int(m), int(n)
        self.dtype = np.dtype(dtype)
        self.device = device
        self._keep = []
        self._cb = None
        self._h = C.c_void_p()
        w = KrylovWorkspaceOptions(memory, window)
        rc = lib().krylov_workspace_create(SOLVER_IDS[self.solver], self.m, self.n, _dtype_id(self.dtype),
                                           KRYLOV_CUDA if device == "cuda" else KRYLOV_CPU, C.byref(w), C.byref(self._h))
        if rc != 0:
            raise B200Error(f"krylov_workspace_create({self.solver}) -> {rc}: {_lib.last_error()}")
        self._ext = lib().krylov_b200_default_options()
        self._op_id = None
        if A is not None and not callable(A):
            self.set_operator(A)

    # -- lifetime -----------------------------------------------------------
    def free(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().krylov_workspace_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def _order_after(self, *arrays):
        """Stream contract of device inputs (include/krylov_b200.h): torch tensors are produced on torch's current
        stream, the library works on its own non-blocking stream -- make the latter wait for the former."""
        for a in arrays:
            if a is not None and _is_torch(a) and a.is_cuda:
                import torch
                s = torch.cuda.current_stream(a.device)
                if lib().krylov_b200_wait_stream(self._h, C.c_void_p(s.cuda_stream)) != 0:
                    raise B200Error(_lib.last_error())
                return

    # -- operator -----------------------------------------------------------
    def set_operator(self, A):
        """Upload A as the device-resident CSR operator (krylov_b200_set_operator_csr)."""
        if id(A) == self._op_id:
            return
        if isinstance(A, CsrOperator):  # a device-resident operator (Matrix Market file, transposed operator, ...)
            if A.shape != (self.m, self.n):
                raise B200Error(f"(workspace.m, workspace.n) = ({self.m}, {self.n}) is inconsistent with size(A) = {A.shape}")
            if lib().krylov_b200_attach_csr(self._h, A._csr) != 0:
                raise B200Error(_lib.last_error())
            self._keep.append(A)
            self._op_id = id(A)
            return
        if isinstance(A, tuple):        # (rowptr, colind, values[, index_base]) NumPy or torch
            rp, ci, va = A[:3]
            base = A[3] if len(A) > 3 else 0
            loc = 1 if _is_torch(va) else 0
            if loc:
                ib = rp.element_size()
                nnz = int(va.numel())
            else:
                rp = np.ascontiguousarray(rp)
                ci = np.ascontiguousarray(ci, dtype=rp.dtype)
                va = np.ascontiguousarray(va, dtype=self.dtype)
                ib = rp.dtype.itemsize
                nnz = int(va.shape[0])
            n = int(rp.shape[0]) - 1
        else:
            import scipy.sparse as sp
            M = sp.csr_matrix(A)
            M.sort_indices()
            if M.shape != (self.m, self.n):
                raise B200Error(f"(workspace.m, workspace.n) = ({self.m}, {self.n}) is inconsistent with size(A) = {M.shape}")
            rp = np.ascontiguousarray(M.indptr)
            ci = np.ascontiguousarray(M.indices, dtype=rp.dtype)
            va = np.ascontiguousarray(M.data, dtype=self.dtype)
            ib, base, loc, nnz, n = rp.dtype.itemsize, 0, 0, int(M.nnz), M.shape[0]
        p_rp, k1 = _ptr(rp)
        p_ci, k2 = _ptr(ci)
        p_va, k3 = _ptr(va)
        self._order_after(va)
        rc = lib().krylov_b200_set_operator_csr(self._h, n, nnz, p_rp, p_ci, p_va, int(base), int(ib), loc)
        if rc != 0:
            raise B200Error(_lib.last_error())
        self._op_id = id(A)

    def share_operator(self, other: "KrylovWorkspace"):
        if lib().krylov_b200_share_operator(self._h, other._h) != 0:
            raise B200Error(_lib.last_error())
        self._op_id = other._op_id

    def _set_diag(self, which: int, d):
        attached = getattr(self, "_precond_attached", None)
        if attached is None:
            attached = self._precond_attached = [False, False]
        if d is None:
            if not attached[which]:             # nothing to detach: no library call on the per-solve path
                return
            lib().krylov_b200_set_preconditioner_diag(self._h, which, None, 0)
            if not isinstance(self, BlockGmresWorkspace):
                lib().krylov_b200_set_preconditioner_blockdiag(self._h, which, 0, None, 0)
            attached[which] = False
            return
        attached[which] = True
        if getattr(d, "ndim", 1) == 3:      # block-Jacobi: (nblocks, bs, bs) dense diagonal blocks
            nb, bs, bs2 = d.shape
            if bs != bs2 or nb != (self.n + bs - 1) // bs:
                raise B200Error(f"block-diagonal preconditioner: expected ({(self.n + bs - 1) // bs}, {bs}, {bs}) blocks, got {tuple(d.shape)}")
            if not _is_torch(d):
                d = np.ascontiguousarray(d, dtype=self.dtype)
            p, keep = _ptr(d)
            self._order_after(d)
            lib().krylov_b200_set_preconditioner_diag(self._h, which, None, 0)
            if lib().krylov_b200_set_preconditioner_blockdiag(self._h, which, int(bs), p, 1 if _is_torch(d) else 0) != 0:
                raise B200Error(_lib.last_error())
            return
        if not isinstance(self, BlockGmresWorkspace):
            lib().krylov_b200_set_preconditioner_blockdiag(self._h, which, 0, None, 0)
        if not _is_torch(d):
            d = np.ascontiguousarray(d, dtype=self.dtype)
        p, keep = _ptr(d)
        self._order_after(d)
        if lib().krylov_b200_set_preconditioner_diag(self._h, which, p, 1 if _is_torch(d) else 0) != 0:
            raise B200Error(_lib.last_error())

    # -- solve --------------------------------------------------------------
    def _wrap_matvec(self, f: Optional[Callable]):
        if f is None:
            return _lib.MATVEC(), None
        n, dt = self.n, self.dtype
        if self.device == "cuda":
            raise B200Error("Python callables are host operators; create the workspace with device='host'")

        def tramp(xp, yp, _ud):
            x = np.ctypeslib.as_array(C.cast(xp, C.POINTER(C.c_byte)), shape=(n * dt.itemsize,)).view(dt)
            y = np.ctypeslib.as_array(C.cast(yp, C.POINTER(C.c_byte)), shape=(n * dt.itemsize,)).view(dt)
            y[:] = f(x)
        cb = _lib.MATVEC(tramp)
        return cb, cb

    def solve(self, A, b, *, c=None, M=None, N=None, atol=None, rtol=None, itmax=0, timemax=math.inf, verbose=0,
              history=False, callback=None, radius=0.0, linesearch=False, lambda_=0.0, etol=None, conlim=None,
              restart=False, reorthogonalization=False, ldiv=False, fused=True, batch=0, time_kernels=False,
              check_curvature=False, gamma=None):
        """solver!(ws, A, b; kwargs...)  -- kwargs as in cg.jl:100-111, gmres.jl:96-108,
        bicgstab.jl:105-116, minres.jl:138-151.  M / N: None (identity), a 1-D array
        (Diagonal preconditioner) or a host callable."""
        o = lib().krylov_default_options()
        if atol is not None:
            o.atol = float(atol)
        if rtol is not None:
            o.rtol = float(rtol)
        o.itmax, o.verbose = int(itmax), int(verbose)
        o.timemax = math.nan if math.isinf(timemax) else float(timemax)
        o.radius, o.linesearch, o.lambda_ = float(radius), int(linesearch), float(lambda_)
        o.restart, o.reorthogonalization = int(restart), int(reorthogonalization)
        e = lib().krylov_b200_default_options()
        e.history, e.ldiv, e.fused, e.batch = int(history), int(ldiv), int(fused), int(batch)
        e.time_kernels = int(time_kernels)
        e.check_curvature = int(check_curvature)
        if gamma is not None:
            e.cr_gamma = float(gamma)
        if etol is not None:
            e.etol = float(etol)
        if conlim is not None:
            e.conlim = float(conlim)
        keep = []
        if callback is not None:
            wsref = self

            def cb_tramp(_ws, _user):
                r = callback(wsref)
                if not isinstance(r, (bool, np.bool_)):
                    wsref._cb_error = TypeError(f"callback must return Bool, got {type(r).__name__}")   # cg.jl:264
                    return 1
                return int(r)
            e.callback = _lib.CALLBACK(cb_tramp)
            keep.append(e.callback)
        self._cb_error = None
        lib().krylov_b200_set_options(self._h, C.byref(e))

        fA = None
        if callable(A) and not hasattr(A, "shape"):
            fA, k = self._wrap_matvec(A)
            keep.append(k)
        elif A is not None:
            self.set_operator(A)
        fM = fN = None
        for which, P in ((0, M), (1, N)):
            if P is None:
                self._set_diag(which, None)
            elif callable(P) and not hasattr(P, "shape"):
                f, k = self._wrap_matvec(P)
                keep.append(k)
                if which == 0:
                    fM = f
                else:
                    fN = f
                self._set_diag(which, None)
            else:
                self._set_diag(which, P)
        if not _is_torch(b):
            b = np.ascontiguousarray(b, dtype=self.dtype)
            if self.device == "cuda":
                raise B200Error("ktypeof(b) must be a device vector for a device workspace")
        elif self.device != "cuda":
            raise B200Error("ktypeof(b) must be a host vector for a host workspace")
        if b.shape[0] != self.n:
            raise B200Error("Inconsistent problem size")
        pb, kb = _ptr(b)
        if c is not None and not _is_torch(c):
            c = np.ascontiguousarray(c, dtype=self.dtype)
        pc, kc = _ptr(c)
        self._order_after(kb, kc)
        null = _lib.MATVEC()
        rc = lib().krylov_solve(self._h, fA or null, null, fM or null, fN or null, pb, pc, None, C.byref(o))
        del keep
        if self._cb_error is not None:
            raise self._cb_error
        if rc != 0:
            raise B200Error(_lib.last_error())
        return self

    def warm_start(self, x0):
        if not _is_torch(x0):
            x0 = np.ascontiguousarray(x0, dtype=self.dtype)
        if x0.shape[0] != self.n:
            raise B200Error(f"x0 should have size {self.n}")
        p, k = _ptr(x0)
        self._order_after(k)
        rc = lib().krylov_warm_start(self._h, p, self.n)
        if rc != 0:
            raise B200Error(_lib.last_error())
        return self

    # -- accessors ------------------------------------------------------------
    @property
    def x(self):
        """solution(ws): a host copy (or a torch CUDA tensor for device workspaces)."""
        if self.device == "cuda":
            import torch
            out = torch.empty(self.n, dtype=torch.float64 if self.dtype == np.float64 else torch.float32, device="cuda")
            lib().krylov_get_x(self._h, C.c_void_p(out.data_ptr()), self.n)
            return out
        out = np.empty(self.n, dtype=self.dtype)
        if lib().krylov_get_x(self._h, out.ctypes.data_as(C.c_void_p), self.n) != 0:
            raise B200Error(_lib.last_error())
        return out

    def vector(self, name: str) -> np.ndarray:
        """Host copy of a workspace vector by its reference field name (x, r, p, Ap, npc_dir, ...)."""
        p = C.c_void_p()
        if lib().krylov_b200_get_vector(self._h, name.encode(), C.byref(p)) != 0 or not p.value:
            raise B200Error(f"workspace has no vector {name!r}")
        out = np.empty(self.n, dtype=self.dtype)
        lib().kb200_d2h(out.ctypes.data_as(C.c_void_p), p, out.nbytes)
        return out

    @property
    def stats(self) -> SimpleStats:
        s = KrylovB200Stats()
        if lib().krylov_b200_get_stats(self._h, C.byref(s)) != 0:
            raise B200Error(_lib.last_error())

        def hist(which, cnt):
            buf = (C.c_double * max(cnt, 1))()
            k = lib().krylov_b200_get_history(self._h, which, buf, cnt)
            return list(buf[:max(k, 0)])
        out = SimpleStats(s.niter, bool(s.solved), bool(s.inconsistent), bool(s.indefinite), s.npcCount,
                          hist(0, s.nresiduals), hist(1, s.nAresiduals), hist(2, s.nAcond), s.allocation_timer, s.timer,
                          s.status.decode("utf-8"))
        out.Anorm = s.Anorm          # LanczosStats.Anorm (cg_lanczos!), NaN otherwise
        return out

    @property
    def launches(self) -> int:
        return int(lib().krylov_b200_launch_count(self._h))

    @property
    def kernel_times(self):
        """(K1 ms, K2 ms, timed iterations) of the last solve run with time_kernels=True."""
        out = (C.c_double * 3)()
        lib().krylov_b200_get_kernel_times(self._h, out)
        return float(out[0]), float(out[1]), int(out[2])

    @property
    def npc_dir(self):
        return self.vector("npc_dir")


class CgWorkspace(KrylovWorkspace):
    solver = "cg"


class MinresWorkspace(KrylovWorkspace):
    solver = "minres"


class GmresWorkspace(KrylovWorkspace):
    solver = "gmres"


class BicgstabWorkspace(KrylovWorkspace):
    solver = "bicgstab"
    nA = 2


# sibling solvers on the same kernels
class FomWorkspace(KrylovWorkspace):
    solver = "fom"


class FgmresWorkspace(KrylovWorkspace):
    solver = "fgmres"


class CgsWorkspace(KrylovWorkspace):
    solver = "cgs"
    nA = 2


class CgLanczosWorkspace(KrylovWorkspace):
    solver = "cg_lanczos"


class CrWorkspace(KrylovWorkspace):
    solver = "cr"


class DiomWorkspace(KrylovWorkspace):
    solver = "diom"


class DqgmresWorkspace(KrylovWorkspace):
    solver = "dqgmres"


class BlockGmresWorkspace(KrylovWorkspace):
    """BlockGmresWorkspace(m, n, p, dtype; memory=5) (src/block_krylov_workspaces.jl:108-171): block_gmres! on
    n x p blocks of right-hand sides.  B, X0 and X are n x p arrays (any layout on the Python
    side; the C ABI exchanges the reference's column-major blocks)."""

    solver = "block_gmres"

    def __init__(self, m, n, p, dtype=np.float64, *, memory: int = 0, device: str = "host"):
        self.m, self.n, self.p = int(m), int(n), int(p)
        self.dtype = np.dtype(dtype)
        self.device = device
        self._keep = []
        self._cb = None
        self._h = C.c_void_p()
        w = KrylovWorkspaceOptions(memory, 0)
        rc = lib().krylov_block_workspace_create(0, self.m, self.n, self.p, _dtype_id(self.dtype),
                                                 KRYLOV_CUDA if device == "cuda" else KRYLOV_CPU, C.byref(w), C.byref(self._h))
        if rc != 0:
            raise B200Error(f"krylov_block_workspace_create -> {rc}: {_lib.last_error()}")
        self._ext = lib().krylov_b200_default_options()
        self._op_id = None

    def free(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().krylov_block_workspace_free(self._h)
            self._h = C.c_void_p()

    def _block_cb(self, f):
        if f is None:
            return _lib.BLOCK_MATVEC(), None
        n, p, dt = self.n, self.p, self.dtype
        if self.device == "cuda":
            raise B200Error("Python callables are host operators; create the workspace with device='host'")

        def tramp(xp, yp, pcols, _ud):
            X = np.ctypeslib.as_array(C.cast(xp, C.POINTER(C.c_byte)), shape=(n * pcols * dt.itemsize,)).view(dt)
            Y = np.ctypeslib.as_array(C.cast(yp, C.POINTER(C.c_byte)), shape=(n * pcols * dt.itemsize,)).view(dt)
            Y.reshape((pcols, n)).T[:] = f(X.reshape((pcols, n)).T)       # column-major n x p views
        cb = _lib.BLOCK_MATVEC(tramp)
        return cb, cb

    def _colmajor(self, B):
        if _is_torch(B):
            return B.t().contiguous()                                     # p x n row-major == n x p column-major
        return np.asfortranarray(B, dtype=self.dtype)

    def solve(self, A, B, *, M=None, N=None, atol=None, rtol=None, itmax=0, timemax=math.inf, verbose=0, history=False,
              callback=None, restart=False, reorthogonalization=False, ldiv=False):
        """block_gmres!(ws, A, B; kwargs...)  (src/block_gmres.jl:85-98)."""
        o = lib().krylov_default_options()
        if atol is not None:
            o.atol = float(atol)
        if rtol is not None:
            o.rtol = float(rtol)
        o.itmax, o.verbose = int(itmax), int(verbose)
        o.timemax = math.nan if math.isinf(timemax) else float(timemax)
        o.restart, o.reorthogonalization = int(restart), int(reorthogonalization)
        e = lib().krylov_b200_default_options()
        e.history, e.ldiv = int(history), int(ldiv)
        keep = []
        if callback is not None:
            wsref = self

            def cb_tramp(_ws, _user):
                r = callback(wsref)
                if not isinstance(r, (bool, np.bool_)):
                    wsref._cb_error = TypeError(f"callback must return Bool, got {type(r).__name__}")
                    return 1
                return int(r)
            e.callback = _lib.CALLBACK(cb_tramp)
            keep.append(e.callback)
        self._cb_error = None
        lib().krylov_b200_set_options(self._h, C.byref(e))
        fA = None
        if callable(A) and not hasattr(A, "shape"):
            fA, k = self._block_cb(A)
            keep.append(k)
        elif A is not None:
            self.set_operator(A)
        fM = fN = None
        for which, P in ((0, M), (1, N)):
            if P is None:
                self._set_diag(which, None)
            elif callable(P) and not hasattr(P, "shape"):
                f, k = self._block_cb(P)
                keep.append(k)
                if which == 0:
                    fM = f
                else:
                    fN = f
                self._set_diag(which, None)
            else:
                self._set_diag(which, P)
        if tuple(B.shape) != (self.n, self.p):
            raise B200Error("Inconsistent problem size")
        if _is_torch(B) != (self.device == "cuda"):
            raise B200Error("ktypeof(B) must match the workspace storage (host array / device tensor)")
        Bc = self._colmajor(B)
        pb, kb_ = _ptr(Bc)
        self._order_after(kb_)
        null = _lib.BLOCK_MATVEC()
        rc = lib().krylov_block_solve(self._h, fA or null, fM or null, fN or null, pb, None, C.byref(o))
        del keep
        if self._cb_error is not None:
            raise self._cb_error
        if rc != 0:
            raise B200Error(_lib.last_error())
        return self

    def warm_start(self, X0):
        if tuple(X0.shape) != (self.n, self.p):
            raise B200Error(f"X0 should have size {self.n} x {self.p}")
        Xc = self._colmajor(X0)
        p, k = _ptr(Xc)
        self._order_after(k)
        if lib().krylov_block_warm_start(self._h, p, self.n, self.p) != 0:
            raise B200Error(_lib.last_error())
        return self

    @property
    def x(self):
        """solution(ws): n x p (host array, or a torch CUDA tensor for device workspaces)."""
        if self.device == "cuda":
            import torch
            out = torch.empty((self.p, self.n), dtype=torch.float64 if self.dtype == np.float64 else torch.float32, device="cuda")
            if lib().krylov_block_get_X(self._h, C.c_void_p(out.data_ptr()), self.n, self.p) != 0:
                raise B200Error(_lib.last_error())
            return out.t()
        out = np.empty((self.n, self.p), dtype=self.dtype, order="F")
        if lib().krylov_block_get_X(self._h, out.ctypes.data_as(C.c_void_p), self.n, self.p) != 0:
            raise B200Error(_lib.last_error())
        return out

    X = x

    @property
    def qr_fallbacks(self) -> int:
        """Panel QR factorizations that took the slow Householder path (rank-deficient blocks); 0 normally."""
        return int(lib().krylov_b200_block_qr_fallbacks(self._h))


def block_gmres(A, B, X0=None, *, memory=0, **kw):
    """(X, stats) = block_gmres(A, B[, X0]; memory=5, kwargs...)  (src/block_gmres.jl:1-60)"""
    n, p = B.shape
    dt = B.cpu().numpy().dtype if _is_torch(B) else np.asarray(B).dtype
    if dt not in (np.float32, np.float64):
        dt = np.float64
    ws = BlockGmresWorkspace(n, n, p, dt, memory=memory, device="cuda" if _is_torch(B) else "host")
    try:
        if X0 is not None:
            ws.warm_start(X0)
        ws.solve(A, B if _is_torch(B) else np.asarray(B, dtype=dt), **kw)
        return ws.x, ws.stats
    finally:
        ws.free()


def block_gmres_(ws: BlockGmresWorkspace, A, B, X0=None, **kw):
    """block_gmres!(workspace, A, B[, X0]; kwargs...)"""
    if X0 is not None:
        ws.warm_start(X0)
    return ws.solve(A, B, **kw)


_WS = {"cg": CgWorkspace, "minres": MinresWorkspace, "gmres": GmresWorkspace, "bicgstab": BicgstabWorkspace,
       "fom": FomWorkspace, "fgmres": FgmresWorkspace, "cgs": CgsWorkspace, "cg_lanczos": CgLanczosWorkspace,
       "cr": CrWorkspace, "diom": DiomWorkspace, "dqgmres": DqgmresWorkspace}


def krylov_workspace(method: str, *args, **kw) -> KrylovWorkspace:
    """krylov_workspace(Val(method), ...)  (src/interface.jl:248-348)"""
    if method not in _WS:
        raise B200Error(f"method {method!r} is outside the B200 path ({', '.join(sorted(_WS))})")
    return _WS[method](*args, **kw)


def krylov_solve_(ws: KrylovWorkspace, A, b, x0=None, **kw) -> KrylovWorkspace:
    """krylov_solve!(ws, A, b[, x0]; kw...)"""
    if x0 is not None:
        ws.warm_start(x0)
    return ws.solve(A, b, **kw)


def _make_inplace(name):
    def f(ws, A, b, x0=None, **kw):
        if ws.solver != name:
            raise B200Error(f"{name}! needs a {_WS[name].__name__}")
        return krylov_solve_(ws, A, b, x0, **kw)
    f.__name__ = name + "_"
    f.__doc__ = f"{name}!(workspace, A, b[, x0]; kwargs...)"
    return f


def _make_outofplace(name):
    def f(A, b, x0=None, *, memory=0, window=0, **kw):
        n = b.shape[0]
        dt = b.cpu().numpy().dtype if _is_torch(b) else np.asarray(b).dtype
        if dt not in (np.float32, np.float64):
            dt = np.float64
        ws = _WS[name](n, n, dt, memory=memory, window=window, device="cuda" if _is_torch(b) else "host")
        try:
            krylov_solve_(ws, A, b, x0, **kw)
            return ws.x, ws.stats
        finally:
            ws.free()
    f.__name__ = name
    f.__doc__ = f"(x, stats) = {name}(A, b[, x0]; kwargs...)"
    return f


cg_, gmres_, bicgstab_, minres_ = (_make_inplace(s) for s in ("cg", "gmres", "bicgstab", "minres"))
cg, gmres, bicgstab, minres = (_make_outofplace(s) for s in ("cg", "gmres", "bicgstab", "minres"))
fom_, fgmres_, cgs_, cg_lanczos_ = (_make_inplace(s) for s in ("fom", "fgmres", "cgs", "cg_lanczos"))
fom, fgmres, cgs, cg_lanczos = (_make_outofplace(s) for s in ("fom", "fgmres", "cgs", "cg_lanczos"))
cr_, diom_, dqgmres_ = (_make_inplace(s) for s in ("cr", "diom", "dqgmres"))
cr, diom, dqgmres = (_make_outofplace(s) for s in ("cr", "diom", "dqgmres"))


def krylov_solve(method: str, A, b, x0=None, **kw):
    return {"cg": cg, "gmres": gmres, "bicgstab": bicgstab, "minres": minres, "fom": fom, "fgmres": fgmres, "cgs": cgs,
            "cg_lanczos": cg_lanczos, "cr": cr, "diom": diom, "dqgmres": dqgmres}[method](A, b, x0, **kw)


# workspace_accessors.jl:140-152
def solution(ws): return ws.x
def statistics(ws): return ws.stats
def results(ws): return (ws.x, ws.stats)
def issolved(ws): return bool(lib().krylov_is_solved(ws._h) == 1)
def iteration_count(ws): return int(lib().krylov_niter(ws._h))
def elapsed_time(ws): return float(lib().krylov_elapsed_time(ws._h))
def Aprod_count(ws): return ws.nA * iteration_count(ws)
def warm_start_(ws, x0): return ws.warm_start(x0)
