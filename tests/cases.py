"""Named parity cases shared by the golden generator, the oracle regression test and the GPU parity tests."""
import numpy as np
import scipy.sparse as sp

from krylov_b200 import problems as P


def _mat(csr):
    rp, ci, va = csr
    n = len(rp) - 1
    return sp.csr_matrix((va, ci, rp), shape=(n, n))


def build(name):
    """-> (solver, A (scipy csr), b, kwargs, dtype)"""
    f64, f32 = np.float64, np.float32
    if name == "cg_divgrad16_default":
        A = _mat(P.div_grad_csr(16)); return "cg", A, np.ones(A.shape[0]), {}, f64
    if name == "cg_divgrad32_bench":            # BASELINE config 1: benchmark/benchmarks.jl:14-21
        A = _mat(P.div_grad_csr(32)); return "cg", A, np.ones(A.shape[0]), dict(atol=0.0, rtol=1e-8, itmax=A.shape[0]), f64
    if name == "cg_divgrad12_f32":
        A = _mat(P.div_grad_csr(12, dtype=f32)); return "cg", A, np.ones(A.shape[0], f32), {}, f32
    if name == "cg_ragged_7x5x3":
        A = _mat(P.div_grad_csr(7, 5, 3)); return "cg", A, np.arange(1.0, A.shape[0] + 1), dict(rtol=1e-10, atol=0.0), f64
    if name == "gmres_kron10_restart30":        # BASELINE config 3 family
        A = _mat(P.kron_unsymmetric_csr(10)); return "gmres", A, A @ np.ones(A.shape[0]), dict(memory=30, restart=True), f64
    if name == "gmres_divgrad16_mem10_restart":  # test/test_gmres.jl:93-101
        A = _mat(P.div_grad_csr(16)); return "gmres", A, np.ones(A.shape[0]), dict(memory=10, restart=True), f64
    if name == "gmres_divgrad16_mem10_norestart":
        A = _mat(P.div_grad_csr(16)); return "gmres", A, np.ones(A.shape[0]), dict(memory=10, restart=False), f64
    if name == "gmres_kron8_reorth":
        A = _mat(P.kron_unsymmetric_csr(8)); return "gmres", A, A @ np.ones(A.shape[0]), dict(memory=20, reorthogonalization=True), f64
    if name == "bicgstab_kron10":
        A = _mat(P.kron_unsymmetric_csr(10)); return "bicgstab", A, A @ np.ones(A.shape[0]), {}, f64
    if name == "bicgstab_random3000_f32":        # BASELINE config 4 family
        A = _mat(P.random_csr(3000, 20, seed=1234, dtype=f32)); return "bicgstab", A, (A @ np.ones(3000, f32)).astype(f32), {}, f32
    if name == "minres_divgrad16":
        A = _mat(P.div_grad_csr(16)); return "minres", A, np.ones(A.shape[0]), {}, f64
    if name == "minres_shift":
        A = _mat(P.div_grad_csr(10)); return "minres", A, np.ones(A.shape[0]), dict(lambda_=0.5, atol=1e-10, rtol=1e-10), f64
    if name == "minres_indefinite":              # symmetric_indefinite(40), test/test_utils.jl:26-31
        n = 40
        A = sp.csr_matrix(sp.diags([np.ones(n - 1), np.ones(n), np.ones(n - 1)], [-1, 0, 1]))
        return "minres", A, A @ np.arange(1.0, n + 1), {}, f64
    # sibling solvers
    if name == "cgs_kron10":
        A = _mat(P.kron_unsymmetric_csr(10)); return "cgs", A, A @ np.ones(A.shape[0]), {}, f64
    if name == "fom_kron10_mem30_restart":
        A = _mat(P.kron_unsymmetric_csr(10)); return "fom", A, A @ np.ones(A.shape[0]), dict(memory=30, restart=True), f64
    if name == "fgmres_kron10_mem30_restart":
        A = _mat(P.kron_unsymmetric_csr(10)); return "fgmres", A, A @ np.ones(A.shape[0]), dict(memory=30, restart=True), f64
    if name == "cg_lanczos_divgrad16":
        A = _mat(P.div_grad_csr(16)); return "cg_lanczos", A, np.ones(A.shape[0]), {}, f64
    if name == "dqgmres_kron10_mem8":
        A = _mat(P.kron_unsymmetric_csr(10)); return "dqgmres", A, A @ np.ones(A.shape[0]), dict(memory=8), f64
    if name == "diom_kron10_mem8":
        A = _mat(P.kron_unsymmetric_csr(10)); return "diom", A, A @ np.ones(A.shape[0]), dict(memory=8), f64
    if name == "cr_divgrad16":
        A = _mat(P.div_grad_csr(16)); return "cr", A, np.ones(A.shape[0]), {}, f64
    raise KeyError(name)


NAMES = ["cg_divgrad16_default", "cg_divgrad32_bench", "cg_divgrad12_f32", "cg_ragged_7x5x3",
         "gmres_kron10_restart30", "gmres_divgrad16_mem10_restart", "gmres_divgrad16_mem10_norestart",
         "gmres_kron8_reorth", "bicgstab_kron10", "bicgstab_random3000_f32", "minres_divgrad16", "minres_shift",
         "minres_indefinite", "cgs_kron10", "fom_kron10_mem30_restart", "fgmres_kron10_mem30_restart",
         "cg_lanczos_divgrad16", "dqgmres_kron10_mem8", "diom_kron10_mem8", "cr_divgrad16"]


def run_oracle(O, name):
    solver, A, b, kw, dt = build(name)
    x, st = getattr(O, solver)(A, b, dtype=dt, **kw)
    return x, st
