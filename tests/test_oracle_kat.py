"""Pins the CPU oracle against the reference's own known-answer tests.

Every case cites the reference test it restates (paths relative to Krylov.jl v0.10.8).
"""
import math

import numpy as np
import pytest
import scipy.sparse as sp


def resid(A, x, b):
    return np.linalg.norm(b - A @ x) / np.linalg.norm(b)


# ---- test/test_aux.jl:5-21 -------------------------------------------------
def test_sym_givens_corner_cases(O):
    assert O.sym_givens(0.0, 0.0) == (1.0, 0.0, 0.0)
    a = 3.14
    assert O.sym_givens(a, 0.0) == (1.0, 0.0, a)
    assert O.sym_givens(-a, 0.0) == (-1.0, 0.0, a)
    assert O.sym_givens(0.0, a) == (0.0, 1.0, a)
    assert O.sym_givens(0.0, -a) == (0.0, -1.0, a)
    c, s, r = O.sym_givens(3.0, 4.0)
    assert abs(c * 3.0 + s * 4.0 - r) < 1e-15 and abs(s * 3.0 - c * 4.0) < 1e-15


# ---- test/test_aux.jl:38-77 ------------------------------------------------
def test_roots_quadratic(O):
    assert O.roots_quadratic(0.0, 0.0, 0.0) == (0.0, 0.0)
    with pytest.raises(ArithmeticError):
        O.roots_quadratic(0.0, 0.0, 1.0)
    assert O.roots_quadratic(0.0, 3.14, -1.0) == (1.0 / 3.14, 1.0 / 3.14)
    with pytest.raises(ArithmeticError):
        O.roots_quadratic(1.0, 0.0, 1.0)
    assert O.roots_quadratic(1.0, 0.0, 0.0) == (0.0, 0.0)
    r = O.roots_quadratic(1.0, 3.0, 2.0)
    assert r[0] == pytest.approx(-2.0) and r[1] == pytest.approx(-1.0)
    with pytest.raises(ArithmeticError):
        O.roots_quadratic(1.0e8, 1.0, 1.0)
    assert O.roots_quadratic(-1.0e-8, 1.0e5, 1.0, nitref=0) == (1.0e13, 0.0)
    assert O.roots_quadratic(-1.0e-8, 1.0e5, 1.0, nitref=1) == (1.0e13, -1.0e-05)
    for nitref in (0, 1):
        r = O.roots_quadratic(-1.0e-7, 1.0, 1.0, nitref=nitref)
        assert r[0] == pytest.approx(1.0e7, rel=1e-6) and r[1] == pytest.approx(-1.0, rel=1e-6)


# ---- test/test_aux.jl:103-117 ----------------------------------------------
def test_to_boundary(O):
    n = 5
    x = np.ones(n)
    d = np.ones(n)
    d[0::2] = -1
    for bad in (-1.0, 0.5):
        with pytest.raises(ArithmeticError):
            O.to_boundary(x, d, bad)
    with pytest.raises(ArithmeticError):
        O.to_boundary(x, np.zeros(n), 1.0)
    assert max(O.to_boundary(x, d, 5.0)) == pytest.approx(2.209975124224178)
    assert min(O.to_boundary(x, d, 5.0)) == pytest.approx(-1.8099751242241782)
    assert max(O.to_boundary(x, d, 5.0, flip=True)) == pytest.approx(1.8099751242241782)
    assert min(O.to_boundary(x, d, 5.0, flip=True)) == pytest.approx(-2.209975124224178)


# ---- interfaces/examples/C/basic_cg.c:13-15 ---------------------------------
def test_basic_cg_example(O):
    A = sp.diags([-np.ones(4), 2 * np.ones(5), -np.ones(4)], [-1, 0, 1], format="csr")
    x, st = O.cg(A, np.array([1.0, 0, 0, 0, 1.0]))
    assert st["solved"] and st["niter"] == 3
    assert np.allclose(x, 1.0, atol=1e-12)


# ---- test/test_cg.jl ---------------------------------------------------------
def test_cg_reference_cases(O):
    tol = 1e-6
    A, b = O.symmetric_definite()                                    # :8-13
    x, st = O.cg(A, b, itmax=10)
    assert resid(A, x, b) <= tol and st["solved"]
    radius = 0.75 * np.linalg.norm(x)                                 # :15-20
    x, st = O.cg(A, b, radius=radius, itmax=10)
    assert st["solved"] and abs(radius - np.linalg.norm(x)) <= tol * radius
    assert st["status"] == "on trust-region boundary"
    A, b = O.sparse_laplacian()                                       # :22-28
    x, st = O.cg(A, b)
    assert resid(A, x, b) <= tol and st["solved"]
    radius = 0.75 * np.linalg.norm(x)                                 # :30-35
    x, st = O.cg(A, b, radius=radius, itmax=10)
    assert st["solved"] and abs(radius - np.linalg.norm(x)) <= tol * radius
    A, b = O.zero_rhs()                                               # :37-41
    x, st = O.cg(A, b)
    assert np.linalg.norm(x) == 0 and st["status"] == "x is a zero-residual solution"
    A, b, M = O.square_preconditioned()                               # :43-49
    x, st = O.cg(A, b, M=M)
    r = b - A @ x
    assert math.sqrt(r @ (M * r)) / math.sqrt(b @ (M * b)) <= tol and st["solved"]
    A, b = O.symmetric_indefinite(shift=10)                           # :51-62
    x, st = O.cg(A, b, linesearch=True)
    assert st["status"] == "nonpositive curvature" and not st["inconsistent"] and st["niter"] == 0
    assert st["indefinite"] and st["npcCount"] == 1
    assert st["npc_dir"] @ (A @ st["npc_dir"]) <= 0 and np.all(st["npc_dir"] == b)
    A, b = O.zero_rhs()                                               # :64-71, :73-80
    for kw in (dict(linesearch=True), dict(radius=10.0)):
        x, st = O.cg(A, b, **kw)
        assert st["status"] == "x is a zero-residual solution" and np.linalg.norm(x) == 0 and st["niter"] == 0
    A4 = sp.csr_matrix(np.diag([10.0, 8.0, 5.0, -1.0]))               # :82-96
    b4 = np.array([1.0, 1.0, 1.0, 0.1])
    x, st = O.cg(A4, b4, radius=10.0)
    assert st["npcCount"] == 1 and st["status"] == "nonpositive curvature" and st["indefinite"]
    assert st["npc_dir"] @ (A4 @ st["npc_dir"]) <= 0.01
    A, b = O.singular_consistent()                                    # :98-103
    x, st = O.cg(A, b)
    assert resid(A, x, b) <= tol and not st["inconsistent"]
    A, b = O.square_inconsistent()                                    # :105-109
    x, st = O.cg(A, b)
    assert st["inconsistent"]
    A, b = O.cartesian_poisson()                                      # :111-117 (negative definite: alpha < 0 throughout, same iterates)
    x, st = O.cg(A, b)
    assert resid(A, x, b) <= tol and st["solved"]
    A, b = O.symmetric_indefinite(shift=5)                            # :132-134
    x, st = O.cg(A, b, radius=1.0, linesearch=True)
    assert st["error"] == 1
    x, st = O.cg(A4, b4, linesearch=True)                             # :136-170
    assert st["npcCount"] == 1 and st["indefinite"] and st["status"] == "nonpositive curvature"
    x, st = O.cg(sp.csr_matrix(np.diag([10.0, 8.0, 5.0, 1.0])), np.ones(4), linesearch=True)
    assert st["npcCount"] == 0 and not st["indefinite"] and st["solved"]


# ---- test/test_gmres.jl (thresholds 1e-6; restart :93-129) -------------------
def test_gmres_reference_cases(O):
    tol = 1e-6
    A, b = O.symmetric_definite()
    x, st = O.gmres(A, b)
    assert resid(A, x, b) <= tol and st["solved"]
    A, b = O.symmetric_indefinite()
    x, st = O.gmres(A, b)
    assert resid(A, x, b) <= tol and st["solved"]
    A, b = O.sparse_laplacian()
    d = A.diagonal()
    for restart in (False, True):
        x, st = O.gmres(A, b, restart=restart, memory=10)
        assert resid(A, x, b) <= tol and st["niter"] > 10 and st["solved"]
        M = 1.0 / d
        x, st = O.gmres(A, b, M=M, restart=restart, memory=10)
        r = b - A @ x
        assert np.linalg.norm(M * r) / np.linalg.norm(M * b) <= tol and st["niter"] > 10 and st["solved"]
        x, st = O.gmres(A, b, N=M, restart=restart, memory=10)
        assert resid(A, x, b) <= tol and st["niter"] > 10 and st["solved"]
        Ns = 1.0 / np.sqrt(d)
        x, st = O.gmres(A, b, M=M, N=Ns, restart=restart, memory=10)
        r = b - A @ x
        assert np.linalg.norm(M * r) / np.linalg.norm(M * b) <= tol and st["niter"] > 10 and st["solved"]
    A, b = O.zero_rhs()
    x, st = O.gmres(A, b)
    assert np.linalg.norm(x) == 0 and st["status"] == "x is a zero-residual solution"
    A, b = O.kron_unsymmetric(8)
    x, st = O.gmres(A, b)
    assert resid(A, x, b) <= tol and st["solved"]


# ---- test/test_bicgstab.jl ----------------------------------------------------
def test_bicgstab_reference_cases(O):
    tol = 1e-6
    for gen in (O.symmetric_definite, O.sparse_laplacian):
        A, b = gen()
        x, st = O.bicgstab(A, b)
        assert resid(A, x, b) <= tol and st["solved"]
    A, b = O.kron_unsymmetric(8)
    x, st = O.bicgstab(A, b)
    assert resid(A, x, b) <= tol and st["solved"]
    A, b = O.zero_rhs()
    x, st = O.bicgstab(A, b)
    assert np.linalg.norm(x) == 0 and st["status"] == "x is a zero-residual solution"
    A = sp.csr_matrix(np.array([[1.0, 2.0], [3.0, 4.0]]))             # bc_breakdown, test_bicgstab.jl:85-88
    x, st = O.bicgstab(A, np.array([0.0, 1.0]), c=np.array([1.0, 0.0]))
    assert st["status"] == "Breakdown bᴴc = 0" and not st["solved"] and st["niter"] == 0
    A, b = O.sparse_laplacian()
    M = 1.0 / A.diagonal()
    x, st = O.bicgstab(A, b, M=M)
    assert resid(A, x, b) <= tol and st["solved"]
    x, st = O.bicgstab(A, b, N=M)
    assert resid(A, x, b) <= tol and st["solved"]


# ---- test/test_minres.jl -------------------------------------------------------
def test_minres_reference_cases(O):
    tol = 1e-5
    A, b = O.symmetric_definite()
    x, st = O.minres(A, b)
    assert resid(A, x, b) <= tol and st["solved"]
    A, b = O.symmetric_indefinite()
    x, st = O.minres(A, b)
    assert resid(A, x, b) <= tol and st["solved"]
    A, b = O.sparse_laplacian()
    x, st = O.minres(A, b)
    assert resid(A, x, b) <= tol and st["solved"]
    lam = 0.5                                                          # shifted system
    x, st = O.minres(A, b, lambda_=lam, atol=1e-10, rtol=1e-10)
    assert np.linalg.norm(b - (A @ x + lam * x)) / np.linalg.norm(b) <= tol
    A, b = O.zero_rhs()                                                # test_minres.jl:45-52: niter == 1 quirk
    x, st = O.minres(A, b)
    assert np.linalg.norm(x) == 0 and st["status"] == "x is a zero-residual solution" and st["niter"] == 1
    A, b = O.symmetric_indefinite(shift=5)                             # linesearch: npc at first iteration
    x, st = O.minres(A, b, linesearch=True)
    assert st["status"] == "nonpositive curvature" and st["indefinite"] and st["npcCount"] >= 1
    A, b = O.sparse_laplacian()
    x, st = O.minres(A, b, M=1.0 / A.diagonal())
    assert resid(A, x, b) <= tol and st["solved"]


# ---- expected magnitudes of the benchmark family --------
def test_cg_iteration_counts_div_grad(O):
    for N, default, tight in ((16, 38, 39), (32, 78, 79)):
        A, b = O.sparse_laplacian(N)
        assert A.nnz == 7 * N ** 3 - 6 * N ** 2
        assert O.cg(A, b)[1]["niter"] == default
        assert O.cg(A, b, atol=0.0, rtol=1e-8)[1]["niter"] == tight


def test_float32_instantiation(O):
    A, b = O.sparse_laplacian(8)
    x, st = O.cg(A, b, dtype=np.float32)
    assert st["solved"] and x.dtype == np.float32
    assert resid(A, x.astype(np.float64), b) <= 1e-3
    x, st = O.bicgstab(A, b, dtype=np.float32)
    assert st["solved"]


# ---- SURVEY.md 8(f)-3 sibling solvers (oracle/krylov_oracle_siblings.h) ---------------------------------------
# test/test_cgs.jl
def test_cgs_reference_cases(O):
    tol = 1e-6
    for gen in (O.symmetric_definite, O.symmetric_indefinite, O.sparse_laplacian):
        A, b = gen()
        x, st = O.cgs(A, b)
        assert resid(A, x, b) <= tol and st["solved"]
    A, b = O.kron_unsymmetric(8)
    x, st = O.cgs(A, b)
    assert resid(A, x, b) <= tol and st["solved"]
    A, b = O.zero_rhs()
    x, st = O.cgs(A, b)
    assert np.linalg.norm(x) == 0 and st["status"] == "x is a zero-residual solution"
    A, b, M = O.square_preconditioned()
    x, st = O.cgs(A, b, M=M)
    assert resid(A, x, b) <= tol and st["solved"]
    x, st = O.cgs(A, b, N=M)
    assert resid(A, x, b) <= tol and st["solved"]
    A = sp.csr_matrix(np.array([[1.0, 2.0], [3.0, 4.0]]))             # bc_breakdown, test_cgs.jl
    x, st = O.cgs(A, np.array([0.0, 1.0]), c=np.array([1.0, 0.0]))
    assert st["status"] == "Breakdown bᴴc = 0" and not st["solved"] and st["niter"] == 0


# test/test_cg_lanczos.jl
def test_cg_lanczos_reference_cases(O):
    tol = 1e-6
    n = 10
    A, b = O.symmetric_definite(n)
    x, st = O.cg_lanczos(A, b, itmax=n)
    assert resid(A, x, b) <= tol and st["solved"] and st["Anorm"] > 0
    A = sp.lil_matrix(A)
    A[n - 2, n - 2] = -4.0                                             # test_cg_lanczos.jl: negative curvature detection
    x, st = O.cg_lanczos(sp.csr_matrix(A), b, check_curvature=True)
    assert st["status"] == "negative curvature" and st["indefinite"]
    A, b = O.zero_rhs()
    x, st = O.cg_lanczos(A, b)
    assert np.linalg.norm(x) == 0 and st["status"] == "x is a zero-residual solution"
    A, b, M = O.square_preconditioned()
    x, st = O.cg_lanczos(A, b, M=M)
    assert resid(A, x, b) <= tol and st["solved"]
    A, b = O.sparse_laplacian()                                        # same Krylov iterates as cg! (Frommer & Maass)
    x, st = O.cg_lanczos(A, b)
    xc, sc = O.cg(A, b)
    assert st["niter"] == sc["niter"] and np.allclose(st["residuals"], sc["residuals"], rtol=1e-8)


# test/test_fom.jl, test/test_fgmres.jl
@pytest.mark.parametrize("name", ["fom", "fgmres"])
def test_fom_fgmres_reference_cases(O, name):
    f = getattr(O, name)
    tol = 1e-6
    for gen in (O.symmetric_definite, O.symmetric_indefinite):
        A, b = gen()
        x, st = f(A, b)
        assert resid(A, x, b) <= tol and st["solved"]
    A, b = O.kron_unsymmetric(8)
    x, st = f(A, b)
    assert resid(A, x, b) <= tol and st["solved"]
    A, b = O.almost_singular()
    x, st = f(A, b)
    assert resid(A, x, b) <= 100 * tol and st["solved"]
    A, b = O.square_inconsistent()
    x, st = f(A, b)
    assert st["inconsistent"]
    A, b = O.zero_rhs()
    x, st = f(A, b)
    assert np.linalg.norm(x) == 0 and st["status"] == "x is a zero-residual solution"
    A, b, M = O.square_preconditioned()
    x, st = f(A, b, M=M)
    assert np.linalg.norm(M * (b - A @ x)) / np.linalg.norm(M * b) <= tol and st["solved"]
    x, st = f(A, b, N=M)
    assert resid(A, x, b) <= tol and st["solved"]
    A, b = O.sparse_laplacian()
    d = 1.0 / A.diagonal()
    for restart in (False, True):                                      # restart block of both test files
        memory = 10
        x, st = f(A, b, restart=restart, memory=memory)
        assert resid(A, x, b) <= tol and st["niter"] > memory and st["solved"]
        x, st = f(A, b, M=d, restart=restart, memory=memory)
        assert np.linalg.norm(d * (b - A @ x)) / np.linalg.norm(d * b) <= tol and st["niter"] > memory and st["solved"]
        x, st = f(A, b, N=d, restart=restart, memory=memory)
        assert resid(A, x, b) <= tol and st["niter"] > memory and st["solved"]
        x, st = f(A, b, M=d, N=1.0 / np.sqrt(A.diagonal()), restart=restart, memory=memory)
        assert np.linalg.norm(d * (b - A @ x)) / np.linalg.norm(d * b) <= tol and st["niter"] > memory and st["solved"]
    A, b = O.cartesian_poisson(12, 12)
    x, st = f(A, b, reorthogonalization=True)
    assert resid(A, x, b) <= tol and st["solved"]


def test_fgmres_with_identity_N_is_gmres(O):
    """fgmres! with N = I runs gmres!'s arithmetic exactly (Z[k] is a copy of V[k])."""
    A, b = O.kron_unsymmetric(8)
    for kw in (dict(memory=20), dict(memory=10, restart=True)):
        xg, sg = O.gmres(A, b, **kw)
        xf, sf = O.fgmres(A, b, **kw)
        assert sg["niter"] == sf["niter"] and np.array_equal(sg["residuals"], sf["residuals"]) and np.array_equal(xg, xf)


# ---- SURVEY.md 8(f)-2 block_gmres (oracle/krylov_oracle_block.h) -----------------------------------------------
def _test_block_problem(n=20, p=3):
    """interfaces/test/C/test_block.c:36-87: A = tridiag(-1, 8, -1), X_true columns 1, t, t^2, B = A X_true."""
    A = sp.diags([-np.ones(n - 1), 8.0 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1], format="csr")
    t = np.arange(1, n + 1) / n
    Xt = np.stack([np.ones(n), t, t * t][:p], axis=1)
    return A, Xt, A @ Xt


def test_householder_matches_lapack(O):
    """householder! (src/block_krylov_utils.jl:201-208) calls LAPACK geqrf/orgqr; numpy.linalg.qr is the same LAPACK
    pair, so the restated dgeqr2/dorg2r must agree to rounding, signs included."""
    rng = np.random.default_rng(0)
    for m, k in ((50, 6), (7, 7), (40, 1), (64, 16)):
        Am = rng.standard_normal((m, k))
        Q, R, tau = O.householder(Am)
        Qn, Rn = np.linalg.qr(Am)
        assert np.abs(Q - Qn).max() <= 1e-13 and np.abs(R - Rn).max() <= 1e-13
        assert np.abs(Q.T @ Q - np.eye(k)).max() <= 1e-14


def test_block_gmres_reference_c_cases(O):
    """interfaces/test/C/test_block.c: solver, preconditioner (left / right Jacobi), warm start, memory=4."""
    A, Xt, B = _test_block_problem()
    kw = dict(atol=1e-10, rtol=1e-10, itmax=200)
    X, st = O.block_gmres(A, B, **kw)
    assert st["solved"] and st["niter"] > 0 and np.abs(X - Xt).max() < 1e-6
    niter_cold = st["niter"]
    d = 1.0 / A.diagonal()
    for pre in (dict(M=d), dict(N=d)):
        X, st = O.block_gmres(A, B, **pre, **kw)
        assert st["solved"] and np.abs(X - Xt).max() < 1e-6
    X, st = O.block_gmres(A, B, X0=Xt, **kw)
    assert st["solved"] and st["niter"] < niter_cold
    X, st = O.block_gmres(A, B, memory=4, **kw)
    assert st["solved"] and np.abs(X - Xt).max() < 1e-6


def test_block_gmres_with_one_column_is_gmres(O):
    """p = 1: the block method is GMRES (Householder QR of one column is a normalisation up to sign)."""
    A, b = O.kron_unsymmetric(6)
    X, st = O.block_gmres(A, b[:, None], memory=20)
    x, sg = O.gmres(A, b, memory=20)
    assert st["niter"] == sg["niter"] and np.allclose(st["residuals"], sg["residuals"], rtol=1e-9)
    assert np.allclose(X[:, 0], x, rtol=1e-9, atol=1e-12)
    for kw in (dict(restart=True, memory=5), dict(reorthogonalization=True)):
        X, st = O.block_gmres(A, np.stack([b, b[::-1].copy(), np.cos(np.arange(len(b)))], axis=1), **kw)
        assert st["solved"]


# ---- remaining siblings of SURVEY.md 8(f)-3: dqgmres!, diom!, cr! --------------------------------------------
# test/test_dqgmres.jl, test/test_diom.jl
@pytest.mark.parametrize("name", ["dqgmres", "diom"])
def test_dqgmres_diom_reference_cases(O, name):
    f = getattr(O, name)
    tol = 1e-6
    for gen in (O.symmetric_definite, O.symmetric_indefinite, O.sparse_laplacian):
        A, b = gen()
        x, st = f(A, b)
        assert resid(A, x, b) <= tol and st["solved"]
    A, b = O.kron_unsymmetric(8)
    x, st = f(A, b)
    assert resid(A, x, b) <= tol and st["solved"]
    A, b = O.zero_rhs()
    x, st = f(A, b)
    assert np.linalg.norm(x) == 0 and st["status"] == "x is a zero-residual solution"
    A, b, M = O.square_preconditioned()
    x, st = f(A, b, M=M)
    assert resid(A, x, b) <= tol and st["solved"]
    x, st = f(A, b, N=M)
    assert resid(A, x, b) <= tol and st["solved"]
    A, b = O.cartesian_poisson(12, 12)
    x, st = f(A, b, memory=100, reorthogonalization=True)
    assert resid(A, x, b) <= tol and st["solved"]


def test_truncated_methods_with_full_memory_are_the_full_methods(O):
    """DQGMRES / DIOM with memory >= niter run the arithmetic of GMRES / FOM (same Arnoldi, same rotations / LU)."""
    A, b = O.kron_unsymmetric(8)
    for trunc, full in (("dqgmres", "gmres"), ("diom", "fom")):
        xt, st = getattr(O, trunc)(A, b, memory=40)
        xf, sf = getattr(O, full)(A, b, memory=40)
        assert st["niter"] == sf["niter"] and np.allclose(st["residuals"], sf["residuals"], rtol=1e-12)
        assert np.allclose(xt, xf, rtol=1e-9, atol=1e-12)


# test/test_cr.jl
def test_cr_reference_cases(O):
    tol = 1e-6
    A, b = O.symmetric_definite()
    x, st = O.cr(A, b)
    assert resid(A, x, b) <= tol and st["solved"] and not st["indefinite"]
    radius = 0.75 * np.linalg.norm(x)
    x, st = O.cr(A, b, radius=radius)
    assert st["solved"] and abs(np.linalg.norm(x) - radius) <= tol * radius
    A, _ = O.sparse_laplacian()
    b = np.random.default_rng(0).standard_normal(A.shape[0])
    x, st = O.cr(A, b, radius=10.0)                                   # ||x*|| > radius
    assert abs(np.linalg.norm(x) - 10.0) <= tol * 10.0 and st["solved"]
    x, st = O.cr(A, b, radius=30.0)                                   # ||x*|| < radius
    assert resid(A, x, b) <= tol and st["solved"]
    radius = 0.75 * np.linalg.norm(x)
    x, st = O.cr(A, b, radius=radius)
    assert st["solved"] and abs(radius - np.linalg.norm(x)) <= tol * radius
    A, b = O.zero_rhs()
    x, st = O.cr(A, b)
    assert np.linalg.norm(x) == 0 and st["status"] == "x is a zero-residual solution"
    A, b, M = O.square_preconditioned()
    x, st = O.cr(A, b, M=M, atol=1e-5, rtol=0.0)
    r = b - A @ x
    assert np.sqrt(r @ (M * r)) / np.sqrt(b @ (M * b)) <= 10 * tol and st["solved"]
    A, b = O.symmetric_indefinite(shift=10)                           # linesearch stops at the first call
    x, st = O.cr(A, b, linesearch=True)
    npc = st["npc_dir"]
    assert st["status"] == "nonpositive curvature" and st["niter"] == 0 and st["solved"] and st["indefinite"]
    assert npc @ (A @ npc) <= 0 and np.array_equal(x, b)
    A2 = sp.csr_matrix(np.array([[1.0, 0.0], [0.0, 0.0]]))            # 2 negative-curvature directions
    x, st = O.cr(A2, np.ones(2), linesearch=True)
    assert st["npcCount"] == 2
    A3 = sp.csr_matrix(-np.eye(2))                                    # only -p negative curvature
    x, st = O.cr(A3, np.ones(2), linesearch=True)
    assert st["status"] == "nonpositive curvature" and st["npcCount"] == 1
    A4 = sp.csr_matrix(np.array([[0.0, 1.0], [1.0, 0.0]]))            # system_zero_quad: b'Ab == 0
    x, st = O.cr(A4, np.array([1.0, 0.0]), linesearch=True)
    assert st["niter"] == 0 and st["status"] == "b is a zero-curvature direction" and st["npc_dir"] @ (A4 @ st["npc_dir"]) == 0
    x, st = O.cr(A4, np.array([1.0, 0.0]))
    assert st["status"] == "b is a zero-curvature direction" and np.linalg.norm(x) == 0 and st["solved"] and st["niter"] == 0
    with pytest.raises(ArithmeticError):
        O.cr(A, b, linesearch=True, radius=1.0)


# ---- remaining generator-driven cases of the reference's solver tests -----------------------------------------
@pytest.mark.parametrize("name", ["gmres", "bicgstab", "cgs", "fom", "fgmres", "dqgmres", "diom"])
def test_unsymmetric_family_on_the_reference_generators(O, name):
    """test_gmres.jl / test_bicgstab.jl / test_cgs.jl / test_fom.jl / test_fgmres.jl / test_dqgmres.jl / test_diom.jl:
    nonsymmetric definite and indefinite systems, split preconditioning, Poisson in polar coordinates."""
    f = getattr(O, name)
    tol = 1e-6
    for gen in (O.nonsymmetric_definite, O.nonsymmetric_indefinite):
        A, b = gen()
        x, st = f(A, b)
        assert resid(A, x, b) <= tol and st["solved"], gen.__name__
    A, b, M, N = O.two_preconditioners()
    x, st = f(A, b, M=M, N=N)
    assert np.linalg.norm(M * (b - A @ x)) / np.linalg.norm(M * b) <= tol and st["solved"]
    if name != "cgs":                                   # test_cgs.jl has no polar_poisson case (CGS diverges on it)
        A, b = O.polar_poisson(12, 12)
        kw = dict(reorthogonalization=True) if name in ("gmres", "fom", "fgmres", "dqgmres", "diom") else {}
        if name in ("dqgmres", "diom"):
            kw["memory"] = 100                          # test_dqgmres.jl / test_diom.jl: memory = 100
        x, st = f(A, b, **kw)
        assert resid(A, x, b) <= tol and st["solved"]
    A, b, c = O.bc_breakdown()
    if name in ("bicgstab", "cgs"):
        x, st = f(A, b, c=c)
        assert st["status"] == "Breakdown bᴴc = 0"


def test_cg_cr_on_system_zero_quad(O):
    """test_cr.jl (system_zero_quad) and the same system through cg!: b'Ab = 0 at the first iteration."""
    A, b = O.system_zero_quad()
    x, st = O.cr(A, b, linesearch=True)
    assert st["niter"] == 0 and st["status"] == "b is a zero-curvature direction" and st["npc_dir"] @ (A @ st["npc_dir"]) == 0
    x, st = O.cg(A, b, linesearch=True)                 # cg.jl:198-209: both flags set, the later status wins (:272-275)
    assert st["status"] == "zero curvature detected" and st["niter"] == 0 and st["indefinite"] and st["npcCount"] == 1
    assert not st["inconsistent"] and np.array_equal(x, b)
    x, st = O.cg(A, b)                                  # zero curvature without linesearch: inconsistent = true
    assert st["status"] == "zero curvature detected" and st["inconsistent"] and not st["indefinite"]


def test_block_jacobi_twin_is_the_dense_block_operator(O):
    """The oracle's block-preconditioner knob (bdiagmul) against NumPy: mul! = block product, ldiv! = block solve,
    and CG preconditioned with the inverted diagonal blocks converges to the solution in fewer iterations."""
    import scipy.sparse as sp
    A, b = O.sparse_laplacian(6)
    A = sp.csr_matrix(A + sp.diags(np.linspace(0.0, 2.0, A.shape[0])))
    n, bs = A.shape[0], 4
    nb = n // bs
    D = np.stack([A[k * bs:(k + 1) * bs, k * bs:(k + 1) * bs].toarray() for k in range(nb)])
    Dinv = np.linalg.inv(D)
    x0, s0 = O.cg(A, b, atol=0.0, rtol=1e-10)
    with O.precond_block(bs):
        x1, s1 = O.cg(A, b, M=Dinv.reshape(-1), atol=0.0, rtol=1e-10)
        x2, s2 = O.cg(A, b, M=D.reshape(-1), ldiv=True, atol=0.0, rtol=1e-10)
    xs = sp.linalg.spsolve(sp.csc_matrix(A), b)
    assert s1["solved"] and s2["solved"] and s1["niter"] < s0["niter"] and abs(s1["niter"] - s2["niter"]) <= 1
    assert np.linalg.norm(x1 - xs) <= 1e-8 * np.linalg.norm(xs) and np.linalg.norm(x2 - xs) <= 1e-8 * np.linalg.norm(xs)
    assert np.allclose(s1["residuals"][:5], s2["residuals"][:5], rtol=1e-10)
