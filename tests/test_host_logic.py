"""CPU: host-side logic of the product library that makes no CUDA call -- the Matrix Market parser and the small
dense algebra of the block path -- checked against scipy / numpy / the oracle."""
import ctypes as C

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

from krylov_b200 import _lib


def _mtx_read(path):
    L = _lib.lib()
    n, nnz = C.c_int(), C.c_longlong()
    if L.kb200_mtx_read(path.encode(), C.byref(n), C.byref(nnz), None, None, None) != 0:
        raise RuntimeError(_lib.last_error())
    rp, ci, va = np.empty(n.value + 1, np.int32), np.empty(nnz.value, np.int32), np.empty(nnz.value)
    assert L.kb200_mtx_read(path.encode(), None, None, rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p),
                            va.ctypes.data_as(C.c_void_p)) == 0
    return sp.csr_matrix((va, ci, rp), shape=(n.value, n.value))


def test_matrix_market_parser_matches_scipy(tmp_path):
    rng = np.random.default_rng(0)
    A = sp.random(60, 60, density=0.08, random_state=2, format="coo") + sp.identity(60)
    variants = {
        "general": (sp.coo_matrix(A), {}),
        "symmetric": (sp.coo_matrix(A + A.T), dict(symmetry="symmetric")),
        "skew": (sp.coo_matrix(sp.triu(A, 1) - sp.triu(A, 1).T), dict(symmetry="skew-symmetric")),
        "pattern": (sp.coo_matrix(A), dict(field="pattern")),
        "integer": (sp.coo_matrix((rng.integers(-9, 10, A.nnz), (sp.coo_matrix(A).row, sp.coo_matrix(A).col)), shape=A.shape),
                    dict(field="integer")),
    }
    for name, (M, kw) in variants.items():
        path = str(tmp_path / f"{name}.mtx")
        scipy.io.mmwrite(path, M, **kw)
        got = _mtx_read(path)
        ref = sp.csr_matrix(scipy.io.mmread(path))
        ref.sum_duplicates(); ref.sort_indices()
        assert got.nnz == ref.nnz and np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices), name
        assert np.array_equal(got.data, ref.data.astype(float)), name
    p = tmp_path / "dups.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real general\n% c\n\n3 3 5\n1 1 1.5\n3 2 -2\n1 1 0.25\n2 3 4e0\n3 2 1\n")
    assert np.array_equal(_mtx_read(str(p)).toarray(), np.array([[1.75, 0, 0], [0, 0, 4.0], [0, -1.0, 0]]))
    for name, text in {"complex": "%%MatrixMarket matrix coordinate complex general\n2 2 1\n1 1 1 0\n",
                       "array": "%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n",
                       "rect": "%%MatrixMarket matrix coordinate real general\n2 3 1\n1 1 1\n",
                       "trunc": "%%MatrixMarket matrix coordinate real general\n2 2 3\n1 1 1\n",
                       "range": "%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1\n"}.items():
        q = tmp_path / f"{name}.mtx"
        q.write_text(text)
        with pytest.raises(RuntimeError):
            _mtx_read(str(q))


def _householder(Am, compact=False):
    L = _lib.lib()
    Q = np.array(Am, dtype=float, order="F")
    m, k = Q.shape
    R, tau = np.zeros((k, k), order="F"), np.zeros(k)
    assert L.kb200_host_householder(m, k, Q.ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p), tau.ctypes.data_as(C.c_void_p),
                                    int(compact)) == 0
    return Q, R, tau


def test_product_householder_matches_lapack_and_oracle(O):
    rng = np.random.default_rng(1)
    for m, k in ((50, 6), (8, 8), (40, 1), (64, 16), (16, 8)):
        Am = rng.standard_normal((m, k))
        Q, R, tau = _householder(Am)
        Qn, Rn = np.linalg.qr(Am)                       # LAPACK geqrf + orgqr
        assert np.abs(Q - Qn).max() <= 1e-13 and np.abs(R - Rn).max() <= 1e-13
        Qo, Ro, tauo = O.householder(Am)                # the oracle's restatement
        assert np.abs(Q - Qo).max() <= 1e-14 and np.abs(R - Ro).max() <= 1e-14 and np.abs(tau - tauo).max() <= 1e-14
        Qc, Rc, tauc = _householder(Am, compact=True)   # reflectors kept below the diagonal, R on and above it
        assert np.allclose(np.triu(Qc[:k]), R) and np.array_equal(tauc, tau)


def test_cholqr2_with_sign_reconstruction_reproduces_householder():
    """What panel_qr does on the device, replayed on the host with the library's own small-matrix code: two CholQR
    passes + the Householder signs of the top block give LAPACK's Q and R, signs included."""
    L = _lib.lib()
    rng = np.random.default_rng(2)
    for trial, (n, p) in enumerate(((200, 6), (300, 8), (100, 3), (64, 16), (50, 2))):
        A = rng.standard_normal((n, p))
        if trial == 1:
            A[:p, :p] = 0.0                             # zero top block
        Q = A.copy()
        Rs = []
        for _ in range(2):
            G = np.asfortranarray(Q.T @ Q)
            R, Rinv = np.zeros((p, p), order="F"), np.zeros((p, p), order="F")
            assert L.kb200_host_cholqr_factors(p, G.ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p),
                                               Rinv.ctypes.data_as(C.c_void_p)) == 0
            assert np.allclose(R.T @ R, G, rtol=1e-12) and np.allclose(R @ Rinv, np.eye(p), atol=1e-12)
            Q = Q @ Rinv
            Rs.append(R)
        top = np.asfortranarray(Q[:p, :p])
        s = np.zeros(p)
        assert L.kb200_host_householder_signs(p, top.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p)) == 0
        Qh, Rh = np.linalg.qr(A)
        assert np.abs(Q * s - Qh).max() <= 1e-12
        assert np.abs((s[:, None] * (Rs[1] @ Rs[0])) - Rh).max() <= 1e-11
    # a rank-deficient Gram matrix is refused (the device path then takes Householder on the host)
    B = rng.standard_normal((40, 3))
    B[:, 2] = B[:, 0]
    G = np.asfortranarray(B.T @ B)
    R, Rinv = np.zeros((3, 3), order="F"), np.zeros((3, 3), order="F")
    assert L.kb200_host_cholqr_factors(3, G.ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p), Rinv.ctypes.data_as(C.c_void_p)) == 1
