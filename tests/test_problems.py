"""The closed-form CSR generators (product side) equal the literal Kronecker
transcriptions of the reference's generators (oracle side), entry by entry."""
import numpy as np
import scipy.sparse as sp

from krylov_b200 import problems as P


def _same(csr, A):
    rp, ci, va = csr
    A = sp.csr_matrix(A)
    A.sort_indices()
    assert rp.dtype == np.int32 and ci.dtype == np.int32
    assert np.array_equal(rp, A.indptr) and np.array_equal(ci, A.indices) and np.array_equal(va, A.data)


def test_div_grad_matches_literal(O):
    for dims in ((4, 4, 4), (5, 3, 2), (1, 1, 1), (2, 7, 3)):
        _same(P.div_grad_csr(*dims), O.get_div_grad(*dims))


def test_kron_unsymmetric_matches_literal(O):
    for n in (2, 3, 6):
        A, b = O.kron_unsymmetric(n)
        csr = P.kron_unsymmetric_csr(n)
        _same(csr, A)
        assert np.allclose(P.csr_matvec_ones(*csr), b, atol=1e-13)


def test_row_slabs_concatenate():
    n = 6
    full = P.div_grad_csr(n)
    parts = [P.div_grad_csr(n, k_lo=a, k_hi=b) for a, b in ((0, 2), (2, 5), (5, 6))]
    ci = np.concatenate([p[1] for p in parts])
    va = np.concatenate([p[2] for p in parts])
    assert np.array_equal(ci, full[1]) and np.array_equal(va, full[2])
    assert sum(len(p[0]) - 1 for p in parts) == n ** 3


def test_random_csr_is_deterministic():
    a = P.random_csr(2000, 20, seed=1234)
    b = P.random_csr(2000, 20, seed=1234)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    A = sp.csr_matrix((a[2], a[1], a[0]), shape=(2000, 2000))
    assert A.has_sorted_indices and abs(A.diagonal().mean() - 3.0) < 0.1
