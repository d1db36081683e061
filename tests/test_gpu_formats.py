"""GPU: the data formats either side of the path -- Matrix Market ingestion and the transposed
operator.  The checker is scipy.io.mmread / scipy's transpose on the same files (bit-exact: integer and index work,
values copied or summed in file order)."""
import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def _same(op, ref):
    got = op.to_scipy()
    ref = sp.csr_matrix(ref)
    ref.sum_duplicates()
    ref.sort_indices()
    assert got.shape == ref.shape and got.nnz == ref.nnz
    assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)
    assert np.array_equal(got.data, ref.data.astype(got.dtype))


def test_read_matrix_market_variants(kb, O, tmp_path):
    rng = np.random.default_rng(0)
    A = sp.random(40, 40, density=0.1, random_state=1, format="coo") + sp.identity(40)
    scipy.io.mmwrite(str(tmp_path / "general.mtx"), sp.coo_matrix(A))
    S = sp.coo_matrix(A + A.T)
    scipy.io.mmwrite(str(tmp_path / "symmetric.mtx"), S, symmetry="symmetric")
    K = sp.coo_matrix(sp.triu(A, 1) - sp.triu(A, 1).T)
    scipy.io.mmwrite(str(tmp_path / "skew.mtx"), K, symmetry="skew-symmetric")
    Pm = sp.coo_matrix((np.ones(A.nnz), (sp.coo_matrix(A).row, sp.coo_matrix(A).col)), shape=A.shape)
    scipy.io.mmwrite(str(tmp_path / "pattern.mtx"), Pm, field="pattern")
    Ii = sp.coo_matrix((rng.integers(-5, 6, A.nnz), (sp.coo_matrix(A).row, sp.coo_matrix(A).col)), shape=A.shape)
    scipy.io.mmwrite(str(tmp_path / "integer.mtx"), Ii, field="integer")
    for name in ("general", "symmetric", "skew", "pattern", "integer"):
        path = str(tmp_path / f"{name}.mtx")
        op = kb.CsrOperator.read_mtx(path)
        _same(op, scipy.io.mmread(path))
        op.free()
    # comments, blank lines, 1-based indices, duplicates summed in file order (SparseArrays.sparse semantics)
    path = _write(tmp_path, "dups.mtx", "%%MatrixMarket matrix coordinate real general\n% a comment\n\n3 3 5\n"
                  "1 1 1.5\n3 2 -2\n1 1 0.25\n2 3 4e0\n3 2 1\n")
    op = kb.CsrOperator.read_mtx(path)
    assert np.array_equal(op.to_scipy().toarray(), np.array([[1.75, 0, 0], [0, 0, 4.0], [0, -1.0, 0]]))
    op32 = kb.CsrOperator.read_mtx(path, dtype=np.float32)
    assert op32.to_scipy().dtype == np.float32 and np.array_equal(op32.to_scipy().toarray(), op.to_scipy().toarray())


def test_read_matrix_market_errors(kb, tmp_path):
    cases = {
        "complex.mtx": "%%MatrixMarket matrix coordinate complex general\n2 2 1\n1 1 1 0\n",
        "array.mtx": "%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n",
        "rect.mtx": "%%MatrixMarket matrix coordinate real general\n2 3 1\n1 1 1\n",
        "trunc.mtx": "%%MatrixMarket matrix coordinate real general\n2 2 3\n1 1 1\n",
        "range.mtx": "%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1\n",
        "banner.mtx": "2 2 1\n1 1 1\n",
    }
    for name, text in cases.items():
        with pytest.raises(kb.B200Error):
            kb.CsrOperator.read_mtx(_write(tmp_path, name, text))
    with pytest.raises(kb.B200Error):
        kb.CsrOperator.read_mtx(str(tmp_path / "missing.mtx"))


def test_transpose_and_solve_from_file(kb, O, tmp_path):
    A, b = O.kron_unsymmetric(7)
    A = sp.csr_matrix(A)
    path = str(tmp_path / "kron.mtx")
    scipy.io.mmwrite(path, sp.coo_matrix(A), precision=17)
    op = kb.CsrOperator.read_mtx(path)
    _same(op, A)
    At = op.transpose()
    _same(At, A.T)
    x = np.cos(np.arange(A.shape[0]))
    assert np.array_equal(At.matvec(x), O.spmv(sp.csr_matrix(A.T), x))      # bit-identical row sums (column order)
    assert np.array_equal(op.matvec(x), O.spmv(A, x))
    _same(At.transpose(), A)
    # solvers take the device-resident operator like a SciPy matrix: same iterates as the upload path
    x1, s1 = kb.gmres(op, b, memory=20, history=True)
    x2, s2 = kb.gmres(A, b, memory=20, history=True)
    assert s1.niter == s2.niter and s1.residuals == s2.residuals and np.array_equal(x1, x2)
    xt, st = kb.bicgstab(At, b, history=True)
    xo, so = O.bicgstab(sp.csr_matrix(A.T), b)
    assert st.niter == so["niter"] and np.allclose(st.residuals, so["residuals"], rtol=1e-5)
    S = kb.CsrOperator.from_scipy(O.sparse_laplacian(8)[0])
    xs, ss = kb.cg(S, O.sparse_laplacian(8)[1], history=True)
    xo, so = O.cg(*O.sparse_laplacian(8))
    assert ss.niter == so["niter"] and np.allclose(xs, xo, rtol=1e-7)
