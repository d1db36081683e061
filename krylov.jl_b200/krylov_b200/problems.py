"""Synthetic benchmark matrices, generated directly as int32 / 0-based CSR.

These are closed-form equivalents of the reference's test generators
(test/get_div_grad.jl:8-25, test/test_utils.jl:153-169) that scale to n ~ 1e8
without forming Kronecker products; tests/test_problems.py checks them entry
by entry against the literal transcriptions in oracle/oracle.py at small N.
`xp` may be numpy or torch (torch lets the bench build the matrix on the GPU).
"""
from __future__ import annotations

import numpy as np


def _stencil_csr(n1, n2, n3, diag, lo, hi, dtype, k_lo=0, k_hi=None, xp=np, device=None):
    """7-point stencil on an n1 x n2 x n3 grid, unknown (i,j,k) -> i + n1*(j + n2*k).

    lo[d] / hi[d]: coefficient of the neighbour at -1 / +1 along axis d
    (d = 0 fastest).  Rows of planes k_lo <= k < k_hi only (row slab), with
    GLOBAL column indices.  Returns (rowptr, colind, values), columns ascending.
    """
    if k_hi is None:
        k_hi = n3
    is_t = xp is not np
    kw = dict(device=device) if is_t else {}
    i64 = xp.int64
    nloc = n1 * n2 * (k_hi - k_lo)
    rows = xp.arange(n1 * n2 * k_lo, n1 * n2 * k_hi, dtype=i64, **kw)
    i = rows % n1
    j = (rows // n1) % n2
    k = rows // (n1 * n2)
    strides = (1, n1, n1 * n2)
    # candidate slots in ascending column order: -s2, -s1, -s0, 0, +s0, +s1, +s2
    offs = [-strides[2], -strides[1], -strides[0], 0, strides[0], strides[1], strides[2]]
    vals = [lo[2], lo[1], lo[0], diag, hi[0], hi[1], hi[2]]
    masks = [k > 0, j > 0, i > 0, None, i < n1 - 1, j < n2 - 1, k < n3 - 1]
    if is_t:
        import torch
        mask = torch.ones((nloc, 7), dtype=torch.bool, **kw)
        for s, m in enumerate(masks):
            if m is not None:
                mask[:, s] = m
        cols = rows[:, None] + torch.tensor(offs, dtype=i64, **kw)[None, :]
        tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
        v = torch.tensor(vals, dtype=tdt, **kw)[None, :].expand(nloc, 7)
        counts = mask.sum(dim=1)
        rowptr = torch.zeros(nloc + 1, dtype=i64, **kw)
        rowptr[1:] = torch.cumsum(counts, 0)
        return rowptr.to(torch.int32), cols[mask].to(torch.int32), v[mask].contiguous()
    mask = np.ones((nloc, 7), dtype=bool)
    for s, m in enumerate(masks):
        if m is not None:
            mask[:, s] = m
    cols = rows[:, None] + np.asarray(offs, dtype=np.int64)[None, :]
    v = np.broadcast_to(np.asarray(vals, dtype=dtype)[None, :], (nloc, 7))
    rowptr = np.zeros(nloc + 1, dtype=np.int64)
    np.cumsum(mask.sum(axis=1), out=rowptr[1:])
    return rowptr.astype(np.int32), cols[mask].astype(np.int32), np.ascontiguousarray(v[mask])


def div_grad_csr(n1, n2=None, n3=None, dtype=np.float64, k_lo=0, k_hi=None, xp=np, device=None):
    """get_div_grad(n1,n2,n3) = Div*Div' (test/get_div_grad.jl:8-19): diagonal 6, neighbours -1."""
    n2 = n1 if n2 is None else n2
    n3 = n1 if n3 is None else n3
    return _stencil_csr(n1, n2, n3, 6.0, (-1.0, -1.0, -1.0), (-1.0, -1.0, -1.0), dtype, k_lo, k_hi, xp, device)


def kron_unsymmetric_csr(n, dtype=np.float64, k_lo=0, k_hi=None, xp=np, device=None):
    """kron_unsymmetric(n) (test/test_utils.jl:160-169): with T = tridiag(-1, 3, -2),
    A = T(x)I(x)I + 2 I(x)T(x)I + I(x)I(x)T; the first Kronecker factor is the slowest index."""
    return _stencil_csr(n, n, n, 12.0, (-1.0, -2.0, -1.0), (-2.0, -4.0, -2.0), dtype, k_lo, k_hi, xp, device)


def random_csr(n, per_row=20, seed=1234, dtype=np.float32, shift=3.0):
    """BASELINE config 4: per row `per_row` iid uniform columns with U(-1,1) values
    (indices drawn first, then values, numpy default_rng(seed)), duplicates summed, +shift on the diagonal."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    cols = rng.integers(0, n, size=(n, per_row), dtype=np.int64)
    vals = rng.uniform(-1.0, 1.0, size=(n, per_row)).astype(dtype)
    rows = np.repeat(np.arange(n, dtype=np.int64), per_row)
    A = sp.coo_matrix((vals.ravel(), (rows, cols.ravel())), shape=(n, n)).tocsr()   # sums duplicates
    A = (A + shift * sp.identity(n, dtype=dtype, format="csr")).tocsr()
    A.sort_indices()
    return A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(dtype)


def csr_matvec_ones(rowptr, colind, values):
    """b = A * ones (row sums), as the reference builds b for kron_unsymmetric."""
    if isinstance(values, np.ndarray):
        return np.add.reduceat(values, rowptr[:-1].astype(np.int64)) if len(values) else np.zeros(len(rowptr) - 1, values.dtype)
    import torch
    n = rowptr.numel() - 1
    out = torch.zeros(n, dtype=values.dtype, device=values.device)
    rows = torch.repeat_interleave(torch.arange(n, device=values.device), (rowptr[1:] - rowptr[:-1]).long())
    out.index_add_(0, rows, values)
    return out
