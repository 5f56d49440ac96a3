"""GPU (-m gpu): DoF numberings of the step before the path as device
permutations: deal.II's distribute_dofs order on a refined subdivided_hyper_cube against an
independent sequential traversal, Q A Q^T against scipy, Cuthill-McKee end to end."""
import numpy as np
import pytest

import amg_ann_b200 as ab
from helpers import device_data, poisson

pytestmark = pytest.mark.gpu


def _traverse(c, r):
    """Sequential restatement: coarse cells lexicographic, children of a cell in lexicographic
    order (x fastest), vertices of an active cell in lexicographic order, numbered when first met."""
    m = c << r
    N = m + 1
    number = -np.ones(N ** 3, dtype=np.int64)
    order = []

    def visit(x0, y0, z0, size):
        if size == 1:
            for v in range(8):
                vx, vy, vz = x0 + (v & 1), y0 + ((v >> 1) & 1), z0 + ((v >> 2) & 1)
                lex = vx + N * (vy + N * vz)
                if number[lex] < 0:
                    number[lex] = len(order)
                    order.append(lex)
            return
        h = size // 2
        for ch in range(8):
            visit(x0 + (ch & 1) * h, y0 + ((ch >> 1) & 1) * h, z0 + ((ch >> 2) & 1) * h, h)

    for Z in range(c):
        for Y in range(c):
            for X in range(c):
                visit(X << r, Y << r, Z << r, 1 << r)
    return np.array(order, dtype=np.int32)


@pytest.mark.parametrize("c,r", [(1, 0), (3, 0), (1, 2), (2, 1), (3, 2), (4, 3), (5, 1)])
def test_dealii_numbering_matches_a_sequential_traversal(gpu_ctx, c, r):
    got = ab.numbering_dealii_q1(gpu_ctx, c, r)
    want = _traverse(c, r)
    assert np.array_equal(got, want)
    assert np.array_equal(np.sort(got), np.arange(len(got)))       # a permutation
    if r == 0 and c == 1:
        assert list(got) == list(range(8))                          # one cell: the lexicographic vertices


def test_permuted_matrix_equals_scipy_and_solves_the_same_system(gpu_ctx):
    import scipy.sparse as sp
    c, r = 3, 2
    m = c << r
    s = poisson(m, contrast=3.0, ps=3)
    new_to_lex = ab.numbering_dealii_q1(gpu_ctx, c, r)
    A = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
    B = A.permuted(new_to_lex)
    rp, cl, vl = B.download()
    M = sp.csr_matrix((s.val, s.col, s.rowptr), shape=(s.n, s.n))
    want = M[new_to_lex][:, new_to_lex].tocsr()
    want.sort_indices()
    assert np.array_equal(rp, want.indptr) and np.array_equal(cl, want.indices) and np.array_equal(vl, want.data)
    # same system in another numbering: the permuted solution is the solution, to the solver tolerance
    x1 = s.x0.copy()
    row1 = ab.amg_solve(device_data(0.25), 1e-10, A, s.rhs, x1)
    x2 = s.x0[new_to_lex].copy()
    row2 = ab.amg_solve(device_data(0.25), 1e-10, B, s.rhs[new_to_lex], x2)
    assert np.abs(x2 - x1[new_to_lex]).max() <= 1e-7 * np.abs(x1).max()
    assert abs(row1["niters"] - row2["niters"]) <= 6               # PMIS tie breaks depend on the numbering
    # the pooled image does depend on the numbering: that is why the numbering matters (row f1)
    v1 = ab.ViewMaker(10).make_view(A).count
    v2 = ab.ViewMaker(10).make_view(B).count
    assert v1.sum() == v2.sum() == s.nnz and not np.array_equal(v1, v2)
    with pytest.raises(ab.AmgbError):
        A.permuted(np.zeros(s.n, dtype=np.int32))                   # not a permutation


def test_cuthill_mckee_renumbering_end_to_end(gpu_ctx):
    s = poisson(10, contrast=2.0)
    perm = ab.gen.cuthill_mckee(s.rowptr, s.col)
    assert np.array_equal(np.sort(perm), np.arange(s.n))
    deg = np.diff(s.rowptr)
    assert deg[perm[0]] == deg.min()                                # starts at a point of least coordination
    A = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
    B = A.permuted(perm)
    rp, cl, vl = B.download()
    inv = np.empty_like(perm)
    inv[perm] = np.arange(s.n)
    rows = np.repeat(np.arange(s.n), np.diff(rp))
    # level structure: an entry never couples points more than two BFS levels apart, and the
    # reverse ordering is its mirror image
    rperm = ab.gen.cuthill_mckee(s.rowptr, s.col, reversed=True)
    assert np.array_equal(rperm, perm[::-1])
    x = s.x0[perm].copy()
    row = ab.amg_solve(device_data(0.25), 1e-8, B, s.rhs[perm], x)
    assert row["p_res"][-1] <= 1e-8
    r = s.rhs[perm] - B.vmult(x)
    assert np.linalg.norm(r) <= 1e-6 * np.linalg.norm(s.rhs)
    del rows, cl, vl, inv
