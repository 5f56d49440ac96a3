"""Shared problem builders for the tests."""
import numpy as np

import amg_ann_b200 as ab

R = ab.RelaxationType


def device_data(theta=0.25, **kw):
    """AdditionalData as the reference passes it (sym=1, theta, mrs=0.9, agg=0,
    details) with the device-capable smoother made explicit."""
    kw.setdefault("relaxation_type_up", R.l1scaledJacobi)
    kw.setdefault("relaxation_type_down", R.l1scaledJacobi)
    kw.setdefault("keep_setup_intermediates", True)
    return ab.AdditionalData(True, theta, 0.9, 0, True, **kw)


def poisson(m, contrast=0.0, ps=2, mode=3):
    epsv = ab.gen.checkerboard_epsv(ps, mode, contrast) if contrast else None
    if epsv is None:
        return ab.gen.poisson_q1(m)
    return ab.gen.poisson_q1(m, ps, mode, epsv)


def random_spd_csr(n, density, seed):
    """Small unstructured SPD-ish M-matrix (ragged rows, includes an isolated row)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    B = sp.random(n, n, density=density, random_state=rng, format="csr")
    B = -(abs(B) + abs(B.T))
    B.setdiag(0)
    B.eliminate_zeros()
    d = -np.asarray(B.sum(axis=1)).ravel() + rng.random(n) * 0.1 + 1e-3
    A = (B + sp.diags(d)).tocsr()
    A.sort_indices()
    return A


def spd_laplacian(m, seed=0, decades=2.0):
    """Symmetric positive definite test system: 7-point Laplacian on an m^3 grid, symmetrically
    scaled by 10^U(0, decades) per point.  (The Q1 generators keep deal.II's row-only
    elimination of Dirichlet values for PETSc matrices, so their matrices are not symmetric;
    smoothers that estimate a spectrum with CG need a symmetric operator.)"""
    import scipy.sparse as sp
    from types import SimpleNamespace
    rng = np.random.default_rng(seed)
    eye = sp.identity(m)
    t = sp.diags([-np.ones(m - 1), 2 * np.ones(m), -np.ones(m - 1)], [-1, 0, 1])
    A = sp.kron(sp.kron(t, eye), eye) + sp.kron(sp.kron(eye, t), eye) + sp.kron(sp.kron(eye, eye), t)
    w = sp.diags(10.0 ** rng.uniform(0, decades, m ** 3))
    A = (w @ A @ w).tocsr()
    A.sort_indices()
    n = A.shape[0]
    return SimpleNamespace(n=n, nnz=A.nnz, rowptr=A.indptr.astype(np.int64), col=A.indices.astype(np.int32),
                           val=A.data.astype(np.float64), rhs=rng.normal(size=n), x0=np.zeros(n), csr=A,
                           rowptr32=lambda: A.indptr.astype(np.int32))


def hub_leaf_csr(nh, nl, per_leaf, leaf_leaf, seed):
    """SPD M-matrix on nh hubs + nl leaves: every leaf is coupled to `per_leaf` hubs and to about
    `leaf_leaf` other leaves, hubs are not coupled to each other.  PMIS picks the hubs (largest
    measures) as C points, so every leaf's interpolation row has `per_leaf` entries."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    n = nh + nl
    rows, cols, vals = [], [], []
    for leaf in range(nl):
        for h in rng.choice(nh, size=per_leaf, replace=False):
            w = -rng.uniform(0.5, 1.0)
            rows += [nh + leaf, h]
            cols += [h, nh + leaf]
            vals += [w, w]
        for o in rng.choice(nl, size=leaf_leaf, replace=False):
            if o != leaf:
                w = -rng.uniform(0.5, 1.0)
                rows += [nh + leaf, nh + o]
                cols += [nh + o, nh + leaf]
                vals += [w, w]
    B = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    B.sum_duplicates()
    d = -np.asarray(B.sum(axis=1)).ravel() + rng.random(n) * 0.1 + 1e-3
    A = (B + sp.diags(d)).tocsr()
    A.sort_indices()
    return A
