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
