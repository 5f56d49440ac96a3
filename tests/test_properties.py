"""Property tests (hypothesis) of the CPU oracle on random sparse systems -- SURVEY.md
section 4, item 3: invariants the domain offers independently of any particular matrix.
The device path is compared with the oracle in tests/test_gpu_parity.py; here the oracle
itself is held to the properties, on ragged unstructured matrices that the FE generators
never produce."""
import numpy as np
import scipy.sparse as sp
from hypothesis import HealthCheck, given, settings, strategies as st

from helpers import device_data, random_spd_csr
from oracle import binding as orc

# derandomize: the same examples on every run (a CI gate must not depend on the draw)
SETTINGS = dict(max_examples=12, deadline=None, derandomize=True, database=None,
                suppress_health_check=[HealthCheck.too_slow])


def _csr(n, density, seed):
    A = random_spd_csr(n, density, seed)
    return A, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)


@settings(**SETTINGS)
@given(n=st.integers(20, 160), density=st.floats(0.02, 0.2), seed=st.integers(0, 10 ** 6),
       t1=st.floats(0.05, 0.5), dt=st.floats(0.0, 0.45))
def test_strength_mask_is_monotone_in_theta_and_inside_the_pattern(n, density, seed, t1, dt):
    A, rp, col, val = _csr(n, density, seed)
    m1 = orc.strength(rp, col, val, t1, 0.9)
    m2 = orc.strength(rp, col, val, t1 + dt, 0.9)
    assert m1.shape == val.shape and set(np.unique(m1)) <= {0, 1}
    assert (m2 <= m1).all()                       # S(theta2) is a subset of S(theta1) for theta2 > theta1
    rows = np.repeat(np.arange(n), np.diff(rp))
    assert not m1[rows == col].any()              # the diagonal is never a strong connection


@settings(**SETTINGS)
@given(n=st.integers(30, 160), density=st.floats(0.03, 0.15), seed=st.integers(0, 10 ** 6),
       theta=st.sampled_from([0.1, 0.25, 0.5, 0.8]))
def test_hierarchy_invariants(n, density, seed, theta):
    A, rp, col, val = _csr(n, density, seed)
    H = orc.Hierarchy(rp, col, val, device_data(theta).to_struct())
    for l in range(H.num_levels - 1):
        arp, acl, avl = H.A(l)
        Al = sp.csr_matrix((avl, acl, arp))
        cf = H.cf_marker(l)
        mask = H.strength_mask(l)
        prp, pcl, pvl, nc = H.P(l)
        P = sp.csr_matrix((pvl, pcl, prp), shape=(Al.shape[0], nc))
        assert nc == int((cf > 0).sum())
        # C points interpolate from themselves with weight one
        c_rows = np.flatnonzero(cf > 0)
        assert (np.diff(prp)[c_rows] == 1).all() and np.allclose(P[c_rows].data, 1.0)
        # the interpolation row of an F point holds exactly its strong C neighbours
        S = sp.csr_matrix((mask.astype(np.float64), acl, arp), shape=Al.shape)
        strong_c = np.asarray((S @ sp.diags((cf > 0).astype(np.float64))).sum(axis=1)).ravel()
        f_rows = np.flatnonzero(cf <= 0)
        assert np.array_equal(np.diff(prp)[f_rows], strong_c[f_rows].astype(np.int64))
        # Galerkin: the next level is P^T A P, symmetric when A is
        nrp, ncl, nvl = H.A(l + 1)
        An = sp.csr_matrix((nvl, ncl, nrp), shape=(nc, nc))
        G = (P.T @ Al @ P).tocsr()
        assert abs(An - G).max() <= 1e-12 * max(1.0, abs(G).max())
        assert abs(An - An.T).max() <= 1e-12 * max(1.0, abs(An).max())
    H.close()


@settings(**SETTINGS)
@given(n=st.integers(10, 300), density=st.floats(0.01, 0.3), seed=st.integers(0, 10 ** 6),
       V=st.integers(1, 40))
def test_pooled_image_conserves_count_and_sum(n, density, seed, V):
    A, rp, col, val = _csr(n, density, seed)
    V = min(V, n)
    view, count, max_pp, max_np = orc.make_view(rp, col, val, V)
    assert count.sum() == len(val)                                   # every stored entry lands in one bin
    assert np.isclose(view.sum(), val.sum(), rtol=1e-12, atol=1e-12 * np.abs(val).sum())
    assert (max_pp >= 0).all() and (max_np >= 0).all()
    assert np.isclose(max_pp.max(), max(val.max(), 0.0)) and np.isclose(max_np.max(), max((-val).max(), 0.0))


@settings(**SETTINGS)
@given(n=st.integers(30, 200), density=st.floats(0.03, 0.15), seed=st.integers(0, 10 ** 6))
def test_pcg_reaches_the_direct_solution(n, density, seed):
    import scipy.sparse.linalg as spl
    A, rp, col, val = _csr(n, density, seed)
    rng = np.random.default_rng(seed)
    b = rng.normal(size=n)
    H = orc.Hierarchy(rp, col, val, device_data(0.25).to_struct())
    rc, x, nit, hist = H.cg_solve(b, np.zeros(n), abs_tol=1e-10)
    assert rc == 0 and nit <= n
    assert (np.diff(np.log(hist[hist > 0])) < 5).all()   # no blow-up of the preconditioned residual
    xd = spl.spsolve(A.tocsc(), b)
    assert np.abs(x - xd).max() <= 1e-6 * max(1.0, np.abs(xd).max())
    H.close()


@settings(**SETTINGS)
@given(n=st.integers(30, 200), density=st.floats(0.02, 0.2), seed=st.integers(0, 10 ** 6))
def test_multicolouring_is_valid_and_the_sweep_is_gauss_seidel_in_colour_order(n, density, seed):
    """smoother_policy = SMOOTHER_MULTICOLOR on random symmetric patterns: coupled points never share a
    colour, no more colours than the largest degree + 1, and one forward sweep from zero on a one-level
    hierarchy equals the lower-triangular solve of the colour-ordered system."""
    import amg_ann_b200 as ab
    R = ab.RelaxationType
    A, rp, col, val = _csr(n, density, seed)
    data = device_data(0.25, relaxation_type_up=R.backwardSORJacobi, relaxation_type_down=R.SORJacobi,
                       relaxation_type_coarse=R.SORJacobi, smoother_policy=ab.SMOOTHER_MULTICOLOR, max_levels=1)
    data.symmetric_operator = False  # (deal.II maps SORJacobi to the symmetric sweep for symmetric operators)
    H = orc.Hierarchy(rp, col, val, data.to_struct())
    assert H.effective_relax() == (103, 104, 103)
    colors, nc = H.colors(0)
    rows = np.repeat(np.arange(n), np.diff(rp))
    off = rows != col
    assert not (colors[rows[off]] == colors[col[off]]).any()
    assert nc <= int(np.diff(rp).max()) + 1 and set(np.unique(colors)) == set(range(nc))
    r = np.random.default_rng(seed).standard_normal(n)
    z = H.vmult(r)                                   # one level: the coarse relaxation = one forward sweep
    order = np.lexsort((np.arange(n), colors))
    Ap = A.toarray()[np.ix_(order, order)]
    want = np.linalg.solve(np.tril(Ap), r[order])
    assert np.abs(z[order] - want).max() <= 1e-10 * max(1.0, np.abs(want).max())


def test_multicolouring_rejects_a_pattern_that_is_not_symmetric():
    import amg_ann_b200 as ab
    import pytest
    R = ab.RelaxationType
    A = random_spd_csr(200, 0.05, 1).tolil()
    rows, cols = A.nonzero()
    dropped = 0
    for i, j in zip(rows, cols):
        if i > j and dropped < 60:
            A[i, j] = 0
            dropped += 1
    A = A.tocsr()
    A.eliminate_zeros()
    A.sort_indices()
    data = device_data(0.25, relaxation_type_up=R.symmetricSORJacobi, relaxation_type_down=R.symmetricSORJacobi,
                       smoother_policy=ab.SMOOTHER_MULTICOLOR)
    with pytest.raises(RuntimeError):
        orc.Hierarchy(A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data, data.to_struct())
