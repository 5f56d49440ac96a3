"""CPU suite: the oracle against independent scipy/numpy computations and the
structural properties SURVEY.md section 4 lists.  No GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

import amg_ann_b200 as ab
from helpers import device_data, poisson, random_spd_csr
from oracle import binding as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- generators
def test_generator_sizes_match_survey_formulas():
    # n = (m+1)^3, nnz = (3m+1)^3
    for m in (3, 7, 12):
        s = ab.gen.poisson_q1(m)
        assert s.n == (m + 1) ** 3 and s.nnz == (3 * m + 1) ** 3
    assert ab.gen.sizes(0, 100) == (1030301, 27270901)
    assert ab.gen.sizes(0, 200) == (8120601, 217081801)
    assert ab.gen.sizes(0, 464) == (100544625, 2703045457)


def test_poisson_stencil_and_dirichlet_rows():
    m = 6
    s = ab.gen.poisson_q1(m)
    A = s.to_scipy()
    h = 2.0 / m
    N = m + 1
    c = (N // 2) * (1 + N + N * N)  # interior node
    # diag 8h/3, faces 0, edges -h/6, corners -h/12
    assert A[c, c] == pytest.approx(8 * h / 3, rel=1e-13)
    assert abs(A[c, c + 1]) < 1e-15
    assert A[c, c + 1 + N] == pytest.approx(-h / 6, rel=1e-12)
    assert A[c, c + 1 + N + N * N] == pytest.approx(-h / 12, rel=1e-12)
    # explicit zeros are stored (A.5): every interior row has 27 entries
    assert s.rowptr[c + 1] - s.rowptr[c] == 27
    # Dirichlet row: pattern kept, off-diagonals zero, diag = |first diagonal|
    r0 = A.getrow(0)
    assert r0.nnz == 8 and A[0, 0] == pytest.approx(h / 3)
    assert abs(r0).sum() == pytest.approx(h / 3)
    # columns are NOT eliminated: interior rows keep couplings to boundary nodes
    i = 1 + N + N * N
    assert A[i, 0] != 0.0


def test_poisson_row_range_equals_slice_of_whole():
    m = 5
    s = ab.gen.poisson_q1(m)
    part = ab.gen.poisson_q1(m, row_begin=50, row_end=140)
    lo, hi = s.rowptr[50], s.rowptr[140]
    assert np.array_equal(part.rowptr, s.rowptr[50:141] - lo)
    assert np.array_equal(part.col, s.col[lo:hi])
    assert np.array_equal(part.val, s.val[lo:hi])
    assert np.array_equal(part.rhs, s.rhs[50:140])


def test_discontinuous_coefficients_contrast():
    s = poisson(8, contrast=6.0, ps=2, mode=3)
    d = s.to_scipy().diagonal()
    assert d.max() / d.min() > 1e5


def test_elasticity_is_symmetric_with_diagonal_constrained_rows():
    s = ab.gen.elasticity_q1(4)
    A = s.to_scipy()
    assert s.n == 3 * 5 ** 3 and s.nnz == ab.gen.sizes(1, 4)[1]
    assert abs(A - A.T).max() < 1e-9 * abs(A).max()
    lens = np.diff(s.rowptr)
    assert lens[0] == 1  # boundary DoF: diagonal only (t3 main.cpp:264-272)
    w = spl.eigsh(A, k=1, which="SA", return_eigenvectors=False)
    assert w[0] > 0


def test_theta_sweep_is_built_by_accumulation():
    # ref t2 datagen.py:47 "0.05,0.96,0.05" -> 19 values, t2 main.cpp:443
    th = ab.gen.theta_sweep(0.05, 0.96, 0.05)
    assert len(th) == 19
    assert th[2] == 0.05 + 0.05 + 0.05 and th[2] != 0.15
    # deal.II forwards theta through std::to_string (A.1, H4)
    assert orc.option_roundtrip(th[2]) == 0.15


def test_random_vec_matches_minstd_generate_canonical():
    v = ab.gen.random_vec(0, 4, 6.0)
    assert ((0 <= v) & (v < 6)).all() and len(set(v)) == 4
    assert np.array_equal(v, ab.gen.random_vec(0, 4, 6.0))


# ------------------------------------------------------------------- oracle
def test_hypre_rand_known_answers():
    # multiplicative LCG a=16807, m=2^31-1, seed 2747 (A.3): by hand
    m = 2147483647
    s = 2747
    for i in range(5):
        s = (16807 * s) % m
        assert orc.lib().orc_hypre_rand(i) == s / m


def test_strength_known_answer_uniform_stencil():
    # H3: corner/edge = 0.5, so theta = 0.25 keeps edges+corners, theta = 0.6 only edges
    m = 6
    s = ab.gen.poisson_q1(m)
    N = m + 1
    c = (N // 2) * (1 + N + N * N)
    rp = s.rowptr32()
    row = slice(rp[c], rp[c + 1])
    m25 = orc.strength(rp, s.col, s.val, 0.25)[row]
    m60 = orc.strength(rp, s.col, s.val, 0.6)[row]
    assert m25.sum() == 20 and m60.sum() == 12
    # Dirichlet rows: |row_sum| = diag > 0.9 diag -> no strong connections
    assert orc.strength(rp, s.col, s.val, 0.25)[rp[0]:rp[1]].sum() == 0


def test_strength_monotone_in_theta_and_subset_of_pattern():
    s = poisson(7, contrast=3.0)
    rp = s.rowptr32()
    prev = None
    for th in (0.05, 0.25, 0.5, 0.75, 0.95):
        mk = orc.strength(rp, s.col, s.val, th)
        diag = s.col == np.repeat(np.arange(s.n), np.diff(rp))
        assert not mk[diag].any()
        if prev is not None:
            assert (mk <= prev).all()  # S(theta2) subset of S(theta1)
        prev = mk


def test_pmis_splitting_properties():
    s = poisson(8)
    rp = s.rowptr32()
    mk = orc.strength(rp, s.col, s.val, 0.25)
    cf = orc.coarsen(rp, s.col, mk, "pmis")
    assert set(np.unique(cf)) <= {1, -1, -3}
    S = sp.csr_matrix((mk.astype(np.float64), s.col, rp), shape=(s.n, s.n))
    S.eliminate_zeros()
    isC = cf > 0
    # independent set: no two C points strongly connected
    assert (S[isC][:, isC]).nnz == 0
    # every F point (-1) with strong connections that somebody depends on has a strong C neighbour
    strongC = np.asarray(S[:, isC].sum(axis=1)).ravel()
    influence = np.asarray(S.sum(axis=0)).ravel()
    f = (cf == -1) & (influence >= 1)
    assert (strongC[f] > 0).all()
    # special F points are exactly the rows without strong connections
    assert np.array_equal(cf == -3, np.diff(S.indptr) == 0)


def _setup(s, theta=0.25, **kw):
    return orc.Hierarchy(s.rowptr32(), s.col, s.val, device_data(theta, **kw).to_struct())


def test_galerkin_product_equals_scipy_triple_product():
    s = poisson(8, contrast=2.0)
    H = _setup(s)
    A = s.to_scipy()
    for l in range(H.num_levels - 1):
        rp, cl, vl, nc = H.P(l)
        n = len(rp) - 1
        P = sp.csr_matrix((vl, cl, rp), shape=(n, nc))
        rpa, cla, vla = H.A(l + 1)
        Ac = sp.csr_matrix((vla, cla, rpa), shape=(nc, nc))
        ref = (P.T @ A @ P).tocsr()
        assert abs(Ac - ref).max() <= 1e-12 * abs(ref).max()
        # columns sorted, diagonal present
        assert all(np.all(np.diff(cla[rpa[i]:rpa[i + 1]]) > 0) for i in range(nc))
        assert (Ac.diagonal() != 0).all()
        A = Ac


def test_interpolation_properties():
    s = ab.gen.poisson_q1(8)
    H = _setup(s)
    rp, cl, vl, nc = H.P(0)
    cf = H.cf_marker(0)
    P = sp.csr_matrix((vl, cl, rp), shape=(s.n, nc))
    # C rows are unit vectors, numbered by ascending fine index
    cidx = np.flatnonzero(cf > 0)
    assert np.array_equal(P[cidx].indices, np.arange(nc)) and (P[cidx].data == 1).all()
    # interior F rows away from the boundary reproduce constants (zero row sum rows)
    A = s.to_scipy()
    rowsum = np.asarray(A.sum(axis=1)).ravel()
    touches_boundary = np.asarray((A != 0) @ (np.diff(s.rowptr) < 27).astype(float)).ravel() > 0
    good = (cf == -1) & (abs(rowsum) < 1e-12) & ~touches_boundary & (np.diff(rp) > 0)
    assert good.sum() > 0
    assert np.allclose(np.asarray(P.sum(axis=1)).ravel()[good], 1.0, atol=1e-12)
    # special F points have empty rows
    assert (np.diff(rp)[cf == -3] == 0).all()


def test_level_stats_definitions():
    s = ab.gen.poisson_q1(10)
    H = _setup(s)
    st = H.stats()
    assert st["rows"][0] == s.n and st["nnz"][0] == s.nnz
    assert st["rows"][-1] <= 9 or H.num_levels == 25
    nnzp = sum(H.level_dims(l)[3] for l in range(H.num_levels))
    assert st["grid"] == pytest.approx(st["rows"].sum() / s.n)
    assert st["operator"] == pytest.approx(st["nnz"].sum() / s.nnz)
    assert st["memory"] == pytest.approx((st["nnz"].sum() + nnzp) / s.nnz)
    assert np.allclose(st["sparsity"], st["nnz"] / st["rows"].astype(float) ** 2)


def test_vcycle_is_a_symmetric_contraction_on_the_interior():
    s = ab.gen.poisson_q1(8)
    H = _setup(s)
    A = s.to_scipy()
    interior = np.diff(s.rowptr) == 27
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal(s.n) * interior, rng.standard_normal(s.n) * interior
    Mx, My = H.vmult(x), H.vmult(y)
    # symmetric smoother pairing (C-F down, F-C up) => M symmetric on the interior block
    assert abs(y @ Mx - x @ My) <= 1e-10 * abs(y @ Mx)
    # error propagation I - M A contracts in the A-norm
    e = x.copy()
    Aii = A[interior][:, interior]
    def anorm(v):
        return np.sqrt(v[interior] @ (Aii @ v[interior]))
    e1 = e - H.vmult((A @ e) * interior) * interior
    assert anorm(e1) < 0.6 * anorm(e)


@pytest.mark.parametrize("theta", [0.25, 0.5, 0.8])
def test_pcg_converges_to_the_direct_solution(theta):
    s = poisson(10, contrast=4.0)
    H = _setup(s, theta)
    rc, x, nit, hist = H.cg_solve(s.rhs, s.x0, abs_tol=1e-8)
    assert rc == 0 and 1 <= nit < 60 and len(hist) == nit + 1
    assert hist[-1] <= 1e-8 < hist[-2]
    xs = spl.spsolve(s.to_scipy().tocsc(), s.rhs)
    assert abs(x - xs).max() <= 1e-6 * abs(xs).max()


def test_pcg_absolute_tolerance_and_no_convergence():
    s = ab.gen.poisson_q1(6)
    H = _setup(s)
    rc, _, nit, hist = H.cg_solve(s.rhs, s.x0, max_steps=2, abs_tol=1e-30)
    assert rc == -6 and nit == 2 and len(hist) == 3
    # tol is absolute (A.4): a huge tolerance stops at iteration 0
    rc, x, nit, hist = H.cg_solve(s.rhs, s.x0, abs_tol=1e30)
    assert rc == 0 and nit == 0 and np.array_equal(x, s.x0)


def test_reference_flavour_falgout_symgs_runs():
    # the reference's own configuration: Falgout + hybrid symmetric GS (A.2)
    s = ab.gen.poisson_q1(8)
    d = ab.AdditionalData(True, 0.5, 0.9, 0, True, coarsen_type=ab.COARSEN_FALGOUT)
    H = orc.Hierarchy(s.rowptr32(), s.col, s.val, d.to_struct())
    assert H.effective_relax() == (6, 6, 9)
    rc, x, nit, _ = H.cg_solve(s.rhs, s.x0)
    assert rc == 0 and nit < 25
    cf = H.cf_marker(0)
    mk = H.strength_mask(0)
    S = sp.csr_matrix((mk.astype(float), s.col, s.rowptr32()), shape=(s.n, s.n))
    S.eliminate_zeros()
    strongC = np.asarray(S[:, cf > 0].sum(axis=1)).ravel()
    assert (strongC[cf == -1] > 0).all()


def test_unstructured_matrix_with_isolated_and_ragged_rows():
    A = random_spd_csr(300, 0.03, 1)
    H = orc.Hierarchy(A.indptr, A.indices, A.data, device_data(0.25).to_struct())
    b = np.ones(300)
    rc, x, nit, _ = H.cg_solve(b, np.zeros(300))
    assert rc == 0
    assert np.linalg.norm(A @ x - b) < 1e-6


# ------------------------------------------------------------------ pooling
def _view_numpy(A, V):
    """Independent numpy transcription of ref common/view_maker.h:41-65."""
    n = A.shape[0]
    q, p = divmod(n, V)
    q1, t = q + 1, (q + 1) * p

    def b(i):
        i = np.asarray(i)
        out = np.empty_like(i)
        lo = i < t
        out[lo] = i[lo] // q1
        if q:
            out[~lo] = (i[~lo] - t) // q + p
        return out
    rows = np.repeat(np.arange(n), np.diff(A.indptr))
    flat = V * b(rows) + b(A.indices)
    s = np.zeros(V * V)
    cnt = np.zeros(V * V, dtype=np.int64)
    pp = np.zeros(V * V)
    np_ = np.zeros(V * V)
    np.add.at(s, flat, A.data)
    np.add.at(cnt, flat, 1)
    np.maximum.at(pp, flat, np.maximum(A.data, 0))
    np.maximum.at(np_, flat, np.maximum(-A.data, 0))
    return s, cnt, pp, np_


@pytest.mark.parametrize("m,V", [(6, 5), (6, 75), (9, 50), (3, 100)])
def test_pooling_oracle_matches_numpy_transcription(m, V):
    s = poisson(m, contrast=2.0)
    A = s.to_scipy()
    got = orc.make_view(s.rowptr32(), s.col, s.val, V)
    ref = _view_numpy(A, V)
    assert np.array_equal(got[1], ref[1]) and got[1].sum() == s.nnz  # stored zeros count
    assert np.array_equal(got[2], ref[2]) and np.array_equal(got[3], ref[3])
    assert np.allclose(got[0], ref[0], rtol=0, atol=1e-12 * abs(A).sum())
    assert got[0].sum() == pytest.approx(A.sum(), abs=1e-10 * abs(A).sum())


# ------------------------------------------------------------------ C ABI
def test_library_exports_every_symbol_the_header_declares():
    hdr = open(os.path.join(ROOT, "include", "amgb.h")).read()
    declared = sorted(set(re.findall(r"\b(amgb_[A-Za-z0-9_]+)\s*\(", hdr)))
    lib = C.CDLL(os.path.join(ROOT, "amg-ann_b200", "libamgb.so"))
    missing = [f for f in declared if not hasattr(lib, f)]
    assert not missing, missing
    assert sorted(ab.AMGB_SYMBOLS) == declared


def test_generator_library_exports_header_symbols():
    hdr = open(os.path.join(ROOT, "include", "amgb_gen.h")).read()
    declared = set(re.findall(r"\b(amgb_gen_[a-z0-9_]+)\s*\(", hdr))
    lib = C.CDLL(os.path.join(ROOT, "amg-ann_b200", "libamgb_gen.so"))
    assert all(hasattr(lib, f) for f in declared) and len(declared) >= 6


def test_defaults_struct_matches_python_mirror():
    from amg_ann_b200._native import BoomerAMGDataStruct, amgb_lib
    s = BoomerAMGDataStruct()
    assert amgb_lib().amgb_boomeramg_data_default(C.byref(s)) == 0
    py = ab.AdditionalData().to_struct()
    for name, _ in BoomerAMGDataStruct._fields_:
        if name != "reserved":
            assert getattr(s, name) == getattr(py, name), name
    # deal.II defaults (A.1)
    assert (s.strong_threshold, s.max_row_sum, s.max_iter) == (0.25, 0.9, 1)
    assert s.max_levels == 25 and s.max_coarse_size == 9 and s.relax_order == 1


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "amg-ann_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle/" not in txt.replace("oracle/amg_oracle.cpp", "") or f.endswith((".cu", ".cuh")), f
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_aggressive_coarsening_oracle_properties():
    """Second-stage PMIS + multipass interpolation (oracle): far fewer coarse points, rows of
    P reproduce constants where the operator has zero row sum, PCG still converges."""
    import amg_ann_b200 as ab
    from oracle import binding as orc
    from helpers import device_data, poisson
    s = poisson(14, contrast=2.0)
    d0, d2 = device_data(0.25), device_data(0.25)
    d2.aggressive_coarsening_num_levels = 2
    H0 = orc.Hierarchy(s.rowptr32(), s.col, s.val, d0.to_struct())
    H2 = orc.Hierarchy(s.rowptr32(), s.col, s.val, d2.to_struct())
    r0, r2 = H0.stats()["rows"], H2.stats()["rows"]
    assert r2[1] * 4 < r0[1] and H2.stats()["operator"] < H0.stats()["operator"]
    cf0, cf2 = H0.cf_marker(0), H2.cf_marker(0)
    assert set(np.flatnonzero(cf2 > 0)) <= set(np.flatnonzero(cf0 > 0))   # second stage only removes C points
    rp, cl, vl, nc = H2.P(0)
    rows = np.flatnonzero(np.diff(rp) > 0)
    sums = np.add.reduceat(vl, rp[rows])
    A = s.to_scipy()
    zero_sum = np.abs(np.asarray(A.sum(axis=1)).ravel()[rows]) < 1e-9 * np.abs(A.diagonal()[rows])
    assert zero_sum.any() and np.allclose(sums[zero_sum], 1.0, atol=1e-12)
    rc, x, nit, hist = H2.cg_solve(s.rhs, s.x0, abs_tol=1e-8)
    assert rc == 0 and hist[-1] <= 1e-8
    # Falgout + aggressive levels is not restated
    d = device_data(0.25, coarsen_type=ab.COARSEN_FALGOUT)
    d.aggressive_coarsening_num_levels = 1
    with pytest.raises(Exception):
        orc.Hierarchy(s.rowptr32(), s.col, s.val, d.to_struct())


# ---- Chebyshev smoother (hypre relax type 16) ----
def test_tql1_matches_lapack_on_random_tridiagonals():
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 7, 10, 25):
        d, e = rng.normal(size=n), rng.normal(size=n)
        T = np.diag(d) + np.diag(e[1:], 1) + np.diag(e[1:], -1)
        assert np.allclose(orc.tql1(d, e), np.linalg.eigvalsh(T), rtol=0, atol=1e-13 * max(1.0, abs(T).max()))


def test_chebyshev_spectrum_estimates_bracket_like_lanczos():
    """10 CG/Lanczos steps: Ritz values lie inside the spectrum of D^-1/2 A D^-1/2 and the
    largest one is close to its top."""
    from helpers import spd_laplacian
    s = spd_laplacian(10, seed=1)
    R = ab.RelaxationType
    d = device_data(0.25, relaxation_type_up=R.Chebyshev, relaxation_type_down=R.Chebyshev)
    H = orc.Hierarchy(s.rowptr32(), s.col, s.val, d.to_struct())
    assert H.effective_relax()[:2] == (16, 16)
    ds = 1.0 / np.sqrt(s.csr.diagonal())
    ev = np.linalg.eigvalsh((sp.diags(ds) @ s.csr @ sp.diags(ds)).toarray())
    mx, mn, coefs = H.level_cheby(0)
    assert ev[0] - 1e-12 <= mn <= mx <= ev[-1] + 1e-12
    assert mx > 0.9 * ev[-1]
    # order 2: p(t) = c0 + c1 t with the residual polynomial 1 - t p(t) small on [lower, upper]
    assert len(coefs) == 2 and coefs[0] > 0 > coefs[1]
    upper = 1.1 * mx
    lower = (upper - mn) * 0.3 + mn
    t = np.linspace(lower, upper, 101)
    assert np.abs(1 - t * (coefs[0] + coefs[1] * t)).max() < 0.5


def test_chebyshev_smoothed_pcg_converges_to_the_direct_solution():
    from helpers import spd_laplacian
    s = spd_laplacian(12, seed=2)
    R = ab.RelaxationType
    for kw in (dict(), dict(w_cycle=True), dict(n_sweeps=2)):
        d = device_data(0.25, relaxation_type_up=R.Chebyshev, relaxation_type_down=R.Chebyshev, **kw)
        H = orc.Hierarchy(s.rowptr32(), s.col, s.val, d.to_struct())
        rc, x, nit, hist = H.cg_solve(s.rhs, s.x0, abs_tol=1e-10)
        assert rc == 0 and nit < 60
        xd = spl.spsolve(s.csr.tocsc(), s.rhs)
        assert np.abs(x - xd).max() <= 1e-6 * np.abs(xd).max()


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU restatement timed on the host cores) on a tiny
    sample: one JSON line with the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-m", "8", "--cells", "12"], capture_output=True, text=True,
                         timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "s/system" and line["higher_is_better"] is False
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"]
    assert line["metric"] == "AMG-PCG setup+solve seconds per system" and "workload" in line["config"]


def test_chebyshev_coefficients_are_the_scaled_chebyshev_polynomial():
    """The coefficients restated from hypre's par_cheby.c are checked against the definition:
    the residual polynomial 1 - t p(t) of the order-2 smoother is T_2((theta - t)/delta) /
    T_2(theta/delta) on [lower, upper] = [(1.1 max - min) 0.3 + min, 1.1 max]."""
    from helpers import spd_laplacian
    s = spd_laplacian(8, seed=7)
    R = ab.RelaxationType
    d = device_data(0.25, relaxation_type_up=R.Chebyshev, relaxation_type_down=R.Chebyshev)
    H = orc.Hierarchy(s.rowptr32(), s.col, s.val, d.to_struct())
    for l in range(H.num_levels):
        mx, mn, c = H.level_cheby(l)
        upper = 1.1 * mx
        lower = (upper - mn) * 0.3 + mn
        theta, delta = (upper + lower) / 2, (upper - lower) / 2
        t = np.linspace(0.0, upper, 201)
        T2 = lambda x: 2 * x * x - 1          # noqa: E731
        want = T2((theta - t) / delta) / T2(theta / delta)
        got = 1 - t * (c[0] + c[1] * t)
        assert np.allclose(got, want, rtol=0, atol=1e-12)
        assert np.abs(got[t >= lower]).max() <= 1.0 / T2(theta / delta) + 1e-12   # equioscillation bound


@pytest.mark.parametrize("m,theta,contrast", [(8, 0.25, 0.0), (10, 0.5, 4.0), (12, 0.25, 6.0)])
def test_cljp_splitting_properties(m, theta, contrast):
    """CLJP (coarsen type 0, the parallel stage of the Falgout family) in the oracle: every point is
    decided; rows without strong connections are special F points; every ordinary F point keeps
    a strong C neighbour to interpolate from; the coarse grid is denser than the PMIS one."""
    import amg_ann_b200 as ab
    from oracle import binding as orc
    epsv = ab.gen.checkerboard_epsv(2, 3, contrast) if contrast else None
    s = ab.gen.poisson_q1(m, 2, 3, epsv) if contrast else ab.gen.poisson_q1(m)
    rp = s.rowptr32()
    mask = orc.strength(rp, s.col, s.val, theta)
    cf = orc.coarsen(rp, s.col, mask, "cljp")
    assert set(np.unique(cf)) <= {1, -1, -3}
    strong_rows = np.add.reduceat(mask.astype(np.int64), rp[:-1]) > 0
    assert ((cf == -3) == ~strong_rows).all()
    is_c = cf > 0
    has_c = np.add.reduceat((mask.astype(bool) & is_c[s.col]).astype(np.int64), rp[:-1]) > 0
    assert has_c[(cf == -1)].all()          # every ordinary F point interpolates from at least one C point
    pm = orc.coarsen(rp, s.col, mask, "pmis")
    assert (cf > 0).sum() > (pm > 0).sum()  # denser coarse grids than PMIS: that is what buys the convergence


def test_cuthill_mckee_is_a_level_ordering():
    """include/amgb_gen.h amgb_gen_cuthill_mckee on a path graph and on a 2-component graph."""
    import amg_ann_b200 as ab
    import scipy.sparse as sp
    n = 9
    P = sp.diags([np.ones(n - 1), np.ones(n), np.ones(n - 1)], [-1, 0, 1]).tocsr()
    perm = ab.gen.cuthill_mckee(P.indptr, P.indices)
    assert list(perm) == list(range(n))               # starts at an end point (degree 2), walks the path
    B = sp.block_diag([P, P]).tocsr()
    perm = ab.gen.cuthill_mckee(B.indptr, B.indices)
    assert sorted(perm) == list(range(2 * n)) and list(perm[:n]) == list(range(n))
    assert list(ab.gen.cuthill_mckee(B.indptr, B.indices, reversed=True)) == list(perm[::-1])


def test_oracle_multicolour_gauss_seidel():
    """smoother_policy = SMOOTHER_MULTICOLOR in the oracle: a valid greedy colouring on every level,
    the multicolour sweep is the Gauss-Seidel sweep of the colour-ordered system (checked against a
    dense triangular solve), and PCG needs fewer iterations than with the l1-Jacobi substitute."""
    import amg_ann_b200 as ab
    from oracle import binding as orc
    from helpers import device_data, poisson
    R = ab.RelaxationType
    s = poisson(10, contrast=3.0)
    data = device_data(0.25, relaxation_type_up=R.symmetricSORJacobi, relaxation_type_down=R.symmetricSORJacobi,
                       smoother_policy=ab.SMOOTHER_MULTICOLOR, max_levels=1, relaxation_type_coarse=R.SORJacobi,
                       n_sweeps_coarse=1)
    H = orc.Hierarchy(s.rowptr32(), s.col, s.val, data.to_struct())
    assert H.effective_relax() == (106, 106, 106)
    col, nc = H.colors(0)
    rp = s.rowptr32()
    rows = np.repeat(np.arange(s.n), np.diff(rp))
    off = rows != s.col
    assert col.min() == 0 and col.max() == nc - 1 and not (col[rows[off]] == col[s.col[off]]).any()
    # one level, one symmetric sweep from zero = (forward then backward Gauss-Seidel) in colour order
    import scipy.sparse as sp
    A = sp.csr_matrix((s.val, s.col, rp), shape=(s.n, s.n)).toarray()
    order = np.lexsort((np.arange(s.n), col))
    Ap = A[np.ix_(order, order)]
    r = np.random.default_rng(0).standard_normal(s.n)
    f = r[order]
    Lo, Up, D = np.tril(Ap, -1), np.triu(Ap, 1), np.diag(np.diag(Ap))
    u1 = np.linalg.solve(D + Lo, f)                      # forward from zero
    u2 = np.linalg.solve(D + Up, f - Lo @ u1)            # backward
    z = H.vmult(r)
    assert np.abs(z[order] - u2).max() <= 1e-11 * np.abs(u2).max()
    # full hierarchy: fewer iterations than C/F l1-Jacobi
    s = poisson(14, contrast=3.0)
    d_mc = device_data(0.25, relaxation_type_up=R.symmetricSORJacobi, relaxation_type_down=R.symmetricSORJacobi,
                       smoother_policy=ab.SMOOTHER_MULTICOLOR)
    it_mc = orc.Hierarchy(s.rowptr32(), s.col, s.val, d_mc.to_struct()).cg_solve(s.rhs, s.x0, abs_tol=1e-8)[2]
    it_j = orc.Hierarchy(s.rowptr32(), s.col, s.val, device_data(0.25).to_struct()).cg_solve(s.rhs, s.x0, abs_tol=1e-8)[2]
    assert it_mc < it_j
