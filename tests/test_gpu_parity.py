"""GPU parity suite (-m gpu): the CUDA path, called through the C ABI, against
the CPU oracle on the same seeded inputs.  Integer outputs bit-exact, operators
bit-identical, residual histories to 1e-10 relative, iterations within +-1
(BASELINE.json north_star)."""
import numpy as np
import pytest

import amg_ann_b200 as ab
from helpers import device_data, poisson, random_spd_csr
from oracle import binding as orc

pytestmark = pytest.mark.gpu

RES_RTOL = 1e-10  # north_star: residual history agrees to 1e-10 relative


def _both(ctx, s, data):
    A = ab.SparseMatrix(ctx, s.rowptr32(), s.col, s.val)
    P = ab.PreconditionBoomerAMG()
    P.initialize(A, data)
    H = orc.Hierarchy(s.rowptr32(), s.col, s.val, data.to_struct())
    return A, P, H


def _assert_hierarchy_identical(P, H):
    assert P.num_levels == H.num_levels
    for l in range(H.num_levels):
        assert P.level_dims(l) == H.level_dims(l), l
        rp, cl, vl = P.A(l)
        rpo, clo, vlo = H.A(l)
        assert np.array_equal(rp, rpo) and np.array_equal(cl, clo), f"A pattern level {l}"
        assert np.array_equal(vl, vlo), f"A values level {l}: max diff {abs(vl - vlo).max()}"
        if l + 1 < H.num_levels:
            assert np.array_equal(P.strength_mask(l), H.strength_mask(l)), f"mask level {l}"
            assert np.array_equal(P.cf_marker(l), H.cf_marker(l)), f"cf level {l}"
            prp, pcl, pvl, nc = P.P(l)
            orp, ocl, ovl, onc = H.P(l)
            assert nc == onc and np.array_equal(prp, orp) and np.array_equal(pcl, ocl), f"P pattern {l}"
            assert np.array_equal(pvl, ovl), f"P values level {l}"


def test_spmv_matches_oracle(gpu_ctx):
    s = poisson(12, contrast=3.0)
    A = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
    x = np.random.default_rng(0).standard_normal(s.n)
    y = A.vmult(x)
    yo = orc.spmv(s.rowptr32(), s.col, s.val, x)
    assert np.abs(y - yo).max() <= 1e-13 * np.abs(yo).max()
    # 64-bit row pointers take the same path
    A64 = ab.SparseMatrix(gpu_ctx, s.rowptr, s.col, s.val)
    assert np.array_equal(A64.vmult(x), y)


@pytest.mark.parametrize("m,theta,contrast", [(8, 0.25, 0.0), (12, 0.5, 0.0), (12, 0.25, 6.0),
                                              (16, 0.7, 3.0), (10, 0.05, 0.0), (10, 0.95, 6.0)])
def test_setup_is_bit_exact(gpu_ctx, m, theta, contrast):
    s = poisson(m, contrast=contrast)
    A, P, H = _both(gpu_ctx, s, device_data(theta))
    _assert_hierarchy_identical(P, H)
    st, so = P.level_stats(), H.stats()
    assert np.array_equal(st["rows"], so["rows"]) and np.array_equal(st["nnz"], so["nnz"])
    assert (st["grid"], st["operator"], st["memory"]) == (so["grid"], so["operator"], so["memory"])


def test_setup_theta_at_exact_tie(gpu_ctx):
    # H3: corner/edge couplings of the uniform stencil are in ratio 0.5 exactly
    s = ab.gen.poisson_q1(10)
    A, P, H = _both(gpu_ctx, s, device_data(0.5))
    _assert_hierarchy_identical(P, H)


def test_setup_unstructured_ragged_rows(gpu_ctx):
    M = random_spd_csr(2000, 0.004, 3)
    class S:  # minimal System look-alike
        n = M.shape[0]; col = M.indices; val = M.data
        def rowptr32(self): return M.indptr.astype(np.int32)
    A, P, H = _both(gpu_ctx, S(), device_data(0.25))
    _assert_hierarchy_identical(P, H)


def test_setup_elasticity(gpu_ctx):
    young = 10.0 ** ab.gen.checkerboard_epsv(2, 3, 2.0)
    s = ab.gen.elasticity_q1(6, 2, 3, young)
    A, P, H = _both(gpu_ctx, s, device_data(0.5))
    _assert_hierarchy_identical(P, H)


def test_vmult_matches_oracle(gpu_ctx):
    s = poisson(12, contrast=2.0)
    A, P, H = _both(gpu_ctx, s, device_data(0.25))
    r = np.random.default_rng(1).standard_normal(s.n)
    z = np.empty(s.n)
    P.vmult(z, r)
    zo = H.vmult(r)
    assert np.abs(z - zo).max() <= 1e-12 * np.abs(zo).max()


@pytest.mark.parametrize("m,theta,contrast", [(12, 0.25, 0.0), (16, 0.5, 6.0), (20, 0.25, 3.0)])
def test_pcg_residual_history_and_iterations(gpu_ctx, m, theta, contrast):
    s = poisson(m, contrast=contrast)
    A, P, H = _both(gpu_ctx, s, device_data(theta))
    ctl = ab.SolverControl(s.n, 1e-8)
    x = s.x0.copy()
    ab.SolverCG(ctl).solve(A, x, s.rhs, P)
    rc, xo, nit, hist = H.cg_solve(s.rhs, s.x0, abs_tol=1e-8)
    assert rc == 0 and abs(ctl.last_step() - nit) <= 1
    k = min(len(hist), len(ctl.history))
    assert (np.abs(ctl.history[:k] - hist[:k]) <= RES_RTOL * hist[:k]).all()
    assert np.abs(x - xo).max() <= 1e-9 * np.abs(xo).max()
    assert ctl.last_value() <= 1e-8


def test_jacobi_smoother_and_no_cf_relaxation(gpu_ctx):
    s = ab.gen.poisson_q1(10)
    R = ab.RelaxationType
    data = device_data(0.25, relaxation_type_up=R.Jacobi, relaxation_type_down=R.Jacobi,
                       relax_weight=0.6, relax_order=0)
    A, P, H = _both(gpu_ctx, s, data)
    assert P.effective_relax() == H.effective_relax() == (0, 0, 9)
    ctl = ab.SolverControl(s.n, 1e-8)
    x = s.x0.copy()
    ab.SolverCG(ctl).solve(A, x, s.rhs, P)
    rc, xo, nit, hist = H.cg_solve(s.rhs, s.x0, abs_tol=1e-8)
    assert abs(ctl.last_step() - nit) <= 1
    k = min(len(hist), len(ctl.history))
    assert (np.abs(ctl.history[:k] - hist[:k]) <= RES_RTOL * hist[:k]).all()


def test_reference_default_smoother_is_substituted_or_rejected(gpu_ctx):
    s = ab.gen.poisson_q1(8)
    A = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
    P = ab.PreconditionBoomerAMG()
    # the reference's 5-argument construction (t2 main.cpp:447-453)
    P.initialize(A, ab.AdditionalData(True, 0.25, 0.9, 0, True))
    assert P.effective_relax() == (18, 18, 9)
    with pytest.raises(ab.AmgbError) as e:
        P.initialize(A, ab.AdditionalData(True, 0.25, 0.9, 0, True, smoother_policy=ab.SMOOTHER_STRICT))
    assert e.value.status == -5
    # aggressive levels need the parallel coarsening; with the reference's serial Falgout they are refused
    with pytest.raises(ab.AmgbError):
        P.initialize(A, ab.AdditionalData(True, 0.25, 0.9, 2, True, coarsen_type=ab.COARSEN_FALGOUT))


def test_no_convergence_is_reported(gpu_ctx):
    s = ab.gen.poisson_q1(8)
    A, P, H = _both(gpu_ctx, s, device_data(0.25))
    ctl = ab.SolverControl(2, 1e-30)
    x = s.x0.copy()
    with pytest.raises(ab.NoConvergence):
        ab.SolverCG(ctl).solve(A, x, s.rhs, P)
    assert ctl.last_step() == 2 and len(ctl.history) == 3
    ctl = ab.SolverControl(10, 1e30)  # absolute tolerance: stops before the first step
    ab.SolverCG(ctl).solve(A, x, s.rhs, P)
    assert ctl.last_step() == 0


@pytest.mark.parametrize("m,V,contrast", [(6, 5, 0.0), (10, 75, 2.0), (14, 50, 6.0), (3, 100, 0.0)])
def test_pooling_parity(gpu_ctx, m, V, contrast):
    s = poisson(m, contrast=contrast)
    A = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
    vm = ab.ViewMaker(V).make_view(A)
    so, co, ppo, npo = orc.make_view(s.rowptr32(), s.col, s.val, V)
    assert np.array_equal(vm.count, co) and vm.count.sum() == s.nnz
    assert np.array_equal(vm.max_pp, ppo) and np.array_equal(vm.max_np, npo)
    assert np.allclose(vm.view, so, rtol=0, atol=1e-13 * np.abs(s.val).sum())


def test_theta_sweep_reuses_resident_matrix(gpu_ctx):
    # ref t2 main.cpp:440-467: independent solves sharing one matrix
    s = poisson(10, contrast=3.0)
    A = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
    for th in ab.gen.theta_sweep(0.05, 0.96, 0.15):
        x = s.x0.copy()
        row = ab.amg_solve(device_data(th), 1e-8, A, s.rhs, x)
        H = orc.Hierarchy(s.rowptr32(), s.col, s.val, device_data(th).to_struct())
        rc, xo, nit, hist = H.cg_solve(s.rhs, s.x0, abs_tol=1e-8)
        assert abs(row["niters"] - nit) <= 1
        assert list(row["nrows"]) == list(H.stats()["rows"])
        assert list(row["nze"]) == list(H.stats()["nnz"])


def test_full_size_properties_config1(gpu_ctx):
    """BASELINE config 1 (m=100, 1.03M DoFs): size-independent properties."""
    s = ab.gen.poisson_q1(100)
    A = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
    P = ab.PreconditionBoomerAMG()
    P.initialize(A, device_data(0.25))
    st = P.level_stats()
    assert st["rows"][0] == 1030301 and st["nnz"][0] == 27270901
    assert (np.diff(st["rows"]) < 0).all() and st["rows"][-1] <= 9
    # level-0 mask: strong count per interior row = 20 (edges + corners), H3
    mk = P.strength_mask(0)
    rp = s.rowptr32()
    N = 101
    c = (N // 2) * (1 + N + N * N)
    assert mk[rp[c]:rp[c + 1]].sum() == 20
    cf = P.cf_marker(0)
    assert (cf > 0).sum() == st["rows"][1]
    # coarse operator symmetric away from Dirichlet couplings: A_c row sums bounded
    ctl = ab.SolverControl(s.n, 1e-8)
    x = s.x0.copy()
    ab.SolverCG(ctl).solve(A, x, s.rhs, P)
    assert ctl.last_step() < 60 and ctl.history[-1] <= 1e-8
    # true residual of the returned solution
    r = s.rhs - A.vmult(x)
    assert np.linalg.norm(r) <= 1e-6 * np.linalg.norm(s.rhs)
    # pooled image: count sums to nnz, sum channel sums to A.sum()
    vm = ab.ViewMaker(75).make_view(A)
    assert vm.count.sum() == s.nnz
    assert vm.view.sum() == pytest.approx(s.val.sum(), abs=1e-9 * np.abs(s.val).sum())
    assert vm.max_pp.max() == s.val.max() and vm.max_np.max() == -s.val.min()


@pytest.mark.parametrize("m,max_levels", [(12, 2), (16, 3), (12, 1), (6, 2)])
def test_truncated_hierarchy_dense_or_relaxed_coarsest(gpu_ctx, m, max_levels):
    """max_levels cuts the hierarchy: the coarsest grid is then large.  <= 1024 rows:
    Gaussian elimination (grid-parallel factorisation + blocked substitution must
    reproduce the sequential hypre_gselim order); larger: smoother sweeps."""
    s = poisson(m, contrast=3.0)
    A, P, H = _both(gpu_ctx, s, device_data(0.25, max_levels=max_levels))
    assert P.num_levels == H.num_levels == max_levels
    r = np.random.default_rng(5).standard_normal(s.n)
    z = np.empty(s.n)
    P.vmult(z, r)
    zo = H.vmult(r)
    assert np.abs(z - zo).max() <= 1e-11 * np.abs(zo).max()
    ctl = ab.SolverControl(400, 1e-8)
    x = s.x0.copy()
    try:
        ab.SolverCG(ctl).solve(A, x, s.rhs, P)
    except ab.NoConvergence:
        pass
    rc, xo, nit, hist = H.cg_solve(s.rhs, s.x0, max_steps=400, abs_tol=1e-8)
    assert abs(ctl.last_step() - nit) <= 1
    k = min(len(hist), len(ctl.history), 60)
    assert (np.abs(ctl.history[:k] - hist[:k]) <= 1e-9 * hist[:k]).all()


@pytest.mark.parametrize("kw", [dict(w_cycle=True), dict(n_sweeps=2), dict(max_iter=2),
                                dict(relaxation_type_coarse=ab.RelaxationType.l1scaledJacobi, n_sweeps_coarse=3),
                                dict(relaxation_type_coarse=ab.RelaxationType.Jacobi, n_sweeps_coarse=2, relax_weight=0.8),
                                dict(max_coarse_size=60), dict(max_row_sum=1.0)])
def test_cycle_options_of_additional_data(gpu_ctx, kw):
    """The remaining deal.II AdditionalData / PCHYPRE knobs: W-cycle,
    several sweeps, several cycles per application, relaxed coarsest grid, larger coarsest
    grid, no dependency weakening."""
    s = poisson(12, contrast=3.0)
    mrs = kw.pop("max_row_sum", 0.9)
    data = device_data(0.25, **kw)
    data.max_row_sum = mrs
    A, P, H = _both(gpu_ctx, s, data)
    _assert_hierarchy_identical(P, H)
    r = np.random.default_rng(7).standard_normal(s.n)
    z = np.empty(s.n)
    P.vmult(z, r)
    zo = H.vmult(r)
    assert np.abs(z - zo).max() <= 1e-11 * np.abs(zo).max()
    ctl = ab.SolverControl(s.n, 1e-8)
    x = s.x0.copy()
    ab.SolverCG(ctl).solve(A, x, s.rhs, P)
    rc, xo, nit, hist = H.cg_solve(s.rhs, s.x0, abs_tol=1e-8)
    assert rc == 0 and abs(ctl.last_step() - nit) <= 1
    k = min(len(hist), len(ctl.history))
    assert (np.abs(ctl.history[:k] - hist[:k]) <= RES_RTOL * hist[:k]).all()


def test_bad_arguments_are_reported_not_crashed(gpu_ctx):
    s = poisson(6)
    A = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
    P = ab.PreconditionBoomerAMG()
    with pytest.raises(ab.AmgbError) as e:
        P.initialize(A, device_data(0.25, coarsen_type=ab.COARSEN_FALGOUT))   # serial algorithm: not on device
    assert e.value.status == -5
    with pytest.raises(ab.AmgbError):
        P.initialize(A, device_data(0.25, interp_type=6))
    with pytest.raises(ab.AmgbError):
        P.initialize(A, device_data(0.25, max_levels=0))
    with pytest.raises(ab.AmgbError):
        ab.ViewMaker(100000).make_view(A)
    # a preconditioner built for another matrix is refused by cg.solve
    s2 = poisson(5)
    A2 = ab.SparseMatrix(gpu_ctx, s2.rowptr32(), s2.col, s2.val)
    P.initialize(A2, device_data(0.25))
    with pytest.raises(ab.AmgbError):
        ab.SolverCG(ab.SolverControl(10, 1e-8)).solve(A, s.x0.copy(), s.rhs, P)


@pytest.mark.parametrize("m,theta,contrast,agg", [(12, 0.25, 0.0, 1), (16, 0.25, 3.0, 2), (14, 0.5, 6.0, 2),
                                                  (20, 0.25, 0.0, 2), (10, 0.7, 0.0, 3)])
def test_aggressive_coarsening_and_multipass_interpolation(gpu_ctx, m, theta, contrast, agg):
    """aggressive_coarsening_num_levels > 0 (ref t3 main.cpp:456 passes 2): second PMIS on
    the distance-two strength graph + multipass interpolation, bit-exact vs the oracle."""
    s = poisson(m, contrast=contrast)
    data = device_data(theta)
    data.aggressive_coarsening_num_levels = agg
    A, P, H = _both(gpu_ctx, s, data)
    _assert_hierarchy_identical(P, H)
    std = ab.PreconditionBoomerAMG()
    std.initialize(A, device_data(theta))
    assert P.level_stats()["operator"] < std.level_stats()["operator"]   # that is the point of it
    ctl = ab.SolverControl(s.n, 1e-8)
    x = s.x0.copy()
    ab.SolverCG(ctl).solve(A, x, s.rhs, P)
    rc, xo, nit, hist = H.cg_solve(s.rhs, s.x0, abs_tol=1e-8)
    assert rc == 0 and abs(ctl.last_step() - nit) <= 1
    k = min(len(hist), len(ctl.history))
    assert (np.abs(ctl.history[:k] - hist[:k]) <= RES_RTOL * hist[:k]).all()


def test_aggressive_coarsening_elasticity_like_the_reference_case(gpu_ctx):
    # testcase 3: vector elasticity, scalar AMG, agg_nl = 2 (ref t3 main.cpp:454-464)
    young = 10.0 ** ab.gen.checkerboard_epsv(2, 3, 2.0)
    s = ab.gen.elasticity_q1(6, 2, 3, young)
    data = device_data(0.5)
    data.aggressive_coarsening_num_levels = 2
    A, P, H = _both(gpu_ctx, s, data)
    _assert_hierarchy_identical(P, H)


def test_level_row_stats_match_the_level_operators(gpu_ctx):
    s = poisson(10, contrast=2.0)
    A, P, H = _both(gpu_ctx, s, device_data(0.25))
    for l in range(P.num_levels):
        rp, cl, vl = P.A(l)
        lens = np.diff(rp)
        sums = np.add.reduceat(vl, rp[:-1])
        mn, mx, smin, smax = P.level_row_stats(l)
        assert (mn, mx) == (lens.min(), lens.max())
        assert smin == pytest.approx(sums.min(), abs=1e-12 * np.abs(vl).max())
        assert smax == pytest.approx(sums.max(), abs=1e-12 * np.abs(vl).max())


# ---- Chebyshev smoother (RelaxationType::Chebyshev, hypre relax type 16) ----
def _cheby_data(**kw):
    R = ab.RelaxationType
    return device_data(0.25, relaxation_type_up=R.Chebyshev, relaxation_type_down=R.Chebyshev, **kw)


@pytest.mark.parametrize("kw", [dict(), dict(w_cycle=True), dict(n_sweeps=2), dict(max_levels=2)])
def test_chebyshev_smoother_matches_the_oracle(gpu_ctx, kw):
    from helpers import spd_laplacian
    s = spd_laplacian(14, seed=5, decades=1.0)
    data = _cheby_data(**kw)
    A, P, H = _both(gpu_ctx, s, data)
    assert P.effective_relax()[:2] == (16, 16)
    _assert_hierarchy_identical(P, H)
    # the spectrum estimates are computed in a fixed arithmetic order: identical bits
    for l in range(P.num_levels):
        mx, mn, co = P.level_cheby(l)
        omx, omn, oco = H.level_cheby(l)
        assert (mx, mn) == (omx, omn), (l, mx - omx, mn - omn)
        assert np.array_equal(co, oco)
    ctl = ab.SolverControl(s.n, 1e-9)
    x = s.x0.copy()
    ab.SolverCG(ctl).solve(A, x, s.rhs, P)
    rc, xo, nit, hist = H.cg_solve(s.rhs, s.x0, abs_tol=1e-9)
    assert rc == 0 and abs(ctl.last_step() - nit) <= 1
    k = min(len(hist), len(ctl.history))
    d = np.abs(ctl.history[:k] - hist[:k])
    assert (d[:12] <= RES_RTOL * hist[:12]).all()   # north_star: 1e-10 relative
    # rounding differences grow with the iteration count on this badly scaled system
    assert (d <= RES_RTOL * hist[0]).all() and (d <= 1e-8 * hist[:k]).all()
    assert np.abs(x - xo).max() <= 1e-8 * np.abs(xo).max()


def test_chebyshev_on_symmetric_elasticity_and_vmult(gpu_ctx):
    """testcase-3 style system (constraints eliminated symmetrically): one V-cycle applied to
    a vector agrees with the oracle's."""
    s = ab.gen.elasticity_q1(6, 2, 3, 10.0 ** ab.gen.checkerboard_epsv(2, 3, 1.0))
    A, P, H = _both(gpu_ctx, s, _cheby_data())
    for l in range(P.num_levels):
        assert P.level_cheby(l)[:2] == H.level_cheby(l)[:2]
    r = np.random.default_rng(0).normal(size=s.n)
    z = np.empty_like(r)
    P.vmult(z, r)
    zo = H.vmult(r)
    assert np.abs(z - zo).max() <= 1e-12 * np.abs(zo).max()


def test_chebyshev_is_rejected_on_the_coarse_grid_and_when_partitioned(gpu_ctx):
    from helpers import spd_laplacian
    R = ab.RelaxationType
    s = spd_laplacian(6)
    A = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
    with pytest.raises(ab.AmgbError):
        ab.PreconditionBoomerAMG().initialize(A, _cheby_data(relaxation_type_coarse=R.Chebyshev))


def test_one_uploaded_matrix_serves_several_contexts_at_once(gpu_ctx):
    """The independent systems of a theta sweep share one read-only matrix and run on
    different contexts (streams, host threads) at the same time; every one of them gets
    the bits of a sequential run."""
    import threading
    s = poisson(16, contrast=3.0)
    A = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
    thetas = [0.25, 0.5, 0.75, 0.9]
    want = {}
    for th in thetas:
        x = s.x0.copy()
        want[th] = (ab.amg_solve(device_data(th), 1e-8, A, s.rhs, x), x)
    got, errs = {}, []

    def lane(th):
        try:
            ctx = ab.Context(0)
            x = s.x0.copy()
            got[th] = (ab.amg_solve(device_data(th), 1e-8, A, s.rhs, x, ctx), x)
            ctx.close()
        except BaseException as e:  # noqa: BLE001
            errs.append(e)
    ts = [threading.Thread(target=lane, args=(th,)) for th in thetas]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for th in thetas:
        assert got[th][0]["niters"] == want[th][0]["niters"]
        assert np.array_equal(got[th][0]["p_res"], want[th][0]["p_res"])
        assert np.array_equal(got[th][1], want[th][1])
        assert np.array_equal(got[th][0]["nrows"], want[th][0]["nrows"])


@pytest.mark.parametrize("tier", ["1", "2"])
def test_spgemm_table_tiers_give_the_same_hierarchy(gpu_ctx, monkeypatch, tier):
    """AMGB_SPGEMM_TIER forces the wider first-stage hash tables of A*P (picked for wide
    operators such as 3-DoF elasticity) on a small system: same bits as the oracle."""
    monkeypatch.setenv("AMGB_SPGEMM_TIER", tier)
    s = ab.gen.elasticity_q1(5, 2, 3, 10.0 ** ab.gen.checkerboard_epsv(2, 3, 2.0))
    A, P, H = _both(gpu_ctx, s, device_data(0.25))
    _assert_hierarchy_identical(P, H)


@pytest.mark.parametrize("mode", ["0", "1"])
@pytest.mark.parametrize("kind", ["poisson", "elasticity", "hub"])
def test_spgemm_row_per_thread_and_sub_warp_routes_agree_with_the_oracle(gpu_ctx, monkeypatch, mode, kind):
    """AMGB_SPGEMM_ROWTHREAD=1 sends every unsorted product (A*P) through the row-per-thread
    kernels, =0 through the sub-warp hash kernels.  Elasticity and hub/leaf rows outgrow the
    48-entry rows (and the hub rows the 112-column count table) of the row-per-thread tables, so
    the overflow lists of both passes are exercised.  Same bits as the oracle either way."""
    from helpers import hub_leaf_csr
    from types import SimpleNamespace
    monkeypatch.setenv("AMGB_SPGEMM_ROWTHREAD", mode)
    if kind == "poisson":
        s = poisson(14, contrast=3.0)
        theta = 0.25
    elif kind == "elasticity":
        s = ab.gen.elasticity_q1(6, 2, 3, 10.0 ** ab.gen.checkerboard_epsv(2, 3, 2.0))
        theta = 0.25
    else:
        M = hub_leaf_csr(300, 500, 60, 8, 7)
        s = SimpleNamespace(n=M.shape[0], col=M.indices.astype(np.int32), val=M.data.astype(np.float64),
                            rowptr32=lambda: M.indptr.astype(np.int32))
        theta = 0.05
    gpu_ctx.reset_routes()
    A, P, H = _both(gpu_ctx, s, device_data(theta))
    r = gpu_ctx.routes()
    assert (r["spgemm_rowreg"] > 0) == (mode == "1"), r
    if mode == "1" and kind != "poisson":
        assert r["spgemm_num_big"] > 0, r          # rows past 48 entries went to the second stage
    _assert_hierarchy_identical(P, H)


@pytest.mark.parametrize("mode", ["0", "1", "2", "4", "auto"])
@pytest.mark.parametrize("kind", ["poisson", "elasticity", "hub", "aggressive"])
def test_spgemm_flattened_and_entrywise_first_stages_agree_with_the_oracle(gpu_ctx, monkeypatch, mode, kind):
    """AMGB_SPGEMM_FLAT: 1 = the flattened first-stage kernels for A*P-type products (products
    of a 32-entry chunk numbered by scans, match.any batches), 0 = the entry-by-entry sub-warp
    kernels, 2 = flattened with one warp per row for R*(AP) too, 4 = flattened count pass and
    the entry-by-entry numeric pass inside the warp-uniform kernel, unset = picked per product.  Hub/leaf rows span
    several chunks and outgrow the product stage (overflow list); the aggressive levels send
    SORTED products through the same kernels.  Same bits as the oracle every way."""
    from helpers import hub_leaf_csr
    from types import SimpleNamespace
    if mode == "auto":
        monkeypatch.delenv("AMGB_SPGEMM_FLAT", raising=False)
    else:
        monkeypatch.setenv("AMGB_SPGEMM_FLAT", mode)
    agg = 0
    if kind == "poisson":
        s = poisson(14, contrast=3.0)
        theta = 0.25
    elif kind == "aggressive":
        s = poisson(12, contrast=2.0)
        theta, agg = 0.25, 1
    elif kind == "elasticity":
        s = ab.gen.elasticity_q1(6, 2, 3, 10.0 ** ab.gen.checkerboard_epsv(2, 3, 2.0))
        theta = 0.25
    else:
        M = hub_leaf_csr(300, 500, 60, 8, 7)
        s = SimpleNamespace(n=M.shape[0], col=M.indices.astype(np.int32), val=M.data.astype(np.float64),
                            rowptr32=lambda: M.indptr.astype(np.int32))
        theta = 0.05
    gpu_ctx.reset_routes()
    data = device_data(theta)
    data.aggressive_coarsening_num_levels = agg
    A, P, H = _both(gpu_ctx, s, data)
    r = gpu_ctx.routes()
    assert (r["spgemm_flat"] > 0) == (mode != "0"), r
    _assert_hierarchy_identical(P, H)


@pytest.mark.parametrize("nh,nl,per_leaf,leaf_leaf", [(100, 300, 30, 4), (160, 400, 70, 6)])
def test_setup_long_interpolation_rows(gpu_ctx, nh, nl, per_leaf, leaf_leaf):
    """Hub/leaf systems: every leaf depends strongly on 30 (70) hubs, which PMIS makes the C
    points, and on a few other leaves.  Interpolation rows longer than the lane-parallel limit
    (16) and than the shared-memory stage (48), rows of A spanning several 32-entry chunks,
    transposed rows past the two-entries-per-lane rank sort.  Still the oracle's bits."""
    from helpers import hub_leaf_csr
    from types import SimpleNamespace
    M = hub_leaf_csr(nh, nl, per_leaf, leaf_leaf, 5)
    s = SimpleNamespace(n=M.shape[0], col=M.indices.astype(np.int32), val=M.data.astype(np.float64),
                        rowptr32=lambda: M.indptr.astype(np.int32))
    A, P, H = _both(gpu_ctx, s, device_data(0.05))
    prp, _, _, nc = P.P(0)
    assert nc == nh and np.diff(prp).max() == per_leaf      # the long-row routes are really taken
    _assert_hierarchy_identical(P, H)
    rhs = np.random.default_rng(1).normal(size=s.n)
    ctl = ab.SolverControl(s.n, 1e-9)
    x = np.zeros(s.n)
    ab.SolverCG(ctl).solve(A, x, rhs, P)
    rc, xo, nit, hist = H.cg_solve(rhs, np.zeros(s.n), abs_tol=1e-9)
    assert rc == 0 and abs(ctl.last_step() - nit) <= 1
    k = min(len(hist), len(ctl.history))
    # rows of 200 entries are summed in different orders by the SELL kernels and the oracle:
    # agreement to 1e-10 of the initial residual, and to 1e-6 of each entry down to 1e-10
    d = np.abs(ctl.history[:k] - hist[:k])
    assert (d <= RES_RTOL * hist[0]).all() and (d <= 1e-6 * hist[:k]).all()


def test_pcg_as_a_while_graph_gives_the_same_history(gpu_ctx, monkeypatch):
    """AMGB_PCG_GRAPH_LOOP=1: prologue, WHILE(not converged) { PCG step }, scatter as ONE graph
    whose condition is set on the device; same residual history, iteration count and solution
    bits as the host-driven loop, NoConvergence and immediate convergence included."""
    s = poisson(16, contrast=4.0)
    A = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
    out = {}
    for mode in ("host", "graph"):
        if mode == "graph":
            monkeypatch.setenv("AMGB_PCG_GRAPH_LOOP", "1")
        gpu_ctx.reset_routes()
        P = ab.PreconditionBoomerAMG()
        P.initialize(A, device_data(0.5))
        ctl = ab.SolverControl(s.n, 1e-8)
        x = s.x0.copy()
        ab.SolverCG(ctl).solve(A, x, s.rhs, P)
        assert (gpu_ctx.routes()["pcg_device_loop"] > 0) == (mode == "graph")
        short = ab.SolverControl(3, 1e-30)
        x3 = s.x0.copy()
        with pytest.raises(ab.NoConvergence):
            ab.SolverCG(short).solve(A, x3, s.rhs, P)
        easy = ab.SolverControl(10, 1e30)
        ab.SolverCG(easy).solve(A, s.x0.copy(), s.rhs, P)
        out[mode] = (ctl.last_step(), ctl.history.copy(), x, short.last_step(), len(short.history), x3, easy.last_step())
        P.close()
    h, g = out["host"], out["graph"]
    assert h[0] == g[0] and np.array_equal(h[1], g[1]) and np.array_equal(h[2], g[2])
    assert (h[3], h[4]) == (g[3], g[4]) == (3, 4) and np.array_equal(h[5], g[5])
    assert h[6] == g[6] == 0


@pytest.mark.parametrize("m,theta,contrast,relax", [(16, 0.25, 3.0, "l1"), (20, 0.5, 6.0, "l1"), (24, 0.25, 0.0, "jacobi"),
                                                    (12, 0.9, 6.0, "l1")])
def test_coarse_tail_in_one_kernel_matches_the_per_level_cycle(gpu_ctx, monkeypatch, m, theta, contrast, relax):
    """Levels of a few hundred rows and below run in ONE kernel from shared memory
    (amgb_tail.cu).  Same hierarchy; V-cycle and residual history agree with the per-level
    launches (AMGB_NO_TAIL=1) to rounding and with the oracle to the north-star tolerance."""
    s = poisson(m, contrast=contrast)
    R = ab.RelaxationType
    kw = dict(relaxation_type_up=R.Jacobi, relaxation_type_down=R.Jacobi, relax_weight=0.7) if relax == "jacobi" else {}
    data = device_data(theta, **kw)
    A = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
    H = orc.Hierarchy(s.rowptr32(), s.col, s.val, data.to_struct())
    r = np.random.default_rng(11).standard_normal(s.n)
    zo = H.vmult(r)
    rc, xo, nit, hist = H.cg_solve(s.rhs, s.x0, abs_tol=1e-8)
    got = {}
    for mode in ("tail", "levels"):
        if mode == "levels":
            monkeypatch.setenv("AMGB_NO_TAIL", "1")
        gpu_ctx.reset_routes()
        P = ab.PreconditionBoomerAMG()
        P.initialize(A, data)
        _assert_hierarchy_identical(P, H)
        z = np.empty(s.n)
        P.vmult(z, r)
        ctl = ab.SolverControl(s.n, 1e-8)
        x = s.x0.copy()
        ab.SolverCG(ctl).solve(A, x, s.rhs, P)
        assert (gpu_ctx.routes()["cycle_tail_fused"] > 0) == (mode == "tail"), gpu_ctx.routes()
        got[mode] = (z, ctl.history.copy(), ctl.last_step(), x)
        P.close()
    for mode in got:
        z, h, it, x = got[mode]
        assert np.abs(z - zo).max() <= 1e-12 * np.abs(zo).max(), mode
        assert abs(it - nit) <= 1
        k = min(len(hist), len(h))
        assert (np.abs(h[:k] - hist[:k]) <= RES_RTOL * hist[:k]).all(), mode
    k = min(len(got["tail"][1]), len(got["levels"][1]))
    assert (np.abs(got["tail"][1][:k] - got["levels"][1][:k]) <= 1e-11 * got["levels"][1][:k]).all()


# ---- CLJP coarsening (hypre coarsen type 0): the parallel member of the Falgout family ----
@pytest.mark.parametrize("m,theta,contrast", [(8, 0.25, 0.0), (12, 0.5, 3.0), (16, 0.25, 6.0), (14, 0.7, 0.0), (10, 0.05, 6.0)])
def test_cljp_coarsening_is_bit_exact(gpu_ctx, m, theta, contrast):
    """coarsen_type = CLJP on the device against the oracle's CLJP: C/F splittings, operators and
    interpolation bit-identical on every level; PCG history to the north-star tolerance."""
    s = poisson(m, contrast=contrast)
    data = device_data(theta, coarsen_type=ab.COARSEN_CLJP)
    A, P, H = _both(gpu_ctx, s, data)
    _assert_hierarchy_identical(P, H)
    cf = P.cf_marker(0)
    assert set(np.unique(cf)) <= {1, -1, -3}
    ctl = ab.SolverControl(s.n, 1e-8)
    x = s.x0.copy()
    ab.SolverCG(ctl).solve(A, x, s.rhs, P)
    rc, xo, nit, hist = H.cg_solve(s.rhs, s.x0, abs_tol=1e-8)
    assert rc == 0 and abs(ctl.last_step() - nit) <= 1
    k = min(len(hist), len(ctl.history))
    assert (np.abs(ctl.history[:k] - hist[:k]) <= RES_RTOL * hist[:k]).all()
    # the point of this family: fewer iterations than PMIS at a higher operator complexity
    Pm = ab.PreconditionBoomerAMG()
    Pm.initialize(A, device_data(theta))
    assert P.level_stats()["operator"] > Pm.level_stats()["operator"]


def test_cljp_on_an_unstructured_system(gpu_ctx):
    M = random_spd_csr(3000, 0.003, 11)
    class S:
        n = M.shape[0]; col = M.indices; val = M.data
        def rowptr32(self): return M.indptr.astype(np.int32)
    A, P, H = _both(gpu_ctx, S(), device_data(0.25, coarsen_type=ab.COARSEN_CLJP))
    _assert_hierarchy_identical(P, H)


def test_cljp_is_refused_where_it_is_not_implemented(gpu_ctx):
    s = poisson(6)
    A = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
    d = device_data(0.25, coarsen_type=ab.COARSEN_CLJP)
    d.aggressive_coarsening_num_levels = 1
    with pytest.raises(ab.AmgbError) as e:
        ab.PreconditionBoomerAMG().initialize(A, d)
    assert e.value.status == -5


@pytest.mark.parametrize("kind", ["poisson", "elasticity", "hub"])
def test_warp_per_row_interpolation_kernel_alone_agrees_with_the_oracle(gpu_ctx, monkeypatch, kind):
    """AMGB_INTERP_WARP=1 builds every interpolation row with the warp-per-row kernel (otherwise the
    second stage of the grouped kernel, which only sees rows with more than 32 interpolation points
    or a neighbour matching more than 8 of them).  Same bits as the oracle."""
    from helpers import hub_leaf_csr
    from types import SimpleNamespace
    monkeypatch.setenv("AMGB_INTERP_WARP", "1")
    if kind == "poisson":
        s, theta = poisson(14, contrast=3.0), 0.25
    elif kind == "elasticity":
        s, theta = ab.gen.elasticity_q1(6, 2, 3, 10.0 ** ab.gen.checkerboard_epsv(2, 3, 2.0)), 0.25
    else:
        M = hub_leaf_csr(160, 400, 70, 6, 5)
        s = SimpleNamespace(n=M.shape[0], col=M.indices.astype(np.int32), val=M.data.astype(np.float64),
                            rowptr32=lambda: M.indptr.astype(np.int32))
        theta = 0.05
    A, P, H = _both(gpu_ctx, s, device_data(theta))
    _assert_hierarchy_identical(P, H)


@pytest.mark.parametrize("m,theta,contrast", [(24, 0.25, 3.0), (30, 0.6, 6.0)])
def test_window_sorted_solve_numbering_agrees_with_the_oracle(gpu_ctx, monkeypatch, m, theta, contrast):
    """AMGB_ROW_SORT=1: inside the C block and the F block of the solve numbering, windows of 1024
    rows are ordered by (row length of A, row length of P) -- SELL-C-sigma.  The hierarchy is
    untouched (it lives in the original numbering); V-cycle and residual history agree with the
    oracle to the north-star tolerance and with the unsorted numbering to rounding."""
    s = poisson(m, contrast=contrast)
    data = device_data(theta)
    A = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
    H = orc.Hierarchy(s.rowptr32(), s.col, s.val, data.to_struct())
    r = np.random.default_rng(5).standard_normal(s.n)
    zo = H.vmult(r)
    rc, xo, nit, hist = H.cg_solve(s.rhs, s.x0, abs_tol=1e-8)
    hs = {}
    for mode in ("plain", "sorted"):
        if mode == "sorted":
            monkeypatch.setenv("AMGB_ROW_SORT", "1")
        P = ab.PreconditionBoomerAMG()
        P.initialize(A, data)
        _assert_hierarchy_identical(P, H)
        z = np.empty(s.n)
        P.vmult(z, r)
        assert np.abs(z - zo).max() <= 1e-12 * np.abs(zo).max(), mode
        ctl = ab.SolverControl(s.n, 1e-8)
        x = s.x0.copy()
        ab.SolverCG(ctl).solve(A, x, s.rhs, P)
        assert abs(ctl.last_step() - nit) <= 1
        k = min(len(hist), len(ctl.history))
        assert (np.abs(ctl.history[:k] - hist[:k]) <= RES_RTOL * hist[:k]).all(), mode
        assert np.abs(x - xo).max() <= 1e-9 * np.abs(xo).max()
        hs[mode] = ctl.history.copy()
        P.close()
    k = min(len(hs["plain"]), len(hs["sorted"]))
    assert (np.abs(hs["plain"][:k] - hs["sorted"][:k]) <= 1e-11 * hs["plain"][:k]).all()
