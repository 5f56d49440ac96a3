"""GPU parity tests of the variants added last in round 2 (multicolour Gauss-Seidel, the
warp-uniform interpolation kernel).  In a file of their own that sorts after the full-size and the
reference-harness suites: under `pytest -x` a failure here cannot hide those."""
import numpy as np
import pytest

import amg_ann_b200 as ab
from helpers import device_data, poisson
from oracle import binding as orc
from test_gpu_parity import RES_RTOL, _assert_hierarchy_identical, _both

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["poisson", "elasticity", "hub"])
def test_warp_uniform_grouped_interpolation_kernel_agrees_with_the_oracle(gpu_ctx, monkeypatch, kind):
    """AMGB_INTERP_UNIFORM=1: the grouped interpolation kernel with warp-uniform control flow
    (full-warp collectives, loops to the warp's maximum trip count) instead of group masks.
    Same bits as the oracle; hub rows overflow into the warp-per-row second stage."""
    from helpers import hub_leaf_csr
    from types import SimpleNamespace
    monkeypatch.setenv("AMGB_INTERP_UNIFORM", "1")
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


@pytest.mark.parametrize("m,theta,contrast,kinds", [(12, 0.25, 0.0, "sym"), (20, 0.5, 6.0, "sym"), (24, 0.25, 3.0, "fb"),
                                                    (16, 0.9, 6.0, "sym")])
def test_multicolour_gauss_seidel_agrees_with_the_oracle(gpu_ctx, m, theta, contrast, kinds):
    """smoother_policy = SMOOTHER_MULTICOLOR: the reference's Gauss-Seidel sweeps (deal.II
    symmetricSORJacobi / SORJacobi / backwardSORJacobi -> hypre 6 / 3 / 4, the PCHYPRE default
    smoother of every reference call site, ref common/amg_solver.h:48) run in multicolour order
    instead of being substituted by l1-Jacobi.  Colours bit-identical to the oracle's greedy
    colouring on every level, a valid colouring, V-cycle and residual history to the
    north-star tolerance, and fewer PCG iterations than the l1-Jacobi substitute."""
    s = poisson(m, contrast=contrast)
    R = ab.RelaxationType
    if kinds == "sym":
        kw = dict(relaxation_type_up=R.symmetricSORJacobi, relaxation_type_down=R.symmetricSORJacobi)
        want_types = (106, 106, 9)
    else:  # forward on the way down, backward on the way up
        kw = dict(relaxation_type_up=R.backwardSORJacobi, relaxation_type_down=R.SORJacobi)
        want_types = (103, 104, 9)
    data = device_data(theta, smoother_policy=ab.SMOOTHER_MULTICOLOR, **kw)
    if kinds == "fb":
        data.symmetric_operator = False  # (deal.II maps SORJacobi to hypre 6 for symmetric operators)
    A, P, H = _both(gpu_ctx, s, data)
    assert P.effective_relax() == H.effective_relax() == want_types
    _assert_hierarchy_identical(P, H)
    for l in range(H.num_levels - 1):
        cd, ncd = P.colors(l)
        co, nco = H.colors(l)
        assert ncd == nco and np.array_equal(cd, co), f"colours level {l}"
        rp, cl, _ = H.A(l)
        rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
        off = rows != cl
        assert not (cd[rows[off]] == cd[cl[off]]).any(), f"invalid colouring on level {l}"
    r = np.random.default_rng(3).standard_normal(s.n)
    z = np.empty(s.n)
    P.vmult(z, r)
    zo = H.vmult(r)
    assert np.abs(z - zo).max() <= 1e-12 * np.abs(zo).max()
    ctl = ab.SolverControl(s.n, 1e-8)
    x = s.x0.copy()
    ab.SolverCG(ctl).solve(A, x, s.rhs, P)
    rc, xo, nit, hist = H.cg_solve(s.rhs, s.x0, abs_tol=1e-8)
    assert rc == 0 and abs(ctl.last_step() - nit) <= 1
    k = min(len(hist), len(ctl.history))
    assert (np.abs(ctl.history[:k] - hist[:k]) <= RES_RTOL * hist[:k]).all()
    assert np.abs(x - xo).max() <= 1e-9 * np.abs(xo).max()
    # against the substitute: same system, C/F l1-Jacobi
    P2 = ab.PreconditionBoomerAMG()
    P2.initialize(A, device_data(theta))
    ctl2 = ab.SolverControl(s.n, 1e-8)
    x2 = s.x0.copy()
    ab.SolverCG(ctl2).solve(A, x2, s.rhs, P2)
    if kinds == "sym":
        assert ctl.last_step() < ctl2.last_step(), (ctl.last_step(), ctl2.last_step())
