"""GPU (-m gpu): the row-partitioned path against the single-device path.  The ranks are
host threads of one process on one GPU (in-process communicator); the NCCL back end runs
the same code with one process per GPU (bench.py --partitioned, tests/test_dist_host.py for
the host logic).  Integer outputs and operator values must be IDENTICAL for any number of
ranks (SURVEY.md 8e "Determinism"); residual histories agree to 1e-10."""
import numpy as np
import pytest

import amg_ann_b200 as ab
from amg_ann_b200 import dist
from helpers import device_data, poisson

pytestmark = pytest.mark.gpu


def _single(gpu_ctx, s, data):
    A = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
    P = ab.PreconditionBoomerAMG()
    P.initialize(A, data)
    ctl = ab.SolverControl(s.n, 1e-8)
    x = s.x0.copy()
    ab.SolverCG(ctl).solve(A, x, s.rhs, P)
    return A, P, ctl, x


def _run_partitioned(m, contrast, starts, data, solve=True, device_ids=None):
    nranks = len(starts) - 1

    def fn(rank, comm):
        b, e = starts[rank], starts[rank + 1]
        epsv = ab.gen.checkerboard_epsv(2, 3, contrast) if contrast else None
        sl = ab.gen.poisson_q1(m, 2, 3, epsv, row_begin=b, row_end=e) if contrast else \
            ab.gen.poisson_q1(m, row_begin=b, row_end=e)
        A = dist.DistSparseMatrix(comm, sl.n, b, e, sl.rowptr, sl.col, sl.val)
        P = dist.DistPreconditionBoomerAMG()
        P.initialize(A, data)
        out = dict(stats=P.level_stats(), levels=[], replicated_from=P.replicated_from)
        for l in range(P.num_levels):
            if l >= P.replicated_from:      # held whole on every rank
                out["levels"].append(dict(full=P.full_level(l)))
                continue
            d = P.level_dims(l)
            lev = dict(dims=d, A=P.A_rows(l))
            if l + 1 < P.num_levels:
                lev["cf"] = P.cf_marker(l)
                lev["P"] = P.P_rows(l)
            out["levels"].append(lev)
        if solve:
            ctl = ab.SolverControl(sl.n, 1e-8)
            x = sl.x0.copy()
            dist.DistSolverCG(ctl).solve(A, x, sl.rhs, P)
            out.update(x=x, hist=ctl.history, niters=ctl.last_step())
        P.close()
        A.close()
        return out

    return dist.run_local_group(nranks, fn, device_ids)


def _assert_same_hierarchy(parts, P1):
    nl = P1.num_levels
    assert all(len(p["levels"]) == nl for p in parts)
    for l in range(nl):
        rp, cl, vl = P1.A(l)
        n = len(rp) - 1
        if "full" in parts[0]["levels"][l]:     # replicated level: every rank holds the single-device level
            for p in parts:
                f = p["levels"][l]["full"]
                assert all(np.array_equal(a, b) for a, b in zip(f["A"], (rp, cl, vl))), f"replicated A level {l}"
                if l + 1 < nl:
                    assert np.array_equal(f["cf"], P1.cf_marker(l)), f"replicated cf level {l}"
                    prp1, pcl1, pvl1, nc = P1.P(l)
                    assert f["P"][3] == nc and all(np.array_equal(a, b) for a, b in zip(f["P"][:3], (prp1, pcl1, pvl1)))
            continue
        assert sum(p["levels"][l]["dims"]["n_local"] for p in parts) == n
        for p in parts:
            d = p["levels"][l]["dims"]
            b, k = d["row_begin"], d["n_local"]
            prp, pcl, pvl = p["levels"][l]["A"]
            assert np.array_equal(prp, rp[b:b + k + 1] - rp[b]), f"A rowptr level {l}"
            assert np.array_equal(pcl, cl[rp[b]:rp[b + k]]), f"A pattern level {l}"
            assert np.array_equal(pvl, vl[rp[b]:rp[b + k]]), f"A values level {l}"
        if l + 1 < nl:
            cf = P1.cf_marker(l)
            prp1, pcl1, pvl1, nc = P1.P(l)
            for p in parts:
                d = p["levels"][l]["dims"]
                b, k = d["row_begin"], d["n_local"]
                assert np.array_equal(p["levels"][l]["cf"], cf[b:b + k]), f"cf level {l}"
                assert d["n_coarse_global"] == nc
                qrp, qcl, qvl = p["levels"][l]["P"]
                assert np.array_equal(qrp, prp1[b:b + k + 1] - prp1[b]), f"P rowptr level {l}"
                assert np.array_equal(qcl, pcl1[prp1[b]:prp1[b + k]]), f"P pattern level {l}"
                assert np.array_equal(qvl, pvl1[prp1[b]:prp1[b + k]]), f"P values level {l}"


@pytest.mark.parametrize("replicate_below", [0, 300, 1 << 20])
@pytest.mark.parametrize("m,nranks,theta,contrast", [(8, 2, 0.25, 0.0), (12, 3, 0.5, 3.0), (10, 4, 0.25, 6.0),
                                                     (12, 2, 0.7, 6.0)])
def test_partitioned_equals_single_device(gpu_ctx, m, nranks, theta, contrast, replicate_below):
    """replicate_below: 0 = every level partitioned; 300 = the small levels gathered on every
    rank; 2^20 = everything below the finest level replicated."""
    s = poisson(m, contrast=contrast)
    data = device_data(theta, dist_replicate_below=replicate_below)
    A1, P1, ctl1, x1 = _single(gpu_ctx, s, data)
    parts = _run_partitioned(m, contrast, dist.slab_partition(m, nranks), data)
    _assert_same_hierarchy(parts, P1)
    st1 = P1.level_stats()
    for p in parts:
        assert np.array_equal(p["stats"]["rows"], st1["rows"]) and np.array_equal(p["stats"]["nnz"], st1["nnz"])
        assert p["stats"]["memory"] == st1["memory"]
        assert abs(p["niters"] - ctl1.last_step()) <= 1
        k = min(len(p["hist"]), len(ctl1.history))
        assert (np.abs(p["hist"][:k] - ctl1.history[:k]) <= 1e-10 * ctl1.history[:k]).all()
        assert np.array_equal(p["hist"], parts[0]["hist"])  # every rank sees the same scalars
    x = np.concatenate([p["x"] for p in parts])
    assert np.abs(x - x1).max() <= 1e-9 * np.abs(x1).max()


def test_partition_not_aligned_with_planes_and_single_rank(gpu_ctx):
    m, theta = 9, 0.25
    s = poisson(m, contrast=2.0)
    data = device_data(theta, dist_replicate_below=0)
    A1, P1, ctl1, x1 = _single(gpu_ctx, s, data)
    for starts in ([0, 137, 600, s.n], [0, s.n]):  # ragged ranges cutting through planes; one rank
        parts = _run_partitioned(m, 2.0, starts, data)
        _assert_same_hierarchy(parts, P1)
        assert all(abs(p["niters"] - ctl1.last_step()) <= 1 for p in parts)


@pytest.mark.parametrize("kw", [dict(w_cycle=True), dict(n_sweeps=2, max_iter=2),
                                dict(relaxation_type_coarse=ab.RelaxationType.l1scaledJacobi, n_sweeps_coarse=2),
                                dict(max_levels=3)])
def test_partitioned_cycle_options(gpu_ctx, kw):
    m = 10
    s = poisson(m, contrast=3.0)
    data = device_data(0.25, dist_replicate_below=100, **kw)
    A1, P1, ctl1, x1 = _single(gpu_ctx, s, data)
    parts = _run_partitioned(m, 3.0, dist.slab_partition(m, 3), data)
    _assert_same_hierarchy(parts, P1)
    for p in parts:
        assert abs(p["niters"] - ctl1.last_step()) <= 1
        k = min(len(p["hist"]), len(ctl1.history))
        assert (np.abs(p["hist"][:k] - ctl1.history[:k]) <= 1e-10 * ctl1.history[:k]).all()


@pytest.mark.parametrize("starts_kind", ["slabs", "ragged"])
@pytest.mark.parametrize("replicate_below", [0, 300])
def test_overlapped_halo_exchange_equals_single_device(gpu_ctx, monkeypatch, starts_kind, replicate_below):
    """AMGB_OVERLAP_MIN_ROWS=0 forces the interior/boundary row split (put, interior rows, wait,
    boundary rows) on every partitioned level of a small system; ragged ranges give interior
    ranges that do not line up with mesh planes, or none at all."""
    monkeypatch.setenv("AMGB_OVERLAP_MIN_ROWS", "0")
    m, contrast = 12, 3.0
    s = poisson(m, contrast=contrast)
    data = device_data(0.25, dist_replicate_below=replicate_below)
    A1, P1, ctl1, x1 = _single(gpu_ctx, s, data)
    starts = dist.slab_partition(m, 3) if starts_kind == "slabs" else [0, 5, 911, 1400, s.n]
    parts = _run_partitioned(m, contrast, starts, data)
    _assert_same_hierarchy(parts, P1)
    for p in parts:
        assert abs(p["niters"] - ctl1.last_step()) <= 1
        k = min(len(p["hist"]), len(ctl1.history))
        assert (np.abs(p["hist"][:k] - ctl1.history[:k]) <= 1e-10 * ctl1.history[:k]).all()
        assert np.array_equal(p["hist"], parts[0]["hist"])
    x = np.concatenate([p["x"] for p in parts])
    assert np.abs(x - x1).max() <= 1e-9 * np.abs(x1).max()


def test_library_exchanges_still_work_without_peer_windows(gpu_ctx, monkeypatch):
    """AMGB_PEER=0: the communicator's alltoallv / allreduce carry the solve-phase exchanges
    (the path taken when a rank cannot map its peers' memory)."""
    monkeypatch.setenv("AMGB_PEER", "0")
    m, contrast = 10, 3.0
    s = poisson(m, contrast=contrast)
    data = device_data(0.25, dist_replicate_below=100)
    A1, P1, ctl1, x1 = _single(gpu_ctx, s, data)
    parts = _run_partitioned(m, contrast, dist.slab_partition(m, 3), data)
    for p in parts:
        assert abs(p["niters"] - ctl1.last_step()) <= 1
        k = min(len(p["hist"]), len(ctl1.history))
        assert (np.abs(p["hist"][:k] - ctl1.history[:k]) <= 1e-10 * ctl1.history[:k]).all()


@pytest.mark.parametrize("replicate_below", [0, 400])
def test_chebyshev_smoother_partitioned_equals_single_device(gpu_ctx, replicate_below):
    """Chebyshev (hypre relax type 16) on the row-partitioned path: the spectrum estimate is a
    collective CG/Lanczos run, the sweeps exchange the halo of their gather source.  The
    hierarchy is identical; coefficients and residuals agree to rounding (the inner products
    are summed per rank, so the last bits depend on the partition)."""
    from helpers import spd_laplacian
    R = ab.RelaxationType
    s = spd_laplacian(12, seed=4, decades=1.0)
    data = device_data(0.25, relaxation_type_up=R.Chebyshev, relaxation_type_down=R.Chebyshev,
                       dist_replicate_below=replicate_below)
    A1, P1, ctl1, x1 = _single(gpu_ctx, s, data)
    starts = [0, 500, 1100, s.n]
    rp = s.rowptr

    def fn(rank, comm):
        b, e = starts[rank], starts[rank + 1]
        A = dist.DistSparseMatrix(comm, s.n, b, e, (rp[b:e + 1] - rp[b]).astype(np.int64), s.col[rp[b]:rp[e]],
                                  s.val[rp[b]:rp[e]])
        P = dist.DistPreconditionBoomerAMG()
        P.initialize(A, data)
        ctl = ab.SolverControl(s.n, 1e-8)
        x = s.x0[b:e].copy()
        dist.DistSolverCG(ctl).solve(A, x, s.rhs[b:e], P)
        out = dict(rows=P.level_stats()["rows"], x=x, hist=ctl.history, niters=ctl.last_step())
        P.close()
        A.close()
        return out

    parts = dist.run_local_group(3, fn)
    for p in parts:
        assert np.array_equal(p["rows"], P1.level_stats()["rows"])
        assert abs(p["niters"] - ctl1.last_step()) <= 1
        k = min(len(p["hist"]), len(ctl1.history))
        d = np.abs(p["hist"][:k] - ctl1.history[:k])
        assert (d[:10] <= 1e-10 * ctl1.history[:10]).all() and (d <= 1e-8 * ctl1.history[:k]).all()
        assert np.array_equal(p["hist"], parts[0]["hist"])
    x = np.concatenate([p["x"] for p in parts])
    assert np.abs(x - x1).max() <= 1e-8 * np.abs(x1).max()


@pytest.mark.parametrize("m,V,nranks", [(10, 20, 3), (14, 75, 2), (6, 50, 4)])
def test_pooled_image_of_a_partitioned_matrix(gpu_ctx, m, V, nranks):
    """amgb_dist_make_view: local pooling of the owned rows + combination over the ranks = the image
    of the assembled matrix (count / maxima exact, sum to rounding), the same on every rank."""
    s = poisson(m, contrast=4.0)
    A1 = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
    vm = ab.ViewMaker(V).make_view(A1)
    starts = dist.slab_partition(m, nranks) if nranks < 4 else [0, 17, 100, 101, s.n]

    def fn(rank, comm):
        b, e = starts[rank], starts[rank + 1]
        sl = ab.gen.poisson_q1(m, 2, 3, ab.gen.checkerboard_epsv(2, 3, 4.0), row_begin=b, row_end=e)
        A = dist.DistSparseMatrix(comm, sl.n, b, e, sl.rowptr, sl.col, sl.val)
        out = A.make_view(V)
        A.close()
        return out

    parts = dist.run_local_group(nranks, fn)
    for sm, cnt, pp, npv, _ in parts:
        assert np.array_equal(cnt, vm.count) and cnt.sum() == s.nnz
        assert np.array_equal(pp, vm.max_pp) and np.array_equal(npv, vm.max_np)
        assert np.allclose(sm, vm.view, rtol=0, atol=1e-13 * np.abs(s.val).sum())
        assert np.array_equal(sm, parts[0][0])


def test_a_rank_with_bad_arguments_fails_on_every_rank_together(gpu_ctx):
    """amgb_dist_matrix_create is collective including its failures: the rank whose arguments are
    wrong and its peers all come back with an error from the same all-gather; nobody is left
    waiting in a collective (the group is NOT aborted by the test: the library must agree)."""
    m = 6
    starts = dist.slab_partition(m, 2)

    def fn(rank, comm):
        b, e = starts[rank], starts[rank + 1]
        sl = ab.gen.poisson_q1(m, row_begin=b, row_end=e)
        try:
            if rank == 1:   # row range past the end of the matrix
                dist.DistSparseMatrix(comm, sl.n, b, sl.n + 5, sl.rowptr, sl.col, sl.val)
            else:
                dist.DistSparseMatrix(comm, sl.n, b, e, sl.rowptr, sl.col, sl.val)
        except ab.AmgbError as err:
            return err.status
        return 0

    out = dist.run_local_group(2, fn)
    assert out == [-9, -1]     # peer: AMGB_ERR_COMM naming the rank; offender: AMGB_ERR_BAD_ARG


# ---- real multi-GPU variants (skipped on a one-GPU box; run with `gpurun --gpus 2`) ----
def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("replicate_below", [0, 300])
@pytest.mark.parametrize("m,theta,contrast", [(12, 0.25, 3.0), (16, 0.5, 6.0)])
def test_partitioned_on_distinct_gpus_equals_single_device(gpu_ctx, m, theta, contrast, replicate_below):
    """The ranks are host threads again, but every rank owns its own GPU: halo values, the
    replication all-gather and the PCG scalars cross NVLink through the peer windows."""
    ng = min(_device_count(), 4)
    if ng < 2:
        pytest.skip("needs at least 2 GPUs")
    s = poisson(m, contrast=contrast)
    data = device_data(theta, dist_replicate_below=replicate_below)
    A1, P1, ctl1, x1 = _single(gpu_ctx, s, data)
    parts = _run_partitioned(m, contrast, dist.slab_partition(m, ng), data, device_ids=list(range(ng)))
    _assert_same_hierarchy(parts, P1)
    for p in parts:
        assert abs(p["niters"] - ctl1.last_step()) <= 1
        k = min(len(p["hist"]), len(ctl1.history))
        assert (np.abs(p["hist"][:k] - ctl1.history[:k]) <= 1e-10 * ctl1.history[:k]).all()
        assert np.array_equal(p["hist"], parts[0]["hist"])
    x = np.concatenate([p["x"] for p in parts])
    assert np.abs(x - x1).max() <= 1e-9 * np.abs(x1).max()


@pytest.mark.parametrize("peer", ["1", "0"])
def test_partitioned_nccl_processes_equal_single_device(peer):
    """One PROCESS per GPU over NCCL (the way bench.py --gpus N runs config 5): tools/dist_nccl.py
    --check compares level sizes, iteration count and residual history with the single-device
    run on rank 0.  peer=1: NVLink peer windows for the solve-phase exchanges; peer=0: NCCL."""
    import json
    import os
    import subprocess
    import sys
    ng = min(_device_count(), 4)
    if ng < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, AMGB_PEER=peer, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ng}",
           "--master-addr", "127.0.0.1", "--master-port", "29547" if peer == "1" else "29548",
           os.path.join(root, "tools", "dist_nccl.py"), "--cells", "40", "--theta", "0.25", "--contrast", "3",
           "--check"]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{\"check\"")]
    assert line, out.stdout[-2000:]
    chk = json.loads(line[-1])
    assert chk["rows_equal"] and abs(chk["iters_single"] - chk["iters_dist"]) <= 1
    assert chk["max_rel_hist_diff"] <= 1e-10
