"""GPU (-m gpu): pooling against the reference's golden vectors, and the reference's OWN
harness code (amg_solver.h + view_maker.h compiled unmodified against the deal.II-compat
layer, oracle/_ref/ref_harness_gpu, built in the container where /root/reference exists)
running on libamgb.so."""
import csv
import os
import subprocess
import sys

import numpy as np
import pytest

import amg_ann_b200 as ab
from helpers import poisson

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
REFDIR = os.path.join(os.path.dirname(HERE), "oracle", "_ref")
sys.path.insert(0, GOLD)
import make_golden  # noqa: E402

csv.field_size_limit(1 << 30)


@pytest.mark.parametrize("name", sorted(make_golden.VIEW_CASES))
def test_device_pooling_equals_reference_viewmaker_golden(gpu_ctx, name):
    kind, kw, V = make_golden.VIEW_CASES[name]
    rp, col, val = make_golden.build_case(kind, kw)
    g = np.load(os.path.join(GOLD, f"view_{name}.npz"))
    A = ab.SparseMatrix(gpu_ctx, rp.astype(np.int32), col, val)
    vm = ab.ViewMaker(V).make_view(A)
    assert np.array_equal(vm.count, g["count"])                       # integer: bit-exact
    assert np.array_equal(vm.max_pp, g["max_pp"]) and np.array_equal(vm.max_np, g["max_np"])
    # fp64 sum: fixed tree on the device vs the reference's serial order
    assert np.allclose(vm.view, g["view"], rtol=0, atol=1e-13 * np.abs(val).sum())


@pytest.mark.skipif(not os.path.exists(os.path.join(REFDIR, "ref_harness_gpu")),
                    reason="oracle/_ref/ref_harness_gpu not built (needs /root/reference at build time)")
def test_reference_harness_runs_unmodified_on_the_device(gpu_ctx, tmp_path):
    m, ps, mode, contrast, V = 12, 2, 3, 3.0, 20
    out = tmp_path / "stats.csv"
    r = subprocess.run([os.path.join(REFDIR, "ref_harness_gpu"), str(m), str(ps), str(mode), str(contrast),
                        "0.05,0.96,0.3", str(V), str(out)], capture_output=True, text=True, cwd=tmp_path,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = list(csv.reader(open(out)))
    assert rows[0][0] == "view" and [x[0] for x in rows[1:]] == ["solve"] * 4
    s = poisson(m, contrast=contrast, ps=ps, mode=mode)
    A = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
    # pooled image: the reference's CPU loop (MatGetRow on the compat matrix) vs the device pass
    vm = ab.ViewMaker(V).make_view(A)
    assert int(rows[0][2]) == V
    assert np.array_equal(np.array([int(float(x)) for x in rows[0][4].split(",")]), vm.count)
    assert np.array_equal(np.array([float(x) for x in rows[0][5].split(",")]), vm.max_pp)
    assert np.array_equal(np.array([float(x) for x in rows[0][6].split(",")]), vm.max_np)
    assert np.allclose(np.array([float(x) for x in rows[0][3].split(",")]), vm.view, rtol=0,
                       atol=1e-13 * np.abs(s.val).sum())
    # solve rows: theta,maxrowsum,symop,agg,tol,t_setup,t_solve,nrows,nze,sparsity,grid,operator,memory,niters,p_res
    for row, th in zip(rows[1:], ab.gen.theta_sweep(0.05, 0.96, 0.3)):
        f = row[1:]
        assert float(f[0]) == th and float(f[1]) == 0.9 and f[2] == "1" and f[3] == "0"
        x = s.x0.copy()
        mine = ab.amg_solve(ab.AdditionalData(True, th, 0.9, 0, True), 1e-8, A, s.rhs, x)
        assert [float(v) for v in f[7].split(",")] == [float(v) for v in mine["nrows"]]
        assert [float(v) for v in f[8].split(",")] == [float(v) for v in mine["nze"]]
        assert int(f[13]) == mine["niters"]
        pres = np.array([float(v) for v in f[14].split(",")])
        assert len(pres) == mine["niters"] + 1
        # scraped from "%14.12e": 13 significant digits
        assert np.allclose(pres, mine["p_res"], rtol=2e-12, atol=0)
        assert float(f[10]) == pytest.approx(mine["grid"], abs=1e-6)
        assert float(f[11]) == pytest.approx(mine["operator"], abs=1e-6)
