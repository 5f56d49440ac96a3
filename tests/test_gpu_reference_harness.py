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
    out = tmp_path / "stats.csv"  # relative name: the reference's redirector prepends "./" (redirector.h:30)
    r = subprocess.run([os.path.join(REFDIR, "ref_harness_gpu"), str(m), str(ps), str(mode), str(contrast),
                        "0.05,0.96,0.3", str(V), "stats.csv"], capture_output=True, text=True, cwd=tmp_path,
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


def test_cpp_datagen_driver_writes_the_reference_csv_schema(gpu_ctx, tmp_path):
    """amg-ann_b200/host/amgb_datagen (C++ host code on the C ABI) vs the Python mirror."""
    exe = os.path.join(os.path.dirname(HERE), "amg-ann_b200", "host", "amgb_datagen")
    out = tmp_path / "stats.csv"
    r = subprocess.run([exe, "--m", "12", "--pattern-size", "2", "--mode", "3", "--contrast", "3",
                        "--theta", "0.05,0.96,0.3", "--out", str(out)], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = list(csv.reader(open(out)))
    assert rows[0][:3] == ["setting", "dim", "ndof"] and rows[0][-3:] == ["t_solve", "niters", "p_res"]
    s = poisson(12, contrast=3.0)
    A = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
    assert len(rows) == 5
    for row, th in zip(rows[1:], ab.gen.theta_sweep(0.05, 0.96, 0.3)):
        assert int(row[2]) == s.n
        f = row[10:]
        assert float(f[0]) == th
        x = s.x0.copy()
        mine = ab.amg_solve(ab.AdditionalData(True, th, 0.9, 0, True), 1e-8, A, s.rhs, x)
        assert [int(float(v)) for v in f[7].split(",")] == [int(v) for v in mine["nrows"]]
        assert int(f[13]) == mine["niters"]
        assert np.array_equal(np.array([float(v) for v in f[14].split(",")]), mine["p_res"])  # %.17e round trip
    # pooled image mode + several systems over host threads
    out2 = tmp_path / "views.csv"
    r = subprocess.run([exe, "--m", "8", "--pattern-size", "2", "--mode", "3", "--seed", "5", "--make-view", "1",
                        "--view-size", "10", "--systems", "4", "--threads", "2", "--out", str(out2)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = list(csv.reader(open(out2)))
    assert len(rows) == 5
    for row in rows[1:]:
        cnt = np.array([int(float(v)) for v in row[13].split(",")])
        assert cnt.sum() == (3 * 8 + 1) ** 3 and int(row[11]) == 10


@pytest.mark.parametrize("mode", make_golden.NORM_MODES)
def test_device_view_normalisation_equals_reference_norm_view_golden(gpu_ctx, mode):
    """f3: pooled image normalised + stacked on the device vs the reference's own norm_view
    (golden generated from ref data-modeling/train_ann.py:133-172 by tests/golden/make_golden.py)."""
    kind, kw, V = make_golden.NORM_CASE
    rp, col, val = make_golden.build_case(kind, kw)
    A = ab.SparseMatrix(gpu_ctx, rp.astype(np.int32), col, val)
    g = np.load(os.path.join(GOLD, "view_norm_poisson_m10_c2_V12.npz"))[mode].reshape(V, V, 4)
    out = ab.ViewMaker(V).make_model_input(A, mode, count_channel_as_reference=True)
    assert out.shape == (V, V, 4)
    # fp64: the pooled sum differs from the serial sum in the last bits, log() by <= 1 ulp
    assert np.allclose(out, g, rtol=1e-12, atol=1e-13)
    if mode in ("pure", "pure_log", "resc", "resc_log"):
        assert np.abs(out).max() <= 1.0 and np.isclose(np.abs(out[..., 0]).max(), 1.0)
    # the true-count channel (library extension): counts normalised under the same mode
    own = ab.ViewMaker(V).make_model_input(A, mode, count_channel_as_reference=False)
    cnt = ab.ViewMaker(V).make_view(A).count.reshape(V, V).astype(float)
    expect = {"nothing": cnt, "pure": cnt / cnt.max(), "mean": (cnt > 0).astype(float)}.get(mode)
    if expect is not None:
        assert np.allclose(own[..., 3], expect, rtol=1e-15, atol=0)
    assert np.array_equal(own[..., :3], out[..., :3])


def test_cpp_driver_with_device_assembly_gives_identical_rows(tmp_path):
    exe = os.path.join(os.path.dirname(HERE), "amg-ann_b200", "host", "amgb_datagen")
    outs = []
    for dev in ("0", "1"):
        out = tmp_path / f"stats{dev}.csv"
        r = subprocess.run([exe, "--m", "10", "--pattern-size", "2", "--mode", "3", "--seed", "7", "--theta",
                            "0.25,0.6,0.25", "--device-assembly", dev, "--out", str(out)], capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(list(csv.reader(open(out))))
    assert len(outs[0]) == len(outs[1]) == 3
    for a, b in zip(outs[0][1:], outs[1][1:]):
        # everything but the timestamp / timing columns is identical (the matrices are bit-identical)
        assert a[:9] == b[:9] and a[10:15] == b[10:15] and a[17:] == b[17:]
