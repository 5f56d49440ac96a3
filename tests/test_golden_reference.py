"""Pins against the REFERENCE's own source (CPU, no GPU): the committed golden vectors of
tests/golden/ were produced by the reference's ViewMaker and output scrapers compiled from
/root/reference (tests/golden/make_golden.py, oracle/ref/Makefile).  Where oracle/_ref/ is
present (this container) the same comparisons are also made live."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
REFDIR = os.path.join(os.path.dirname(HERE), "oracle", "_ref")

import sys
sys.path.insert(0, GOLD)
import make_golden  # noqa: E402  (case table + builders; nothing runs at import)


def _have(tool):
    return os.path.exists(os.path.join(REFDIR, tool))


@pytest.mark.parametrize("name", sorted(make_golden.VIEW_CASES))
def test_oracle_pooling_equals_reference_viewmaker_golden(orc, name):
    kind, kw, V = make_golden.VIEW_CASES[name]
    rp, col, val = make_golden.build_case(kind, kw)
    g = np.load(os.path.join(GOLD, f"view_{name}.npz"))
    digest = hashlib.sha256(rp.tobytes() + col.tobytes() + val.tobytes()).hexdigest()
    assert digest == str(g["input_sha256"]), "generator output changed: regenerate the golden fixtures"
    s, c, pp, npn = orc.make_view(rp.astype(np.int32), col, val, V)
    # the oracle walks the CSR in the reference's order: every channel is bit-identical
    assert np.array_equal(c, g["count"])
    assert np.array_equal(pp, g["max_pp"]) and np.array_equal(npn, g["max_np"])
    assert np.array_equal(s, g["view"])


@pytest.mark.skipif(not _have("ref_view_cpu"), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["poisson_m6_V5", "random_n300_V7"])
def test_live_reference_viewmaker_equals_golden(name):
    kind, kw, V = make_golden.VIEW_CASES[name]
    rp, col, val = make_golden.build_case(kind, kw)
    view, count, mpp, mnp = make_golden.run_ref_view(rp, col, val, V)
    g = np.load(os.path.join(GOLD, f"view_{name}.npz"))
    assert np.array_equal(view, g["view"]) and np.array_equal(count, g["count"])
    assert np.array_equal(mpp, g["max_pp"]) and np.array_equal(mnp, g["max_np"])


def test_scrape_golden_is_self_consistent():
    with open(os.path.join(GOLD, "scrape.json")) as f:
        g = json.load(f)
    rows = [r for r, _ in g["levels"]]
    nnz = [z for _, z in g["levels"]]
    assert g["parsed_rows"] == [float(r) for r in rows]
    assert g["parsed_nze"] == [float(z) for z in nnz]
    # hypre prints sparsity with 3 decimals; that is what reaches the reference's CSV
    assert g["parsed_sparsity"] == [round(z / r / r, 3) for r, z in g["levels"]]
    grid = sum(rows) / rows[0]
    op = sum(nnz) / nnz[0]
    assert g["parsed_complexities"][0] == pytest.approx(grid, abs=5e-7)
    assert g["parsed_complexities"][1] == pytest.approx(op, abs=5e-7)
    # -ksp_monitor prints 12 digits after the point (parser.h:153)
    for a, b in zip(g["parsed_residuals"], g["residuals"]):
        assert a == pytest.approx(b, rel=1e-12)
    assert len(g["parsed_residuals"]) == len(g["residuals"])


@pytest.mark.skipif(not (_have("ref_parse") and _have("format_probe")),
                    reason="oracle/_ref not built (needs /root/reference)")
def test_compat_layer_text_is_scraped_by_the_reference_parsers():
    with open(os.path.join(GOLD, "scrape.json")) as f:
        g = json.load(f)
    lv = "".join(f"{r} {z}\n" for r, z in g["levels"])
    text = make_golden.run_tool("format_probe", ["boomeramg", "0.25", "0.9", "25"], lv)
    assert text == g["stats_text"]
    parsed = make_golden.run_tool("ref_parse", ["boomeramg"], text).split("\n")
    assert [float(x) for x in parsed[0].split()] == g["parsed_rows"]
    assert [float(x) for x in parsed[1].split()] == g["parsed_nze"]
    ksp_text = make_golden.run_tool("format_probe", ["ksp"], "".join(f"{r!r}\n" for r in g["residuals"]))
    assert ksp_text == g["ksp_text"]
    assert [float(x) for x in make_golden.run_tool("ref_parse", ["ksp"], ksp_text).split()] == g["parsed_residuals"]
    # a malformed printout must be rejected by the reference's regex, not half-parsed
    bad = subprocess.run([os.path.join(REFDIR, "ref_parse"), "boomeramg"], input=text.replace("Complexity", "Cmplx"),
                         capture_output=True, text=True)
    assert bad.returncode == 1


def test_oracle_regression_pins(orc):
    from helpers import device_data, poisson
    with open(os.path.join(GOLD, "oracle_regression.json")) as f:
        reg = json.load(f)
    for name, g in reg.items():
        s = poisson(g["m"], contrast=g["contrast"])
        H = orc.Hierarchy(s.rowptr32(), s.col, s.val, device_data(g["theta"]).to_struct())
        st = H.stats()
        assert [int(v) for v in st["rows"]] == g["rows"] and [int(v) for v in st["nnz"]] == g["nnz"], name
        for l in range(H.num_levels - 1):
            assert hashlib.sha256(H.cf_marker(l).tobytes()).hexdigest() == g["cf_sha256"][l], (name, l)
            assert hashlib.sha256(H.strength_mask(l).tobytes()).hexdigest() == g["mask_sha256"][l], (name, l)
        rc, x, nit, hist = H.cg_solve(s.rhs, s.x0, abs_tol=1e-8)
        assert rc == 0 and nit == g["niters"]
        assert hist[0] == pytest.approx(g["res0"], rel=1e-12)
        H.close()
