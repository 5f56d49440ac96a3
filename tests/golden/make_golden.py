#!/usr/bin/env python
"""Generates the committed golden fixtures of tests/golden/.  Run in the build
container, where /root/reference exists and `make -C oracle/ref` has produced
oracle/_ref/ (the reference's own view_maker.h / parser.h compiled from where they lie).

  view_*.npz   outputs of the REFERENCE's ViewMaker (ref common/view_maker.h:26-85) on
               seeded systems that the tests regenerate; parsed from the CSV text the
               reference prints with setprecision(17) (ref t2 main.cpp:503).
  scrape.json  hypre-format statistics / -ksp_monitor text printed by the dealii_compat
               layer and what the REFERENCE's scrapers (ref common/parser.h) read from it.
  oracle_regression.json  integer outputs of the oracle itself on two small systems (a
               regression pin of the restatement, not a reference pin: hypre is not available).
"""
import csv
csv.field_size_limit(1 << 30)
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REFDIR = os.path.join(ROOT, "oracle", "_ref")

VIEW_CASES = {  # name -> (builder, args, view_size)
    "poisson_m6_V5": ("poisson", dict(m=6, contrast=0.0), 5),
    "poisson_m10_c2_V75": ("poisson", dict(m=10, contrast=2.0), 75),
    "poisson_m2_V50_bins_exceed_rows": ("poisson", dict(m=2, contrast=0.0), 50),
    "elasticity_m4_V13": ("elasticity", dict(m=4), 13),
    "random_n300_V7": ("random", dict(n=300, density=0.03, seed=11), 7),
}


def build_case(kind, kw):
    import amg_ann_b200 as ab
    from helpers import poisson, random_spd_csr
    if kind == "poisson":
        s = poisson(kw["m"], contrast=kw["contrast"])
        return s.rowptr.astype(np.int64), s.col, s.val
    if kind == "elasticity":
        s = ab.gen.elasticity_q1(kw["m"], 2, 3, 10.0 ** ab.gen.checkerboard_epsv(2, 3, 1.0))
        return s.rowptr.astype(np.int64), s.col, s.val
    M = random_spd_csr(kw["n"], kw["density"], kw["seed"])
    return M.indptr.astype(np.int64), M.indices.astype(np.int32), M.data.astype(np.float64)


def run_ref_view(rowptr, col, val, V):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "csr.bin")
        with open(p, "wb") as f:
            np.array([len(rowptr) - 1, len(col)], dtype=np.int64).tofile(f)
            rowptr.astype(np.int64).tofile(f)
            col.astype(np.int32).tofile(f)
            val.astype(np.float64).tofile(f)
        out = os.path.join(d, "out.csv")
        subprocess.run([os.path.join(REFDIR, "ref_view_cpu"), p, str(V), out], check=True)
        with open(out) as f:
            row = next(csv.reader(f))
    assert int(row[1]) == V
    view = np.array([float(x) for x in row[2].split(",")])
    count = np.array([int(float(x)) for x in row[3].split(",")], dtype=np.int64)
    mpp = np.array([float(x) for x in row[4].split(",")])
    mnp = np.array([float(x) for x in row[5].split(",")])
    return view, count, mpp, mnp


def reference_norm_view():
    """The reference's own norm_view (ref code/data-modeling/train_ann.py:133-172), taken from
    where it lies: the function's source is cut out of the file with ast and executed with
    numpy only (the module itself imports tensorflow, which is not installed)."""
    import ast
    path = "/root/reference/code/data-modeling/train_ann.py"
    src = open(path).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "norm_view")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["norm_view"]


NORM_CASE = ("poisson", dict(m=10, contrast=2.0), 12)   # the image every mode is applied to
NORM_MODES = ["nothing", "pure", "resc", "pure_log", "resc_log", "mean"]


def make_norm_golden():
    norm_view = reference_norm_view()
    kind, kw, V = NORM_CASE
    rp, col, val = build_case(kind, kw)
    view, count, mpp, mnp = run_ref_view(rp, col, val, V)
    row = {"view": view, "view_count": count, "view_max_pp": mpp, "view_max_np": mnp}
    out = {}
    for mode in NORM_MODES:
        ch = [norm_view(row, mode, None, vt) for vt in ("", "_max_pp", "_max_np")]
        # normalize_view_df (:189-193) fills view_count_<mode> from the LAST view type (max_np)
        ch.append(norm_view(row, mode, None, "_max_np"))
        out[mode] = np.stack(ch, axis=-1)          # "sum+max+c" stacking, :247-256
    np.savez_compressed(os.path.join(HERE, "view_norm_poisson_m10_c2_V12.npz"), **out)
    print("norm golden:", {k: v.shape for k, v in out.items()})


def run_tool(name, args, text):
    return subprocess.run([os.path.join(REFDIR, name)] + args, input=text, capture_output=True, text=True,
                          check=True).stdout


SCRAPE_LEVELS = [(1030301, 27270901), (146952, 8012345), (17904, 912345), (2418, 101234), (327, 9001), (38, 500),
                 (9, 81)]
SCRAPE_RES = [1.2345678901234e+03, 4.5e+01, 3.25e-02, 9.87654321e-06, 7.43212345678e-09]


def main():
    from oracle import binding as orc
    from helpers import device_data, poisson
    for name, (kind, kw, V) in VIEW_CASES.items():
        rp, col, val = build_case(kind, kw)
        view, count, mpp, mnp = run_ref_view(rp, col, val, V)
        digest = hashlib.sha256(rp.tobytes() + col.tobytes() + val.tobytes()).hexdigest()
        np.savez_compressed(os.path.join(HERE, f"view_{name}.npz"), view=view, count=count, max_pp=mpp,
                            max_np=mnp, view_size=V, input_sha256=digest)
        print(name, "nnz", len(col), "count sum", count.sum())
    make_norm_golden()
    lv = "".join(f"{r} {z}\n" for r, z in SCRAPE_LEVELS)
    text = run_tool("format_probe", ["boomeramg", "0.25", "0.9", "25"], lv)
    parsed = run_tool("ref_parse", ["boomeramg"], text).split("\n")
    ksp_text = run_tool("format_probe", ["ksp"], "".join(f"{r!r}\n" for r in SCRAPE_RES))
    ksp = [float(x) for x in run_tool("ref_parse", ["ksp"], ksp_text).split()]
    scrape = {"levels": SCRAPE_LEVELS, "stats_text": text,
              "parsed_rows": [float(x) for x in parsed[0].split()],
              "parsed_nze": [float(x) for x in parsed[1].split()],
              "parsed_sparsity": [float(x) for x in parsed[2].split()],
              "parsed_complexities": [float(x) for x in parsed[3].split()],
              "residuals": SCRAPE_RES, "ksp_text": ksp_text, "parsed_residuals": ksp}
    with open(os.path.join(HERE, "scrape.json"), "w") as f:
        json.dump(scrape, f, indent=1)
    reg = {}
    for name, (m, theta, contrast) in {"poisson_m10_t0.25": (10, 0.25, 0.0), "diffusion_m12_t0.5_c6": (12, 0.5, 6.0)}.items():
        s = poisson(m, contrast=contrast)
        H = orc.Hierarchy(s.rowptr32(), s.col, s.val, device_data(theta).to_struct())
        rc, x, nit, hist = H.cg_solve(s.rhs, s.x0, abs_tol=1e-8)
        st = H.stats()
        reg[name] = {"m": m, "theta": theta, "contrast": contrast, "rows": [int(v) for v in st["rows"]],
                     "nnz": [int(v) for v in st["nnz"]], "niters": int(nit),
                     "cf_sha256": [hashlib.sha256(H.cf_marker(l).tobytes()).hexdigest()
                                   for l in range(H.num_levels - 1)],
                     "mask_sha256": [hashlib.sha256(H.strength_mask(l).tobytes()).hexdigest()
                                     for l in range(H.num_levels - 1)],
                     "res0": float(hist[0]), "res_last": float(hist[-1])}
        H.close()
    with open(os.path.join(HERE, "oracle_regression.json"), "w") as f:
        json.dump(reg, f, indent=1)


if __name__ == "__main__":
    main()
