#!/usr/bin/env python
"""Run one (matrix, theta) pass of the hot path on cuda:0 -- the small command
profiled under ncu (profiles/README.md).  --mode setup|solve|full|pool"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amg_ann_b200 as ab  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=100)
ap.add_argument("--theta", type=float, default=0.25)
ap.add_argument("--contrast", type=float, default=6.0)
ap.add_argument("--mode", default="full")
ap.add_argument("--repeat", type=int, default=1)
ap.add_argument("--max-steps", type=int, default=0)
ap.add_argument("--timers", action="store_true", help="per-family CUDA-event timers (disables the V-cycle graph)")
ap.add_argument("--kind", default="poisson", choices=["poisson", "elasticity"])
ap.add_argument("--smoother", default="l1jacobi", choices=["l1jacobi", "chebyshev", "mcgs", "mcfb"],
                help="mcgs: multicolour symmetric Gauss-Seidel (SMOOTHER_MULTICOLOR); mcfb: forward down, backward up")
ap.add_argument("--agg", type=int, default=0, help="aggressive_coarsening_num_levels (testcase 3 passes 2)")
args = ap.parse_args()

epsv = ab.gen.checkerboard_epsv(4, 3, args.contrast)
t_gen = time.perf_counter()
if args.kind == "elasticity":   # BASELINE config 3: Q1 vector elasticity, 3 DoF/node, scalar AMG
    s = ab.gen.elasticity_q1(args.m, 4, 3, 10.0 ** epsv)
else:
    s = ab.gen.poisson_q1(args.m, 4, 3, epsv)
print(f"{args.kind} m={args.m}: n={s.n} nnz={s.nnz} generated in {time.perf_counter() - t_gen:.1f} s", flush=True)
R = ab.RelaxationType
sm = R.Chebyshev if args.smoother == "chebyshev" else R.l1scaledJacobi
data = ab.AdditionalData(True, args.theta, 0.9, args.agg, True, relaxation_type_up=sm, relaxation_type_down=sm)
if args.smoother == "mcgs":
    data = ab.AdditionalData(True, args.theta, 0.9, args.agg, True, relaxation_type_up=R.symmetricSORJacobi,
                             relaxation_type_down=R.symmetricSORJacobi, smoother_policy=ab.SMOOTHER_MULTICOLOR)
elif args.smoother == "mcfb":
    data = ab.AdditionalData(False, args.theta, 0.9, args.agg, True, relaxation_type_up=R.backwardSORJacobi,
                             relaxation_type_down=R.SORJacobi, smoother_policy=ab.SMOOTHER_MULTICOLOR)
ctx = ab.Context(0)
A = ab.SparseMatrix(ctx, s.rowptr32(), s.col, s.val)
for rep in range(args.repeat):
    if args.timers and rep == args.repeat - 1:
        ctx.enable_timers(True)
        ctx.reset_timers()
    if args.mode == "pool":
        vm = ab.ViewMaker(75).make_view(A)
        print(f"pool: device {vm.t_device_us:.1f} us, count sum {vm.count.sum()}")
        continue
    P = ab.PreconditionBoomerAMG()
    t0 = time.perf_counter()
    P.initialize(A, data)
    ctx.synchronize()
    t1 = time.perf_counter()
    st = P.level_stats()
    msg = f"setup {1e3 * (t1 - t0):.1f} ms levels {list(st['rows'])} opcx {st['operator']:.3f}"
    if args.mode in ("solve", "full"):
        ctl = ab.SolverControl(args.max_steps or s.n, 1e-8)
        x = s.x0.copy()
        t2 = time.perf_counter()
        try:
            ab.SolverCG(ctl).solve(A, x, s.rhs, P)
        except ab.NoConvergence:
            pass
        t3 = time.perf_counter()
        msg += f" | solve {1e3 * (t3 - t2):.1f} ms iters {ctl.last_step()} res {ctl.last_value():.3e}"
    print(msg)
    P.close()
print("kernel launches", ctx.kernel_launches())
if args.timers:
    t = ctx.timers()
    tot = sum(v["ms"] for v in t.values())
    for k, v in sorted(t.items(), key=lambda kv: -kv[1]["ms"]):
        if v["launches"]:
            print(f"  {k:12s} {v['ms']:9.2f} ms {100 * v['ms'] / tot:5.1f}% {v['launches']:6d} launches "
                  f"{v['bytes'] / max(v['ms'], 1e-9) / 1e6:8.1f} GB/s  {1e3 * v['ms'] / v['launches']:8.1f} us/launch")
    print("per level (ms | GB/s | launches):")
    for fam, rows in ctx.level_timers().items():
        if fam.endswith("_l0"):
            continue
        cells = [f"L{l}: {r['ms']:.2f}|{r['bytes'] / max(r['ms'], 1e-9) / 1e6:.0f}|{r['launches']}"
                 for l, r in enumerate(rows) if r["launches"]]
        print(f"  {fam:10s} " + "  ".join(cells))
