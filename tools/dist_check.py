#!/usr/bin/env python
"""Row-partitioned path vs single device on one GPU (ranks = host threads): prints level
sizes, iteration counts and timings.  python tools/dist_check.py --m 40 --ranks 4"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amg_ann_b200 as ab  # noqa: E402
from amg_ann_b200 import dist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=40)
ap.add_argument("--ranks", type=int, default=2)
ap.add_argument("--theta", type=float, default=0.25)
ap.add_argument("--contrast", type=float, default=6.0)
ap.add_argument("--skip-single", action="store_true")
args = ap.parse_args()
R = ab.RelaxationType
data = ab.AdditionalData(True, args.theta, 0.9, 0, True, relaxation_type_up=R.l1scaledJacobi,
                         relaxation_type_down=R.l1scaledJacobi)
epsv = ab.gen.checkerboard_epsv(4, 3, args.contrast)
if not args.skip_single:
    s = ab.gen.poisson_q1(args.m, 4, 3, epsv)
    ctx = ab.Context(0)
    A = ab.SparseMatrix(ctx, s.rowptr32(), s.col, s.val)
    for rep in range(2):
        x = s.x0.copy()
        row = ab.amg_solve(data, 1e-8, A, s.rhs, x)
    print("single:", list(row["nrows"]), "iters", row["niters"], "setup ms", row["t_amg_setup"] / 1e3, "solve ms",
          row["t_solve"] / 1e3, "res", row["p_res"][-1])
starts = dist.slab_partition(args.m, args.ranks)


def fn(rank, comm):
    b, e = starts[rank], starts[rank + 1]
    sl = ab.gen.poisson_q1(args.m, 4, 3, epsv, row_begin=b, row_end=e)
    A = dist.DistSparseMatrix(comm, sl.n, b, e, sl.rowptr, sl.col, sl.val)
    res = None
    for rep in range(2):
        P = dist.DistPreconditionBoomerAMG()
        t0 = time.perf_counter()
        P.initialize(A, data)
        comm.ctx.synchronize()
        t1 = time.perf_counter()
        ctl = ab.SolverControl(sl.n, 1e-8)
        x = sl.x0.copy()
        dist.DistSolverCG(ctl).solve(A, x, sl.rhs, P)
        t2 = time.perf_counter()
        res = dict(rows=list(P.level_stats()["rows"]), iters=ctl.last_step(), setup_ms=1e3 * (t1 - t0),
                   solve_ms=1e3 * (t2 - t1), res=ctl.last_value(),
                   local=[P.level_dims(l)["n_local"] for l in range(min(P.num_levels, P.replicated_from))],
                   x=x, hist=ctl.history)
        P.close()
    A.close()
    return res


out = dist.run_local_group(args.ranks, fn)
for r, o in enumerate(out):
    print(f"rank {r}: rows {o['rows']} local {o['local']} iters {o['iters']} setup {o['setup_ms']:.1f} ms "
          f"solve {o['solve_ms']:.1f} ms res {o['res']:.3e}")
if not args.skip_single:
    xx = np.concatenate([o["x"] for o in out])
    k = min(len(row["p_res"]), len(out[0]["hist"]))
    print("max rel hist diff", np.max(np.abs(out[0]["hist"][:k] - row["p_res"][:k]) / row["p_res"][:k]),
          "x diff", np.abs(xx - x).max() / np.abs(x).max())
