#!/usr/bin/env python
"""Row-partitioned AMG-PCG over NCCL, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/dist_nccl.py --cells 200 --theta 0.25 [--check]
Generates this rank's z-slab of the Q1 diffusion system, runs setup + PCG, prints sizes,
iterations and device-synchronised wall times (max over ranks).  --check: rank 0 also runs
the single-device path on the whole system and compares level sizes / residual history."""
import argparse
import json
import os
import sys
import time

import numpy as np

# torchrun pins OMP_NUM_THREADS=1; the slab generator is OpenMP-parallel
os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // max(1, int(os.environ.get("WORLD_SIZE", 1)))))

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as tdist  # noqa: E402

import amg_ann_b200 as ab  # noqa: E402
from amg_ann_b200 import dist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", dest="m", type=int, default=200)
ap.add_argument("--theta", type=str, default="0.25", help="comma-separated list; the slab is generated once")
ap.add_argument("--contrast", type=float, default=0.0)
ap.add_argument("--repeat", type=int, default=2)
ap.add_argument("--check", action="store_true")
ap.add_argument("--timers", action="store_true")
ap.add_argument("--device-assembly", action="store_true", help="assemble the slab on the GPU instead of the host")
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29511")
tdist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
ctx = ab.Context(local)
comm = dist.Communicator.nccl_from_torch(ctx)
R = ab.RelaxationType
thetas = [float(t) for t in args.theta.split(",")]


def options(theta):
    return ab.AdditionalData(True, theta, 0.9, 0, True, relaxation_type_up=R.l1scaledJacobi,
                             relaxation_type_down=R.l1scaledJacobi)


data = options(thetas[0])
epsv = ab.gen.checkerboard_epsv(4, 3, args.contrast)
starts = dist.slab_partition(args.m, world)
b, e = starts[rank], starts[rank + 1]
t0 = time.perf_counter()
if args.device_assembly:
    from types import SimpleNamespace
    d_rhs = torch.empty(e - b, dtype=torch.float64, device="cuda")
    d_x0 = torch.empty(e - b, dtype=torch.float64, device="cuda")
    A = dist.DistSparseMatrix.assemble_poisson_q1(comm, args.m, b, e, 4, 3, epsv, d_rhs.data_ptr(), d_x0.data_ptr())
    ctx.synchronize()
    sl = SimpleNamespace(n=(args.m + 1) ** 3, rhs=d_rhs.cpu().numpy(), x0=d_x0.cpu().numpy())
    t_gen = time.perf_counter() - t0
else:
    sl = ab.gen.poisson_q1(args.m, 4, 3, epsv, row_begin=b, row_end=e)
    t_gen = time.perf_counter() - t0
    A = dist.DistSparseMatrix(comm, sl.n, b, e, sl.rowptr, sl.col, sl.val)


def tmax(v):
    t = torch.tensor([v], device="cuda", dtype=torch.float64)
    tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
    return float(t.item())


res = None
for rep, theta in [(r, t) for t in thetas for r in range(args.repeat)]:
    data = options(theta)
    if args.timers and rep == args.repeat - 1 and theta == thetas[-1]:
        ctx.enable_timers(True)
        ctx.reset_timers()
    tdist.barrier()
    torch.cuda.synchronize()
    P = dist.DistPreconditionBoomerAMG()
    t0 = time.perf_counter()
    P.initialize(A, data)
    ctx.synchronize()
    t1 = time.perf_counter()
    ctl = ab.SolverControl(sl.n, 1e-8)
    x = sl.x0.copy()
    dist.DistSolverCG(ctl).solve(A, x, sl.rhs, P)
    t2 = time.perf_counter()
    st = P.level_stats()
    res = dict(theta=theta, n=sl.n, nnz_global=int(st["nnz"][0]), ranks=world, rows=[int(v) for v in st["rows"]],
               operator_complexity=st["operator"], iters=ctl.last_step(), res=ctl.last_value(),
               setup_s=tmax(t1 - t0), solve_s=tmax(t2 - t1), gen_s=tmax(t_gen), rep=rep)
    hist = ctl.history
    P.close()
    if rank == 0:
        print(json.dumps(res), flush=True)
if args.timers:
    t = ctx.timers()
    tot = sum(v["ms"] for k, v in t.items() if not k.endswith("_l0"))
    if rank == 0:
        for k, v in sorted(t.items(), key=lambda kv: -kv[1]["ms"]):
            if v["launches"]:
                print(f"  {k:12s} {v['ms']:9.2f} ms {100 * v['ms'] / tot:5.1f}% {v['launches']:6d} launches "
                      f"{v['bytes'] / max(v['ms'], 1e-9) / 1e6:8.1f} GB/s")
if args.check and rank == 0:
    s = ab.gen.poisson_q1(args.m, 4, 3, epsv)
    A1 = ab.SparseMatrix(ctx, s.rowptr32(), s.col, s.val)
    x1 = s.x0.copy()
    row = ab.amg_solve(data, 1e-8, A1, s.rhs, x1)
    k = min(len(hist), len(row["p_res"]))
    print(json.dumps(dict(check="single-device", rows_equal=[int(v) for v in row["nrows"]] == res["rows"],
                          iters_single=int(row["niters"]), iters_dist=res["iters"],
                          max_rel_hist_diff=float(np.max(np.abs(hist[:k] - row["p_res"][:k]) / row["p_res"][:k])),
                          x_diff=float(np.abs(x - x1[b:e]).max() / np.abs(x1).max()),
                          setup_ms_single=row["t_amg_setup"] / 1e3, solve_ms_single=row["t_solve"] / 1e3)))
A.close()
comm.close()
tdist.barrier()
tdist.destroy_process_group()
