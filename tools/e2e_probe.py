#!/usr/bin/env python
"""Where the end-to-end (host buffers) time of one config-2 sweep goes: CSR upload, the host copy of
the initial guess, and per system the H2D/D2H around setup + solve, on one lane."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import amg_ann_b200 as ab
import bench

m = int(sys.argv[1]) if len(sys.argv) > 1 else 200
s = bench.make_system(ab, m)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
h_rp, h_col, h_val, h_b, h_x0, h_x = pin(s.rowptr32()), pin(s.col), pin(s.val), pin(s.rhs), pin(s.x0), pin(s.x0)
ctx = ab.Context(0)
thetas = ab.gen.theta_sweep(0.05, 0.96, 0.05)
for rep in range(2):
    t0 = time.perf_counter()
    A = ab.SparseMatrix(ctx, h_rp, h_col, h_val)
    t_up = time.perf_counter() - t0
    t_copy = t_call = t_setup = t_solve = 0.0
    for th in thetas:
        t1 = time.perf_counter()
        h_x[...] = h_x0
        t2 = time.perf_counter()
        row = ab.amg_solve(bench.device_options(ab, th), bench.TOL, A, h_b, h_x, ctx)
        t3 = time.perf_counter()
        t_copy += t2 - t1
        t_call += t3 - t2
        t_setup += row["t_amg_setup"] * 1e-6
        t_solve += row["t_solve"] * 1e-6
    A.close()
    tot = time.perf_counter() - t0
    print(f"rep {rep}: sweep {tot:.3f} s = upload {t_up:.3f} + x0 host copies {t_copy:.3f} + amg_solve calls {t_call:.3f} "
          f"(t_amg_setup {t_setup:.3f} + t_solve incl. x,b H2D and x D2H {t_solve:.3f} + rest {t_call - t_setup - t_solve:.3f}); "
          f"per system {tot / len(thetas):.4f} s")
# the transfers alone
d = torch.empty(s.n, dtype=torch.float64, device="cuda")
hb = torch.from_numpy(h_b)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    d.copy_(hb, non_blocking=True)
torch.cuda.synchronize()
print(f"H2D of one vector ({s.n * 8 / 1e6:.0f} MB, pinned): {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms")
t0 = time.perf_counter()
for _ in range(10):
    hb.copy_(d, non_blocking=True)
torch.cuda.synchronize()
print(f"D2H of one vector: {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms")
