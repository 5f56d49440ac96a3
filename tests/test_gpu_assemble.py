"""GPU (-m gpu): on-device assembly of the Q1 diffusion system must
equal the host generator bit for bit -- matrix, right-hand side and initial guess -- for
whole matrices and for the z-slabs of the row-partitioned path."""
import numpy as np
import pytest
import torch

import amg_ann_b200 as ab
from amg_ann_b200 import dist
from helpers import device_data

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,ps,mode,kind", [(1, 1, 1, "zero"), (4, 2, 3, "checker"), (9, 3, 2, "random"),
                                            (12, 4, 3, "random"), (16, 2, 1, "checker"), (7, 5, 3, "random")])
def test_device_assembly_equals_host_generator(gpu_ctx, m, ps, mode, kind):
    if kind == "zero":
        epsv = np.zeros(ps ** mode)
    elif kind == "checker":
        epsv = ab.gen.checkerboard_epsv(ps, mode, 6.0)
    else:
        epsv = ab.gen.random_vec(3, ps ** mode, 6.0)
    s = ab.gen.poisson_q1(m, ps, mode, epsv)
    rhs = torch.empty(s.n, dtype=torch.float64, device="cuda")
    x0 = torch.empty(s.n, dtype=torch.float64, device="cuda")
    A = ab.SparseMatrix.assemble_poisson_q1(gpu_ctx, m, ps, mode, epsv, rhs.data_ptr(), x0.data_ptr())
    rp, cl, vl = A.download()
    assert np.array_equal(rp, s.rowptr32()) and np.array_equal(cl, s.col)
    assert np.array_equal(vl, s.val), np.abs(vl - s.val).max()
    assert np.array_equal(rhs.cpu().numpy(), s.rhs)
    assert np.array_equal(x0.cpu().numpy(), s.x0)
    # the assembled matrix feeds the path like an uploaded one
    if m >= 4:
        x = s.x0.copy()
        row = ab.amg_solve(device_data(0.25), 1e-8, A, s.rhs, x)
        A2 = ab.SparseMatrix(gpu_ctx, s.rowptr32(), s.col, s.val)
        x2 = s.x0.copy()
        row2 = ab.amg_solve(device_data(0.25), 1e-8, A2, s.rhs, x2)
        assert row["niters"] == row2["niters"] and np.array_equal(row["p_res"], row2["p_res"])


def test_device_assembly_of_slabs_feeds_the_partitioned_path(gpu_ctx):
    m, ps, mode = 10, 2, 3
    epsv = ab.gen.checkerboard_epsv(ps, mode, 3.0)
    starts = dist.slab_partition(m, 3)
    whole = ab.gen.poisson_q1(m, ps, mode, epsv)

    def fn(rank, comm):
        b, e = starts[rank], starts[rank + 1]
        rhs = torch.empty(e - b, dtype=torch.float64, device="cuda")
        x0 = torch.empty(e - b, dtype=torch.float64, device="cuda")
        A = dist.DistSparseMatrix.assemble_poisson_q1(comm, m, b, e, ps, mode, epsv, rhs.data_ptr(), x0.data_ptr())
        P = dist.DistPreconditionBoomerAMG()
        P.initialize(A, device_data(0.25))
        rows = P.A_rows(0)
        ctl = ab.SolverControl(whole.n, 1e-8)
        x = x0.cpu().numpy().copy()
        dist.DistSolverCG(ctl).solve(A, x, rhs.cpu().numpy(), P)
        out = dict(rows=rows, rhs=rhs.cpu().numpy(), x0=x0.cpu().numpy(), niters=ctl.last_step(), hist=ctl.history)
        P.close()
        A.close()
        return out

    parts = dist.run_local_group(3, fn)
    for r, p in enumerate(parts):
        b, e = starts[r], starts[r + 1]
        rp, cl, vl = p["rows"]
        assert np.array_equal(rp, (whole.rowptr[b:e + 1] - whole.rowptr[b]).astype(np.int32))
        assert np.array_equal(cl, whole.col[whole.rowptr[b]:whole.rowptr[e]])
        assert np.array_equal(vl, whole.val[whole.rowptr[b]:whole.rowptr[e]])
        assert np.array_equal(p["rhs"], whole.rhs[b:e]) and np.array_equal(p["x0"], whole.x0[b:e])
    A1 = ab.SparseMatrix(gpu_ctx, whole.rowptr32(), whole.col, whole.val)
    x1 = whole.x0.copy()
    row = ab.amg_solve(device_data(0.25), 1e-8, A1, whole.rhs, x1)
    assert all(abs(p["niters"] - row["niters"]) <= 1 for p in parts)


@pytest.mark.parametrize("m,ps,mode", [(2, 1, 1), (5, 2, 3), (8, 4, 3), (11, 3, 2)])
def test_device_elasticity_assembly_equals_host_generator(gpu_ctx, m, ps, mode):
    """Q1 elasticity (ref t3 main.cpp:320-342) assembled on the device: pattern, values and the
    initial guess are the host generator's bit for bit; the right-hand side (device sin/cos in
    the body force) agrees to rounding."""
    young = 10.0 ** ab.gen.random_vec(5, ps ** mode, 2.0)
    s = ab.gen.elasticity_q1(m, ps, mode, young)
    rhs = torch.empty(s.n, dtype=torch.float64, device="cuda")
    x0 = torch.empty(s.n, dtype=torch.float64, device="cuda")
    A = ab.SparseMatrix.assemble_elasticity_q1(gpu_ctx, m, ps, mode, young, rhs.data_ptr(), x0.data_ptr())
    rp, cl, vl = A.download()
    assert np.array_equal(rp, s.rowptr32()) and np.array_equal(cl, s.col)
    assert np.array_equal(vl, s.val), np.abs(vl - s.val).max()
    assert np.array_equal(x0.cpu().numpy(), s.x0)
    r = rhs.cpu().numpy()
    assert np.abs(r - s.rhs).max() <= 1e-11 * max(1.0, np.abs(s.rhs).max())
