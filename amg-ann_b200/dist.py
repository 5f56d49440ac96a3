"""Row-partitioned (multi-GPU) front end: contiguous global row ranges, one rank per GPU
(include/amgb.h "row-partitioned path").

  partition_rows / slab_partition   host logic: who owns which rows
  Communicator                      NCCL (one process per GPU, unique id broadcast through
                                    torch.distributed) or an in-process thread group
  DistSparseMatrix                  this rank's slab, resident on its GPU
  DistPreconditionBoomerAMG         initialize() / level_stats() / owned-part accessors
  DistSolverCG                      solve(x_local, b_local)
  run_local_group                   run an SPMD function on nranks host threads
"""
import ctypes as C
import threading

import numpy as np

from . import (AMGB_OK, AdditionalData, AmgbError, Context, NoConvergence, SolverControl, _chk, _p)
from ._native import amgb_lib, c_f64p, c_i32p, c_i64p

AMGB_ERR_NO_CONVERGENCE = -6


def partition_rows(n_global, nranks):
    """Balanced contiguous ranges: starts[r]..starts[r+1] (PETSc's default ownership split)."""
    base, rem = divmod(int(n_global), int(nranks))
    starts = [0]
    for r in range(nranks):
        starts.append(starts[-1] + base + (1 if r < rem else 0))
    return starts


def slab_partition(m, nranks):
    """z-slab partition of the lexicographically numbered (m+1)^3 Q1 grid: every rank owns
    whole xy-planes, so a rank only talks to ranks r-1 and r+1 on the finest level."""
    planes = partition_rows(m + 1, nranks)
    return [p * (m + 1) * (m + 1) for p in planes]


def owner_of(starts, gid):
    """Rank owning global row gid (vectorised)."""
    return np.searchsorted(np.asarray(starts), np.asarray(gid), side="right") - 1


def halo_columns(starts, rank, col_global):
    """Sorted distinct non-owned columns referenced by a slab, grouped by owner: the halo
    exchange pattern of one SpMV (host mirror of the device plan, used by the CPU tests)."""
    col = np.unique(np.asarray(col_global))
    ext = col[(col < starts[rank]) | (col >= starts[rank + 1])]
    own = owner_of(starts, ext)
    return {int(q): ext[own == q] for q in np.unique(own)}


class LocalGroup:
    """In-process communicator group: ranks are host threads (one Context each)."""

    def __init__(self, nranks):
        self._h = C.c_void_p()
        rc = amgb_lib().amgb_local_group_create(int(nranks), C.byref(self._h))
        if rc != AMGB_OK:
            raise AmgbError(rc, "amgb_local_group_create")
        self.size = int(nranks)

    def abort(self):
        """Called by a failing rank: the others' collectives return an error instead of waiting."""
        if self._h:
            amgb_lib().amgb_local_group_abort(self._h)

    def close(self):
        if self._h:
            amgb_lib().amgb_local_group_destroy(self._h)
            self._h = C.c_void_p()


class Communicator:
    def __init__(self, handle, ctx):
        self._h, self.ctx = handle, ctx
        self.rank = amgb_lib().amgb_comm_rank(handle)
        self.size = amgb_lib().amgb_comm_size(handle)

    @classmethod
    def local(cls, group, rank, ctx):
        h = C.c_void_p()
        rc = amgb_lib().amgb_comm_create_local(group._h, int(rank), C.byref(h))
        if rc != AMGB_OK:
            raise AmgbError(rc, "amgb_comm_create_local")
        return cls(h, ctx)

    @classmethod
    def nccl_from_torch(cls, ctx):
        """One process per GPU: rank 0 creates the NCCL unique id, torch.distributed (any
        backend) broadcasts it."""
        import torch
        import torch.distributed as dist
        rank, size = dist.get_rank(), dist.get_world_size()
        buf = (C.c_char * 128)()
        if rank == 0:
            rc = amgb_lib().amgb_nccl_unique_id(buf, 128)
            if rc != AMGB_OK:
                raise AmgbError(rc, "amgb_nccl_unique_id")
        t = torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, 0)
        raw = bytes(t.cpu().numpy().tobytes())
        h = C.c_void_p()
        rc = amgb_lib().amgb_comm_create_nccl(ctx._h, size, rank, raw, C.byref(h))
        _chk(ctx._h, rc, "amgb_comm_create_nccl")
        return cls(h, ctx)

    def close(self):
        if self._h:
            amgb_lib().amgb_comm_destroy(self._h)
            self._h = C.c_void_p()


class DistSparseMatrix:
    """Rows [row_begin,row_end) of the global matrix with GLOBAL column ids (what
    gen.poisson_q1(row_begin=, row_end=) returns)."""

    def __init__(self, comm, n_global, row_begin, row_end, rowptr_local, col_global, val):
        self.comm, self.ctx = comm, comm.ctx
        self.n_global, self.row_begin, self.row_end = int(n_global), int(row_begin), int(row_end)
        rp = np.ascontiguousarray(rowptr_local, dtype=np.int64)
        col = np.ascontiguousarray(col_global, dtype=np.int32)
        v = np.ascontiguousarray(val, dtype=np.float64)
        self._h = C.c_void_p()
        rc = amgb_lib().amgb_dist_matrix_create(self.ctx._h, comm._h, self.n_global, self.row_begin,
                                                self.row_end, _p(rp, c_i64p), _p(col, c_i32p),
                                                _p(v, c_f64p), C.byref(self._h))
        _chk(self.ctx._h, rc, "amgb_dist_matrix_create")

    @classmethod
    def assemble_poisson_q1(cls, comm, m, row_begin, row_end, pattern_size=1, mode=1, epsv=None, rhs_ptr=0,
                            x0_ptr=0):
        """This rank's slab assembled on its device (bit-identical to gen.poisson_q1 with the
        same row range).  rhs_ptr / x0_ptr: device pointers to row_end-row_begin doubles, or 0."""
        if epsv is None:
            epsv = np.zeros(pattern_size ** mode)
        epsv = np.ascontiguousarray(epsv, dtype=np.float64)
        self = cls.__new__(cls)
        self.comm, self.ctx = comm, comm.ctx
        self.n_global, self.row_begin, self.row_end = (m + 1) ** 3, int(row_begin), int(row_end)
        self._h = C.c_void_p()
        rc = amgb_lib().amgb_dist_matrix_assemble_poisson_q1(
            self.ctx._h, comm._h, m, pattern_size, mode, _p(epsv, c_f64p), len(epsv), self.row_begin, self.row_end,
            C.byref(self._h), C.c_void_p(rhs_ptr), C.c_void_p(x0_ptr))
        _chk(self.ctx._h, rc, "amgb_dist_matrix_assemble_poisson_q1")
        return self

    @property
    def n_local(self):
        return self.row_end - self.row_begin

    def make_view(self, view_size):
        """Pooled image of the whole partitioned matrix (collective; ref common/view_maker.h:26-74):
        returns (sum, count, max_pp, max_np, device_us), each V*V row-major."""
        vv = int(view_size) ** 2
        s, c = np.zeros(vv), np.zeros(vv, dtype=np.int64)
        pp, npv = np.zeros(vv), np.zeros(vv)
        t = C.c_double()
        rc = amgb_lib().amgb_dist_make_view(self.ctx._h, self._h, int(view_size), _p(s, c_f64p), _p(c, c_i64p),
                                            _p(pp, c_f64p), _p(npv, c_f64p), C.byref(t))
        _chk(self.ctx._h, rc, "amgb_dist_make_view")
        return s, c, pp, npv, t.value

    def close(self):
        if self._h:
            amgb_lib().amgb_dist_matrix_destroy(self._h)
            self._h = C.c_void_p()


class DistPreconditionBoomerAMG:
    def __init__(self):
        self._h = C.c_void_p()
        self.ctx = None

    def initialize(self, matrix, data=None):
        self.close()
        s = (data or AdditionalData()).to_struct()
        self.ctx = matrix.ctx
        rc = amgb_lib().amgb_dist_precond_initialize(self.ctx._h, matrix._h, C.byref(s), C.byref(self._h))
        _chk(self.ctx._h, rc, "amgb_dist_precond_initialize")

    def close(self):
        if self._h:
            amgb_lib().amgb_precond_destroy(self._h)
            self._h = C.c_void_p()

    @property
    def num_levels(self):
        v = C.c_int32()
        _chk(self.ctx._h, amgb_lib().amgb_precond_num_levels(self._h, C.byref(v)), "num_levels")
        return v.value

    def level_stats(self):
        cap = 64
        nl = C.c_int32()
        rows = np.empty(cap, dtype=np.int64)
        nnz = np.empty(cap, dtype=np.int64)
        sp = np.empty(cap)
        g, o, m = C.c_double(), C.c_double(), C.c_double()
        _chk(self.ctx._h, amgb_lib().amgb_precond_level_stats(
            self._h, cap, C.byref(nl), _p(rows, c_i64p), _p(nnz, c_i64p), _p(sp, c_f64p),
            C.byref(g), C.byref(o), C.byref(m)), "level_stats")
        k = nl.value
        return dict(rows=rows[:k].copy(), nnz=nnz[:k].copy(), sparsity=sp[:k].copy(), grid=g.value,
                    operator=o.value, memory=m.value)

    @property
    def replicated_from(self):
        """First level held whole on every rank (= num_levels if none)."""
        v = C.c_int32()
        _chk(self.ctx._h, amgb_lib().amgb_dist_precond_replicated_from(self._h, C.byref(v)), "replicated_from")
        return v.value

    def full_level(self, level):
        """(A, cf, P) of a replicated level through the single-device accessors."""
        L = amgb_lib()
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        _chk(self.ctx._h, L.amgb_precond_level_dims(self._h, level, C.byref(a), C.byref(b), C.byref(c), C.byref(d)),
             "level_dims")
        n, nnz, nc, nnzp = a.value, b.value, c.value, d.value
        rp, cl, vl = np.empty(n + 1, dtype=np.int32), np.empty(nnz, dtype=np.int32), np.empty(nnz)
        _chk(self.ctx._h, L.amgb_precond_get_A_csr(self._h, level, _p(rp, c_i32p), _p(cl, c_i32p), _p(vl, c_f64p)),
             "get_A_csr")
        out = dict(A=(rp, cl, vl))
        if level + 1 < self.num_levels:
            cf = np.empty(n, dtype=np.int32)
            _chk(self.ctx._h, L.amgb_precond_get_cf_marker(self._h, level, _p(cf, c_i32p)), "get_cf_marker")
            prp, pcl, pvl = np.empty(n + 1, dtype=np.int32), np.empty(nnzp, dtype=np.int32), np.empty(nnzp)
            _chk(self.ctx._h, L.amgb_precond_get_P_csr(self._h, level, _p(prp, c_i32p), _p(pcl, c_i32p),
                                                       _p(pvl, c_f64p)), "get_P_csr")
            out.update(cf=cf, P=(prp, pcl, pvl, nc))
        return out

    def level_dims(self, level):
        v = [C.c_int64() for _ in range(7)]
        _chk(self.ctx._h, amgb_lib().amgb_dist_precond_level_dims(self._h, level, *[C.byref(x) for x in v]),
             "dist level_dims")
        keys = ("n_global", "row_begin", "n_local", "nnz_local", "n_coarse_global", "coarse_begin", "nnz_P_local")
        return dict(zip(keys, (x.value for x in v)))

    def cf_marker(self, level):
        d = self.level_dims(level)
        out = np.empty(d["n_local"], dtype=np.int32)
        _chk(self.ctx._h, amgb_lib().amgb_dist_precond_get_cf_marker(self._h, level, _p(out, c_i32p)), "dist cf")
        return out

    def A_rows(self, level):
        d = self.level_dims(level)
        rp = np.empty(d["n_local"] + 1, dtype=np.int32)
        cl = np.empty(d["nnz_local"], dtype=np.int32)
        vl = np.empty(d["nnz_local"])
        _chk(self.ctx._h, amgb_lib().amgb_dist_precond_get_A_rows(self._h, level, _p(rp, c_i32p), _p(cl, c_i32p),
                                                                  _p(vl, c_f64p)), "dist A rows")
        return rp, cl, vl

    def P_rows(self, level):
        d = self.level_dims(level)
        rp = np.empty(d["n_local"] + 1, dtype=np.int32)
        cl = np.empty(d["nnz_P_local"], dtype=np.int32)
        vl = np.empty(d["nnz_P_local"])
        _chk(self.ctx._h, amgb_lib().amgb_dist_precond_get_P_rows(self._h, level, _p(rp, c_i32p), _p(cl, c_i32p),
                                                                  _p(vl, c_f64p)), "dist P rows")
        return rp, cl, vl


class DistSolverCG:
    def __init__(self, solver_control):
        self.control = solver_control

    def solve(self, A, x_local, b_local, preconditioner):
        assert x_local.dtype == np.float64 and x_local.flags.c_contiguous
        b = np.ascontiguousarray(b_local, dtype=np.float64)
        cap = min(self.control.max_steps, 65535) + 1   # the library keeps at most 65 536 entries
        hist = np.zeros(cap)
        nit = C.c_int64()
        rc = amgb_lib().amgb_dist_cg_solve(A.ctx._h, _p(x_local, c_f64p), _p(b, c_f64p), preconditioner._h,
                                           self.control.max_steps, self.control.tol, _p(hist, c_f64p), cap,
                                           C.byref(nit))
        k = min(cap, nit.value + 1)
        self.control._last_step = nit.value
        self.control.history = hist[:k].copy()
        self.control._last_value = float(hist[k - 1]) if k else float("nan")
        _chk(A.ctx._h, rc, "amgb_dist_cg_solve")


def run_local_group(nranks, fn, device_ids=None):
    """Run fn(rank, comm) on nranks host threads that form one in-process communicator
    (ctypes releases the GIL inside the library calls).  Returns the list of results;
    re-raises the first exception."""
    group = LocalGroup(nranks)
    out, err = [None] * nranks, [None] * nranks

    def body(r):
        ctx = comm = None
        try:
            ctx = Context(device_ids[r] if device_ids else 0)
            comm = Communicator.local(group, r, ctx)
            out[r] = fn(r, comm)
        except BaseException as e:  # noqa: BLE001 - reported to the caller below
            err[r] = e
            group.abort()
        finally:
            if comm:
                comm.close()
            if ctx:
                ctx.close()

    ts = [threading.Thread(target=body, args=(r,)) for r in range(nranks)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    group.close()
    for e in err:
        if e is not None:
            raise e
    return out
