"""amg-ann_b200: B200-native drop-in for the AMG-PCG theta-sweep hot path of
MatteoCaldana/AMG-ANN (see DESIGN.md for scope, include/amgb.h for the C ABI).

This Python layer is a thin mirror of the reference-facing interface on top of
the C ABI, used by the tests and bench.py; the host-side product code for the
reference's C++ drivers is host/amgb_shim.hpp.  Names follow the reference:

  AdditionalData            deal.II PreconditionBoomerAMG::AdditionalData
                            (ref common/amg_solver.h:20, t2 main.cpp:447-453)
  PreconditionBoomerAMG     .initialize(A, data) / .vmult(dst, src)
                            (ref common/amg_solver.h:39,48)
  SolverControl, SolverCG   (ref common/amg_solver.h:33,38,54)
  ViewMaker                 (ref common/view_maker.h:17-92)
  amg_solve                 (ref common/amg_solver.h:22-92)

Import name: the directory is `amg-ann_b200`; import it as `amg_ann_b200`
through the loader module at the repository root.
"""
import ctypes as C
import time

import numpy as np

from . import gen
from ._native import (AMGB_SYMBOLS, BoomerAMGDataStruct, amgb_lib, amgb_lib_path, c_f64p,
                      c_i32p, c_i64p, c_u8p, gen_lib)

__all__ = [
    "AdditionalData", "RelaxationType", "Context", "SparseMatrix", "PreconditionBoomerAMG",
    "SolverControl", "SolverCG", "NoConvergence", "ViewMaker", "amg_solve", "AmgbError", "gen",
]

AMGB_OK = 0
AMGB_ERR_NO_CONVERGENCE = -6


class AmgbError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        name = amgb_lib().amgb_status_string(status).decode()
        super().__init__(f"{where}: {name} ({status}) {detail}")


class NoConvergence(AmgbError):
    """deal.II SolverControl::NoConvergence equivalent."""


class RelaxationType:
    """deal.II RelaxationType enum, same order (include/amgb.h)."""
    Jacobi = 0
    sequentialGaussSeidel = 1
    seqboundaryGaussSeidel = 2
    SORJacobi = 3
    backwardSORJacobi = 4
    symmetricSORJacobi = 5
    l1scaledSORJacobi = 6
    GaussianElimination = 7
    l1GaussSeidel = 8
    backwardl1GaussSeidel = 9
    CG = 10
    Chebyshev = 11
    FCFJacobi = 12
    l1scaledJacobi = 13
    none = 14


COARSEN_CLJP, COARSEN_FALGOUT, COARSEN_PMIS = 0, 6, 8
INTERP_CLASSICAL = 0
SMOOTHER_SUBSTITUTE, SMOOTHER_STRICT, SMOOTHER_MULTICOLOR = 0, 1, 2


class AdditionalData:
    """Same positional order as deal.II's constructor; the
    reference passes the first five.  Keyword-only extras are the PCHYPRE
    defaults made explicit (A.2)."""

    def __init__(self, symmetric_operator=False, strong_threshold=0.25, max_row_sum=0.9,
                 aggressive_coarsening_num_levels=0, output_details=False,
                 relaxation_type_up=RelaxationType.SORJacobi,
                 relaxation_type_down=RelaxationType.SORJacobi,
                 relaxation_type_coarse=RelaxationType.GaussianElimination,
                 n_sweeps_coarse=1, tol=0.0, max_iter=1, w_cycle=False, *,
                 coarsen_type=COARSEN_PMIS, interp_type=INTERP_CLASSICAL, relax_order=1,
                 n_sweeps=1, max_levels=25, max_coarse_size=9, relax_weight=1.0,
                 smoother_policy=SMOOTHER_SUBSTITUTE, options_via_string=True,
                 keep_setup_intermediates=False, dist_replicate_below=262144):
        self.symmetric_operator = bool(symmetric_operator)
        self.strong_threshold = float(strong_threshold)
        self.max_row_sum = float(max_row_sum)
        self.aggressive_coarsening_num_levels = int(aggressive_coarsening_num_levels)
        self.output_details = bool(output_details)
        self.relaxation_type_up = int(relaxation_type_up)
        self.relaxation_type_down = int(relaxation_type_down)
        self.relaxation_type_coarse = int(relaxation_type_coarse)
        self.n_sweeps_coarse = int(n_sweeps_coarse)
        self.tol = float(tol)
        self.max_iter = int(max_iter)
        self.w_cycle = bool(w_cycle)
        self.coarsen_type = int(coarsen_type)
        self.interp_type = int(interp_type)
        self.relax_order = int(relax_order)
        self.n_sweeps = int(n_sweeps)
        self.max_levels = int(max_levels)
        self.max_coarse_size = int(max_coarse_size)
        self.relax_weight = float(relax_weight)
        self.smoother_policy = int(smoother_policy)
        self.options_via_string = bool(options_via_string)
        self.keep_setup_intermediates = bool(keep_setup_intermediates)
        self.dist_replicate_below = int(dist_replicate_below)

    def to_struct(self):
        s = BoomerAMGDataStruct()
        for name, _ in BoomerAMGDataStruct._fields_:
            if name == "reserved":
                continue
            setattr(s, name, int(getattr(self, name)) if not isinstance(getattr(self, name), float)
                    else getattr(self, name))
        return s


def _p(a, t):
    return a.ctypes.data_as(t)


def _chk(ctx_handle, rc, where):
    if rc != AMGB_OK:
        detail = ""
        if ctx_handle:
            detail = amgb_lib().amgb_last_error(ctx_handle).decode()
        if rc == AMGB_ERR_NO_CONVERGENCE:
            raise NoConvergence(rc, where, detail)
        raise AmgbError(rc, where, detail)


class Context:
    """amgb_ctx: one CUDA device + stream.  stream: raw cudaStream_t (int) or None."""

    def __init__(self, device_id=0, stream=None):
        self._h = C.c_void_p()
        rc = amgb_lib().amgb_ctx_create(C.byref(self._h), device_id, C.c_void_p(stream or 0))
        if rc != AMGB_OK:
            self._h = C.c_void_p()
            raise AmgbError(rc, "amgb_ctx_create",
                            "(a CUDA device is required: this library has no CPU path)")

    def close(self):
        if self._h:
            amgb_lib().amgb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _chk(self._h, amgb_lib().amgb_ctx_synchronize(self._h), "amgb_ctx_synchronize")

    def reserve(self, nbytes):
        """Grow the context's memory pool to nbytes up front (include/amgb.h amgb_ctx_reserve)."""
        _chk(self._h, amgb_lib().amgb_ctx_reserve(self._h, int(nbytes)), "amgb_ctx_reserve")

    def kernel_launches(self):
        v = C.c_int64()
        _chk(self._h, amgb_lib().amgb_ctx_kernel_launches(self._h, C.byref(v)), "kernel_launches")
        return v.value

    def reset_kernel_launches(self):
        amgb_lib().amgb_ctx_reset_kernel_launches(self._h)

    def enable_timers(self, on=True):
        _chk(self._h, amgb_lib().amgb_ctx_enable_timers(self._h, int(on)), "enable_timers")

    def reset_timers(self):
        amgb_lib().amgb_ctx_reset_timers(self._h)

    def timers(self):
        L = amgb_lib()
        out = {}
        for f in range(L.amgb_timer_count()):
            ms, cnt, by = C.c_double(), C.c_int64(), C.c_double()
            _chk(self._h, L.amgb_ctx_get_timer(self._h, f, C.byref(ms), C.byref(cnt), C.byref(by)),
                 "get_timer")
            out[L.amgb_timer_name(f).decode()] = dict(ms=ms.value, launches=cnt.value,
                                                       bytes=by.value)
        return out


    def routes(self):
        """{route name: times taken} since creation / reset_routes (include/amgb.h)."""
        L = amgb_lib()
        out = {}
        for r in range(L.amgb_route_count()):
            v = C.c_int64()
            _chk(self._h, L.amgb_ctx_get_route(self._h, r, C.byref(v)), "get_route")
            out[L.amgb_route_name(r).decode()] = v.value
        return out

    def reset_routes(self):
        amgb_lib().amgb_ctx_reset_routes(self._h)

    def level_timers(self, max_levels=32):
        """{family: [dict(ms, launches, bytes) per level]} for launches timed with timers on."""
        L = amgb_lib()
        out = {}
        for f in range(L.amgb_timer_count()):
            rows = []
            for lv in range(max_levels):
                ms, cnt, by = C.c_double(), C.c_int64(), C.c_double()
                L.amgb_ctx_get_timer_level(self._h, f, lv, C.byref(ms), C.byref(cnt), C.byref(by))
                rows.append(dict(ms=ms.value, launches=cnt.value, bytes=by.value))
            while rows and rows[-1]["launches"] == 0:
                rows.pop()
            if rows:
                out[L.amgb_timer_name(f).decode()] = rows
        return out


class SparseMatrix:
    """Device-resident CSR system matrix; stays resident across the theta sweep
    (the reference re-converts it for every theta)."""

    def __init__(self, ctx, rowptr, col, val):
        self.ctx = ctx
        self.rowptr = np.ascontiguousarray(rowptr)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.val = np.ascontiguousarray(val, dtype=np.float64)
        self.n = len(self.rowptr) - 1
        self._h = C.c_void_p()
        L = amgb_lib()
        if self.rowptr.dtype == np.int32:
            rc = L.amgb_matrix_upload_csr(ctx._h, self.n, _p(self.rowptr, c_i32p),
                                          _p(self.col, c_i32p), _p(self.val, c_f64p),
                                          C.byref(self._h))
        else:
            self.rowptr = np.ascontiguousarray(self.rowptr, dtype=np.int64)
            rc = L.amgb_matrix_upload_csr64(ctx._h, self.n, _p(self.rowptr, c_i64p),
                                            _p(self.col, c_i32p), _p(self.val, c_f64p),
                                            C.byref(self._h))
        _chk(ctx._h, rc, "amgb_matrix_upload_csr")

    @classmethod
    def wrap_device(cls, ctx, n, nnz, rowptr_ptr, col_ptr, val_ptr):
        self = cls.__new__(cls)
        self.ctx, self.n = ctx, n
        self._h = C.c_void_p()
        rc = amgb_lib().amgb_matrix_wrap_device_csr(ctx._h, n, nnz, C.c_void_p(rowptr_ptr),
                                                    C.c_void_p(col_ptr), C.c_void_p(val_ptr),
                                                    C.byref(self._h))
        _chk(ctx._h, rc, "amgb_matrix_wrap_device_csr")
        return self

    @classmethod
    def assemble_poisson_q1(cls, ctx, m, pattern_size=1, mode=1, epsv=None, rhs_ptr=0, x0_ptr=0):
        """On-device assembly (ref t2 main.cpp:255-320), bit-identical to gen.poisson_q1.
        rhs_ptr / x0_ptr: device pointers to (m+1)^3 doubles (e.g. torch data_ptr()), or 0."""
        if epsv is None:
            epsv = np.zeros(pattern_size ** mode)
        epsv = np.ascontiguousarray(epsv, dtype=np.float64)
        self = cls.__new__(cls)
        self.ctx, self.n = ctx, (m + 1) ** 3
        self._h = C.c_void_p()
        rc = amgb_lib().amgb_matrix_assemble_poisson_q1(ctx._h, m, pattern_size, mode, _p(epsv, c_f64p), len(epsv),
                                                        C.byref(self._h), C.c_void_p(rhs_ptr), C.c_void_p(x0_ptr))
        _chk(ctx._h, rc, "amgb_matrix_assemble_poisson_q1")
        return self

    @classmethod
    def assemble_elasticity_q1(cls, ctx, m, pattern_size=1, mode=1, young=None, rhs_ptr=0, x0_ptr=0):
        """On-device assembly of the Q1 elasticity system (ref t3 main.cpp:320-342): matrix and x0
        bit-identical to gen.elasticity_q1, rhs to rounding.  rhs_ptr / x0_ptr: device pointers to
        3 (m+1)^3 doubles, or 0."""
        if young is None:
            young = np.ones(pattern_size ** mode)
        young = np.ascontiguousarray(young, dtype=np.float64)
        self = cls.__new__(cls)
        self.ctx, self.n = ctx, 3 * (m + 1) ** 3
        self._h = C.c_void_p()
        rc = amgb_lib().amgb_matrix_assemble_elasticity_q1(ctx._h, m, pattern_size, mode, _p(young, c_f64p), len(young),
                                                           C.byref(self._h), C.c_void_p(rhs_ptr), C.c_void_p(x0_ptr))
        _chk(ctx._h, rc, "amgb_matrix_assemble_elasticity_q1")
        return self

    def permuted(self, new_to_old):
        """Q A Q^T on the device for a DoF renumbering given as new -> old (a permutation)."""
        perm = np.ascontiguousarray(new_to_old, dtype=np.int32)
        assert len(perm) == self.n
        out = SparseMatrix.__new__(SparseMatrix)
        out.ctx, out.n = self.ctx, self.n
        out._h = C.c_void_p()
        rc = amgb_lib().amgb_matrix_permute(self.ctx._h, self._h, _p(perm, c_i32p), C.byref(out._h))
        _chk(self.ctx._h, rc, "amgb_matrix_permute")
        return out

    def download(self):
        n, nnz = C.c_int64(), C.c_int64()
        amgb_lib().amgb_matrix_dims(self._h, C.byref(n), C.byref(nnz))
        rp = np.empty(n.value + 1, dtype=np.int32)
        cl = np.empty(nnz.value, dtype=np.int32)
        vl = np.empty(nnz.value)
        _chk(self.ctx._h, amgb_lib().amgb_matrix_download_csr(self._h, _p(rp, c_i32p), _p(cl, c_i32p),
                                                              _p(vl, c_f64p)), "amgb_matrix_download_csr")
        return rp, cl, vl

    def m(self):
        return self.n

    def close(self):
        if self._h:
            amgb_lib().amgb_matrix_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def vmult(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.n)
        _chk(self.ctx._h, amgb_lib().amgb_matrix_vmult(self.ctx._h, self._h, _p(y, c_f64p),
                                                       _p(x, c_f64p)), "amgb_matrix_vmult")
        return y


def numbering_dealii_q1(ctx, coarse_cells, refinements):
    """new -> lexicographic node id of deal.II's distribute_dofs numbering for Q1 on
    subdivided_hyper_cube(coarse_cells) refined `refinements` times (include/amgb.h)."""
    m = coarse_cells << refinements
    out = np.empty((m + 1) ** 3, dtype=np.int32)
    _chk(ctx._h, amgb_lib().amgb_numbering_dealii_q1(ctx._h, coarse_cells, refinements, _p(out, c_i32p)),
         "amgb_numbering_dealii_q1")
    return out


class PreconditionBoomerAMG:
    """ref common/amg_solver.h:39,48 -- `preconditioner.initialize(A, data)`."""

    def __init__(self):
        self._h = C.c_void_p()
        self.ctx = None

    def initialize(self, matrix, data=None, ctx=None):
        """ctx: build (and later apply) the hierarchy on another context of the same device
        than the one that uploaded the matrix.  A matrix is read-only once uploaded, so the
        independent systems of a theta sweep can be in flight on several streams at once."""
        self.close()
        data = data or AdditionalData()
        s = data.to_struct()
        self.ctx = ctx or matrix.ctx
        rc = amgb_lib().amgb_precond_initialize(self.ctx._h, matrix._h, C.byref(s),
                                                C.byref(self._h))
        _chk(self.ctx._h, rc, "amgb_precond_initialize")

    def close(self):
        if self._h:
            amgb_lib().amgb_precond_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def vmult(self, dst, src):
        src = np.ascontiguousarray(src, dtype=np.float64)
        assert dst.dtype == np.float64 and dst.flags.c_contiguous
        _chk(self.ctx._h, amgb_lib().amgb_precond_vmult(self._h, _p(dst, c_f64p), _p(src, c_f64p)),
             "amgb_precond_vmult")

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
        return dict(rows=rows[:k].copy(), nnz=nnz[:k].copy(), sparsity=sp[:k].copy(),
                    grid=g.value, operator=o.value, memory=m.value)

    def level_row_stats(self, level):
        """(min, max entries per row, min, max row sum) of a level: hypre's stats table."""
        a, b = C.c_int32(), C.c_int32()
        c, d = C.c_double(), C.c_double()
        _chk(self.ctx._h, amgb_lib().amgb_precond_level_row_stats(self._h, level, C.byref(a), C.byref(b),
                                                                  C.byref(c), C.byref(d)), "level_row_stats")
        return a.value, b.value, c.value, d.value

    def level_cheby(self, level):
        """(max_eig, min_eig, coefficients) of the Chebyshev smoother of a level."""
        mx, mn, k = C.c_double(), C.c_double(), C.c_int32()
        co = np.zeros(5)
        _chk(self.ctx._h, amgb_lib().amgb_precond_level_cheby(self._h, level, C.byref(mx), C.byref(mn),
                                                              _p(co, c_f64p), C.byref(k)), "level_cheby")
        return mx.value, mn.value, co[:k.value].copy()

    def effective_relax(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        _chk(self.ctx._h, amgb_lib().amgb_precond_effective_relax(self._h, C.byref(a), C.byref(b),
                                                                  C.byref(c)), "effective_relax")
        return a.value, b.value, c.value

    def level_dims(self, level):
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        _chk(self.ctx._h, amgb_lib().amgb_precond_level_dims(
            self._h, level, C.byref(a), C.byref(b), C.byref(c), C.byref(d)), "level_dims")
        return a.value, b.value, c.value, d.value

    def strength_mask(self, level):
        _, nnz, _, _ = self.level_dims(level)
        out = np.empty(nnz, dtype=np.uint8)
        _chk(self.ctx._h, amgb_lib().amgb_precond_get_strength_mask(self._h, level, _p(out, c_u8p)),
             "get_strength_mask")
        return out

    def cf_marker(self, level):
        n, _, _, _ = self.level_dims(level)
        out = np.empty(n, dtype=np.int32)
        _chk(self.ctx._h, amgb_lib().amgb_precond_get_cf_marker(self._h, level, _p(out, c_i32p)),
             "get_cf_marker")
        return out

    def colors(self, level):
        """multicolour Gauss-Seidel (SMOOTHER_MULTICOLOR): (colour of every point, number of colours)"""
        n, _, _, _ = self.level_dims(level)
        out = np.empty(n, dtype=np.int32)
        nc = C.c_int32()
        _chk(self.ctx._h, amgb_lib().amgb_precond_get_colors(self._h, level, _p(out, c_i32p), C.byref(nc)),
             "get_colors")
        return out, nc.value

    def A(self, level):
        n, nnz, _, _ = self.level_dims(level)
        rp = np.empty(n + 1, dtype=np.int32)
        cl = np.empty(nnz, dtype=np.int32)
        vl = np.empty(nnz)
        _chk(self.ctx._h, amgb_lib().amgb_precond_get_A_csr(self._h, level, _p(rp, c_i32p),
                                                            _p(cl, c_i32p), _p(vl, c_f64p)),
             "get_A_csr")
        return rp, cl, vl

    def P(self, level):
        n, _, nc, nnzp = self.level_dims(level)
        rp = np.empty(n + 1, dtype=np.int32)
        cl = np.empty(nnzp, dtype=np.int32)
        vl = np.empty(nnzp)
        _chk(self.ctx._h, amgb_lib().amgb_precond_get_P_csr(self._h, level, _p(rp, c_i32p),
                                                            _p(cl, c_i32p), _p(vl, c_f64p)),
             "get_P_csr")
        return rp, cl, vl, nc


class SolverControl:
    """deal.II SolverControl(max_steps, tol): `tol` is ABSOLUTE on the
    preconditioned residual (ref common/amg_solver.h:33)."""

    def __init__(self, max_steps=100, tol=1e-10):
        self.max_steps = int(max_steps)
        self.tol = float(tol)
        self._last_step = 0
        self._last_value = float("nan")
        self.history = np.zeros(0)

    def last_step(self):
        return self._last_step

    def last_value(self):
        return self._last_value


class SolverCG:
    """ref common/amg_solver.h:38,54 -- `cg.solve(A, x, b, preconditioner)`."""

    def __init__(self, solver_control):
        self.control = solver_control

    def solve(self, A, x, b, preconditioner):
        assert x.dtype == np.float64 and x.flags.c_contiguous
        b = np.ascontiguousarray(b, dtype=np.float64)
        cap = min(self.control.max_steps, 65535) + 1   # the library keeps at most 65 536 entries
        hist = np.zeros(cap)
        nit = C.c_int64()
        ctx = preconditioner.ctx or A.ctx
        rc = amgb_lib().amgb_cg_solve(ctx._h, A._h, _p(x, c_f64p), _p(b, c_f64p),
                                      preconditioner._h, self.control.max_steps,
                                      self.control.tol, _p(hist, c_f64p), cap, C.byref(nit))
        k = min(cap, nit.value + 1)
        self.control._last_step = nit.value
        self.control.history = hist[:k].copy()
        self.control._last_value = float(hist[k - 1]) if k else float("nan")
        _chk(ctx._h, rc, "amgb_cg_solve")


class ViewMaker:
    """ref common/view_maker.h:17-92."""

    def __init__(self, vs):
        self.view_size = int(vs)
        vv = self.view_size ** 2
        self.view = np.zeros(vv)
        self.count = np.zeros(vv, dtype=np.int64)
        self.max_pp = np.zeros(vv)
        self.max_np = np.zeros(vv)
        self.t_view_us = 0.0
        self.t_device_us = 0.0

    def make_view(self, A):
        t = C.c_double()
        t1 = time.perf_counter()
        rc = amgb_lib().amgb_make_view(A.ctx._h, A._h, self.view_size, _p(self.view, c_f64p),
                                       _p(self.count, c_i64p), _p(self.max_pp, c_f64p),
                                       _p(self.max_np, c_f64p), C.byref(t))
        self.t_view_us = (time.perf_counter() - t1) * 1e6
        self.t_device_us = t.value
        _chk(A.ctx._h, rc, "amgb_make_view")
        return self

    NORM_MODES = {"nothing": 0, "pure": 1, "resc": 2, "pure_log": 3, "resc_log": 4, "mean": 5}

    def make_model_input(self, A, normalization_mode="pure_log", count_channel_as_reference=True):
        """ANN-ready [V, V, 4] tensor (sum, max_pp, max_np, count channel), pooled and
        normalised on the device: ref data-modeling/train_ann.py:133-172 (norm_view) and
        :247-256 ("sum+max+c" stacking)."""
        V = self.view_size
        out = np.empty((V, V, 4))
        t = C.c_double()
        rc = amgb_lib().amgb_make_view_normalized(A.ctx._h, A._h, V, self.NORM_MODES[normalization_mode],
                                                  int(bool(count_channel_as_reference)), _p(out, c_f64p),
                                                  C.byref(t))
        self.t_device_us = t.value
        _chk(A.ctx._h, rc, "amgb_make_view_normalized")
        return out


from . import dist  # noqa: E402  (row-partitioned front end; needs the names above)


class _NoLock:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def amg_solve(data, rtol, A, b, x, ctx=None, phase_locks=None):
    """Python mirror of ref common/amg_solver.h:22-92: timed initialize + timed
    cg.solve, returning the CSV fields as a dict instead of scraping stdout.
    ctx: context (stream) to run on when several systems share one uploaded matrix.
    phase_locks: (setup_lock, solve_lock) shared by the host threads that run the systems of
    a sweep side by side on one GPU: at most one system is in its (latency- and
    instruction-bound) setup and one in its (bandwidth-bound) solve at any time, so the two
    phases that overlap are always complementary; the timed intervals exclude the waiting."""
    setup_lock, solve_lock = phase_locks if phase_locks else (_NoLock(), _NoLock())
    control = SolverControl(A.m(), rtol)
    cg = SolverCG(control)
    prec = PreconditionBoomerAMG()
    with setup_lock:
        t1 = time.perf_counter()
        prec.initialize(A, data, ctx)
        prec.ctx.synchronize()
        t2 = time.perf_counter()
    with solve_lock:
        t3 = time.perf_counter()
        cg.solve(A, x, b, prec)
        t4 = time.perf_counter()
    row = dict(theta=data.strong_threshold, maxrowsum=data.max_row_sum,
               symop=int(data.symmetric_operator),
               agg_nl=data.aggressive_coarsening_num_levels, tol=rtol,
               t_amg_setup=int((t2 - t1) * 1e6), t_solve=int((t4 - t3) * 1e6),
               niters=control.last_step(), p_res=control.history)
    if data.output_details:
        st = prec.level_stats()
        row.update(nrows=st["rows"], nze=st["nnz"], sparsity=st["sparsity"], grid=st["grid"],
                   operator=st["operator"], memory=st["memory"])
    prec.close()
    return row
