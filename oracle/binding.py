"""ctypes front end of the CPU oracle (oracle/amg_oracle.h).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)
c_f64p = C.POINTER(C.c_double)


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        vp = C.c_void_p

        def sig(name, res, *args):
            f = getattr(L, name)
            f.restype = res
            f.argtypes = list(args)

        sig("orc_strength", C.c_int, C.c_int64, c_i32p, c_i32p, c_f64p, C.c_double, C.c_double, c_u8p)
        sig("orc_hypre_rand", C.c_double, C.c_int64)
        sig("orc_coarsen_pmis", C.c_int, C.c_int64, c_i32p, c_i32p, c_u8p, c_i32p)
        sig("orc_coarsen_falgout", C.c_int, C.c_int64, c_i32p, c_i32p, c_u8p, c_i32p)
        sig("orc_coarsen_cljp", C.c_int, C.c_int64, c_i32p, c_i32p, c_u8p, c_i32p)
        sig("orc_setup", C.c_int, C.c_int64, c_i32p, c_i32p, c_f64p, vp, C.POINTER(vp))
        sig("orc_destroy", None, vp)
        sig("orc_num_levels", C.c_int, vp)
        sig("orc_level_dims", C.c_int, vp, C.c_int, c_i64p, c_i64p, c_i64p, c_i64p)
        sig("orc_get_strength_mask", C.c_int, vp, C.c_int, c_u8p)
        sig("orc_get_cf_marker", C.c_int, vp, C.c_int, c_i32p)
        sig("orc_get_colors", C.c_int, vp, C.c_int, c_i32p, c_i32p)
        sig("orc_get_A_csr", C.c_int, vp, C.c_int, c_i32p, c_i32p, c_f64p)
        sig("orc_get_P_csr", C.c_int, vp, C.c_int, c_i32p, c_i32p, c_f64p)
        sig("orc_level_stats", C.c_int, vp, C.c_int, c_i32p, c_i64p, c_i64p, c_f64p, c_f64p,
            c_f64p, c_f64p)
        sig("orc_effective_relax", C.c_int, vp, c_i32p, c_i32p, c_i32p)
        sig("orc_level_cheby", C.c_int, vp, C.c_int, c_f64p, c_f64p, c_f64p, c_i32p)
        sig("orc_tql1", C.c_int, C.c_int32, c_f64p, c_f64p)
        sig("orc_vmult", C.c_int, vp, c_f64p, c_f64p)
        sig("orc_cg_solve", C.c_int, vp, C.c_int64, c_i32p, c_i32p, c_f64p, c_f64p, c_f64p,
            C.c_int64, C.c_double, c_f64p, C.c_int64, c_i64p)
        sig("orc_spmv", C.c_int, C.c_int64, c_i32p, c_i32p, c_f64p, c_f64p, c_f64p)
        sig("orc_make_view", C.c_int, C.c_int64, c_i32p, c_i32p, c_f64p, C.c_int32, c_f64p,
            c_i64p, c_f64p, c_f64p)
        sig("orc_option_roundtrip", C.c_double, C.c_double)
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t)


def _csr32(rowptr, col, val):
    rp = np.ascontiguousarray(rowptr, dtype=np.int32)
    cl = np.ascontiguousarray(col, dtype=np.int32)
    vl = np.ascontiguousarray(val, dtype=np.float64)
    return rp, cl, vl


def strength(rowptr, col, val, theta, max_row_sum=0.9):
    rp, cl, vl = _csr32(rowptr, col, val)
    mask = np.empty(len(cl), dtype=np.uint8)
    lib().orc_strength(len(rp) - 1, _p(rp, c_i32p), _p(cl, c_i32p), _p(vl, c_f64p), theta,
                       max_row_sum, _p(mask, c_u8p))
    return mask


def coarsen(rowptr, col, mask, kind="pmis"):
    rp = np.ascontiguousarray(rowptr, dtype=np.int32)
    cl = np.ascontiguousarray(col, dtype=np.int32)
    mk = np.ascontiguousarray(mask, dtype=np.uint8)
    cf = np.empty(len(rp) - 1, dtype=np.int32)
    fn = {"pmis": lib().orc_coarsen_pmis, "cljp": lib().orc_coarsen_cljp}.get(kind, lib().orc_coarsen_falgout)
    fn(len(rp) - 1, _p(rp, c_i32p), _p(cl, c_i32p), _p(mk, c_u8p), _p(cf, c_i32p))
    return cf


def spmv(rowptr, col, val, x):
    rp, cl, vl = _csr32(rowptr, col, val)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty(len(rp) - 1, dtype=np.float64)
    lib().orc_spmv(len(rp) - 1, _p(rp, c_i32p), _p(cl, c_i32p), _p(vl, c_f64p), _p(x, c_f64p),
                   _p(y, c_f64p))
    return y


def make_view(rowptr, col, val, view_size):
    rp, cl, vl = _csr32(rowptr, col, val)
    vv = view_size * view_size
    s = np.empty(vv)
    cnt = np.empty(vv, dtype=np.int64)
    pp = np.empty(vv)
    np_ = np.empty(vv)
    rc = lib().orc_make_view(len(rp) - 1, _p(rp, c_i32p), _p(cl, c_i32p), _p(vl, c_f64p),
                             view_size, _p(s, c_f64p), _p(cnt, c_i64p), _p(pp, c_f64p),
                             _p(np_, c_f64p))
    if rc:
        raise ValueError(f"orc_make_view -> {rc}")
    return s, cnt, pp, np_


def option_roundtrip(v):
    return lib().orc_option_roundtrip(float(v))


class Hierarchy:
    """orc_setup handle.  `data` is a ctypes amgb_boomeramg_data (by reference)."""

    def __init__(self, rowptr, col, val, data):
        self.rp, self.cl, self.vl = _csr32(rowptr, col, val)
        self.n = len(self.rp) - 1
        self._h = C.c_void_p()
        rc = lib().orc_setup(self.n, _p(self.rp, c_i32p), _p(self.cl, c_i32p),
                             _p(self.vl, c_f64p), C.byref(data), C.byref(self._h))
        if rc:
            raise RuntimeError(f"orc_setup -> {rc}")

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_levels(self):
        return lib().orc_num_levels(self._h)

    def level_dims(self, level):
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        rc = lib().orc_level_dims(self._h, level, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        if rc:
            raise IndexError(level)
        return a.value, b.value, c.value, d.value

    def strength_mask(self, level):
        _, nnz, _, _ = self.level_dims(level)
        out = np.empty(nnz, dtype=np.uint8)
        rc = lib().orc_get_strength_mask(self._h, level, _p(out, c_u8p))
        if rc:
            raise IndexError(level)
        return out

    def cf_marker(self, level):
        n, _, _, _ = self.level_dims(level)
        out = np.empty(n, dtype=np.int32)
        rc = lib().orc_get_cf_marker(self._h, level, _p(out, c_i32p))
        if rc:
            raise IndexError(level)
        return out

    def colors(self, level):
        """multicolour Gauss-Seidel: (colour of every point, number of colours)"""
        n, _, _, _ = self.level_dims(level)
        out = np.empty(n, dtype=np.int32)
        nc = C.c_int32()
        rc = lib().orc_get_colors(self._h, level, _p(out, c_i32p), C.byref(nc))
        if rc:
            raise IndexError(level)
        return out, nc.value

    def A(self, level):
        n, nnz, _, _ = self.level_dims(level)
        rp = np.empty(n + 1, dtype=np.int32)
        cl = np.empty(nnz, dtype=np.int32)
        vl = np.empty(nnz)
        lib().orc_get_A_csr(self._h, level, _p(rp, c_i32p), _p(cl, c_i32p), _p(vl, c_f64p))
        return rp, cl, vl

    def P(self, level):
        n, _, nc, nnzp = self.level_dims(level)
        rp = np.empty(n + 1, dtype=np.int32)
        cl = np.empty(nnzp, dtype=np.int32)
        vl = np.empty(nnzp)
        rc = lib().orc_get_P_csr(self._h, level, _p(rp, c_i32p), _p(cl, c_i32p), _p(vl, c_f64p))
        if rc:
            raise IndexError(level)
        return rp, cl, vl, nc

    def stats(self):
        cap = 64
        nl = C.c_int32()
        rows = np.empty(cap, dtype=np.int64)
        nnz = np.empty(cap, dtype=np.int64)
        sp = np.empty(cap)
        g, o, m = C.c_double(), C.c_double(), C.c_double()
        lib().orc_level_stats(self._h, cap, C.byref(nl), _p(rows, c_i64p), _p(nnz, c_i64p),
                              _p(sp, c_f64p), C.byref(g), C.byref(o), C.byref(m))
        k = nl.value
        return dict(rows=rows[:k].copy(), nnz=nnz[:k].copy(), sparsity=sp[:k].copy(),
                    grid=g.value, operator=o.value, memory=m.value)

    def effective_relax(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        lib().orc_effective_relax(self._h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def level_cheby(self, level):
        """(max_eig, min_eig, coefficients) of the Chebyshev smoother of a level"""
        mx, mn, k = C.c_double(), C.c_double(), C.c_int32()
        co = np.zeros(5)
        rc = lib().orc_level_cheby(self._h, level, C.byref(mx), C.byref(mn), _p(co, c_f64p), C.byref(k))
        if rc != 0:
            raise RuntimeError(f"orc_level_cheby -> {rc}")
        return mx.value, mn.value, co[:k.value].copy()

    def vmult(self, r):
        r = np.ascontiguousarray(r, dtype=np.float64)
        z = np.empty_like(r)
        lib().orc_vmult(self._h, _p(z, c_f64p), _p(r, c_f64p))
        return z

    def cg_solve(self, b, x0, max_steps=None, abs_tol=1e-8):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.array(x0, dtype=np.float64, copy=True)
        if max_steps is None:
            max_steps = self.n
        cap = int(min(max_steps, 100000)) + 1
        hist = np.zeros(cap)
        nit = C.c_int64()
        rc = lib().orc_cg_solve(self._h, self.n, _p(self.rp, c_i32p), _p(self.cl, c_i32p),
                                _p(self.vl, c_f64p), _p(x, c_f64p), _p(b, c_f64p), max_steps,
                                abs_tol, _p(hist, c_f64p), cap, C.byref(nit))
        return rc, x, nit.value, hist[:min(cap, nit.value + 1)].copy()


def tql1(diag, offdiag):
    """Eigenvalues (ascending) of the symmetric tridiagonal matrix with diagonal `diag` and
    sub-diagonal offdiag[1:] (EISPACK tql1 argument convention)."""
    d = np.array(diag, dtype=np.float64, copy=True)
    e = np.array(offdiag, dtype=np.float64, copy=True)
    rc = lib().orc_tql1(len(d), _p(d, c_f64p), _p(e, c_f64p))
    if rc != 0:
        raise RuntimeError(f"tql1 did not converge at eigenvalue {rc}")
    return d
