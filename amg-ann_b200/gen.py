"""Synthetic FE systems of the BASELINE.json configs (numpy front end of
include/amgb_gen.h; the arithmetic is in csrc/gen_q1.cpp)."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._native import gen_lib, c_f64p, c_i32p, c_i64p


def _p(a, t):
    return a.ctypes.data_as(t)


@dataclass
class System:
    """One linear system in the reference's conventions (CSR, fp64/int32)."""
    n: int
    rowptr: np.ndarray   # int64, local (starts at 0)
    col: np.ndarray      # int32, global column ids
    val: np.ndarray      # float64
    rhs: np.ndarray
    x0: np.ndarray       # zero + Dirichlet values (ref t2 main.cpp:319,446)
    row_begin: int = 0
    meta: dict = None

    @property
    def nnz(self):
        return int(self.rowptr[-1])

    def rowptr32(self):
        assert self.nnz < 2**31
        return self.rowptr.astype(np.int32)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.val, self.col, self.rowptr), shape=(len(self.rowptr) - 1, self.n))


def sizes(kind, m):
    n, nnz = C.c_int64(), C.c_int64()
    rc = gen_lib().amgb_gen_sizes(kind, m, C.byref(n), C.byref(nnz))
    if rc:
        raise ValueError(f"amgb_gen_sizes({kind},{m}) -> {rc}")
    return n.value, nnz.value


def checkerboard_epsv(pattern_size, mode, contrast_exp):
    out = np.empty(pattern_size ** mode, dtype=np.float64)
    rc = gen_lib().amgb_gen_checkerboard_epsv(pattern_size, mode, float(contrast_exp), _p(out, c_f64p))
    if rc:
        raise ValueError("amgb_gen_checkerboard_epsv")
    return out


def random_vec(seed, length, vmax):
    out = np.empty(length, dtype=np.float64)
    rc = gen_lib().amgb_gen_random_vec(int(seed), int(length), float(vmax), _p(out, c_f64p))
    if rc:
        raise ValueError("amgb_gen_random_vec")
    return out


def poisson_q1(m, pattern_size=1, mode=1, epsv=None, row_begin=0, row_end=None, want_rhs=True):
    """Q1 Poisson / piecewise-constant diffusion on an m^3 cell mesh of [-1,1]^3.

    epsv: exponents, mu = 10^epsv on a pattern_size^mode pattern (None -> mu = 1).
    """
    if epsv is None:
        epsv = np.zeros(pattern_size ** mode)
    epsv = np.ascontiguousarray(epsv, dtype=np.float64)
    n, _ = sizes(0, m)
    if row_end is None:
        row_end = n
    nnz = C.c_int64()
    rc = gen_lib().amgb_gen_poisson_q1_range_sizes(m, row_begin, row_end, C.byref(nnz))
    if rc:
        raise ValueError("bad row range")
    nloc = row_end - row_begin
    rowptr = np.empty(nloc + 1, dtype=np.int64)
    col = np.empty(nnz.value, dtype=np.int32)
    val = np.empty(nnz.value, dtype=np.float64)
    rhs = np.empty(nloc, dtype=np.float64) if want_rhs else None
    x0 = np.empty(nloc, dtype=np.float64) if want_rhs else None
    rc = gen_lib().amgb_gen_poisson_q1(
        m, pattern_size, mode, _p(epsv, c_f64p), len(epsv), row_begin, row_end,
        _p(rowptr, c_i64p), _p(col, c_i32p), _p(val, c_f64p),
        _p(rhs, c_f64p) if want_rhs else None, _p(x0, c_f64p) if want_rhs else None)
    if rc:
        raise ValueError(f"amgb_gen_poisson_q1 -> {rc}")
    return System(n, rowptr, col, val, rhs, x0, row_begin,
                  dict(kind="poisson_q1", m=m, pattern_size=pattern_size, mode=mode))


def elasticity_q1(m, pattern_size=1, mode=1, young=None):
    """Q1 vector elasticity (3 DoF/node, interleaved), nu = 0.29 (ref t3 main.cpp:48-49)."""
    if young is None:
        young = np.ones(pattern_size ** mode)
    young = np.ascontiguousarray(young, dtype=np.float64)
    n, nnz = sizes(1, m)
    rowptr = np.empty(n + 1, dtype=np.int64)
    col = np.empty(nnz, dtype=np.int32)
    val = np.empty(nnz, dtype=np.float64)
    rhs = np.empty(n, dtype=np.float64)
    x0 = np.empty(n, dtype=np.float64)
    rc = gen_lib().amgb_gen_elasticity_q1(m, pattern_size, mode, _p(young, c_f64p), len(young),
                                          _p(rowptr, c_i64p), _p(col, c_i32p), _p(val, c_f64p),
                                          _p(rhs, c_f64p), _p(x0, c_f64p))
    if rc:
        raise ValueError(f"amgb_gen_elasticity_q1 -> {rc}")
    assert rowptr[-1] == nnz
    return System(n, rowptr, col, val, rhs, x0, 0,
                  dict(kind="elasticity_q1", m=m, pattern_size=pattern_size, mode=mode))


def theta_sweep(start, stop, step):
    """theta values exactly as the reference loop builds them: repeated fp
    addition `for (t = start; t <= stop; t += step)` (ref t2 main.cpp:443)."""
    out = []
    t = float(start)
    while t <= stop:
        out.append(t)
        t += step
    return out


def cuthill_mckee(rowptr, col, reversed=False):
    """deal.II-style Cuthill-McKee renumbering (include/amgb_gen.h): new -> old."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    n = len(rowptr) - 1
    out = np.empty(n, dtype=np.int32)
    rc = gen_lib().amgb_gen_cuthill_mckee(n, _p(rowptr, c_i64p), _p(col, c_i32p), int(bool(reversed)), _p(out, c_i32p))
    if rc:
        raise ValueError(f"amgb_gen_cuthill_mckee -> {rc}")
    return out
