"""ctypes loaders for the native libraries of the hot path.

libamgb.so      -- the CUDA product library behind include/amgb.h (sm_100a).
libamgb_gen.so  -- host-only synthetic FE system generators (include/amgb_gen.h).

There is no Python or CPU implementation of the path behind these loaders: if
libamgb.so is missing, or no CUDA device is usable, the calls fail loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)
c_f64p = C.POINTER(C.c_double)


class BoomerAMGDataStruct(C.Structure):
    """Binary mirror of amgb_boomeramg_data (include/amgb.h)."""
    _fields_ = [
        ("symmetric_operator", C.c_int32),
        ("strong_threshold", C.c_double),
        ("max_row_sum", C.c_double),
        ("aggressive_coarsening_num_levels", C.c_uint32),
        ("output_details", C.c_int32),
        ("relaxation_type_up", C.c_int32),
        ("relaxation_type_down", C.c_int32),
        ("relaxation_type_coarse", C.c_int32),
        ("n_sweeps_coarse", C.c_uint32),
        ("tol", C.c_double),
        ("max_iter", C.c_uint32),
        ("w_cycle", C.c_int32),
        ("coarsen_type", C.c_int32),
        ("interp_type", C.c_int32),
        ("relax_order", C.c_int32),
        ("n_sweeps", C.c_uint32),
        ("max_levels", C.c_int32),
        ("max_coarse_size", C.c_int32),
        ("relax_weight", C.c_double),
        ("smoother_policy", C.c_int32),
        ("options_via_string", C.c_int32),
        ("keep_setup_intermediates", C.c_int32),
        ("dist_replicate_below", C.c_int32),
        ("reserved", C.c_int32 * 6),
    ]


def _sig(fn, restype, *argtypes):
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


_gen = None
_amgb = None


def gen_lib():
    global _gen
    if _gen is None:
        path = os.path.join(_HERE, "libamgb_gen.so")
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} not found: build it with `make -C amg-ann_b200` "
                "(or __graft_entry__.build())")
        L = C.CDLL(path)
        _sig(L.amgb_gen_sizes, C.c_int, C.c_int, C.c_int, c_i64p, c_i64p)
        _sig(L.amgb_gen_poisson_q1_range_sizes, C.c_int, C.c_int, C.c_int64, C.c_int64, c_i64p)
        _sig(L.amgb_gen_poisson_q1, C.c_int, C.c_int, C.c_int, C.c_int, c_f64p, C.c_int64,
             C.c_int64, C.c_int64, c_i64p, c_i32p, c_f64p, c_f64p, c_f64p)
        _sig(L.amgb_gen_elasticity_q1, C.c_int, C.c_int, C.c_int, C.c_int, c_f64p, C.c_int64,
             c_i64p, c_i32p, c_f64p, c_f64p, c_f64p)
        _sig(L.amgb_gen_random_vec, C.c_int, C.c_int64, C.c_int64, C.c_double, c_f64p)
        _sig(L.amgb_gen_checkerboard_epsv, C.c_int, C.c_int, C.c_int, C.c_double, c_f64p)
        _sig(L.amgb_gen_cuthill_mckee, C.c_int, C.c_int64, c_i64p, c_i32p, C.c_int, c_i32p)
        _gen = L
    return _gen


AMGB_SYMBOLS = [
    "amgb_ctx_create", "amgb_ctx_destroy", "amgb_ctx_synchronize", "amgb_ctx_reserve", "amgb_last_error",
    "amgb_status_string", "amgb_version", "amgb_ctx_kernel_launches",
    "amgb_ctx_reset_kernel_launches", "amgb_matrix_upload_csr", "amgb_matrix_upload_csr64",
    "amgb_matrix_wrap_device_csr", "amgb_matrix_destroy", "amgb_matrix_dims", "amgb_numbering_dealii_q1", "amgb_matrix_permute",
    "amgb_matrix_assemble_poisson_q1", "amgb_matrix_assemble_elasticity_q1", "amgb_matrix_assemble_poisson_q1_hostvec", "amgb_matrix_download_csr", "amgb_dist_matrix_assemble_poisson_q1",
    "amgb_matrix_vmult", "amgb_boomeramg_data_default", "amgb_precond_initialize",
    "amgb_precond_destroy", "amgb_precond_vmult", "amgb_precond_vmult_device",
    "amgb_precond_num_levels", "amgb_precond_level_stats", "amgb_precond_level_row_stats",
    "amgb_precond_effective_relax", "amgb_precond_level_cheby",
    "amgb_precond_level_dims", "amgb_precond_get_strength_mask", "amgb_precond_get_cf_marker", "amgb_precond_get_colors",
    "amgb_precond_get_A_csr", "amgb_precond_get_P_csr", "amgb_cg_solve",
    "amgb_cg_solve_device", "amgb_make_view", "amgb_make_view_normalized", "amgb_ctx_enable_timers",
    "amgb_ctx_reset_timers", "amgb_timer_count", "amgb_timer_name", "amgb_ctx_get_timer",
    "amgb_ctx_get_timer_level", "amgb_route_count", "amgb_route_name", "amgb_ctx_get_route",
    "amgb_ctx_reset_routes",
    # row-partitioned path
    "amgb_nccl_unique_id", "amgb_comm_create_nccl", "amgb_local_group_create",
    "amgb_local_group_destroy", "amgb_local_group_abort", "amgb_comm_create_local", "amgb_comm_destroy", "amgb_comm_rank",
    "amgb_comm_size", "amgb_dist_matrix_create", "amgb_dist_matrix_destroy", "amgb_dist_make_view",
    "amgb_dist_precond_initialize", "amgb_dist_cg_solve", "amgb_dist_cg_solve_device",
    "amgb_dist_precond_level_dims", "amgb_dist_precond_replicated_from", "amgb_dist_precond_get_cf_marker",
    "amgb_dist_precond_get_A_rows", "amgb_dist_precond_get_P_rows",
]


def amgb_lib_path():
    return os.path.join(_HERE, "libamgb.so")


def amgb_lib():
    """Load the CUDA product library.  Raises if it has not been built."""
    global _amgb
    if _amgb is None:
        path = amgb_lib_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} not found: the CUDA library is required (there is no CPU "
                "fallback). Build it with `make -C amg-ann_b200` or __graft_entry__.build().")
        L = C.CDLL(path)
        vp = C.c_void_p
        _sig(L.amgb_ctx_create, C.c_int, C.POINTER(vp), C.c_int, vp)
        _sig(L.amgb_ctx_destroy, C.c_int, vp)
        _sig(L.amgb_ctx_synchronize, C.c_int, vp)
        _sig(L.amgb_ctx_reserve, C.c_int, vp, C.c_int64)
        _sig(L.amgb_last_error, C.c_char_p, vp)
        _sig(L.amgb_status_string, C.c_char_p, C.c_int)
        _sig(L.amgb_version, C.c_int)
        _sig(L.amgb_ctx_kernel_launches, C.c_int, vp, c_i64p)
        _sig(L.amgb_ctx_reset_kernel_launches, C.c_int, vp)
        _sig(L.amgb_matrix_upload_csr, C.c_int, vp, C.c_int64, c_i32p, c_i32p, c_f64p,
             C.POINTER(vp))
        _sig(L.amgb_matrix_upload_csr64, C.c_int, vp, C.c_int64, c_i64p, c_i32p, c_f64p,
             C.POINTER(vp))
        _sig(L.amgb_matrix_wrap_device_csr, C.c_int, vp, C.c_int64, C.c_int64, vp, vp, vp,
             C.POINTER(vp))
        _sig(L.amgb_matrix_assemble_poisson_q1, C.c_int, vp, C.c_int32, C.c_int32, C.c_int32, c_f64p, C.c_int64,
             C.POINTER(vp), vp, vp)
        _sig(L.amgb_matrix_assemble_elasticity_q1, C.c_int, vp, C.c_int32, C.c_int32, C.c_int32, c_f64p, C.c_int64,
             C.POINTER(vp), vp, vp)
        _sig(L.amgb_matrix_assemble_poisson_q1_hostvec, C.c_int, vp, C.c_int32, C.c_int32, C.c_int32, c_f64p,
             C.c_int64, C.POINTER(vp), c_f64p, c_f64p)
        _sig(L.amgb_matrix_download_csr, C.c_int, vp, c_i32p, c_i32p, c_f64p)
        _sig(L.amgb_dist_matrix_assemble_poisson_q1, C.c_int, vp, vp, C.c_int32, C.c_int32, C.c_int32, c_f64p,
             C.c_int64, C.c_int64, C.c_int64, C.POINTER(vp), vp, vp)
        _sig(L.amgb_matrix_destroy, C.c_int, vp)
        _sig(L.amgb_numbering_dealii_q1, C.c_int, vp, C.c_int32, C.c_int32, c_i32p)
        _sig(L.amgb_matrix_permute, C.c_int, vp, vp, c_i32p, C.POINTER(vp))
        _sig(L.amgb_matrix_dims, C.c_int, vp, c_i64p, c_i64p)
        _sig(L.amgb_matrix_vmult, C.c_int, vp, vp, c_f64p, c_f64p)
        _sig(L.amgb_boomeramg_data_default, C.c_int, C.POINTER(BoomerAMGDataStruct))
        _sig(L.amgb_precond_initialize, C.c_int, vp, vp, C.POINTER(BoomerAMGDataStruct),
             C.POINTER(vp))
        _sig(L.amgb_precond_destroy, C.c_int, vp)
        _sig(L.amgb_precond_vmult, C.c_int, vp, c_f64p, c_f64p)
        _sig(L.amgb_precond_vmult_device, C.c_int, vp, vp, vp)
        _sig(L.amgb_precond_num_levels, C.c_int, vp, c_i32p)
        _sig(L.amgb_precond_level_stats, C.c_int, vp, C.c_int32, c_i32p, c_i64p, c_i64p,
             c_f64p, c_f64p, c_f64p, c_f64p)
        _sig(L.amgb_precond_level_row_stats, C.c_int, vp, C.c_int32, c_i32p, c_i32p, c_f64p, c_f64p)
        _sig(L.amgb_precond_effective_relax, C.c_int, vp, c_i32p, c_i32p, c_i32p)
        _sig(L.amgb_precond_level_cheby, C.c_int, vp, C.c_int32, c_f64p, c_f64p, c_f64p, c_i32p)
        _sig(L.amgb_precond_level_dims, C.c_int, vp, C.c_int32, c_i64p, c_i64p, c_i64p, c_i64p)
        _sig(L.amgb_precond_get_strength_mask, C.c_int, vp, C.c_int32, c_u8p)
        _sig(L.amgb_precond_get_cf_marker, C.c_int, vp, C.c_int32, c_i32p)
        _sig(L.amgb_precond_get_colors, C.c_int, vp, C.c_int32, c_i32p, c_i32p)
        _sig(L.amgb_precond_get_A_csr, C.c_int, vp, C.c_int32, c_i32p, c_i32p, c_f64p)
        _sig(L.amgb_precond_get_P_csr, C.c_int, vp, C.c_int32, c_i32p, c_i32p, c_f64p)
        _sig(L.amgb_cg_solve, C.c_int, vp, vp, c_f64p, c_f64p, vp, C.c_int64, C.c_double,
             c_f64p, C.c_int64, c_i64p)
        _sig(L.amgb_cg_solve_device, C.c_int, vp, vp, vp, vp, vp, C.c_int64, C.c_double,
             c_f64p, C.c_int64, c_i64p)
        _sig(L.amgb_make_view, C.c_int, vp, vp, C.c_int32, c_f64p, c_i64p, c_f64p, c_f64p,
             c_f64p)
        _sig(L.amgb_make_view_normalized, C.c_int, vp, vp, C.c_int32, C.c_int32, C.c_int32, c_f64p, c_f64p)
        _sig(L.amgb_ctx_enable_timers, C.c_int, vp, C.c_int)
        _sig(L.amgb_ctx_reset_timers, C.c_int, vp)
        _sig(L.amgb_timer_count, C.c_int)
        _sig(L.amgb_timer_name, C.c_char_p, C.c_int)
        _sig(L.amgb_ctx_get_timer, C.c_int, vp, C.c_int, c_f64p, c_i64p, c_f64p)
        _sig(L.amgb_ctx_get_timer_level, C.c_int, vp, C.c_int, C.c_int, c_f64p, c_i64p, c_f64p)
        _sig(L.amgb_route_count, C.c_int)
        _sig(L.amgb_route_name, C.c_char_p, C.c_int)
        _sig(L.amgb_ctx_get_route, C.c_int, vp, C.c_int, c_i64p)
        _sig(L.amgb_ctx_reset_routes, C.c_int, vp)
        _sig(L.amgb_nccl_unique_id, C.c_int, vp, C.c_int)
        _sig(L.amgb_comm_create_nccl, C.c_int, vp, C.c_int, C.c_int, vp, C.POINTER(vp))
        _sig(L.amgb_local_group_create, C.c_int, C.c_int, C.POINTER(vp))
        _sig(L.amgb_local_group_destroy, C.c_int, vp)
        _sig(L.amgb_local_group_abort, C.c_int, vp)
        _sig(L.amgb_comm_create_local, C.c_int, vp, C.c_int, C.POINTER(vp))
        _sig(L.amgb_comm_destroy, C.c_int, vp)
        _sig(L.amgb_comm_rank, C.c_int, vp)
        _sig(L.amgb_comm_size, C.c_int, vp)
        _sig(L.amgb_dist_matrix_create, C.c_int, vp, vp, C.c_int64, C.c_int64, C.c_int64, c_i64p,
             c_i32p, c_f64p, C.POINTER(vp))
        _sig(L.amgb_dist_matrix_destroy, C.c_int, vp)
        _sig(L.amgb_dist_make_view, C.c_int, vp, vp, C.c_int32, c_f64p, c_i64p, c_f64p, c_f64p, c_f64p)
        _sig(L.amgb_dist_precond_initialize, C.c_int, vp, vp, C.POINTER(BoomerAMGDataStruct),
             C.POINTER(vp))
        _sig(L.amgb_dist_cg_solve, C.c_int, vp, c_f64p, c_f64p, vp, C.c_int64, C.c_double, c_f64p,
             C.c_int64, c_i64p)
        _sig(L.amgb_dist_cg_solve_device, C.c_int, vp, vp, vp, vp, C.c_int64, C.c_double, c_f64p,
             C.c_int64, c_i64p)
        _sig(L.amgb_dist_precond_level_dims, C.c_int, vp, C.c_int32, c_i64p, c_i64p, c_i64p, c_i64p,
             c_i64p, c_i64p, c_i64p)
        _sig(L.amgb_dist_precond_replicated_from, C.c_int, vp, c_i32p)
        _sig(L.amgb_dist_precond_get_cf_marker, C.c_int, vp, C.c_int32, c_i32p)
        _sig(L.amgb_dist_precond_get_A_rows, C.c_int, vp, C.c_int32, c_i32p, c_i32p, c_f64p)
        _sig(L.amgb_dist_precond_get_P_rows, C.c_int, vp, C.c_int32, c_i32p, c_i32p, c_f64p)
        _amgb = L
    return _amgb
