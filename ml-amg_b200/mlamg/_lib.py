"""ctypes binding of libmlamg_b200.so (include/mlamg.h).

There is NO CPU fallback: if the shared library is missing or cannot be loaded this module raises
ImportError, and every compute wrapper raises RuntimeError when CUDA is unavailable.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MLAMG_LIB_PATH: development aid (A/B of two builds of the extension); the default is the in-tree library
LIB_PATH = os.environ.get("MLAMG_LIB_PATH") or os.path.join(HERE, "libmlamg_b200.so")

F32, F64 = 0, 1
OK, EINVAL, ECUDA, ELIMIT, ESINGULAR, EKEY = 0, 1, 2, 3, 4, 5

I = ctypes.c_int
LL = ctypes.c_longlong
D = ctypes.c_double
P = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/mlamg.h one to one
SIGNATURES = {
    "mlamg_last_error": (ctypes.c_char_p, []),
    "mlamg_version": (I, []),
    "mlamg_launch_count": (LL, []),
    "mlamg_spmv_csr": (I, [I, I, I, P, P, P, P, P, P]),
    "mlamg_spmv_add_csr": (I, [I, I, I, P, P, P, P, P, P]),
    "mlamg_spmv_csr_perm": (I, [I, I, I, P, P, P, P, P, P, P]),
    "mlamg_hierarchy_set_restrict_order": (I, [P, I, P]),
    "mlamg_residual_csr": (I, [I, I, I, P, P, P, P, P, P, P, P]),
    "mlamg_jacobi_csr": (I, [I, I, I, P, P, P, P, P, P, P, P]),
    "mlamg_jacobi_zero": (I, [I, I, P, P, P, P]),
    "mlamg_jacobi_zero_residual_csr": (I, [I, I, I, P, P, P, P, P, P, P, P, P]),
    "mlamg_smoother_diag": (I, [I, I, D, I, P, P, P, P, P]),
    "mlamg_sell_slice_ptr": (I, [I, P, P, P, P]),
    "mlamg_sell_fill": (I, [I, I, P, P, P, P, P, P, P]),
    "mlamg_sell_rowop": (I, [I, I, I, P, P, P, P, P, P, P, P, P]),
    "mlamg_set_csr_lanes": (I, [I]),
    "mlamg_set_csr_batch": (I, [I]),
    "mlamg_hierarchy_set_operator_sell": (I, [P, I, P, P, P]),
    "mlamg_spmm_csr": (I, [I, I, I, P, P, P, P, P, D, D, P]),
    "mlamg_sddmm_csr": (I, [I, I, I, P, P, P, P, P, P]),
    "mlamg_csr_sample_dense": (I, [I, I, I, P, P, P, P, P]),
    "mlamg_agg_product_backward": (I, [I, I, P, P, P, P, P, P, P, P]),
    "mlamg_evolution_step": (I, [I, I, P, P, P, D, P, P, P, P]),
    "mlamg_incomplete_matmul_csr": (I, [I, I, P, P, P, P, P, P, P, P, P, P]),
    "mlamg_evolution_measure": (I, [I, I, P, P, P, P]),
    "mlamg_distance_filter": (I, [I, I, D, P, P, P, P]),
    "mlamg_evolution_symmetrize": (I, [I, I, P, P, P, P, P, I, P, P]),
    "mlamg_invert_scale_rows": (I, [I, I, P, P, P]),
    "mlamg_csr_pattern_add": (I, [I, I, P, P, P, P, P, P, P, P]),
    "mlamg_axpby": (I, [I, I, D, P, D, P, P]),
    "mlamg_dot": (I, [I, I, P, P, P, P]),
    "mlamg_gs_schedule": (I, [I, P, P, P, P, P, P, P]),
    "mlamg_gauss_seidel": (I, [I, I, P, P, P, P, P, P, P, I, P]),
    "mlamg_scan_i32": (I, [P, P, I, P]),
    "mlamg_agg_from_labels": (I, [I, I, P, P, P, P, P, P]),
    "mlamg_center_rank_labels": (I, [I, I, P, P, P, P, P]),
    "mlamg_sa_smoother": (I, [I, I, P, P, P, D, P, P, P]),
    "mlamg_spgemm_symbolic": (I, [I, I, I, P, P, P, P, P, P, P]),
    "mlamg_spgemm_numeric": (I, [I, I, I, I, P, P, P, P, P, P, P, P, P, P]),
    "mlamg_csr_transpose": (I, [I, I, I, I, P, P, P, P, P, P, P]),
    "mlamg_csr_nonzero_count": (I, [I, I, P, P, P, P, P]),
    "mlamg_csr_nonzero_fill": (I, [I, I, P, P, P, P, P, P, P]),
    "mlamg_csr_sort_rows": (I, [I, I, P, P, P, P]),
    "mlamg_csr_to_dense": (I, [I, I, P, P, P, P, P]),
    "mlamg_dense_inverse_f64": (I, [I, P, P, P]),
    "mlamg_gemv": (I, [I, I, P, P, P, P]),
    "mlamg_lambda_max": (I, [I, I, LL, P, P, P, D, I, I, P, P, P, P, P]),
    "mlamg_poisson_nnz": (LL, [I, I, I]),
    "mlamg_poisson_csr": (I, [I, I, I, I, P, P, P, P]),
    "mlamg_poisson_csr_slab": (I, [I, I, I, I, I, I, P, P, P, P, P]),
    "mlamg_rowop_csr": (I, [I, I, I, I, P, P, P, P, P, P, P, P, P, I, P, P]),
    "mlamg_prolong_smooth_csr": (I, [I, I, I, P, P, P, P, P, P, P, P, P]),
    "mlamg_csr_to_w32": (I, [I, I, P, P, P, P, P, P]),
    "mlamg_rowop_w32": (I, [I, I, I, I, I, P, P, P, P, P, P, P, P, P]),
    "mlamg_residual_w32": (I, [I, I, P, P, P, P, P, P, P]),
    "mlamg_prolong_smooth_zero_w32": (I, [I, I, P, P, P, P, P, P, P, P, P]),
    "mlamg_prolong_smooth_zero_csr": (I, [I, I, I, P, P, P, P, P, P, P, P, P]),
    "mlamg_jacobi_zero_residual_scaled_csr": (I, [I, I, I, P, P, P, P, P, P, P, P, P]),
    "mlamg_hierarchy_set_operator_scaled": (I, [P, I, P]),
    "mlamg_gather": (I, [I, I, P, P, P, P]),
    "mlamg_legacy_permutation_head": (I, [ctypes.c_uint, LL, LL, P]),
    "mlamg_agg_stats": (I, [P, I]),
    "mlamg_bellman_ford": (I, [I, I, P, P, P, I, P, P, P, P, P]),
    "mlamg_lloyd_cluster": (I, [I, I, P, P, P, I, P, I, P, P, P, P]),
    "mlamg_modified_bellman_ford": (I, [I, P, P, P, I, P, P, P, P, P]),
    "mlamg_hierarchy_create": (I, [I, I, P]),
    "mlamg_hierarchy_set_operator": (I, [P, I, I, I, P, P, P, P]),
    "mlamg_hierarchy_set_transfer": (I, [P, I, I, P, P, P, P, P, P]),
    "mlamg_hierarchy_set_post_operator": (I, [P, I, I, P, P, P]),
    "mlamg_hierarchy_set_w32": (I, [P, I, P, P, P, P]),
    "mlamg_hierarchy_set_coarse_inverse": (I, [P, P]),
    "mlamg_hierarchy_finalize": (I, [P, P]),
    "mlamg_hierarchy_destroy": (I, [P]),
    "mlamg_hierarchy_cycle_bytes": (D, [P, I, I, I]),
    "mlamg_hierarchy_use_graph": (I, [P, I]),
    "mlamg_vcycle": (I, [P, P, P, I, I, I, P]),
    "mlamg_solve": (I, [P, P, P, I, I, D, I, P, P, P]),
    "mlamg_solve_ex": (I, [P, P, P, I, I, I, D, I, P, P, P]),
    "mlamg_pcg": (I, [P, P, P, I, I, D, I, P, P, P]),
    "mlamg_solver_loop_mode": (I, [P]),
    "mlamg_dloop_state_bytes": (I, []),
    "mlamg_dloop_init": (I, [P, D, I, P]),
    "mlamg_dloop_dot": (I, [I, I, P, P, P, I, P]),
    "mlamg_dloop_scalar": (I, [P, I, P, P]),
    "mlamg_dloop_direction": (I, [I, I, P, P, P, P]),
    "mlamg_dloop_update": (I, [I, I, P, P, P, P, P, P]),
    "mlamg_gmres_orthogonalize": (I, [I, I, I, P, P, P, P, P]),
    "mlamg_vcycle_host": (I, [P, P, P, I, I, I, P]),
    "mlamg_peer_alloc": (I, [LL, P, P]),
    "mlamg_peer_open": (I, [P, P]),
    "mlamg_peer_close": (I, [P]),
    "mlamg_peer_free": (I, [P]),
    "mlamg_channel_slot_bytes": (I, [I]),
    "mlamg_channel_create": (I, [I, P, P, P, I, P, P, P, P]),
    "mlamg_channel_destroy": (I, [P]),
    "mlamg_channel_push": (I, [P, I, P, P, P, P]),
    "mlamg_channel_unpack": (I, [P, I, P, P, P]),
    "mlamg_channel_rowop": (I, [P, I, I, I, I, P, P, P, P, I, P, P, P, P, P, I, P]),
}


class MlamgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mlamg error {code}: {msg}")
        self.code = code


class SingularCoarseError(MlamgError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python ml-amg_b200/build.py` "
            "(there is no CPU fallback for the mlamg hot path)")
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc):
    if rc == OK:
        return
    msg = lib.mlamg_last_error().decode("utf-8", "replace")
    if rc == ESINGULAR:
        raise SingularCoarseError(rc, msg)
    if rc == EKEY:
        raise KeyError(msg)
    raise MlamgError(rc, msg)


def launch_count():
    return int(lib.mlamg_launch_count())
