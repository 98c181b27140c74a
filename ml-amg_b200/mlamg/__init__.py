"""mlamg — B200-native aggregation-AMG hot path (PyTorch host + libmlamg_b200.so C ABI).

Importing this package loads the CUDA extension and fails loudly if it is missing.
"""
from . import _lib                                   # noqa: F401  (raises ImportError when the .so is absent)
from ._lib import MlamgError, SingularCoarseError, launch_count     # noqa: F401
from .core import (DeviceCSR, DeviceSELL, set_csr_lanes, spmv, spmv_perm, spmv_add, residual, jacobi_sweep, jacobi_zero, jacobi_zero_residual, jacobi_zero_residual_scaled, scaled_values, prolong_smooth, prolong_smooth_zero, prolong_smooth_zero_w32, csr_to_w32, smoother_diag, spmm, dot,   # noqa: F401
                   axpby, gemv, GaussSeidelSchedule, scan_i32, agg_from_labels, center_rank_labels, sa_smoother,
                   spgemm, transpose, drop_zeros, sort_rows, lambda_max, dense_inverse, poisson, bellman_ford,
                   lloyd_cluster, modified_bellman_ford, require_cuda)
from .hierarchy import (Hierarchy, Level, build_hierarchy, lloyd_labels, lloyd_seeds, distance_transform,      # noqa: F401
                        sa_prolongator, learned_prolongator, galerkin, post_operator)

__version__ = "0.1.0"
