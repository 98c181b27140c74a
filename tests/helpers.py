"""Shared test helpers (CPU side): golden loading, canonical comparisons, problem generators."""
import os

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["poisson2d_24_unit", "poisson3d_10_unit", "laplace3d_grid_abs", "laplace3d_grid_inv", "randw_same"]
# shapes of BASELINE configs 3 and 4 (Voronoi jump coefficients, Delaunay P1), produced by the unmodified reference too;
# held by the CPU oracle tests (the GPU path meets these shapes in tests/test_gpu_configs.py)
GOLDEN_CASES_CONFIG_SHAPES = ["voronoi_jump_18_inv", "delaunay_600_abs"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"), allow_pickle=False)
    return z


def csr_from(z, prefix):
    return sp.csr_matrix((z[f"{prefix}_data"], z[f"{prefix}_indices"], z[f"{prefix}_indptr"]),
                         shape=tuple(int(v) for v in z[f"{prefix}_shape"]))


def canonical(M, drop_zeros=True):
    M = sp.csr_matrix(M).copy()
    M.sum_duplicates()
    if drop_zeros:
        M.eliminate_zeros()
    M.sort_indices()
    return M


def assert_same_pattern(X, Y):
    X, Y = canonical(X), canonical(Y)
    assert X.shape == Y.shape
    assert np.array_equal(X.indptr, Y.indptr), "row pointers differ"
    assert np.array_equal(X.indices, Y.indices), "column indices differ"


def assert_csr_close(X, Y, rtol):
    """pattern bit-exact (canonical form), values within rtol relative to the largest entry"""
    assert_same_pattern(X, Y)
    X, Y = canonical(X), canonical(Y)
    scale = max(np.abs(Y.data).max(), 1e-300) if Y.nnz else 1.0
    err = np.abs(X.data - Y.data).max() / scale if Y.nnz else 0.0
    assert err <= rtol, f"max relative value error {err:.3e} > {rtol:.1e}"


def assert_csr_bitwise(X, Y):
    """canonical pattern AND values bit-identical (ordered SpGEMM reproduces scipy's summation order)"""
    assert_same_pattern(X, Y)
    X, Y = canonical(X), canonical(Y)
    assert X.data.dtype == Y.data.dtype, (X.data.dtype, Y.data.dtype)
    bad = np.nonzero(X.data != Y.data)[0]
    assert bad.size == 0, f"{bad.size} of {X.nnz} values differ in their bits; first: {X.data[bad[0]]!r} vs {Y.data[bad[0]]!r}"


def rel_hist_err(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    n = min(len(a), len(b))
    return np.max(np.abs(a[:n] - b[:n]) / np.maximum(np.abs(b[:n]), 1e-300))


def hist_err0(a, b):
    """residual-history parity measure: max |a_i - b_i| / b_0.  Deep into convergence the residual
    b - A x is itself only known to eps*||A|| ||x||, so parity is stated relative to the initial
    residual (the 1e-12 north-star bar); early entries are also checked entry-relative."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    n = min(len(a), len(b))
    return np.max(np.abs(a[:n] - b[:n])) / max(abs(b[0]), 1e-300)


def random_csr(n, m, density, seed, dtype=np.float64, empty_rows=True):
    rs = np.random.RandomState(seed)
    A = sp.random(n, m, density=density, random_state=rs, format="csr", dtype=np.float64)
    A.data = rs.randn(A.nnz)
    if empty_rows and n > 3:
        A = sp.csr_matrix(A.multiply(sp.csr_matrix((np.arange(n) % 7 != 3).astype(float)).T))   # some empty rows
    A = sp.csr_matrix(A).astype(dtype)
    A.sort_indices()
    return A


def grid_graph(shape, weights="unit", seed=0, dtype=np.float64, symmetric=True):
    """graph of a Dirichlet Poisson stencil (incl. the diagonal entry, as the reference passes A itself)"""
    from oracle import multilevel as oml
    A = oml.poisson(shape)
    rs = np.random.RandomState(seed)
    if weights == "unit":
        data = np.ones(A.nnz)
    elif weights == "random":
        data = rs.rand(A.nnz) + 0.05
        if symmetric:
            G = sp.csr_matrix((data, A.indices, A.indptr), shape=A.shape)
            G = G.maximum(G.T)
            G.sort_indices()
            return G.astype(dtype)
    elif weights == "relu":      # GNN-like fp32 outputs: many exact zeros and tiny values
        data = np.maximum(rs.randn(A.nnz), 0.0) * (10.0 ** rs.randint(-6, 1, A.nnz))
    else:
        raise ValueError(weights)
    return sp.csr_matrix((data.astype(dtype), A.indices, A.indptr), shape=A.shape)


def load_eval_golden():
    """tests/golden/ref_evaluation_sequences.npz (written by the unmodified reference's utils/common.py drivers)
    -> (npz, {grid name: csr})"""
    z = np.load(os.path.join(GOLDEN, "ref_evaluation_sequences.npz"))
    grids = {}
    for name in z["grid_names"]:
        name = str(name)
        grids[name] = sp.csr_matrix((z[f"{name}_data"], z[f"{name}_indices"], z[f"{name}_indptr"]))
    return z, grids
