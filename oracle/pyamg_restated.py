"""ORACLE (test infrastructure) — restatement of the pyamg 4.x entry points the
reference calls.  **PARITY UNPINNED** (see oracle/__init__.py).

Call sites in the reference:
  pyamg.graph.lloyd_cluster(G, seeds, maxiter)      ns/lib/graph.py:232
  pyamg.graph.bellman_ford(C, centers)              ns/model/agg_interp.py:475
  pyamg.relaxation.relaxation.gauss_seidel(A,x,b,iterations)   ns/lib/multigrid.py:175,184

Two implementations of every loop: the C one (oracle/amg_core_restated.c, fast)
and a pure-Python one (`*_py`, small cases only) used to cross-check the C.
"""
import ctypes
import numpy as np
import scipy.sparse as sp

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _suffix(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise TypeError(f"oracle supports float32/float64 graphs, got {dtype}")


def asgraph(G):
    """pyamg.graph.asgraph: CSR/CSC kept as is (CSC arrays are used as if CSR,
    i.e. the transposed graph), anything else converted with csr_matrix()."""
    if not (sp.isspmatrix_csr(G) or sp.isspmatrix_csc(G)):
        G = sp.csr_matrix(G)
    if G.shape[0] != G.shape[1]:
        raise ValueError("expected square matrix")
    return G


def max_value(dtype):
    return np.finfo(dtype).max


def bellman_ford(G, seeds):
    """-> (distances[N] in G.dtype, nearest_seed[N] int32 = NODE ID of the seed, -1 unreachable)."""
    G = asgraph(G)
    N = G.shape[0]
    if G.dtype == complex:
        raise ValueError("Bellman-Ford algorithm only defined for real weights")
    suf = _suffix(G.dtype)
    seeds = np.ascontiguousarray(np.asarray(seeds, dtype=np.intc))
    Ap = np.ascontiguousarray(G.indptr, dtype=np.intc)
    Aj = np.ascontiguousarray(G.indices, dtype=np.intc)
    Ax = np.ascontiguousarray(G.data)
    dist = np.empty(N, dtype=G.dtype)
    near = np.empty(N, dtype=np.intc)
    getattr(lib(), f"oracle_bellman_ford_{suf}")(
        ctypes.c_int(N), _p(Ap), _p(Aj), _p(Ax), ctypes.c_int(len(seeds)), _p(seeds),
        _p(dist), _p(near))
    return dist, near


def lloyd_cluster(G, seeds, maxiter=10):
    """-> (distances, clusters, seeds) — pyamg 4.x 3-tuple.  `clusters` holds seed
    INDICES (0..k-1), -1 for unreachable nodes; `seeds` are the moved centres."""
    G = asgraph(G)
    N = G.shape[0]
    if G.dtype.kind == "c":
        G = np.abs(G)
    if np.isscalar(seeds):
        seeds = np.random.permutation(N)[:seeds]
        seeds = seeds.astype(np.intc)
    else:
        seeds = np.array(seeds, dtype=np.intc)
    if len(seeds) < 1:
        raise ValueError("at least one seed is required")
    if seeds.min() < 0:
        raise ValueError("invalid seed index (%d)" % seeds.min())
    if seeds.max() >= N:
        raise ValueError("invalid seed index (%d)" % seeds.max())
    suf = _suffix(G.dtype)
    Ap = np.ascontiguousarray(G.indptr, dtype=np.intc)
    Aj = np.ascontiguousarray(G.indices, dtype=np.intc)
    Ax = np.ascontiguousarray(G.data)
    clusters = np.empty(N, dtype=np.intc)
    distances = np.empty(N, dtype=G.dtype)
    getattr(lib(), f"oracle_lloyd_cluster_{suf}")(
        ctypes.c_int(N), _p(Ap), _p(Aj), _p(Ax), ctypes.c_int(len(seeds)),
        ctypes.c_int(maxiter), _p(distances), _p(clusters), _p(seeds))
    return distances, clusters, seeds


def gauss_seidel(A, x, b, iterations=1, sweep="forward"):
    """pyamg.relaxation.relaxation.gauss_seidel — in place on x."""
    A = sp.csr_matrix(A)
    suf = _suffix(A.dtype)
    assert x.dtype == A.dtype and x.flags.c_contiguous
    b = np.ascontiguousarray(b, dtype=A.dtype)
    n = A.shape[0]
    Ap = np.ascontiguousarray(A.indptr, dtype=np.intc)
    Aj = np.ascontiguousarray(A.indices, dtype=np.intc)
    Ax = np.ascontiguousarray(A.data)
    if sweep == "forward":
        rng = [(0, n, 1)]
    elif sweep == "backward":
        rng = [(n - 1, -1, -1)]
    elif sweep == "symmetric":
        rng = [(0, n, 1), (n - 1, -1, -1)]
    else:
        raise ValueError("valid sweep directions: forward, backward, symmetric")
    fn = getattr(lib(), f"oracle_gauss_seidel_{suf}")
    for _ in range(iterations):
        for (s, e, st) in rng:
            fn(_p(Ap), _p(Aj), _p(Ax), _p(x), _p(b), ctypes.c_int(s), ctypes.c_int(e), ctypes.c_int(st))


def jacobi(A, x, b, iterations=1, omega=1.0):
    """pyamg.relaxation.relaxation.jacobi — in place on x."""
    A = sp.csr_matrix(A)
    suf = _suffix(A.dtype)
    n = A.shape[0]
    Ap = np.ascontiguousarray(A.indptr, dtype=np.intc)
    Aj = np.ascontiguousarray(A.indices, dtype=np.intc)
    Ax = np.ascontiguousarray(A.data)
    b = np.ascontiguousarray(b, dtype=A.dtype)
    temp = np.empty_like(x)
    fn = getattr(lib(), f"oracle_jacobi_{suf}")
    om = ctypes.c_double(omega) if suf == "f64" else ctypes.c_float(omega)
    for _ in range(iterations):
        fn(_p(Ap), _p(Aj), _p(Ax), _p(x), _p(b), _p(temp), ctypes.c_int(n), om)


# --------------------------------------------------------------------------
# pure-Python twins (small cases; cross-check of the C restatement)
# --------------------------------------------------------------------------
def _bf_sweep_py(Ap, Aj, Ax, x, z):
    for i in range(len(x)):
        xi, zi = x[i], z[i]
        for jj in range(Ap[i], Ap[i + 1]):
            j = Aj[jj]
            d = Ax[jj] + x[j]
            if d < xi:
                xi, zi = d, z[j]
        x[i], z[i] = xi, zi


def _bf_fixed_point_py(Ap, Aj, Ax, x, z):
    while True:
        old = x.copy()
        _bf_sweep_py(Ap, Aj, Ax, x, z)
        if (old == x).all():
            break


def bellman_ford_py(G, seeds):
    G = asgraph(G)
    N = G.shape[0]
    seeds = np.asarray(seeds, dtype=np.intc)
    x = np.full(N, max_value(G.dtype), dtype=G.dtype)
    x[seeds] = 0
    z = np.full(N, -1, dtype=np.intc)
    z[seeds] = seeds
    _bf_fixed_point_py(G.indptr, G.indices, G.data, x, z)
    return x, z


def lloyd_cluster_py(G, seeds, maxiter=10):
    G = asgraph(G)
    N = G.shape[0]
    Ap, Aj, Ax = G.indptr, G.indices, G.data
    seeds = np.array(seeds, dtype=np.intc)
    big = max_value(G.dtype)
    x = np.empty(N, dtype=G.dtype)
    w = np.empty(N, dtype=np.intc)
    for _ in range(maxiter):
        last = seeds.copy()
        x[:] = big
        w[:] = -1
        for s, node in enumerate(seeds):
            x[node] = 0
            w[node] = s
        _bf_fixed_point_py(Ap, Aj, Ax, x, w)
        x[:] = big
        for i in range(N):
            for jj in range(Ap[i], Ap[i + 1]):
                if w[i] != w[Aj[jj]]:
                    x[i] = 0
                    break
        _bf_fixed_point_py(Ap, Aj, Ax, x, w)
        for i in range(N):
            s = w[i]
            if s == -1:
                continue
            if x[seeds[s]] < x[i]:
                seeds[s] = i
        if (seeds == last).all():
            break
    return x, w, seeds


def gauss_seidel_py(A, x, b, iterations=1):
    A = sp.csr_matrix(A)
    Ap, Aj, Ax = A.indptr, A.indices, A.data
    for _ in range(iterations):
        for i in range(A.shape[0]):
            rsum = A.dtype.type(0)
            diag = A.dtype.type(0)
            for jj in range(Ap[i], Ap[i + 1]):
                j = Aj[jj]
                if i == j:
                    diag = Ax[jj]
                else:
                    rsum += Ax[jj] * x[j]
            if diag != 0:
                x[i] = (b[i] - rsum) / diag
