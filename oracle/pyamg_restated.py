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


# ------------------------------------------------------------------------------------------------
# pyamg.strength.evolution_strength_of_connection and pyamg.util.linalg.approximate_spectral_radius
# (called by the reference at utils/common.py:27,30: the 'evolution' and the DEFAULT 'olson' strength
# measures of evaluate_dataset / evaluate_ref_conv).  Restated from the published pyamg 4.x sources
# (pyamg/strength.py, pyamg/util/linalg.py, amg_core/evolution_strength.h): PARITY UNPINNED.
# ------------------------------------------------------------------------------------------------

def _approximate_eigenvalues(A, tol, maxiter, initial_guess):
    """Arnoldi (modified Gram-Schmidt against every stored vector: pyamg forces symmetric=False inside
    approximate_spectral_radius) -> (eigenvectors of H, eigenvalues of H, H, V, breakdown_flag)"""
    import scipy.linalg
    breakdown = np.finfo(float).eps * 1e6
    breakdown_flag = False
    maxiter = min(A.shape[0], maxiter)
    v0 = initial_guess
    v0 /= np.linalg.norm(v0)
    H = np.zeros((maxiter + 1, maxiter), dtype=np.result_type(v0.dtype, A.dtype))
    V = [v0]
    j = 0
    for j in range(maxiter):
        w = A @ V[-1]
        for i, v in enumerate(V):
            H[i, j] = np.dot(np.conjugate(v.ravel()), w.ravel())
            w = w - H[i, j] * v
        H[j + 1, j] = np.linalg.norm(w)
        if H[j + 1, j] < breakdown:
            breakdown_flag = True
            if H[j + 1, j] != 0.0:
                w = w / H[j + 1, j]
            V.append(w)
            break
        w = w / H[j + 1, j]
        V.append(w)
    Eigs, Vects = scipy.linalg.eig(H[:j + 1, :j + 1], left=False, right=True)
    return Vects, Eigs, H, V, breakdown_flag


def approximate_spectral_radius(A, tol=0.01, maxiter=15, restart=5, return_trace=False):
    """pyamg.util.linalg.approximate_spectral_radius with its defaults: restarted Arnoldi from
    `np.random.rand(n, 1)` (the GLOBAL numpy stream — the reference seeds it with np.random.seed(0) right before,
    utils/common.py:50,88), stop when |H[m,m-1] * y_m| / |theta| < tol.  The result is a 1 %-accurate estimate from
    below, not the spectral radius: the measure is built on THIS number."""
    v0 = np.random.rand(A.shape[1], 1)
    trace = []
    for _ in range(restart + 1):
        evect, ev, H, V, breakdown_flag = _approximate_eigenvalues(A, tol, maxiter, v0)
        nvecs = ev.shape[0]
        max_index = np.abs(ev).argmax()
        error = H[nvecs, nvecs - 1] * evect[-1, max_index]
        v0 = np.dot(np.hstack(V[:-1]), evect[:, max_index].reshape(-1, 1))
        trace.append((float(np.abs(ev[max_index])), float(np.abs(error))))
        if (np.abs(error) / np.abs(ev[max_index]) < tol) or breakdown_flag:
            break
        if np.iscomplexobj(v0) and not np.any(v0.imag):
            v0 = np.ascontiguousarray(v0.real)
    rho = float(np.abs(ev[max_index]))
    return (rho, trace) if return_trace else rho


def evolution_strength_of_connection(A, epsilon=4.0, k=2, symmetrize_measure=True, rho=None):
    """pyamg.strength.evolution_strength_of_connection(A) for a CSR matrix with B = ones, proj_type='l2',
    block_flag=False (the reference passes no other argument).  `rho`: inject rho(D^-1 A) instead of estimating it.
    -> CSR strength matrix (large = strong, rows scaled to a largest entry of 1, unit diagonal)."""
    if epsilon < 1.0:
        raise ValueError("expected epsilon > 1.0")
    if k <= 0:
        raise ValueError("number of time steps must be > 0")
    if not sp.isspmatrix_csr(A):
        raise TypeError("expected csr_matrix")
    A = A.copy()                                    # (pyamg cleans the caller's matrix in place)
    D = A.diagonal()
    Dinv = np.zeros_like(D)
    mask = D != 0.0
    Dinv[mask] = 1.0 / D[mask]
    Dinv[D == 0] = 1.0
    Dinv_A = sp.csr_matrix((A.data * np.repeat(Dinv, np.diff(A.indptr)), A.indices.copy(), A.indptr.copy()), shape=A.shape)   # scale_rows
    A.eliminate_zeros()
    A.sort_indices()
    dimen = A.shape[1]
    rho_DinvA = approximate_spectral_radius(Dinv_A) if rho is None else float(rho)
    nsquare = int(np.log2(k))
    ninc = k - 2 ** nsquare
    eye = sp.eye(dimen, dimen, format="csr", dtype=A.dtype)
    Atilde = eye - (1.0 / rho_DinvA) * Dinv_A
    Atilde = Atilde.T.tocsr()
    mask = A.copy()
    mask.data[:] = 1.0
    if ninc > 0:
        for _ in range(nsquare):
            Atilde = Atilde @ Atilde
        JacobiStep = (eye - (1.0 / rho_DinvA) * Dinv_A).T.tocsr()
        for _ in range(ninc):
            Atilde = Atilde @ JacobiStep
        Atilde = Atilde.multiply(mask).tocsr()
        Atilde.eliminate_zeros()
        Atilde.sort_indices()
    elif nsquare == 0:
        Atilde = Atilde.multiply(mask).tocsr()      # (pyamg masks here only for block systems; the pattern of a
        Atilde.eliminate_zeros()                     #  scalar Atilde^T already lies in pattern(A^T) + diagonal)
        Atilde.sort_indices()
    else:
        Atilde = Atilde.multiply(mask).tocsr()
        Atilde.eliminate_zeros()
        Atilde.sort_indices()
        for _ in range(nsquare - 1):
            Atilde = (Atilde @ Atilde).tocsr()
        AtildeCSC = Atilde.tocsc()
        AtildeCSC.sort_indices()
        mask.sort_indices()
        Atilde.sort_indices()
        L = lib()
        L.oracle_incomplete_mat_mult_csr_f64(_p(Atilde.indptr), _p(Atilde.indices), _p(np.ascontiguousarray(Atilde.data, dtype=np.float64)),
                                             _p(AtildeCSC.indptr), _p(AtildeCSC.indices), _p(np.ascontiguousarray(AtildeCSC.data, dtype=np.float64)),
                                             _p(mask.indptr), _p(mask.indices), _p(mask.data), ctypes.c_int(dimen))
        Atilde = mask
        Atilde.eliminate_zeros()
        Atilde.sort_indices()
    # NullDim == 1 shortcut with B = 1: Strength(i,j) = |1 - z_ii / z_ij|
    DAtilde = Atilde.diagonal()
    data = Atilde.data.copy()
    rows = np.repeat(np.arange(dimen), np.diff(Atilde.indptr))
    Atilde.data[:] = 1.0
    Atilde.data *= DAtilde[rows]                    # scale_rows by DAtilde / B, scale_columns by B = 1
    angle = (np.real(Atilde.data) * np.real(data)) < 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        Atilde.data = Atilde.data / data
    weak_ratio = np.abs(Atilde.data) < 1e-4
    Atilde.data = abs(1.0 - Atilde.data)
    Atilde.data[weak_ratio] = 0.0
    Atilde.data[angle] = 0.0
    Atilde.eliminate_zeros()
    Atilde.data[Atilde.data < np.sqrt(np.finfo(float).eps)] = 1e-4
    Atilde.data = np.array(np.real(Atilde.data), dtype=float)
    if epsilon != np.inf:
        lib().oracle_apply_distance_filter_f64(ctypes.c_int(dimen), ctypes.c_double(epsilon), _p(Atilde.indptr), _p(Atilde.indices),
                                               _p(Atilde.data))
        Atilde.eliminate_zeros()
    if symmetrize_measure:
        Atilde = (0.5 * (Atilde + Atilde.T)).tocsr()
    eye = sp.eye(dimen, dimen, format="csr")
    eye.data -= Atilde.diagonal()
    Atilde = (Atilde + eye).tocsr()
    Atilde.data = 1.0 / Atilde.data
    largest = np.zeros(dimen)
    Atilde.sort_indices()
    lib().oracle_maximum_row_value_f64(ctypes.c_int(dimen), _p(largest), _p(Atilde.indptr), _p(Atilde.indices), _p(Atilde.data))
    largest[largest != 0] = 1.0 / largest[largest != 0]
    Atilde.data *= largest[np.repeat(np.arange(dimen), np.diff(Atilde.indptr))]          # scale_rows
    return Atilde
