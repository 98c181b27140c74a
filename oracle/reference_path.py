"""ORACLE (test infrastructure) — scipy/numpy restatement of the reference-owned
part of the aggregation-AMG path.  Every function cites the reference lines it
follows; pyamg calls are routed to `oracle.pyamg_restated`.

Pinned: `tests/golden/make_golden.py` imports the UNMODIFIED reference modules
(/root/reference/ns/lib/{graph,multigrid}.py) in the build container and the
CPU tests compare this restatement against those outputs.
"""
import ctypes

import numpy as np
import numpy.linalg as la
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import pyamg_restated as pr


# ------------------------------------------------------------------ aggregation
def lloyd_aggregation(C, ratio=0.03, distance="unit", maxiter=10, rand=None):
    """ns/lib/graph.py:156-239.  -> (AggOp csr int8 N x num_seeds, roots, seeds)."""
    if ratio <= 0 or ratio > 1:
        raise ValueError("ratio must be > 0.0 and <= 1.0")
    if not (sp.isspmatrix_csr(C) or sp.isspmatrix_csc(C)):
        raise TypeError("expected csr_matrix or csc_matrix")
    if distance == "unit":
        data = np.ones_like(C.data).astype(float)
    elif distance == "abs":
        data = abs(C.data)
    elif distance == "inv":
        data = 1.0 / abs(C.data)
    elif distance == "same":
        data = C.data
    elif distance == "min":
        data = C.data - C.data.min()
    else:
        raise ValueError(f"Unrecognized value distance={distance}")
    if rand is None:
        rand = np.random
    elif isinstance(rand, int):
        rand = np.random.RandomState(rand)
    elif not isinstance(rand, np.random.RandomState):
        raise TypeError("rand should be an integer seed value or a random state")
    if C.dtype == complex:
        data = np.real(data)
    assert data.min() >= 0
    G = C.__class__((data, C.indices, C.indptr), shape=C.shape)
    N = C.shape[0]
    num_seeds = int(np.ceil(ratio * N))
    seeds = rand.permutation(N)[:num_seeds]
    _, clusters, roots = pr.lloyd_cluster(G, np.copy(seeds), maxiter=maxiter)
    row = (clusters >= 0).nonzero()[0]
    col = clusters[row]
    data = np.ones(len(row), dtype="int8")
    AggOp = sp.coo_matrix((data, (row, col)), shape=(G.shape[0], num_seeds)).tocsr()
    return AggOp, roots, seeds


def modified_bellman_ford(S_coo, centers):
    """ns/lib/graph.py:7-53 on a (coalesced, row-major) scipy COO.
    -> (distance float32[N], nearest_center int64[N])."""
    S = sp.coo_matrix(S_coo)
    S.sum_duplicates()                      # torch .coalesce(): row-major sorted, duplicates summed
    order = np.lexsort((S.col, S.row))
    row = np.ascontiguousarray(S.row[order], dtype=np.int64)
    col = np.ascontiguousarray(S.col[order], dtype=np.int64)
    w = np.ascontiguousarray(S.data[order])
    n = S.shape[0]
    centers = np.ascontiguousarray(np.asarray(centers, dtype=np.int64))
    dist = np.empty(n, dtype=np.float32)
    near = np.empty(n, dtype=np.int64)
    suf = "f32" if w.dtype == np.float32 else "f64"
    if suf == "f64":
        w = w.astype(np.float64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    passes = getattr(pr.lib(), f"oracle_modified_bf_{suf}")(
        ctypes.c_int(n), ctypes.c_longlong(len(w)), p(row), p(col), p(w),
        ctypes.c_int(len(centers)), p(centers), p(dist), p(near))
    if passes < 0:
        raise RuntimeError("modified_bellman_ford: no fixed point after %d passes (float32 distances with float64 weights can be "
                           "rounded up on the store and relax for ever; the reference's callers pass float32 weights)" % -passes)
    return dist, near


def nearest_center_to_agg(top_k, nearest_center):
    """ns/lib/graph.py:56-86.  -> scipy CSR (n x m) float32 of ones; KeyError on a
    label that is not a centre (e.g. -1 for unreachable nodes), as the reference."""
    top_k = np.asarray(top_k)
    nearest_center = np.asarray(nearest_center)
    n, m = len(nearest_center), len(top_k)
    inv = {int(k): i for i, k in enumerate(top_k)}
    cols = np.empty(n, dtype=np.int64)
    for i, c in enumerate(nearest_center):
        cols[i] = inv[int(c)]
    return sp.coo_matrix((np.ones(n, dtype=np.float32), (np.arange(n), cols)), shape=(n, m)).tocsr()


# ------------------------------------------------------------------ prolongators
def lambda_max_dinv_a(A):
    """|lambda_max(D^-1 A)| as ns/lib/multigrid.py:105 (ARPACK, non-deterministic last ulps)."""
    Dinv = sp.diags([1.0 / A.diagonal()], [0])
    return np.abs(spla.eigs(Dinv @ A, k=1, return_eigenvectors=False)).item()


def smoothed_aggregation_jacobi(A, Agg, omega=None):
    """ns/lib/multigrid.py:102-108.  `omega` (=(4/3)/|lambda_max|) may be passed in to
    skip ARPACK; the expression order is the reference's."""
    n = A.shape[0]
    Dinv = sp.diags([1.0 / A.diagonal()], [0])
    if omega is None:
        omega = (4.0 / 3.0) / np.abs(spla.eigs(Dinv @ A, k=1, return_eigenvectors=False)).item()
    smoother = sp.eye(n) - omega * Dinv @ A
    P = smoother @ Agg
    return P


def learned_prolongator(P_hat, Agg):
    """ns/model/agg_interp.py:481-484: P = P_hat . Agg (torch.sparse.mm + coalesce keeps
    explicit zeros; scipy drops them — compare on values, not on stored zeros)."""
    return sp.csr_matrix(P_hat) @ sp.csr_matrix(Agg)


def galerkin(A, P):
    """ns/lib/multigrid.py:165 / MLAMG.py:121: A_H = P.T @ A @ P."""
    return P.T @ A @ P


def canonical_csr(M, drop_zeros=True):
    """Canonical comparison form: CSR, sorted columns, duplicates summed,
    exact zeros dropped."""
    M = sp.csr_matrix(M).copy()
    M.sum_duplicates()
    if drop_zeros:
        M.eliminate_zeros()
    M.sort_indices()
    return M


# ------------------------------------------------------------------ smoothers
def jacobi(A, b, x, Dinv=None, omega=0.666, nu=2):
    """ns/lib/multigrid.py:15-45 (in place on x)."""
    if Dinv is None:
        Dinv = sp.diags(1.0 / A.diagonal())
    for _ in range(nu):
        x += omega * Dinv @ b - omega * Dinv @ A @ x
    return x


def mlamg_jacobi(A, Dinv_w, b, x, nu=2):
    """ns/preconditioner/MLAMG.py:143-146 with Dinv_w = jacobi_weight/diag (MLAMG.py:104)."""
    for _ in range(nu):
        x += Dinv_w @ (b - A @ x)
    return x


def jacobi_torch_like(A, b, x, Dinv, omega=0.666, nu=2):
    """ns/lib/multigrid.py:48-55 with a vector Dinv (callers pass 1/diag, :220)."""
    for _ in range(nu):
        x += omega * (Dinv * b) - omega * Dinv * (A @ x)
    return x


def l1_jacobi(A, b, x, nu=1):
    """NOT in the reference — north-star addition, defined here:
    x += (b - A x) / sum_j |a_ij|.  PARITY UNPINNED."""
    d = np.asarray(abs(A).sum(axis=1)).ravel()
    for _ in range(nu):
        x += (b - A @ x) / d
    return x


# ------------------------------------------------------------------ two-level drivers
def amg_2_v(A, P, b, x, pre_smoothing_steps=1, post_smoothing_steps=1, jacobi_weight=0.666,
            res_tol=None, error_tol=None, max_iter=500, singular=False, smoother="gauss_seidel"):
    """ns/lib/multigrid.py:111-210.  `smoother='gauss_seidel'` is the reference behaviour
    (pyamg GS, `jacobi_weight` ignored).  `smoother='jacobi'` swaps in MLAMG.py:143-146
    (x += w D^-1 (b - A x)) with the same driver — used to pin the GPU Jacobi path."""
    if res_tol is None and error_tol is None:
        raise RuntimeError("One of res_tol or error_tol must be set!")
    tol = res_tol if res_tol is not None else error_tol
    err = np.zeros(max_iter)
    A_H = P.T @ A @ P
    if not singular:
        try:
            A_H_LU = spla.factorized(A_H)
        except Exception:
            return x, np.float64(1.0), err, 0
    x = x.copy()
    if smoother == "jacobi":
        Dw = sp.diags(1.0 / A.diagonal()) * jacobi_weight
    elif smoother == "l1_jacobi":
        pass
    for i in range(max_iter):
        if smoother == "gauss_seidel":
            pr.gauss_seidel(A, x, b, iterations=pre_smoothing_steps)
        elif smoother == "jacobi":
            mlamg_jacobi(A, Dw, b, x, nu=pre_smoothing_steps)
        else:
            l1_jacobi(A, b, x, nu=pre_smoothing_steps)
        if singular:
            x += P @ spla.lsqr(P.T @ A @ P, P.T @ (b - A @ x))[0]
        else:
            x += P @ A_H_LU(P.T @ (b - A @ x))
        if smoother == "gauss_seidel":
            pr.gauss_seidel(A, x, b, iterations=post_smoothing_steps)
        elif smoother == "jacobi":
            mlamg_jacobi(A, Dw, b, x, nu=post_smoothing_steps)
        else:
            l1_jacobi(A, b, x, nu=post_smoothing_steps)
        if singular:
            x -= np.mean(x)
        if res_tol is not None:
            e = la.norm(b - A @ x, 2)
        else:
            e = la.norm(x, 2)
        err[i] = e
        if e <= tol:
            err = err[:i + 1]
            break
    if len(err) != 1:
        try:
            err_n = min(len(err) // 3, 10)
            conv_factor = (err[-1] / err[-err_n]) ** (1 / (err_n - 1))
        except Exception:
            conv_factor = 0
    else:
        conv_factor = 0
    return x, conv_factor, err, len(err)


def amg_2_v_torch_like(A, P, b, x, pre_smoothing_steps=1, post_smoothing_steps=1,
                       jacobi_weight=0.666, error_tol=1e-10, max_iter=20):
    """ns/lib/multigrid.py:213-245 in numpy (dtype of the inputs; dense LU coarse solve)."""
    Dinv = 1.0 / A.diagonal()
    A_H = (P.T @ A @ P).toarray()
    import scipy.linalg as sla
    lu = sla.lu_factor(A_H)
    err = np.zeros(max_iter, dtype=x.dtype)
    x = x.copy()
    for i in range(max_iter):
        x = jacobi_torch_like(A, b, x, Dinv, omega=jacobi_weight, nu=pre_smoothing_steps)
        r_H = P.T @ (b - A @ x)
        e_H = sla.lu_solve(lu, r_H)
        x += P @ e_H
        x = jacobi_torch_like(A, b, x, Dinv, omega=jacobi_weight, nu=post_smoothing_steps)
        err[i] = la.norm(x)
        if err[i] < error_tol:
            break
    n_err = 3
    return (err[i] / err[i - n_err]) ** (1 / (n_err - 1))


def mlamg_amg_2_v(A, P, A_H_lu_solve, Dinv_w, b, x, amg_rtol=1e-8, pre_smoothing_steps=1,
                  post_smoothing_steps=1, max_iter=500):
    """ns/preconditioner/MLAMG.py:148-197 (absolute residual tolerance)."""
    it = 0
    for it in range(max_iter):
        x = mlamg_jacobi(A, Dinv_w, b, x, nu=pre_smoothing_steps)
        x += P @ A_H_lu_solve(P.T @ (b - A @ x))
        x = mlamg_jacobi(A, Dinv_w, b, x, nu=post_smoothing_steps)
        if la.norm(b - A @ x, 2) <= amg_rtol:
            break
    return x, it + 1


def amg_loss_forward(P, A, test_vecs, tot_num_loop=5, no_prerelax=1, no_postrelax=1, neumann_solve_fix=False):
    """ns/model/loss.py:32-96, forward value only (no autograd): fp32 iterates, fp64 coarse solve, softmax-weighted
    per-column convergence factor.  neumann_solve_fix: the coarse operator is bordered with a Lagrange row / column of
    ones (`add_lagrange_rowcols` :11-27, `add_lagrange_vec` :29-30) so that the constant null space is handled."""
    A = sp.csr_matrix(A).astype(np.float32)
    P = sp.csr_matrix(P).astype(np.float32)
    omega = 2.0 / 3.0
    Dinv_v = ((1.0 / A.diagonal().astype(np.float32)) * np.float32(omega)).astype(np.float32)
    A_H = (P.T @ A @ P).astype(np.float64)
    N = A.shape[0]
    if not isinstance(test_vecs, np.ndarray):
        np.random.seed(0)
        x = np.random.normal(0, 1, (N, test_vecs)).astype(np.float32)
        x = x / la.norm(x, 2, axis=0)
    else:
        x = test_vecs.astype(np.float32)
    errs = np.zeros((tot_num_loop + 1, x.shape[1]), dtype=np.float32)
    if neumann_solve_fix:
        k = A_H.shape[0]
        ones = np.ones((k, 1))
        A_H = sp.bmat([[A_H, sp.csr_matrix(ones)], [sp.csr_matrix(ones.T), None]], format="csc")      # (k+1) x (k+1)
    solve = spla.factorized(sp.csc_matrix(A_H))
    for it in range(tot_num_loop + 1):
        for _ in range(no_prerelax):
            x = x - Dinv_v[:, None] * (A @ x)
        r_H = P.T @ (A @ x)
        if neumann_solve_fix:
            r_H = np.vstack([r_H, np.zeros((1, r_H.shape[1]), dtype=r_H.dtype)])
        e_H = np.stack([solve(-r_H[:, c].astype(np.float64)) for c in range(x.shape[1])], axis=1).astype(np.float32)
        if neumann_solve_fix:
            e_H = e_H[:-1]
        x = x + P @ e_H
        for _ in range(no_postrelax):
            x = x - Dinv_v[:, None] * (A @ x)
        x = x - x.mean(0)
        errs[it] = la.norm(x, 2, axis=0)
    n_err = 3
    convs = (errs[-1] / errs[-n_err]) ** (1 / (n_err - 1))
    sm = np.exp(convs - convs.max())
    sm /= sm.sum()
    return float(sm @ convs), errs
