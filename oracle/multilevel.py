"""ORACLE (test infrastructure) — multilevel hierarchy, V-cycle and PCG defined by
recursion over the reference's two-level building blocks.  **PARITY UNPINNED**:
the reference owns only a two-level solver (ns/lib/multigrid.py:111-210); the
multilevel cycle exists there only as a call into pyamg
(ns/preconditioner/PyAMG.py:94,119).  Ordering follows pyamg 4.x
`multilevel_solver.__solve`: presmooth -> r=b-Ax -> b_c=R r -> recurse from
x_c=0 -> x+=P x_c -> postsmooth; coarsest level solved exactly.

Building blocks used per level (all from oracle.reference_path):
  aggregates : lloyd_aggregation (ns/lib/graph.py:156-239) or caller-supplied labels
  P          : (I - w D^-1 A) Agg      (ns/lib/multigrid.py:102-108)
  A_H        : P.T @ A @ P             (ns/lib/multigrid.py:165)
  smoother   : x += w D^-1 (b - A x)   (ns/preconditioner/MLAMG.py:143-146) or L1-Jacobi
"""
import numpy as np
import numpy.linalg as la
import scipy.sparse as sp

from . import reference_path as rp


class Level:
    __slots__ = ("A", "P", "R", "labels", "omega_sa", "dw", "seeds", "roots")

    def __init__(self):
        self.A = self.P = self.R = self.labels = self.omega_sa = self.dw = None
        self.seeds = self.roots = None


def labels_to_agg(labels, ncoarse):
    """Agg as ns/lib/graph.py:234-238 (rows with label < 0 are empty)."""
    labels = np.asarray(labels)
    row = (labels >= 0).nonzero()[0]
    col = labels[row]
    return sp.coo_matrix((np.ones(len(row), dtype="int8"), (row, col)),
                         shape=(len(labels), ncoarse)).tocsr()


def smoother_diag(A, smoother, jacobi_weight):
    if smoother == "jacobi":
        return jacobi_weight / A.diagonal()
    if smoother == "l1_jacobi":
        return 1.0 / np.asarray(abs(A).sum(axis=1)).ravel()
    raise ValueError(smoother)


def build_hierarchy(A, *, labels_per_level=None, ratio=0.1, distance="unit", maxiter=10, rand=0,
                    lam_max=None, max_levels=10, max_coarse=500, smoother="jacobi",
                    jacobi_weight=2.0 / 3.0):
    """Returns list[Level]; the last level has only `.A`.

    lam_max : None -> ARPACK as the reference; float/list/callable -> |lambda_max(D^-1 A)|
              per level supplied by the caller.
    labels_per_level : optional list of (labels, ncoarse) to bypass Lloyd.
    """
    A = sp.csr_matrix(A)
    levels = []
    lvl = 0
    while True:
        L = Level()
        L.A = A
        L.dw = smoother_diag(A, smoother, jacobi_weight).astype(A.dtype)
        levels.append(L)
        if len(levels) >= max_levels or A.shape[0] <= max_coarse:
            break
        if labels_per_level is not None:
            if lvl >= len(labels_per_level):
                break
            labels, nc = labels_per_level[lvl]
            Agg = labels_to_agg(labels, nc)
            L.labels = np.asarray(labels)
        else:
            Agg, roots, seeds = rp.lloyd_aggregation(A, ratio=ratio, distance=distance,
                                                     maxiter=maxiter, rand=rand)
            L.labels = np.full(A.shape[0], -1, dtype=np.int32)
            coo = Agg.tocoo()
            L.labels[coo.row] = coo.col
            L.seeds, L.roots = seeds, roots
        if callable(lam_max):
            lam = lam_max(A)
        elif lam_max is None:
            lam = rp.lambda_max_dinv_a(A)
        elif np.isscalar(lam_max):
            lam = lam_max
        else:
            lam = lam_max[lvl]
        L.omega_sa = (4.0 / 3.0) / lam
        P = sp.csr_matrix(rp.smoothed_aggregation_jacobi(A, Agg, omega=L.omega_sa))
        L.P = rp.canonical_csr(P)
        L.R = sp.csr_matrix(L.P.T)
        A = rp.canonical_csr(rp.galerkin(L.A, L.P))
        lvl += 1
    return levels


def coarse_solve(Ac, b):
    return la.solve(Ac.toarray(), b)


def vcycle(levels, b, x=None, nu1=1, nu2=1, lvl=0, coarse=None):
    """One V(nu1,nu2) cycle, x updated and returned.  `coarse` caches the dense inverse."""
    L = levels[lvl]
    A = L.A
    if x is None:
        x = np.zeros_like(b)
    if lvl == len(levels) - 1:
        return coarse_solve(A, b)
    for _ in range(nu1):
        x += L.dw * (b - A @ x)
    r = b - A @ x
    bc = L.R @ r
    xc = vcycle(levels, bc, None, nu1, nu2, lvl + 1)
    x += L.P @ xc
    for _ in range(nu2):
        x += L.dw * (b - A @ x)
    return x


def solve(levels, b, x0=None, tol=1e-8, maxiter=100, nu1=1, nu2=1):
    """Stationary V-cycle iteration; residual history of ||b - A x||_2 (entry 0 = initial)."""
    A = levels[0].A
    x = np.zeros_like(b) if x0 is None else x0.copy()
    res = [la.norm(b - A @ x)]
    nb = la.norm(b)
    stop = tol * (nb if nb != 0 else 1.0)
    for _ in range(maxiter):
        x = vcycle(levels, b, x, nu1, nu2)
        res.append(la.norm(b - A @ x))
        if res[-1] <= stop:
            break
    return x, np.array(res)


def pcg(levels, b, x0=None, tol=1e-8, maxiter=200, nu1=1, nu2=1):
    """Preconditioned CG, M^-1 = one V-cycle from a zero guess.  Stops when
    ||r||_2 <= tol*||b||_2.  Returns (x, residual history incl. initial, iterations)."""
    A = levels[0].A
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = b - A @ x
    res = [la.norm(r)]
    nb = la.norm(b)
    stop = tol * (nb if nb != 0 else 1.0)
    if res[0] <= stop:
        return x, np.array(res), 0
    z = vcycle(levels, r.copy(), None, nu1, nu2)
    p = z.copy()
    rz = r @ z
    it = 0
    for it in range(1, maxiter + 1):
        Ap = A @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        res.append(la.norm(r))
        if res[-1] <= stop:
            break
        z = vcycle(levels, r.copy(), None, nu1, nu2)
        rz_new = r @ z
        beta = rz_new / rz
        rz = rz_new
        p = z + beta * p
    return x, np.array(res), it


def gmres(levels, b, x0=None, tol=1e-8, maxiter=100, restart=30, nu1=1, nu2=1):
    """Left-preconditioned restarted GMRES (modified Gram-Schmidt, Givens rotations), M^-1 = one V-cycle from a zero
    guess — the shape of `Amg.solve(b, tol, accel='gmres')` at ns/preconditioner/PyAMG.py:119.  PARITY UNPINNED: pyamg's
    own gmres (Householder by default) is not available; this restatement defines the algorithm the GPU path is held to.
    History = norms of the PRECONDITIONED residual (entry 0 = initial); stops when it drops below tol*||M b||_2.
    Returns (x, history, iterations)."""
    A = levels[0].A
    M = lambda v: vcycle(levels, v.copy(), None, nu1, nu2)
    x = np.zeros_like(b) if x0 is None else x0.copy()
    nmb = la.norm(M(b))
    stop = tol * (nmb if nmb != 0 else 1.0)
    res = []
    it = 0
    while True:
        r = M(b - A @ x)
        beta = la.norm(r)
        if not res:
            res.append(beta)
        if beta <= stop or it >= maxiter:
            break
        m = min(restart, maxiter - it)
        V = [r / beta]
        H = np.zeros((m + 1, m))
        cs, sn = np.zeros(m), np.zeros(m)
        g = np.zeros(m + 1)
        g[0] = beta
        k = 0
        for j in range(m):
            w = M(A @ V[j])
            for i in range(j + 1):
                H[i, j] = w @ V[i]
                w = w - H[i, j] * V[i]
            H[j + 1, j] = la.norm(w)
            if H[j + 1, j] != 0.0:
                V.append(w / H[j + 1, j])
            for i in range(j):
                t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                H[i, j] = t
            d = np.hypot(H[j, j], H[j + 1, j])
            cs[j], sn[j] = (1.0, 0.0) if d == 0.0 else (H[j, j] / d, H[j + 1, j] / d)
            H[j, j] = cs[j] * H[j, j] + sn[j] * H[j + 1, j]
            H[j + 1, j] = 0.0
            g[j + 1] = -sn[j] * g[j]
            g[j] = cs[j] * g[j]
            it += 1
            k = j + 1
            res.append(abs(g[j + 1]))
            if res[-1] <= stop or len(V) <= j + 1:
                break
        y = np.linalg.solve(np.triu(H[:k, :k]), g[:k]) if k else np.zeros(0)
        for i in range(k):
            x = x + y[i] * V[i]
        if res[-1] <= stop:
            break
    return x, np.array(res), it


# ------------------------------------------------------------------ problem generators
def poisson(shape, dtype=np.float64):
    """Dirichlet (2*dim+1)-point Laplacian, lexicographic order with shape[0] (x) fastest,
    diagonal 2*dim, off-diagonals -1."""
    dims = [int(s) for s in shape]
    A = None
    for d in range(len(dims)):
        K = None
        for e in reversed(range(len(dims))):          # slowest-varying factor first
            n = dims[e]
            M = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n, n)) if e == d else sp.eye(n)
            K = M if K is None else sp.kron(K, M, format="csr")
        A = K if A is None else A + K
    A = sp.csr_matrix(A).astype(dtype)
    A.sum_duplicates()
    A.sort_indices()
    return A


# ------------------------------------------------------------------ row-partitioned variant
def build_hierarchy_partitioned(A, offsets, *, ratio=0.1, distance="unit", maxiter=10, rand=0, lam_max=None,
                                max_levels=10, max_coarse=500, replicate_below=200000, smoother="jacobi",
                                jacobi_weight=2.0 / 3.0):
    """CPU statement of the multi-GPU algorithm of ml-amg_b200/mlamg/distributed.py (the reference has no
    distributed solve — PARITY UNPINNED by construction; this defines it).  While the level has more than
    `replicate_below` rows: each block [offsets[r], offsets[r+1]) is aggregated on its own diagonal block
    (reference seeding rule per block), coarse dofs are numbered block by block, and P / A_H are the plain
    global expressions of ns/lib/multigrid.py:102-108,165.  Below the threshold the ordinary single-domain
    hierarchy continues.  Returns (levels, list of per-level offsets of the partitioned levels)."""
    A = sp.csr_matrix(A)
    offsets = np.asarray(offsets, dtype=np.int64)
    levels, all_offsets = [], []
    lvl = 0
    lam_list = list(lam_max) if isinstance(lam_max, (list, tuple)) else None
    while A.shape[0] > replicate_below and lvl < max_levels - 1:
        n = A.shape[0]
        labels = np.full(n, -1, dtype=np.int64)
        coffs = [0]
        for r in range(len(offsets) - 1):
            lo, hi = int(offsets[r]), int(offsets[r + 1])
            blk = sp.csr_matrix(A[lo:hi, lo:hi])
            blk.sort_indices()
            Agg_r, _, _ = rp.lloyd_aggregation(blk, ratio=ratio, distance=distance, maxiter=maxiter, rand=rand)
            lab = np.full(hi - lo, -1, dtype=np.int64)
            coo = Agg_r.tocoo()
            lab[coo.row] = coo.col
            labels[lo:hi] = np.where(lab >= 0, lab + coffs[-1], -1)
            coffs.append(coffs[-1] + Agg_r.shape[1])
        L = Level()
        L.A = A
        L.dw = smoother_diag(A, smoother, jacobi_weight).astype(A.dtype)
        L.labels = labels
        Agg = labels_to_agg(labels, coffs[-1])
        if lam_list is not None:
            lam = lam_list[lvl] if lvl < len(lam_list) and lam_list[lvl] is not None else rp.lambda_max_dinv_a(A)
        elif lam_max is None:
            lam = rp.lambda_max_dinv_a(A)
        else:
            lam = lam_max
        L.omega_sa = (4.0 / 3.0) / lam
        L.P = rp.canonical_csr(sp.csr_matrix(rp.smoothed_aggregation_jacobi(A, Agg, omega=L.omega_sa)))
        L.R = sp.csr_matrix(L.P.T)
        levels.append(L)
        all_offsets.append(offsets)
        A = rp.canonical_csr(rp.galerkin(L.A, L.P))
        offsets = np.asarray(coffs, dtype=np.int64)
        lvl += 1
    all_offsets.append(offsets)
    tail_lam = lam_list[lvl:] if lam_list is not None else lam_max
    if lam_list is not None:
        it = iter(tail_lam)
        tail_lam = lambda M, _it=it: next(_it)       # noqa: E731  (one value per tail level, in order)
    tail = build_hierarchy(A, ratio=ratio, distance=distance, maxiter=maxiter, rand=rand, lam_max=tail_lam,
                           max_levels=max(1, max_levels - lvl), max_coarse=max_coarse, smoother=smoother,
                           jacobi_weight=jacobi_weight)
    return levels + tail, all_offsets
