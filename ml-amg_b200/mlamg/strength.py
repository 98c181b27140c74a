"""Evolution strength of connection on the device — `pyamg.strength.evolution_strength_of_connection(A)` with
its defaults (epsilon=4, k=2, B=ones, proj_type='l2', symmetrize_measure=True), which the reference calls for its
'evolution' and DEFAULT 'olson' strength measures (/root/reference/utils/common.py:27,30 -> the distance matrix of
`lloyd_aggregation`, utils/common.py:52-58).

Steps (pyamg/strength.py, 4.x), one kernel of csrc/strength.cu each:
  rho   = approximate_spectral_radius(D^-1 A)        restarted Arnoldi, 15 steps x <= 6 cycles, tol 0.01, start vector
                                                     np.random.rand(n, 1) from numpy's GLOBAL stream (the reference seeds it
                                                     right before, common.py:50,88): SpMV / dot / axpby on the device, the
                                                     15 x 15 Hessenberg eigenproblem on the host
  S     = I - (1/rho) D^-1 A ; T = S^T               mlamg_evolution_step, mlamg_csr_transpose
  Z     = T^k restricted to A's pattern              k = 2: mlamg_incomplete_matmul_csr(T, S); k = 2^m: m-1 ordered SpGEMM squarings first
  M     = |1 - z_ii / z_ij| + weak / angle rules     mlamg_evolution_measure, exact-zero compaction
  drop  off-diagonals >= 4 * row minimum             mlamg_distance_filter, compaction
  sym   0.5 (M + M^T), unit diagonal                 mlamg_evolution_symmetrize (on A's pattern), compaction
  out   1/M, rows scaled by their largest entry      mlamg_invert_scale_rows

The arithmetic is non-fused and in pyamg's order: with the same rho the result has the bits of the CPU evaluation
(oracle.pyamg_restated.evolution_strength_of_connection).  Limits: CSR with a stored diagonal and a symmetric PATTERN
(values may be non-symmetric), real fp64 / fp32, k a power of two (pyamg itself warns about any other k).
"""
import numpy as np
import scipy.linalg
import torch

from . import core
from ._lib import lib, check


def _pattern_is_symmetric(A):
    tags = torch.zeros(A.nnz, dtype=A.val.dtype, device=A.val.device)
    At = core.transpose(core.DeviceCSR(A.rowptr, A.col, tags, A.shape))
    return bool(torch.equal(At.rowptr, A.rowptr)) and bool(torch.equal(At.col, A.col))


def evolution_step(A, rho, want_dinv_a=False):
    """-> (S = I - (1/rho) D^-1 A on A's pattern, D^-1 A or None)"""
    n = A.shape[0]
    s_val = torch.empty_like(A.val)
    d_val = torch.empty_like(A.val) if want_dinv_a else None
    flags = torch.zeros(1, dtype=torch.int32, device=A.val.device)
    check(lib.mlamg_evolution_step(core.dt(A.val), n, core.ptr(A.rowptr), core.ptr(A.col), core.ptr(A.val), 1.0 / float(rho),
                                   core.ptr(s_val), core.ptr(d_val), core.ptr(flags), core.stream()))
    if int(flags.item()) & 1:
        raise ValueError("evolution strength needs a matrix that stores its diagonal")
    return A.with_values(s_val), (A.with_values(d_val) if want_dinv_a else None)


def incomplete_matmul(T, Bt, S):
    """values of (T B) on the pattern of S; Bt = the CSR arrays of B^T (B in CSC), all indices sorted"""
    out = torch.empty(S.nnz, dtype=T.val.dtype, device=T.val.device)
    check(lib.mlamg_incomplete_matmul_csr(core.dt(T.val), S.shape[0], core.ptr(T.rowptr), core.ptr(T.col), core.ptr(T.val),
                                          core.ptr(Bt.rowptr), core.ptr(Bt.col), core.ptr(Bt.val), core.ptr(S.rowptr),
                                          core.ptr(S.col), core.ptr(out), core.stream()))
    return out


def evolution_measure_(Z):
    check(lib.mlamg_evolution_measure(core.dt(Z.val), Z.shape[0], core.ptr(Z.rowptr), core.ptr(Z.col), core.ptr(Z.val),
                                      core.stream()))
    return Z


def distance_filter_(M, epsilon):
    check(lib.mlamg_distance_filter(core.dt(M.val), M.shape[0], float(epsilon), core.ptr(M.rowptr), core.ptr(M.col),
                                    core.ptr(M.val), core.stream()))
    return M


def symmetrize_on(A, M, symmetrize=True):
    out = torch.empty(A.nnz, dtype=M.val.dtype, device=M.val.device)
    check(lib.mlamg_evolution_symmetrize(core.dt(M.val), A.shape[0], core.ptr(A.rowptr), core.ptr(A.col), core.ptr(M.rowptr),
                                         core.ptr(M.col), core.ptr(M.val), 1 if symmetrize else 0, core.ptr(out), core.stream()))
    return core.DeviceCSR(A.rowptr, A.col, out, A.shape)


def invert_scale_rows_(M):
    check(lib.mlamg_invert_scale_rows(core.dt(M.val), M.shape[0], core.ptr(M.rowptr), core.ptr(M.val), core.stream()))
    return M


def pattern_add(A, w, E):
    """E + W for W = (A's pattern, values w), E stored on a sub-pattern of A's -> DeviceCSR on A's pattern"""
    out = torch.empty_like(w)
    check(lib.mlamg_csr_pattern_add(core.dt(w), A.shape[0], core.ptr(A.rowptr), core.ptr(A.col), core.ptr(w), core.ptr(E.rowptr),
                                    core.ptr(E.col), core.ptr(E.val), core.ptr(out), core.stream()))
    return core.DeviceCSR(A.rowptr, A.col, out, A.shape)


def approximate_spectral_radius(M, tol=0.01, maxiter=15, restart=5, return_trace=False):
    """pyamg.util.linalg.approximate_spectral_radius(M) with its defaults for a real DeviceCSR M: the 1 %-accurate
    Arnoldi estimate the reference's measure is built on (NOT the spectral radius — `core.lambda_max` gives that)."""
    n = M.shape[0]
    dev, dtype = M.val.device, M.val.dtype
    breakdown = np.finfo(float).eps * 1e6
    v0 = torch.from_numpy(np.random.rand(n, 1).ravel()).to(device=dev, dtype=dtype)
    trace = []
    rho = 0.0
    for _ in range(restart + 1):
        m = min(n, maxiter)
        v0 = v0 / float(np.sqrt(core.dot(v0, v0)))
        H = np.zeros((m + 1, m))
        V = [v0.contiguous()]
        breakdown_flag = False
        j = 0
        for j in range(m):
            w = core.spmv(M, V[-1])
            for i, v in enumerate(V):                   # modified Gram-Schmidt against every stored vector
                H[i, j] = core.dot(v, w)
                core.axpby(-float(H[i, j]), v, 1.0, w)
            H[j + 1, j] = np.sqrt(core.dot(w, w))
            if H[j + 1, j] < breakdown:
                breakdown_flag = True
                if H[j + 1, j] != 0.0:
                    w = w / float(H[j + 1, j])
                V.append(w)
                break
            V.append(w / float(H[j + 1, j]))
        ev, evect = scipy.linalg.eig(H[:j + 1, :j + 1], left=False, right=True)
        nvecs = ev.shape[0]
        max_index = int(np.abs(ev).argmax())
        error = H[nvecs, nvecs - 1] * evect[-1, max_index]
        rho = float(np.abs(ev[max_index]))
        trace.append((rho, float(np.abs(error))))
        if (np.abs(error) / np.abs(ev[max_index]) < tol) or breakdown_flag:
            break
        # restart from the Ritz vector of the dominant eigenvalue (pyamg forms it before the test and drops it on exit)
        y = evect[:, max_index]
        if np.iscomplexobj(y) and np.any(y.imag != 0):
            raise NotImplementedError("restart from a complex Ritz vector (non-symmetric operator) is not supported on the device path")
        y = np.real(y)
        v0 = torch.zeros(n, dtype=dtype, device=dev)
        for c, v in zip(y, V[:-1]):
            core.axpby(float(c), v, 1.0, v0)
    return (rho, trace) if return_trace else rho


def evolution_strength_of_connection(A, epsilon=4.0, k=2, symmetrize_measure=True, rho=None):
    """A: scipy CSR / DeviceCSR (real, stored diagonal, symmetric pattern) -> DeviceCSR strength matrix (large = strong,
    rows scaled to a largest entry of 1, unit diagonal).  `rho`: inject rho(D^-1 A) instead of the Arnoldi estimate."""
    core.require_cuda()
    if epsilon < 1.0:
        raise ValueError("expected epsilon > 1.0")
    if k <= 0:
        raise ValueError("number of time steps must be > 0")
    nsquare = int(np.log2(k))
    if k != 2 ** nsquare:
        raise NotImplementedError("evolution strength: k must be a power of two")
    A = core.drop_zeros(core.DeviceCSR.wrap(A))                 # A.eliminate_zeros(); A.sort_indices()
    if A.shape[0] != A.shape[1]:
        raise ValueError("expected square matrix")
    if not _pattern_is_symmetric(A):
        raise NotImplementedError("evolution strength: the pattern of A must be symmetric")
    if rho is None:
        _, DinvA = evolution_step(A, 1.0, want_dinv_a=True)
        rho = approximate_spectral_radius(DinvA)
    S, _ = evolution_step(A, rho)
    T = core.transpose(S)                                       # Atilde = (I - D^-1 A / rho)^T, already on A's pattern
    if nsquare == 0:
        Z = core.DeviceCSR(T.rowptr, T.col, T.val.clone(), T.shape)
    else:
        for _ in range(nsquare - 1):
            T = core.spgemm(T, T)
        Bt = S if nsquare == 1 else core.transpose(T)           # CSC of Atilde = CSR of its transpose
        Z = core.DeviceCSR(A.rowptr, A.col, incomplete_matmul(T, Bt, A), A.shape)
    M = core.drop_zeros(evolution_measure_(Z))
    if epsilon != np.inf:
        M = core.drop_zeros(distance_filter_(M, epsilon))
    M = core.drop_zeros(symmetrize_on(A, M, symmetrize_measure))
    return invert_scale_rows_(M)


def olson_measure(A, **kw):
    """utils/common.py:30: evolution_strength_of_connection(A) + 1/|A| on A's pattern -> DeviceCSR"""
    Ad = core.drop_zeros(core.DeviceCSR.wrap(A))
    return pattern_add(Ad, 1.0 / Ad.val.abs(), evolution_strength_of_connection(Ad, **kw))


def evolution_measure_plus_pattern(A, **kw):
    """utils/common.py:27: evolution_strength_of_connection(A) + 0.1 * pattern(A) -> DeviceCSR"""
    Ad = core.drop_zeros(core.DeviceCSR.wrap(A))
    return pattern_add(Ad, torch.ones_like(Ad.val) * 0.1, evolution_strength_of_connection(Ad, **kw))
