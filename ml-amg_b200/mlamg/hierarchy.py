"""Aggregation-AMG hierarchy on the device: setup from A + aggregates (+ optional learned P
weights), V-cycle / stationary / PCG solves, preconditioner apply.

API shape follows what the reference uses of pyamg's multilevel solver
(ns/preconditioner/PyAMG.py:94,119: `smoothed_aggregation_solver(A, max_levels)`,
`.solve(b, tol, accel)`, `.aspreconditioner()`), built from the reference's own blocks:
Lloyd aggregates (ns/lib/graph.py:156-239), P = (I - w D^-1 A) Agg (ns/lib/multigrid.py:102-108)
or P = P_hat Agg (ns/model/agg_interp.py:481-484), A_H = P^T A P (multigrid.py:165), Jacobi
smoothing (MLAMG.py:143-146).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import core
from ._lib import lib, check
from .core import DeviceCSR, ptr, stream


def distance_transform(C, distance):
    """Edge-length transform of ns/lib/graph.py:201-212 on the device."""
    v = C.val
    if distance == "unit":
        data = torch.ones(v.numel(), dtype=torch.float64, device=v.device)
    elif distance == "abs":
        data = v.abs()
    elif distance == "inv":
        data = 1.0 / v.abs()
    elif distance == "same":
        data = v
    elif distance == "min":
        data = v - v.min()
    else:
        raise ValueError(f"Unrecognized value distance={distance}")
    if data.numel() and float(data.min().item()) < 0:
        raise AssertionError("lloyd_aggregation: negative edge length")
    return C.with_values(data)


def lloyd_seeds(N, ratio, rand):
    """Seeding of ns/lib/graph.py:214-231 (host RNG: it defines the reference's seeds)."""
    if ratio <= 0 or ratio > 1:
        raise ValueError("ratio must be > 0.0 and <= 1.0")
    num_seeds = int(np.ceil(ratio * N))
    if isinstance(rand, (int, np.integer)) and not isinstance(rand, bool) and 0 <= int(rand) < 2 ** 32 and N < 2 ** 31:
        # integer seed: the same draws as RandomState(rand).permutation(N), by the extension's host routine (numpy's own
        # loop is the largest single item of the setup at 16.7 M rows)
        out = np.empty(num_seeds, dtype=np.int32)
        check(lib.mlamg_legacy_permutation_head(int(rand), int(N), num_seeds, out.ctypes.data_as(ctypes.c_void_p)))
        return out.astype(np.int64)
    if rand is None:
        rand = np.random
    elif isinstance(rand, (int, np.integer)):
        rand = np.random.RandomState(int(rand))
    elif not isinstance(rand, np.random.RandomState):
        raise TypeError("rand should be an integer seed value or a random state")
    return rand.permutation(N)[:num_seeds]


def lloyd_labels(C, ratio=0.03, distance="unit", maxiter=10, rand=None):
    """-> (labels int32[N] device, num_seeds, roots device, seeds host)."""
    G = distance_transform(C, distance)
    seeds = lloyd_seeds(C.shape[0], ratio, rand)
    _, clusters, roots, _ = core.lloyd_cluster(G, seeds.astype(np.int32), maxiter=maxiter)
    return clusters, len(seeds), roots, seeds


def sa_prolongator(A, Agg, omega, drop=True):
    """P = (I - omega D^-1 A) Agg with scipy's stored-pattern semantics (exact zeros dropped)."""
    P = core.spgemm(core.sa_smoother(A, omega), Agg if Agg.dtype == A.dtype else Agg.astype(A.dtype))
    return core.drop_zeros(P) if drop else P


def learned_prolongator(P_hat, Agg, drop=False):
    """P = P_hat Agg (torch.sparse.mm + coalesce keeps explicit zeros -> drop=False)."""
    P = core.spgemm(P_hat, Agg if Agg.dtype == P_hat.dtype else Agg.astype(P_hat.dtype))
    return core.drop_zeros(P) if drop else P


def galerkin(A, P, R=None, drop=True, symmetric=False):
    """A_H = P^T A P, evaluated in the order scipy evaluates `P.T @ A @ P` (multigrid.py:165): P.T is a CSC
    view, so scipy computes X^T = A^T P row by row, then A_H^T = P^T X^T, each entry summed sequentially
    over the sorted outer index; with the ordered SpGEMM the result is bit-identical to scipy's.
    symmetric=True skips the transpose of A (caller asserts A == A^T bitwise)."""
    if R is None:
        R = core.transpose(P)
    At = A if symmetric else core.transpose(A)
    AHt = core.spgemm(R, core.spgemm(At, P))
    del At
    AH = core.transpose(AHt)
    return core.drop_zeros(AH) if drop else AH


def post_operator(A, P, dw):
    """Q = (I - diag(dw) A) P — the prolongator seen through one smoothing sweep (A must store its diagonal)."""
    n = A.shape[0]
    rows = torch.repeat_interleave(torch.arange(n, device=A.col.device), (A.rowptr[1:] - A.rowptr[:-1]).long())
    sval = -dw[rows] * A.val
    sval += (A.col.long() == rows).to(sval.dtype)
    del rows
    return core.spgemm(A.with_values(sval), P)


def _permuted(M, row_new2old, col_old2new):
    """rows gathered by row_new2old (None = keep), column ids mapped by col_old2new (None = keep), rows re-sorted"""
    if row_new2old is None and col_old2new is None:
        return M
    rowptr, col, val = M.rowptr, M.col, M.val
    if row_new2old is not None:
        lens = (rowptr[1:] - rowptr[:-1]).long()[row_new2old]
        starts = rowptr.long()[row_new2old]
        excl = torch.cumsum(lens, 0) - lens
        ent = torch.repeat_interleave(starts - excl, lens) + torch.arange(int(lens.sum()), device=col.device)
        new_rowptr = torch.zeros(lens.numel() + 1, dtype=torch.int64, device=col.device)
        new_rowptr[1:] = torch.cumsum(lens, 0)
        rowptr, col, val = new_rowptr.to(torch.int32), col[ent], val[ent]
    if col_old2new is not None:
        col = col_old2new[col.long()].to(torch.int32)
    out = DeviceCSR(rowptr.contiguous(), col.contiguous().clone(), val.contiguous().clone(), M.shape)
    return core.sort_rows(out)


class Level:
    __slots__ = ("A", "P", "R", "dw", "labels", "omega_sa", "seeds", "roots", "sell")

    def __init__(self, A):
        self.A = A
        self.P = self.R = self.dw = self.labels = self.omega_sa = self.seeds = self.roots = self.sell = None


class Hierarchy:
    """Owns the device arrays of every level and the C handle that runs the cycles."""

    def __init__(self, levels, smoother="jacobi", jacobi_weight=2.0 / 3.0, use_graph=False, sell="never",
                 sell_max_padding=1.5, restrict_order=True, renumber=True, fuse_post=True, fuse_pre=True, coarse_inv=None,
                 w32_min_rows=100000, w32_fine_operator=False):
        """levels: list[Level] in the REFERENCE numbering (what setup produced and what parity is checked on).

        renumber (default on): the cycle runs on apply copies whose coarse levels are renumbered spatially
        (coarse dof c -> rank of its first fine node, recursively).  The reference numbers aggregates by a random
        seed permutation, so with its numbering every gather of a coarse vector (prolongation, level-1 operator)
        hits a different 32-byte L2 sector: measured at 256^3 the prolongation was L2-transaction bound
        (211 us for 0.88 GB).  Coarse vectors are internal to the cycle, so the renumbering is invisible outside;
        results change only by the re-ordered floating-point sums (within the 1e-12 parity bar).
        fuse_post (default on): per level Q = (I - D_w A) P is built once, so that the prolongation and the first
        post-smoothing sweep run as ONE pass over Q:  x + P e followed by x + dw.*(b - A x) equals
        x + dw.*r + Q e  with the residual r the cycle already computed for the restriction.  Same arithmetic up
        to rounding (well inside the 1e-12 bar), one pass over A less per level and cycle.
        fuse_pre (default on): a copy of every level's values scaled by columns (A D_w) lets the first pass of a
        V(1,*) cycle from a zero guess — x = dw.*b, r = b - A x — gather b alone: r = b - (A D_w) b.
        sell: 'never' (default) -> CSR kernels only: measured on B200 at 256^3 the thread-per-row CSR kernel
        with predicated 4-entry batches (0.322 ms/sweep) is as fast as SELL-32 (0.337 ms), so the second copy
        of A is not worth its HBM; 'auto' / 'always' build SELL-32 copies for the smoother/residual kernels."""
        core.require_cuda()
        self.levels = levels
        self.dtype = levels[0].A.dtype
        self.smoother = smoother
        for lev in levels[:-1]:
            if lev.dw is None:
                lev.dw = core.smoother_diag(lev.A, smoother, jacobi_weight)
        # coarse_inv: caller-supplied dense (pseudo-)inverse of the coarsest operator (the singular mode of amg_2_v)
        self.coarse_inv = core.dense_inverse(levels[-1].A) if coarse_inv is None else coarse_inv.to(self.dtype).contiguous()
        self.renumbered = bool(renumber) and len(levels) > 2
        self._apply = self._renumbered_copies() if self.renumbered else [(l.A, l.P, l.R, l.dw) for l in levels]
        self._h = ctypes.c_void_p()
        check(lib.mlamg_hierarchy_create(core.dt(self.dtype), len(levels), ctypes.byref(self._h)))
        for l, (A, P, R, dw) in enumerate(self._apply):
            check(lib.mlamg_hierarchy_set_operator(self._h, l, A.shape[0], A.nnz, ptr(A.rowptr), ptr(A.col), ptr(A.val),
                                                   ptr(dw)))
        if sell != "never":
            for l, lev in enumerate(levels[:-1]):
                sl = core.DeviceSELL(self._apply[l][0])
                if sell == "always" or sl.padding <= sell_max_padding:
                    lev.sell = sl
                    check(lib.mlamg_hierarchy_set_operator_sell(self._h, l, ptr(sl.slice_ptr), ptr(sl.col), ptr(sl.val)))
        for l, (A, P, R, dw) in enumerate(self._apply[:-1]):
            check(lib.mlamg_hierarchy_set_transfer(self._h, l, P.nnz, ptr(P.rowptr), ptr(P.col), ptr(P.val),
                                                   ptr(R.rowptr), ptr(R.col), ptr(R.val)))
        self._scaled = []
        if fuse_pre:
            for l, (A, P, R, dw) in enumerate(self._apply[:-1]):
                vs = core.scaled_values(A, dw)
                self._scaled.append(vs)
                check(lib.mlamg_hierarchy_set_operator_scaled(self._h, l, ptr(vs)))
        self._Q = []
        if fuse_post:
            for l, (A, P, R, dw) in enumerate(self._apply[:-1]):
                Q = post_operator(A, P, dw)
                self._Q.append(Q)
                check(lib.mlamg_hierarchy_set_post_operator(self._h, l, Q.nnz, ptr(Q.rowptr), ptr(Q.col), ptr(Q.val)))
        # W32 copies (slot-major inside 32-row windows): thread-per-row kernels with a coalesced operator stream for the
        # residual on the scaled copy and for x = dw.*(b + r) + Q e.  Levels with >= w32_min_rows rows (below that one
        # thread per row does not fill the GPU; measured at 256^3: level 1 -18 us per cycle, level 2 slower).  The
        # 7-entry fine operator is DRAM-bound in plain CSR already (w32_fine_operator=False: no second copy of A).
        self._w32 = {}
        for l, (A, P, R, dw) in enumerate(self._apply[:-1]):
            if A.shape[0] < w32_min_rows or not (self._scaled and self._Q):
                continue
            a32 = None
            if A.nnz > 12 * A.shape[0] or w32_fine_operator:
                a32 = core.csr_to_w32(A.with_values(self._scaled[l]))
            q32 = core.csr_to_w32(self._Q[l])
            self._w32[l] = (a32, q32)
            check(lib.mlamg_hierarchy_set_w32(self._h, l, ptr(a32[0]) if a32 else None, ptr(a32[1]) if a32 else None,
                                              ptr(q32[0]), ptr(q32[1])))
        # Without renumbering the restriction rows keep the reference's random seed numbering; then at least
        # VISIT them in the order of their first fine node so neighbouring aggregates share sectors of r.
        self._r_order = []
        if restrict_order:
            for l, (A, P, R, dw) in enumerate(self._apply[:-1]):
                if self.renumbered and l + 1 < len(levels) - 1:
                    continue                       # rows of a renumbered level are already in spatial order
                first = R.col[R.rowptr[:-1].long().clamp(max=max(R.nnz - 1, 0))]
                order = torch.argsort(first, stable=True).to(torch.int32).contiguous()
                self._r_order.append(order)
                check(lib.mlamg_hierarchy_set_restrict_order(self._h, l, ptr(order)))
        check(lib.mlamg_hierarchy_set_coarse_inverse(self._h, ptr(self.coarse_inv)))
        check(lib.mlamg_hierarchy_finalize(self._h, stream()))
        if use_graph:
            self.use_graph(True)

    def _renumbered_copies(self):
        """(A, P, R, dw) per level with levels 1..L-2 renumbered spatially (the coarsest keeps its numbering,
        so the dense inverse is untouched)."""
        levels = self.levels
        L = len(levels)
        dev = levels[0].A.val.device
        old2new = [None] * L                      # None = identity
        new2old = [None] * L
        for l in range(1, L - 1):
            R = levels[l - 1].R                   # rows: level-l dofs, cols: level l-1 dofs (reference numbering)
            cols = R.col.long()
            if old2new[l - 1] is not None:
                cols = old2new[l - 1][cols]
            rows = torch.repeat_interleave(torch.arange(R.shape[0], device=dev), (R.rowptr[1:] - R.rowptr[:-1]).long())
            key = torch.full((R.shape[0],), 2 ** 62, dtype=torch.int64, device=dev)
            key.scatter_reduce_(0, rows, cols, reduce="amin")
            order = torch.argsort(key, stable=True)            # new -> old
            inv = torch.empty_like(order)
            inv[order] = torch.arange(order.numel(), device=dev)
            new2old[l], old2new[l] = order, inv
        out = []
        for l, lev in enumerate(levels):
            A = _permuted(lev.A, new2old[l], old2new[l])
            dw = lev.dw if (new2old[l] is None or lev.dw is None) else lev.dw[new2old[l]].contiguous()
            if l < L - 1:
                P = _permuted(lev.P, new2old[l], old2new[l + 1])
                R = core.transpose(P) if (new2old[l] is not None or old2new[l + 1] is not None) else lev.R
            else:
                P = R = None
            out.append((A, P, R, dw))
        return out

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            lib.mlamg_hierarchy_destroy(h)
            self._h = None

    # -- info -------------------------------------------------------------------------------
    @property
    def shape(self):
        return self.levels[0].A.shape

    def use_graph(self, enable=True):
        check(lib.mlamg_hierarchy_use_graph(self._h, 1 if enable else 0))

    def cycle_bytes(self, nu1=1, nu2=1, zero_guess=True):
        return float(lib.mlamg_hierarchy_cycle_bytes(self._h, nu1, nu2, 1 if zero_guess else 0))

    def operator_complexity(self):
        return sum(l.A.nnz for l in self.levels) / self.levels[0].A.nnz

    def __repr__(self):
        rows = [f"  level {i}: n={l.A.shape[0]:>10d} nnz={l.A.nnz:>12d}" for i, l in enumerate(self.levels)]
        return (f"mlamg Hierarchy({len(self.levels)} levels, {self.dtype}, smoother={self.smoother}, "
                f"operator complexity {self.operator_complexity():.3f})\n" + "\n".join(rows))

    # -- cycles -----------------------------------------------------------------------------
    def vcycle(self, b, x=None, nu1=1, nu2=1):
        """One V(nu1,nu2) cycle.  x=None: zero initial guess (preconditioner apply), result returned;
        otherwise x is updated in place."""
        zero = x is None
        if zero:
            x = torch.empty_like(b)
        check(lib.mlamg_vcycle(self._h, ptr(b), ptr(x), nu1, nu2, 1 if zero else 0, stream()))
        return x

    def solve(self, b, x0=None, tol=1e-8, maxiter=100, nu1=1, nu2=1, cycle="V", accel=None, residuals=None,
              return_residuals=False, restart=30):
        """pyamg-style solve.  accel=None: stationary V-cycles until ||b-Ax|| <= tol*||b|| ;
        accel='cg': V-cycle preconditioned CG; accel='gmres': left-preconditioned restarted GMRES (the call of
        ns/preconditioner/PyAMG.py:119; history = preconditioned residual norms, stop at tol*||M b||).
        Accepts numpy or CUDA tensors, returns the same kind."""
        if cycle != "V":
            raise NotImplementedError("only V cycles are built on this path")
        is_np = isinstance(b, np.ndarray)
        bd = core.as_vec(b, self.dtype)
        xd = torch.zeros_like(bd) if x0 is None else core.as_vec(x0, self.dtype).clone()
        res = (ctypes.c_double * (maxiter + 1))()
        nit = ctypes.c_int(0)
        if accel is None:
            nb = float(torch.linalg.vector_norm(bd).item())
            check(lib.mlamg_solve(self._h, ptr(bd), ptr(xd), nu1, nu2, tol * (nb if nb != 0 else 1.0), maxiter, res,
                                  ctypes.byref(nit), stream()))
        elif accel == "cg":
            check(lib.mlamg_pcg(self._h, ptr(bd), ptr(xd), nu1, nu2, tol, maxiter, res, ctypes.byref(nit), stream()))
        elif accel == "gmres":
            hist = self._gmres(bd, xd, tol, maxiter, nu1, nu2, restart)
            if residuals is not None:
                residuals[:] = list(hist)
            out = xd.cpu().numpy() if is_np else xd
            return (out, hist) if return_residuals else out
        else:
            raise NotImplementedError(f"accel={accel!r}: None, 'cg' and 'gmres' are built")
        hist = np.array(res[:nit.value + 1])
        if residuals is not None:
            residuals[:] = list(hist)
        out = xd.cpu().numpy() if is_np else xd
        return (out, hist) if return_residuals else out

    def solve_device(self, b, x, tol=1e-8, maxiter=100, nu1=1, nu2=1, accel="cg"):
        """solve() on device tensors without any host copy of the vectors: x (initial guess) is updated in place.
        -> (x, residual history)"""
        res = (ctypes.c_double * (maxiter + 1))()
        nit = ctypes.c_int(0)
        if accel == "cg":
            check(lib.mlamg_pcg(self._h, ptr(b), ptr(x), nu1, nu2, tol, maxiter, res, ctypes.byref(nit), stream()))
        elif accel is None:
            nb = float(torch.linalg.vector_norm(b).item())
            check(lib.mlamg_solve(self._h, ptr(b), ptr(x), nu1, nu2, tol * (nb if nb != 0 else 1.0), maxiter, res,
                                  ctypes.byref(nit), stream()))
        else:
            raise NotImplementedError(f"solve_device: accel={accel!r}")
        return x, np.array(res[:nit.value + 1])

    def _gmres(self, b, x, tol, maxiter, nu1, nu2, restart):
        """Left-preconditioned GMRES(restart), modified Gram-Schmidt + Givens; x updated in place.  The operator
        applications (A v, one V-cycle per Krylov vector) run in libmlamg_b200.so and the whole Gram-Schmidt step of a
        Krylov vector is ONE call with one host sync (mlamg_gmres_orthogonalize); the (restart+1) x restart Hessenberg
        lives on the host.  Same algorithm as oracle.multilevel.gmres."""
        A = self.levels[0].A
        n = A.shape[0]
        t = torch.empty(n, dtype=self.dtype, device=b.device)

        def precond_residual():
            core.residual(A, x, b, out=t)
            return self.vcycle(t, None, nu1, nu2)

        nmb = float(torch.linalg.vector_norm(self.vcycle(b, None, nu1, nu2)).item())
        stop = tol * (nmb if nmb != 0 else 1.0)
        res, it = [], 0
        basis = None
        while True:
            r = precond_residual()
            beta = float(np.sqrt(core.dot(r, r)))
            if not res:
                res.append(beta)
            if beta <= stop or it >= maxiter:
                break
            m = min(int(restart), maxiter - it)
            if basis is None or basis.shape[0] < m + 1:
                basis = torch.empty(m + 1, n, dtype=self.dtype, device=b.device)      # Krylov basis, contiguous
            torch.mul(r, 1.0 / beta, out=basis[0])
            nvec = 1
            Hm = np.zeros((m + 1, m))
            hcol = (ctypes.c_double * (m + 2))()
            cs, sn, g = np.zeros(m), np.zeros(m), np.zeros(m + 1)
            g[0] = beta
            k = 0
            for j in range(m):
                core.spmv(A, basis[j], out=t)
                w = self.vcycle(t, None, nu1, nu2)
                # modified Gram-Schmidt against V[0..j], the norm and the next basis vector: one C call, one host sync
                check(lib.mlamg_gmres_orthogonalize(core.dt(self.dtype), n, j, ptr(basis), ptr(w), ptr(basis[j + 1]), hcol,
                                                    stream()))
                Hm[:j + 2, j] = hcol[:j + 2]
                if Hm[j + 1, j] != 0.0:
                    nvec += 1
                for i in range(j):
                    tmp = cs[i] * Hm[i, j] + sn[i] * Hm[i + 1, j]
                    Hm[i + 1, j] = -sn[i] * Hm[i, j] + cs[i] * Hm[i + 1, j]
                    Hm[i, j] = tmp
                d = np.hypot(Hm[j, j], Hm[j + 1, j])
                cs[j], sn[j] = (1.0, 0.0) if d == 0.0 else (Hm[j, j] / d, Hm[j + 1, j] / d)
                Hm[j, j] = cs[j] * Hm[j, j] + sn[j] * Hm[j + 1, j]
                Hm[j + 1, j] = 0.0
                g[j + 1] = -sn[j] * g[j]
                g[j] = cs[j] * g[j]
                it += 1
                k = j + 1
                res.append(abs(g[j + 1]))
                if res[-1] <= stop or nvec <= j + 1:
                    break
            y = np.linalg.solve(np.triu(Hm[:k, :k]), g[:k]) if k else np.zeros(0)
            for i in range(k):
                core.axpby(float(y[i]), basis[i], 1.0, x)
            if res[-1] <= stop:
                break
        return np.array(res)

    SOLVE_XNORM, SOLVE_REMOVE_MEAN, SOLVE_NO_INITIAL_CHECK = 1, 2, 4

    def solve_abs(self, b, x, tol_abs, maxiter, nu1=1, nu2=1, flags=0):
        """Stationary iteration with an ABSOLUTE tolerance (MLAMG.py:189-195, multigrid.py:173-199), device tensors, x
        updated in place.  The whole loop runs on the device (one graph launch, one host sync).  flags: SOLVE_XNORM
        measures ||x|| instead of ||b - A x||, SOLVE_REMOVE_MEAN subtracts the mean after every cycle,
        SOLVE_NO_INITIAL_CHECK always runs one cycle first.  -> (x, history incl. the initial value)."""
        res = (ctypes.c_double * (maxiter + 1))()
        nit = ctypes.c_int(0)
        check(lib.mlamg_solve_ex(self._h, ptr(b), ptr(x), nu1, nu2, int(flags), float(tol_abs), maxiter, res, ctypes.byref(nit),
                                 stream()))
        return x, np.array(res[:nit.value + 1])

    @property
    def loop_mode(self):
        """'while-node' when the solver loops run as a CUDA-graph WHILE node, 'host' for the fallback, None before a solve"""
        return {1: "while-node", -1: "host", 0: None}[int(lib.mlamg_solver_loop_mode(self._h))]

    def apply_host(self, b_host, x_host, nu1=1, nu2=1, cycles=1):
        """Preconditioner apply on HOST arrays (PETSc PC.apply shape): H2D, V-cycle(s), D2H inside the call.
        b_host/x_host: numpy arrays or CPU tensors (pinned memory gives the best copy rate)."""
        bp = b_host.ctypes.data if isinstance(b_host, np.ndarray) else b_host.data_ptr()
        xp = x_host.ctypes.data if isinstance(x_host, np.ndarray) else x_host.data_ptr()
        check(lib.mlamg_vcycle_host(self._h, ctypes.c_void_p(bp), ctypes.c_void_p(xp), nu1, nu2, cycles, stream()))
        return x_host

    def aspreconditioner(self, cycle="V", nu1=1, nu2=1):
        from scipy.sparse.linalg import LinearOperator
        n = self.shape[0]
        npdt = np.float64 if self.dtype == torch.float64 else np.float32

        def matvec(b):
            b = np.ascontiguousarray(b, dtype=npdt).ravel()
            x = np.empty_like(b)
            self.apply_host(b, x, nu1, nu2, 1)
            return x
        return LinearOperator((n, n), matvec=matvec, dtype=npdt)


def build_hierarchy(A, *, aggregates="lloyd", ratio=0.1, distance="unit", maxiter=10, rand=0, lam_max=None,
                    P_hat=None, max_levels=10, max_coarse=500, smoother="jacobi", jacobi_weight=2.0 / 3.0,
                    dtype=None, use_graph=False, keep_labels=True, max_dense=20000, sell="never", fallback=None,
                    fuse_post=True, fuse_pre=True):
    """Build the multilevel hierarchy on the device.

    A           : scipy / torch sparse / DeviceCSR
    aggregates  : 'lloyd' (Lloyd clustering per level, reference seeding) or a list of
                  (labels, ncoarse) per level (learned / external aggregates)
    fallback    : None -> stop coarsening when the list of external aggregates is exhausted;
                  'lloyd' -> continue with Lloyd + smoothed aggregation below the learned levels
    P_hat       : optional list of per-level weights on A_l's pattern (learned P = P_hat Agg)
    lam_max     : |lambda_max(D^-1 A)| per level: None -> on the device (Lanczos, core.lambda_max); float, list or
                  callable(DeviceCSR) -> supplied (parity runs pass the oracle's value)
    """
    core.require_cuda()
    A = DeviceCSR.wrap(A, dtype)
    levels = []
    lvl = 0
    while True:
        L = Level(A)
        levels.append(L)
        n = A.shape[0]
        if len(levels) >= max_levels or n <= max_coarse:
            break
        if isinstance(aggregates, str) or (lvl >= len(aggregates) and fallback == "lloyd"):
            if isinstance(aggregates, str) and aggregates != "lloyd":
                raise ValueError(f"unknown aggregation strategy {aggregates!r}")
            labels, nc, roots, seeds = lloyd_labels(A, ratio=ratio, distance=distance, maxiter=maxiter, rand=rand)
            L.roots, L.seeds = roots, seeds
        else:
            if lvl >= len(aggregates):
                break
            labels, nc = aggregates[lvl]
            labels = core.as_i32(labels)
        Agg = core.agg_from_labels(labels, nc, A.dtype)
        if keep_labels:
            L.labels = labels
        if P_hat is not None and lvl < len(P_hat) and P_hat[lvl] is not None:
            ph = P_hat[lvl]
            ph = ph if isinstance(ph, DeviceCSR) else A.with_values(core.as_vec(ph, A.dtype))
            P = learned_prolongator(ph, Agg)
        else:
            if callable(lam_max):
                lam = lam_max(A)
            elif lam_max is None:
                lam = core.lambda_max(A)
            elif np.isscalar(lam_max):
                lam = float(lam_max)
            else:
                lam = float(lam_max[lvl])
            L.omega_sa = (4.0 / 3.0) / lam
            P = sa_prolongator(A, Agg, L.omega_sa)
        L.P = P
        L.R = core.transpose(P)
        A = galerkin(L.A, P, L.R)
        lvl += 1
    if levels[-1].A.shape[0] > max_dense:
        raise _lib.MlamgError(_lib.ELIMIT, f"coarsest level has {levels[-1].A.shape[0]} rows; raise max_levels or "
                                           f"lower max_coarse (dense coarse solve limit {max_dense})")
    return Hierarchy(levels, smoother=smoother, jacobi_weight=jacobi_weight, use_graph=use_graph, sell=sell,
                     fuse_post=fuse_post, fuse_pre=fuse_pre)
