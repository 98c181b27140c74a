"""Drop-in mirror of the reference's `ns.lib.multigrid` on the B200 kernels.

Reference: /root/reference/ns/lib/multigrid.py
  jacobi :15-45, jacobi_torch :48-55, gauss_seidel :58-90, smoothed_aggregation_jacobi :102-108,
  amg_2_v :111-210, amg_2_v_torch :213-245.
Same names, argument order, defaults, return tuples and error behaviour; numpy/scipy in -> numpy
out, torch in -> torch out.  Extra keyword arguments (always last, optional) expose what the GPU
path adds: `smoother=` on amg_2_v ('gauss_seidel' = the reference's pyamg sweep, exact;
'jacobi' / 'l1_jacobi' = the fused kernels) and `omega=` / `lam_max=` on
smoothed_aggregation_jacobi (the reference calls ARPACK, 60-160 s).
"""
import numpy as np
import scipy.sparse as sp
import torch

import mlamg
from mlamg import core, SingularCoarseError


def _diag_vector(Dinv, n):
    if Dinv is None:
        return None
    if sp.issparse(Dinv):
        return np.asarray(Dinv.diagonal())
    Dinv = np.asarray(Dinv)
    return np.diag(Dinv) if Dinv.ndim == 2 else Dinv


def jacobi(A, b, x, Dinv=None, omega=0.666, nu=2):
    """Weighted Jacobi, x updated in place and returned:  x += omega Dinv b - omega Dinv A x."""
    Ad = core.DeviceCSR.wrap(A)
    dv = _diag_vector(Dinv, Ad.shape[0])
    if dv is None:
        dw = core.smoother_diag(Ad, "jacobi", omega)
    else:
        dw = core.as_vec(np.asarray(dv, dtype=np.float64) * omega, Ad.dtype)
    bd = core.as_vec(np.asarray(b), Ad.dtype)
    xd = core.as_vec(np.asarray(x), Ad.dtype).clone()
    tmp = torch.empty_like(xd)
    for _ in range(nu):
        core.jacobi_sweep(Ad, dw, bd, xd, tmp)
        xd, tmp = tmp, xd
    x[...] = xd.cpu().numpy()
    return x


def jacobi_torch(A, b, x, Dinv=None, omega=0.666, nu=2):
    """Torch twin.  As in the reference (:49-50), Dinv=None uses the diagonal itself, not its inverse;
    callers pass 1/diag (:220)."""
    Ad = core.DeviceCSR.wrap(A)
    dev = x.device
    if Dinv is None:
        Dinv = 1.0 / core.smoother_diag(Ad, "jacobi", 1.0)      # = diag(A)
    dw = (omega * Dinv).to(device="cuda", dtype=Ad.dtype).contiguous()
    bd = b.to(device="cuda", dtype=Ad.dtype).contiguous()
    xd = x.to(device="cuda", dtype=Ad.dtype).contiguous().clone()
    tmp = torch.empty_like(xd)
    for _ in range(nu):
        core.jacobi_sweep(Ad, dw, bd, xd, tmp)
        xd, tmp = tmp, xd
    x.copy_(xd.to(dev))
    return x


def gauss_seidel(A, b, x, L=None, U=None, nu=2):
    """x <- L^-1 (b - U x), nu times (forward Gauss-Seidel); returns the new iterate (reference :58-90).
    L (lower triangle incl. the diagonal) and U (strict upper triangle) default to the halves of A, as in the
    reference; when the caller supplies them the sweep runs on L + U, which is the same update."""
    if L is not None or U is not None:
        A = sp.csr_matrix(A)
        Lm = sp.tril(A).tocsr() if L is None else sp.csr_matrix(L)
        Um = sp.triu(A, k=1).tocsr() if U is None else sp.csr_matrix(U)
        if sp.triu(Lm, k=1).nnz or sp.tril(Um, k=0).nnz:
            raise ValueError('gauss_seidel: L must be lower triangular and U strictly upper triangular')
        A = (Lm + Um).tocsr()
        A.sort_indices()
    Ad = core.DeviceCSR.wrap(A)
    sched = core.GaussSeidelSchedule(Ad)
    bd = core.as_vec(np.asarray(b), Ad.dtype)
    xd = core.as_vec(np.asarray(x), Ad.dtype).clone()
    sched.sweep(bd, xd, iterations=nu)
    return xd.cpu().numpy()


def gauss_seidel_torch(A, b, x, L=None, U=None, nu=2):
    """Torch twin (reference :93-99): x <- tril(A)^-1 (b - U x), nu times, A a dense (or sparse) torch matrix, U its
    strict upper triangle unless given.  Returns a new tensor on x's device; the sweeps run in the level-scheduled
    Gauss-Seidel kernel on tril(A) + U."""
    def host(M):
        M = M.detach().cpu()
        return M.to_dense().numpy() if M.layout != torch.strided else M.numpy()
    Ah = host(A)
    Lm = sp.csr_matrix(np.tril(Ah))
    Um = sp.csr_matrix(np.triu(Ah, 1)) if U is None else sp.csr_matrix(host(U))
    if sp.tril(Um, k=0).nnz:
        raise ValueError('gauss_seidel_torch: U must be strictly upper triangular')
    M = (Lm + Um).tocsr()
    M.sort_indices()
    Ad = core.DeviceCSR.from_scipy(M)
    sched = core.GaussSeidelSchedule(Ad)
    bd = b.detach().to(device="cuda", dtype=Ad.dtype).contiguous()
    xd = x.detach().to(device="cuda", dtype=Ad.dtype).contiguous().clone()
    sched.sweep(bd, xd, iterations=nu)
    return xd.to(device=x.device, dtype=x.dtype)


def smoothed_aggregation_jacobi(A, Agg, omega=None, lam_max=None):
    """P = (I - omega D^-1 A) Agg with omega = (4/3)/|lambda_max(D^-1 A)|; scipy CSR out."""
    Ad = core.DeviceCSR.wrap(A)
    if omega is None:
        if lam_max is None:
            lam_max = core.lambda_max(Ad)
        omega = (4.0 / 3.0) / lam_max
    Aggd = core.DeviceCSR.wrap(sp.csr_matrix(Agg).astype(np.float64 if Ad.dtype == torch.float64 else np.float32))
    return mlamg.sa_prolongator(Ad, Aggd, omega).to_scipy()


MAX_DENSE_COARSE = 8192      # coarse operators up to this size get an explicit dense inverse (2 x 0.5 GB of fp64 at the limit)


class _TwoLevel:
    """Device state of the two-level cycle shared by amg_2_v / amg_2_v_torch / the PC plugin.

    Coarse solve (the reference factorises A_H once with SuperLU, :168 / MLAMG.py:122):
      k <= max_dense : explicit dense inverse (cuSOLVER getrf/getrs once), one GEMV per iteration.  C1 (k = 6 554):
                       0.34 GB, 60 us per apply.
      k >  max_dense : the dense inverse would need 2 k^2 doubles (C2, k = 56 624: 51 GB) — A_H gets its own multilevel
                       hierarchy instead and every coarse solve is an inner V-cycle-preconditioned CG to a relative
                       residual of 1e-14 (an exact solve to rounding; a few MB).  A_H must be SPD for that.
    singular=True (lsqr min-norm solve, :179) needs the dense pseudo-inverse and is limited to k <= max_dense."""

    def __init__(self, A, P, singular=False, max_dense=None):
        max_dense = MAX_DENSE_COARSE if max_dense is None else max_dense
        self.A = core.DeviceCSR.wrap(A)
        self.P = core.DeviceCSR.wrap(P, self.A.dtype)
        self.R = core.transpose(self.P)
        self.AH = mlamg.galerkin(self.A, self.P, self.R)
        k = self.AH.shape[0]
        self.inner = None
        self.AHinv = None
        if singular:
            if k > max_dense:
                raise mlamg.MlamgError(3, f"singular two-level solve: coarse size {k} exceeds the dense pseudo-inverse limit "
                                          f"{max_dense}")
            # lsqr min-norm solve of the singular coarse problem (:179) -> dense pseudo-inverse
            dense = torch.zeros(self.AH.shape, dtype=torch.float64, device="cuda")
            core.check(core.lib.mlamg_csr_to_dense(1, k, core.ptr(self.AH.rowptr), core.ptr(self.AH.col),
                                                   core.ptr(self.AH.astype(torch.float64).val), core.ptr(dense),
                                                   core.stream()))
            self.AHinv = torch.linalg.pinv(dense).to(self.A.dtype).contiguous()
        elif k <= max_dense:
            self.AHinv = core.dense_inverse(self.AH)
        else:
            self.inner = mlamg.build_hierarchy(self.AH, aggregates="lloyd", ratio=0.1, distance="unit", rand=0,
                                               max_coarse=1000, max_levels=12)
        n = self.A.shape[0]
        self.r = torch.empty(n, dtype=self.A.dtype, device="cuda")
        self.rc = torch.empty(k, dtype=self.A.dtype, device="cuda")
        self.ec = torch.empty_like(self.rc)
        self._H = {}

    def coarse_solve(self, rc, ec):
        if self.inner is None:
            return core.gemv(self.AHinv, rc, out=ec)
        ec.zero_()
        _, hist = self.inner.solve_device(rc, ec, tol=1e-14 if self.A.dtype == torch.float64 else 1e-6, maxiter=200, accel="cg")
        if not hist[-1] <= 1e-12 * max(hist[0], 1e-300) and self.A.dtype == torch.float64:
            raise SingularCoarseError(4, f"inner coarse solve stalled at a relative residual of {hist[-1] / hist[0]:.2e} "
                                         "(is P^T A P symmetric positive definite?)")
        return ec

    def coarse_correct(self, b, x):
        core.residual(self.A, x, b, out=self.r)
        core.spmv(self.R, self.r, out=self.rc)
        self.coarse_solve(self.rc, self.ec)
        core.spmv_add(self.P, self.ec, x)

    def hierarchy(self, smoother, jacobi_weight):
        """the two levels as an mlamg Hierarchy (dense coarse level): its stationary solve runs the whole
        amg_2_v / MLAMG loop on the device"""
        key = (smoother, float(jacobi_weight))
        if key not in self._H:
            from mlamg.hierarchy import Level, Hierarchy
            L0, L1 = Level(self.A), Level(self.AH)
            L0.P, L0.R = self.P, self.R
            self._H[key] = Hierarchy([L0, L1], smoother=smoother, jacobi_weight=jacobi_weight, coarse_inv=self.AHinv)
        return self._H[key]


def amg_2_v(A, P, b, x,
            pre_smoothing_steps=1,
            post_smoothing_steps=1,
            jacobi_weight=0.666,
            res_tol=None,
            error_tol=None,
            max_iter=500,
            singular=False,
            smoother='gauss_seidel',
            _state=None):
    """Two-level AMG solver -> (x, conv_factor, err, num_iterations)  (reference :111-210).

    _state: optional dict owned by a caller that solves many times on the SAME A (the GA fitness loop, population x
    grids): the device copy of A and the Gauss-Seidel schedule are built on the first call and reused afterwards.

    Tolerances are absolute; err[i] = ||b - A x||_2 if res_tol is set else ||x||_2; a singular
    coarse operator returns (x, 1.0, err, 0) without raising, as the reference does.
    smoother='gauss_seidel' (the reference's pyamg sweep, bit-exact level-scheduled kernel) runs the loop from the
    host; 'jacobi' / 'l1_jacobi' run the whole loop on the device (one graph launch, mlamg_solve_ex)."""
    if res_tol is None and error_tol is None:
        raise RuntimeError('One of res_tol or error_tol must be set!')
    tol = res_tol if res_tol is not None else error_tol
    err = np.zeros(max_iter)
    if _state is not None:
        if 'A' not in _state:
            _state['A'] = core.DeviceCSR.wrap(A)
        A = _state['A']
    try:
        tl = _TwoLevel(A, P, singular=singular)
    except SingularCoarseError:
        return x, np.float64(1.), err, 0
    Ad = tl.A
    bd = core.as_vec(np.asarray(b), Ad.dtype)
    xd = core.as_vec(np.asarray(x), Ad.dtype).clone()
    if smoother not in ('gauss_seidel', 'jacobi', 'l1_jacobi'):
        raise ValueError(f'unknown smoother {smoother!r}')

    if smoother != 'gauss_seidel' and tl.inner is None:
        H = tl.hierarchy(smoother, jacobi_weight)
        flags = H.SOLVE_NO_INITIAL_CHECK | (H.SOLVE_XNORM if res_tol is None else 0) | (H.SOLVE_REMOVE_MEAN if singular else 0)
        _, hist = H.solve_abs(bd, xd, tol, max_iter, pre_smoothing_steps, post_smoothing_steps, flags)
        nit = len(hist) - 1
        err[:nit] = hist[1:]
        if nit < max_iter or (nit and hist[-1] <= tol):
            err = err[:nit]
    else:
        tmp = torch.empty_like(xd)
        if smoother == 'gauss_seidel':
            if _state is not None:
                if 'gs' not in _state:
                    _state['gs'] = core.GaussSeidelSchedule(Ad)
                sched = _state['gs']
            else:
                sched = core.GaussSeidelSchedule(Ad)
        else:
            dw = core.smoother_diag(Ad, smoother, jacobi_weight)

        def relax(xd, tmp, steps):
            if smoother == 'gauss_seidel':
                sched.sweep(bd, xd, iterations=steps)
                return xd, tmp
            for _ in range(steps):
                core.jacobi_sweep(Ad, dw, bd, xd, tmp)
                xd, tmp = tmp, xd
            return xd, tmp

        for i in range(max_iter):
            xd, tmp = relax(xd, tmp, pre_smoothing_steps)
            tl.coarse_correct(bd, xd)
            xd, tmp = relax(xd, tmp, post_smoothing_steps)
            if singular:
                xd -= xd.mean()
            if res_tol is not None:
                _, e = core.residual(Ad, xd, bd, out=tl.r, norm=True)
            else:
                e = float(np.sqrt(core.dot(xd, xd)))
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
    return xd.cpu().numpy().astype(np.asarray(x).dtype, copy=False), conv_factor, err, len(err)


def amg_2_v_torch(A, P, b, x,
                  pre_smoothing_steps=1,
                  post_smoothing_steps=1,
                  jacobi_weight=0.666,
                  error_tol=1e-10,
                  max_iter=20):
    """Torch twin with Jacobi smoothing and a dense coarse solve (reference :213-245);
    returns the convergence-factor estimate (err[i]/err[i-3])^(1/2) as a 0-dim tensor."""
    device = A.device
    tl = _TwoLevel(A, P)
    Ad = tl.A
    dw = core.smoother_diag(Ad, 'jacobi', jacobi_weight)
    bd = b.to(device="cuda", dtype=Ad.dtype).contiguous()
    xd = x.to(device="cuda", dtype=Ad.dtype).contiguous().clone()
    tmp = torch.empty_like(xd)
    err = torch.zeros(max_iter, dtype=Ad.dtype)
    i = 0
    for i in range(max_iter):
        for _ in range(pre_smoothing_steps):
            core.jacobi_sweep(Ad, dw, bd, xd, tmp)
            xd, tmp = tmp, xd
        tl.coarse_correct(bd, xd)
        for _ in range(post_smoothing_steps):
            core.jacobi_sweep(Ad, dw, bd, xd, tmp)
            xd, tmp = tmp, xd
        err[i] = float(np.sqrt(core.dot(xd, xd)))
        if err[i] < error_tol:
            break
    # the reference's jacobi_torch / `x +=` update the caller's tensor in place (:48-55, :234): keep that side effect
    x.copy_(xd.to(device=x.device, dtype=x.dtype))
    n_err = 3
    return ((err[i] / err[i - n_err]) ** (1 / (n_err - 1))).to(device)
