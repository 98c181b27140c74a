"""PETSc python-PC shaped two-level AMG preconditioner on the B200 kernels.

Mirrors /root/reference/ns/preconditioner/MLAMG.py: initialize :30, update :126, jacobi :143-146,
amg_2_v :148-197, apply :199-212, applyTranspose :214, view :218.  Options keep the reference's
names and defaults (`mlamg_amg_rtol` 1e-8, `mlamg_jacobi_weight` 2/3, :61-67).  The reference builds
P from a greedy C/F splitting + a learned interpolation net (:105-120, classical AMG: out of scope,
SURVEY.md §2.1 rows 8,10); here P comes from the aggregation path (Lloyd + smoothed aggregation, or
`appctx['mlamg_P']` for a learned P), everything after P is the reference's algorithm.
"""
import traceback

import numpy as np
import scipy.sparse as sp
import torch

import mlamg
from mlamg import core
import ns.lib.graph
from ns.lib.multigrid import _TwoLevel
from ._petsc_shim import PCBase, get_option


class MLAMG(PCBase):
    _prefix = 'mlamg_'

    def initialize(self, pc):
        try:
            self._initialize(pc)
        except Exception as e:
            traceback.print_exc()
            raise e

    def _initialize(self, pc):
        if pc.getType() != 'python':
            raise ValueError('Expecting PC type python')
        prefix = (pc.getOptionsPrefix() or '') + self._prefix
        self.amg_rtol = get_option(prefix, 'amg_rtol', 1e-8, float)
        self.jacobi_weight = get_option(prefix, 'jacobi_weight', 2. / 3., float)
        self.agg_ratio = get_option(prefix, 'agg_ratio', 0.1, float)
        self.max_iter = get_option(prefix, 'max_iter', 500, int)
        self.update(pc)

    def _createAmgSolver(self, pc):
        _, Pmat = pc.getOperators()
        row, col, val = Pmat.getValuesCSR()
        self.A = sp.csr_matrix((val, col, row))
        appctx = getattr(pc, 'appctx', None) or {}
        P = appctx.get('mlamg_P')
        if P is None:
            Agg, _, _ = ns.lib.graph.lloyd_aggregation(self.A, ratio=self.agg_ratio, distance='unit', rand=0)
            Ad = core.DeviceCSR.wrap(self.A)
            lam = core.lambda_max(Ad)
            P = mlamg.sa_prolongator(Ad, core.DeviceCSR.wrap(Agg.astype(self.A.dtype)), (4. / 3.) / lam)
        self.two = _TwoLevel(self.A, P)                         # A_H = P^T A P and its factorisation (:121-122)
        self.P_amg = self.two.P
        self.Dinv = core.smoother_diag(self.two.A, 'jacobi', self.jacobi_weight)   # (:104)

    def update(self, pc):
        try:
            self._createAmgSolver(pc)
        except Exception as e:
            traceback.print_exc()
            raise e

    def jacobi(self, b, x, nu=2):
        tmp = torch.empty_like(x)
        for _ in range(nu):
            core.jacobi_sweep(self.two.A, self.Dinv, b, x, tmp)
            x, tmp = tmp, x
        return x

    def amg_2_v(self, P, b, x, pre_smoothing_steps=1, post_smoothing_steps=1, max_iter=500):
        """Reference :148-197 on device tensors; stops when ||b - A x||_2 <= amg_rtol (absolute).  With a dense coarse
        level the whole loop runs on the device (mlamg_solve_ex: one graph launch, one host sync per apply)."""
        if self.two.inner is None:
            H = self.two.hierarchy('jacobi', self.jacobi_weight)
            H.solve_abs(b, x, self.amg_rtol, max_iter, pre_smoothing_steps, post_smoothing_steps, H.SOLVE_NO_INITIAL_CHECK)
            return x
        for _ in range(max_iter):
            x = self.jacobi(b, x, nu=pre_smoothing_steps)
            self.two.coarse_correct(b, x)
            x = self.jacobi(b, x, nu=post_smoothing_steps)
            _, nrm = core.residual(self.two.A, x, b, out=self.two.r, norm=True)
            if nrm <= self.amg_rtol:
                break
        return x

    def apply(self, pc, X, Y):
        try:
            self._apply(pc, X, Y)
        except Exception as e:
            traceback.print_exc()
            raise e

    def _apply(self, pc, X, Y):
        b = core.as_vec(np.asarray(X.array_r), self.two.A.dtype)
        x = core.as_vec(np.random.normal(size=self.A.shape[1]), self.two.A.dtype)    # random guess (:209)
        x = self.amg_2_v(self.P_amg, b, x, max_iter=self.max_iter)
        Y.setArray(x.cpu().numpy())

    def applyTranspose(self, pc, X, Y):
        print('MLAMG applyTranspose!')

    def view(self, pc, viewer=None):
        if viewer is not None:
            viewer.printfASCII('MLAMG (B200) two-level solver:\n')
            viewer.printfASCII(f' coarse size: {self.two.AH.shape[0]}\n')
            viewer.printfASCII(f' amg rtol: {self.amg_rtol}\n')
