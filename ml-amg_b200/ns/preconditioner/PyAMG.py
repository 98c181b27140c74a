"""PETSc python-PC shaped multilevel smoothed-aggregation preconditioner on the B200 kernels.

Mirrors /root/reference/ns/preconditioner/PyAMG.py: _createAmgSolver :79-96
(`pyamg.aggregation.smoothed_aggregation_solver(A, max_levels)`), _apply :118-120
(`Amg.solve(b, tol=amg_rtol, accel='gmres')`).  Options: `pyamg_amg_rtol` 1e-8,
`pyamg_amg_max_levels` 10, `pyamg_amg_precondition_with_gmres` (:52-54).  The Krylov acceleration is the
left-preconditioned restarted GMRES of `mlamg.Hierarchy.solve(accel='gmres')` (pyamg's own gmres is not
available to pin it against; the algorithm is the one restated in oracle/multilevel.py).
"""
import traceback

import numpy as np
import scipy.sparse as sp

import mlamg
from ._petsc_shim import PCBase, get_option


class PyAMG(PCBase):
    _prefix = "pyamg_"

    def initialize(self, pc):
        try:
            self._initialize(pc)
        except Exception as e:
            traceback.print_exc()
            raise e

    def _initialize(self, pc):
        if pc.getType() != "python":
            raise ValueError("Expecting PC type python")
        prefix = (pc.getOptionsPrefix() or '') + self._prefix
        self.amg_rtol = get_option(prefix, 'amg_rtol', 1e-8, float)
        self.amg_max_levels = get_option(prefix, 'amg_max_levels', 10, int)
        self.amg_precon_krylov = get_option(prefix, 'amg_precondition_with_gmres', True, bool)
        self.agg_ratio = get_option(prefix, 'agg_ratio', 0.1, float)
        self.update(pc)

    def _createAmgSolver(self, pc):
        _, Pmat = pc.getOperators()
        row, col, val = Pmat.getValuesCSR()
        self.Pcsr = sp.csr_matrix((val, col, row))
        self.Amg = mlamg.build_hierarchy(self.Pcsr, aggregates='lloyd', ratio=self.agg_ratio, distance='unit', rand=0,
                                         max_levels=self.amg_max_levels)

    def update(self, pc):
        self._createAmgSolver(pc)

    def apply(self, pc, X, Y):
        try:
            self._apply(pc, X, Y)
        except Exception as e:
            traceback.print_exc()
            raise e

    def _apply(self, pc, X, Y):
        y = self.Amg.solve(np.asarray(X.array_r), tol=self.amg_rtol, maxiter=200,
                           accel=('gmres' if self.amg_precon_krylov else None))
        Y.setArray(y)

    def applyTranspose(self, pc, X, Y):
        print('PyAMG applyTranspose!')

    def view(self, pc, viewer=None):
        if viewer is not None:
            viewer.printfASCII('PyAMG (B200) Solver:\n')
            viewer.printfASCII(f' amg solver: {str(self.Amg)}\n')
            viewer.printfASCII(f' amg rtol: {self.amg_rtol}\n')
