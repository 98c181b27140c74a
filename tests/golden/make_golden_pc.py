"""Generate tests/golden/ref_mlamg_pc.npz by running the UNMODIFIED reference class
/root/reference/ns/preconditioner/MLAMG.py::MLAMG — its `jacobi`, `amg_2_v` and `_apply` methods (:143-212), i.e. what
PETSc calls once per Krylov iteration.

Run:  python tests/golden/make_golden_pc.py        (needs /root/reference; CPU only)

The module imports firedrake / PETSc / matplotlib and two of the reference's own network modules at its top; none of them is
installed (or importable without torch_geometric) and none is touched by the three methods exercised, so they are shimmed by
empty modules (`PCBase` = object).  `_initialize` (firedrake assembly, a trained PNet checkpoint, greedy C/F coarsening) is
NOT run: the instance is created bare and given exactly the attributes `_initialize` would leave behind
(`A`, `Dinv = jacobi_weight / diag`, `P_amg`, `A_H_lu = splu(P^T A P, 'COLAMD')`, `amg_rtol`), with a Lloyd + smoothed
aggregation P in place of the PNet's.
"""
import os
import sys
import types

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import multilevel as oml           # noqa: E402
from oracle import reference_path as rp        # noqa: E402
import torch                                   # noqa: E402,F401  (MLAMG.py imports it; load it before the shims go in)


class _Vec:
    def __init__(self, a):
        self.array_r = a
        self.out = None

    def setArray(self, a):
        self.out = np.array(a, copy=True)


def install_shims():
    fd = types.ModuleType("firedrake")
    fd.PCBase = object
    fd.__all__ = ["PCBase"]
    petsc = types.ModuleType("firedrake.petsc")
    petsc.PETSc = types.SimpleNamespace()
    asm = types.ModuleType("firedrake.assemble")
    asm.allocate_matrix = asm.assemble = None
    ali = types.ModuleType("ns.model.ali_interp")
    ali.InterpolationNetwork = None
    greedy = types.ModuleType("ns.lib.greedy")
    greedy.greedy_coarsening = None
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.update({"firedrake": fd, "firedrake.petsc": petsc, "firedrake.assemble": asm, "matplotlib": mpl,
                        "matplotlib.pyplot": plt})
    for pkg in ("ns", "ns.model", "ns.lib"):
        sys.modules[pkg] = types.ModuleType(pkg)
    sys.modules["ns.model.ali_interp"] = ali
    sys.modules["ns.lib.greedy"] = greedy
    # the file itself, loaded by path: the package __init__ would pull in PCDR.py and PyAMG.py (pyamg, more firedrake)
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_MLAMG", os.path.join(REF, "ns", "preconditioner", "MLAMG.py"))
    rm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rm)
    return rm


def main():
    rm = install_shims()
    out = {}
    cases = {"poisson2d_20": (sp.csr_matrix(oml.poisson((20, 20))), 0.1, 2.0 / 3.0, 1e-8),
             "poisson3d_8": (sp.csr_matrix(oml.poisson((8, 8, 7))), 0.06, 0.6, 1e-6)}
    for name, (A, ratio, w, rtol) in cases.items():
        A = A.astype(np.float64)
        n = A.shape[0]
        Agg, _, _ = rp.lloyd_aggregation(A, ratio=ratio, distance="unit", rand=0)
        P = sp.csr_matrix(rp.smoothed_aggregation_jacobi(A, Agg, omega=2.0 / 3.0))
        pc = object.__new__(rm.MLAMG)
        pc.A = A
        pc.jacobi_weight = w
        pc.Dinv = sp.diags(1.0 / A.diagonal()) * w                    # MLAMG.py:104
        pc.P_amg = P
        pc.A_H = P.T @ A @ P
        pc.A_H_lu = spla.splu(pc.A_H, permc_spec="COLAMD")           # :121-122
        pc.amg_rtol = rtol
        b = np.random.RandomState(3).randn(n)
        X, Y = _Vec(b), _Vec(None)
        np.random.seed(0)
        pc.apply(None, X, Y)                                          # :199-212 (random guess from the global stream)
        xj = pc.jacobi(b, np.random.RandomState(4).randn(n), nu=3)    # :143-146
        P.sort_indices()
        out[f"{name}_A_indptr"], out[f"{name}_A_indices"], out[f"{name}_A_data"] = A.indptr, A.indices, A.data
        out[f"{name}_P_indptr"], out[f"{name}_P_indices"], out[f"{name}_P_data"] = P.indptr, P.indices, P.data
        out[f"{name}_P_shape"] = np.array(P.shape)
        out[f"{name}_b"], out[f"{name}_x"], out[f"{name}_jacobi_x"] = b, Y.out, xj
        out[f"{name}_jacobi_weight"], out[f"{name}_amg_rtol"] = w, rtol
        print(name, "n", n, "k", P.shape[1], "residual", np.linalg.norm(b - A @ Y.out))
    np.savez_compressed(os.path.join(HERE, "ref_mlamg_pc.npz"), **out)


if __name__ == "__main__":
    main()
