"""GPU: the mirrored two-level drivers in the modes of tests/golden/ref_two_level_modes.npz — written by the UNMODIFIED
reference (tests/golden/make_golden_modes.py): error_tol measure, singular coarse operator, singular=True, the scipy
Gauss-Seidel form and the fp32 torch twins.  Same bars as tests/test_gpu_hierarchy.py holds against the oracle."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, csr_from, hist_err0, rel_hist_err

pytestmark = pytest.mark.gpu


def test_two_level_driver_modes_vs_reference_golden():
    import ns.lib.multigrid as mg
    import ns.lib.sparse as nsp
    z = np.load(os.path.join(GOLDEN, "ref_two_level_modes.npz"))
    A, P = csr_from(z, "A"), csr_from(z, "P")
    n = A.shape[0]
    x0, b2 = z["x0"], z["b2"]
    x, conv, err, nit = mg.amg_2_v(A, P, np.zeros(n), x0, error_tol=1e-9)            # reference smoother: exact Gauss-Seidel
    assert nit == int(z["errtol_nit"]) and hist_err0(err, z["errtol_err"]) < 1e-12
    assert rel_hist_err(err[:10], z["errtol_err"][:10]) < 1e-11 and abs(conv - float(z["errtol_conv"])) < 1e-9
    assert np.abs(x - z["errtol_x"]).max() < 1e-12
    out = mg.amg_2_v(A, csr_from(z, "Pz"), np.zeros(n), x0, res_tol=1e-10)            # singular coarse operator (:166-170)
    assert out[1] == 1.0 and out[3] == 0 and np.array_equal(np.asarray(out[0]), x0)
    xg = mg.gauss_seidel(A, b2, x0.copy(), nu=3)
    assert np.abs(xg - z["gs_scipy_x"]).max() <= 1e-12 * np.abs(z["gs_scipy_x"]).max()
    A_T, P_T = nsp.to_torch_sparse(A), nsp.to_torch_sparse(P)
    xt = torch.from_numpy(x0.astype(np.float32))
    Dinv = 1.0 / nsp.get_diagonal(A_T)
    xj = mg.jacobi_torch(A_T, torch.from_numpy(z["jacobi_torch_b"]), xt.clone(), Dinv, omega=0.666, nu=3).numpy()
    assert np.abs(xj - z["jacobi_torch_x"]).max() <= 1e-5 * np.abs(z["jacobi_torch_x"]).max()
    cf = mg.amg_2_v_torch(A_T, P_T, torch.zeros(n), xt.clone(), jacobi_weight=2.0 / 3.0)
    assert abs(float(cf) - float(z["amg_2_v_torch_conv"])) <= 1e-3 * float(z["amg_2_v_torch_conv"])
    # singular=True: the exact pseudo-inverse here, lsqr at its default 1e-6 tolerances in the reference
    L, PN = csr_from(z, "L"), csr_from(z, "PN")
    got = mg.amg_2_v(L, PN, np.zeros(L.shape[0]), z["xn0"], res_tol=1e-8, singular=True)
    assert abs(got[3] - int(z["sing_nit"])) <= 1 and hist_err0(got[2][:8], z["sing_err"][:8]) < 1e-4


class _Mat:
    def __init__(self, A):
        self.A = A

    def getValuesCSR(self):
        return self.A.indptr, self.A.indices, self.A.data


class _PC:
    def __init__(self, A, P):
        self.m = _Mat(A)
        self.appctx = {"mlamg_P": P}

    def getType(self):
        return "python"

    def getOptionsPrefix(self):
        return ""

    def getOperators(self):
        return self.m, self.m


class _Vec:
    def __init__(self, a):
        self.array_r = a
        self.out = None

    def setArray(self, a):
        self.out = np.array(a, copy=True)


@pytest.mark.parametrize("name", ["poisson2d_20", "poisson3d_8"])
def test_mlamg_pc_apply_vs_the_reference_class(name):
    """ns.preconditioner.MLAMG.apply against what the UNMODIFIED reference class returned from `apply` on the same A, P, b and
    seeded random guess (tests/golden/make_golden_pc.py)"""
    import scipy.sparse as sp
    from ns.preconditioner import _petsc_shim
    from ns.preconditioner.MLAMG import MLAMG
    z = np.load(os.path.join(GOLDEN, "ref_mlamg_pc.npz"))
    A = sp.csr_matrix((z[f"{name}_A_data"], z[f"{name}_A_indices"], z[f"{name}_A_indptr"]))
    P = sp.csr_matrix((z[f"{name}_P_data"], z[f"{name}_P_indices"], z[f"{name}_P_indptr"]), shape=tuple(z[f"{name}_P_shape"]))
    old = dict(_petsc_shim.OPTIONS)
    _petsc_shim.OPTIONS.update({"mlamg_jacobi_weight": float(z[f"{name}_jacobi_weight"]), "mlamg_amg_rtol": float(z[f"{name}_amg_rtol"])})
    try:
        pc = _PC(A, P)
        p = MLAMG()
        p.initialize(pc)
        X, Y = _Vec(z[f"{name}_b"]), _Vec(None)
        np.random.seed(0)
        p.apply(pc, X, Y)
    finally:
        _petsc_shim.OPTIONS.clear()
        _petsc_shim.OPTIONS.update(old)
    ref = z[f"{name}_x"]
    assert np.linalg.norm(z[f"{name}_b"] - A @ Y.out) <= float(z[f"{name}_amg_rtol"]) * (1 + 1e-6)
    assert np.abs(Y.out - ref).max() <= 1e-11 * np.abs(ref).max(), np.abs(Y.out - ref).max()


class _G:
    def __init__(self, A):
        self.A = A


def test_evaluation_drivers_vs_the_reference_drivers():
    """utils/common.py evaluate_ref_conv / evaluate_dataset on the device against the convergence factors the UNMODIFIED reference
    drivers returned (tests/golden/make_golden_eval.py), ARPACK's recorded lambda_max injected.  The two measures built on the
    Arnoldi estimate of rho are compared on the unstructured grids only: on a structured grid every distance is tied with
    others to the last bit, and the last bits of rho (dot products summed in another order) may legitimately move them."""
    import importlib.util
    from helpers import ROOT, load_eval_golden
    spec = importlib.util.spec_from_file_location("mlamg_utils_common", os.path.join(ROOT, "ml-amg_b200", "utils", "common.py"))
    common = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(common)
    z, grids = load_eval_golden()
    data = [_G(A) for A in grids.values()]

    def lam_of(key):
        table = {A.shape[0]: float(v) for A, v in zip(grids.values(), z[key])}
        return lambda A: table[A.shape[0]]

    for measure in ("abs", "invabs", "unit"):
        got = common.evaluate_ref_conv(data, common.strength_measure_funcs[measure], alpha=0.2, lam_max=lam_of(f"ref_conv_{measure}_lam"))
        assert np.allclose(got, z[f"ref_conv_{measure}"], rtol=0, atol=1e-8), (measure, got, z[f"ref_conv_{measure}"])
    for measure in ("evolution", "olson"):
        got = common.evaluate_ref_conv(data[1:], common.strength_measure_funcs[measure], alpha=0.2, lam_max=lam_of(f"ref_conv_{measure}_lam"))
        assert np.allclose(got, z[f"ref_conv_{measure}"][1:], rtol=0, atol=1e-8), (measure, got, z[f"ref_conv_{measure}"][1:])
    got = common.evaluate_dataset(None, data[1:], alpha=0.2, lam_max=lam_of("dataset_conv_default_lam"))          # 'olson' default
    assert np.allclose(got, z["dataset_conv_default"][1:], rtol=0, atol=1e-8), (got, z["dataset_conv_default"][1:])
    got = common.evaluate_dataset(None, data, S=common.strength_measure_funcs["invabs"], alpha=0.3, omega=0.5,
                                  lam_max=lam_of("dataset_conv_invabs_lam"))
    assert np.allclose(got, z["dataset_conv_invabs"], rtol=0, atol=1e-8), (got, z["dataset_conv_invabs"])
