"""Generate tests/golden/ref_two_level_modes.npz by running the UNMODIFIED reference module
/root/reference/ns/lib/multigrid.py in the modes tests/golden/make_golden.py does not cover:
  * amg_2_v(error_tol=...)                      the ||x|| measure (:190-193)
  * amg_2_v(singular=True, res_tol=...)         lsqr coarse solve + mean removal on a Neumann problem (:178-187)
  * amg_2_v without a tolerance                 RuntimeError text (:155-156)
  * amg_2_v with a singular coarse operator     (x, 1.0, err, 0) without raising (:166-170)
  * jacobi_torch, amg_2_v_torch                 the fp32 torch twins (:48-55, :213-245)
  * gauss_seidel                                the scipy triangular-solve form (:58-90)

Run:  python tests/golden/make_golden_modes.py        (needs /root/reference; CPU only)
Shims: as in make_golden.py (pyamg's Gauss-Seidel -> oracle.pyamg_restated, PARITY UNPINNED for that loop; torch_sparse stub).
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as mg0                      # noqa: E402  (shim installer)
from oracle import multilevel as oml           # noqa: E402
from oracle import reference_path as rp        # noqa: E402


def csr_parts(prefix, M):
    M = sp.csr_matrix(M)
    M.sort_indices()
    return {f"{prefix}_indptr": M.indptr.astype(np.int32), f"{prefix}_indices": M.indices.astype(np.int32),
            f"{prefix}_data": M.data, f"{prefix}_shape": np.array(M.shape, dtype=np.int64)}


def main():
    mg0.install_shims()
    sys.path.insert(0, REF)
    import ns.lib.multigrid as rmg
    import ns.lib.sparse as rsp
    import torch
    assert rmg.__file__.startswith(REF)
    out = {}
    # ---- Dirichlet problem: error_tol mode, scipy gauss_seidel, torch twins
    A = sp.csr_matrix(oml.poisson((14, 12))).astype(np.float64)
    n = A.shape[0]
    Agg, _, _ = rp.lloyd_aggregation(A, ratio=0.12, distance="unit", rand=0)
    P = sp.csr_matrix(rp.smoothed_aggregation_jacobi(A, Agg, omega=2.0 / 3.0))
    out.update(csr_parts("A", A))
    out.update(csr_parts("P", P))
    x0 = np.random.RandomState(0).randn(n)
    x0 /= np.linalg.norm(x0)
    b = np.zeros(n)
    x, conv, err, nit = rmg.amg_2_v(A, P, b, x0, error_tol=1e-9)
    out["errtol_x"], out["errtol_conv"], out["errtol_err"], out["errtol_nit"] = x, conv, err, nit
    try:
        rmg.amg_2_v(A, P, b, x0)
        out["no_tol_message"] = ""
    except RuntimeError as e:
        out["no_tol_message"] = str(e)
    b2 = np.random.RandomState(1).randn(n)
    out["gs_scipy_x"] = rmg.gauss_seidel(A, b2, x0.copy(), nu=3)
    # singular coarse operator: an aggregate whose column of P is zero
    Pz = P.copy().tolil()
    Pz[:, 0] = 0
    Pz = sp.csr_matrix(Pz)
    xs, convs, errs, nits = rmg.amg_2_v(A, Pz, b, x0, res_tol=1e-10)
    out["singcoarse_conv"], out["singcoarse_nit"], out["singcoarse_x_is_x0"] = convs, nits, np.array_equal(xs, x0)
    out.update(csr_parts("Pz", Pz))
    # torch twins (fp32)
    A_T, P_T = rsp.to_torch_sparse(A), rsp.to_torch_sparse(P)
    xt = torch.from_numpy(x0.astype(np.float32))
    bt = torch.from_numpy(b2.astype(np.float32))
    Dinv = 1.0 / rsp.get_diagonal(A_T)
    out["jacobi_torch_x"] = rmg.jacobi_torch(A_T, bt, xt.clone(), Dinv, omega=0.666, nu=3).numpy()
    out["jacobi_torch_b"] = bt.numpy()
    out["x0"] = x0
    out["b2"] = b2
    cf = rmg.amg_2_v_torch(A_T, P_T, torch.zeros(n), xt.clone(), jacobi_weight=2.0 / 3.0)
    out["amg_2_v_torch_conv"] = float(cf)
    cf2 = rmg.amg_2_v_torch(A_T, P_T, torch.zeros(n), xt.clone(), pre_smoothing_steps=2, post_smoothing_steps=2, jacobi_weight=0.5,
                            max_iter=12)
    out["amg_2_v_torch_conv_nu2"] = float(cf2)
    # ---- Neumann problem: singular=True
    G = sp.csr_matrix(oml.poisson((11, 10)))
    G = G - sp.diags(G.diagonal())
    L = (sp.diags(-np.asarray(G.sum(axis=1)).ravel()) + G).tocsr().astype(np.float64)
    AggN, _, _ = rp.lloyd_aggregation(L, ratio=0.15, distance="unit", rand=0)
    PN = sp.csr_matrix(rp.smoothed_aggregation_jacobi(L, AggN, omega=2.0 / 3.0))
    xn0 = np.random.RandomState(2).randn(L.shape[0])
    xn0 -= xn0.mean()
    xn0 /= np.linalg.norm(xn0)
    x, conv, err, nit = rmg.amg_2_v(L, PN, np.zeros(L.shape[0]), xn0, res_tol=1e-8, singular=True)
    out.update(csr_parts("L", L))
    out.update(csr_parts("PN", PN))
    out["xn0"] = xn0
    out["sing_x"], out["sing_conv"], out["sing_err"], out["sing_nit"] = x, conv, err, nit
    np.savez_compressed(os.path.join(HERE, "ref_two_level_modes.npz"), **out)
    print("errtol nit", nit if False else out["errtol_nit"], "conv", float(out["errtol_conv"]), "| singular nit", out["sing_nit"], "conv",
          float(out["sing_conv"]), "| torch conv", out["amg_2_v_torch_conv"], out["amg_2_v_torch_conv_nu2"], "| msg", out["no_tol_message"],
          "| singular coarse", float(out["singcoarse_conv"]), int(out["singcoarse_nit"]), bool(out["singcoarse_x_is_x0"]))


if __name__ == "__main__":
    main()
