"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules
(/root/reference/ns/lib/graph.py, ns/lib/multigrid.py) in the build container.

Run:  python tests/golden/make_golden.py            (needs /root/reference; CPU only)

The reference imports `pyamg` and `torch_sparse` at module top; neither is
installable here.  They are shimmed in `sys.modules`:
  * pyamg.graph.lloyd_cluster / bellman_ford and
    pyamg.relaxation.relaxation.gauss_seidel  -> oracle.pyamg_restated
    (restated third-party algorithm, PARITY UNPINNED for those three loops);
  * torch_sparse -> stub (never called by the functions exercised here).
Everything else that executes is the reference's own code: distance transform,
seeding, AggOp assembly (graph.py:156-239), modified_bellman_ford (graph.py:7-53),
nearest_center_to_agg (graph.py:56-86), jacobi (multigrid.py:15-45),
smoothed_aggregation_jacobi incl. ARPACK (multigrid.py:102-108) and the amg_2_v
driver incl. SuperLU coarse solve (multigrid.py:111-210).  `spla.eigs` is wrapped
by a recorder (not altered) so the omega the reference used is stored with P.
"""
import bz2
import os
import pickle
import sys
import types

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import pyamg_restated as pr          # noqa: E402
from oracle import multilevel as oml             # noqa: E402


def install_shims():
    pyamg = types.ModuleType("pyamg")
    graph = types.ModuleType("pyamg.graph")
    graph.lloyd_cluster = pr.lloyd_cluster
    graph.bellman_ford = pr.bellman_ford
    relaxation = types.ModuleType("pyamg.relaxation")
    relaxation2 = types.ModuleType("pyamg.relaxation.relaxation")
    relaxation2.gauss_seidel = pr.gauss_seidel
    relaxation2.jacobi = pr.jacobi
    relaxation.relaxation = relaxation2
    pyamg.graph = graph
    pyamg.relaxation = relaxation
    sys.modules.update({"pyamg": pyamg, "pyamg.graph": graph, "pyamg.relaxation": relaxation,
                        "pyamg.relaxation.relaxation": relaxation2})
    ts = types.ModuleType("torch_sparse")

    def _absent(*a, **k):
        raise NotImplementedError("torch_sparse is not installed; stub")
    ts.spspmm = ts.spmm = ts.transpose = _absent
    sys.modules["torch_sparse"] = ts


def csr_parts(prefix, M):
    M = sp.csr_matrix(M)
    return {f"{prefix}_indptr": M.indptr.astype(np.int32), f"{prefix}_indices": M.indices.astype(np.int32),
            f"{prefix}_data": M.data, f"{prefix}_shape": np.array(M.shape, dtype=np.int64)}


def main():
    install_shims()
    sys.path.insert(0, REF)
    import ns.lib.graph as rgraph
    import ns.lib.multigrid as rmg
    import scipy.sparse.linalg as spla
    import torch

    recorded = []
    real_eigs = spla.eigs

    def eigs_recorder(*a, **k):
        out = real_eigs(*a, **k)
        recorded.append(np.abs(out).item())
        return out
    rmg.spla.eigs = eigs_recorder

    cases = {}
    # ---- case inputs -------------------------------------------------------------
    A1 = oml.poisson((24, 24))
    A2 = oml.poisson((10, 10, 10))
    d = pickle.load(bz2.open(os.path.join(REF, "demos", "laplace_3d.grid"), "rb"))
    A3 = sp.csr_matrix(d["A"])
    A3.sort_indices()
    rs = np.random.RandomState(7)
    G4 = oml.poisson((20, 15)).astype(np.float64)
    G4.data = rs.rand(G4.nnz) + 0.05          # non-symmetric positive weights, random
    cases["poisson2d_24_unit"] = (A1, A1, dict(ratio=0.1, distance="unit", rand=0))
    cases["poisson3d_10_unit"] = (A2, A2, dict(ratio=0.027, distance="unit", rand=0))
    cases["laplace3d_grid_abs"] = (A3, A3, dict(ratio=0.1, distance="abs", rand=0))
    cases["laplace3d_grid_inv"] = (A3, A3, dict(ratio=0.05, distance="inv", rand=3))
    cases["randw_same"] = (oml.poisson((20, 15)), G4, dict(ratio=0.08, distance="same", rand=1))

    for name, (A, C, kw) in cases.items():
        out = {}
        out.update(csr_parts("A", A))
        out.update(csr_parts("C", C))
        out["ratio"] = kw["ratio"]
        out["distance"] = kw["distance"]
        out["rand"] = kw["rand"]
        Agg, roots, seeds = rgraph.lloyd_aggregation(C, **kw)
        out.update(csr_parts("Agg", Agg))
        out["roots"] = np.asarray(roots)
        out["seeds"] = np.asarray(seeds)
        recorded.clear()
        P = sp.csr_matrix(rmg.smoothed_aggregation_jacobi(A, Agg))
        out["lam_max"] = recorded[-1]
        out.update(csr_parts("P", P))
        AH = sp.csr_matrix(P.T @ A @ P)
        out.update(csr_parts("AH", AH))
        n = A.shape[0]
        x0 = np.random.RandomState(0).randn(n)
        x0 /= np.linalg.norm(x0, 2)
        b = np.zeros(n)
        x, conv, err, nit = rmg.amg_2_v(A, P, b, x0, res_tol=1e-10)
        out["gs_x"], out["gs_conv"], out["gs_err"], out["gs_nit"] = x, conv, err, nit
        b2 = np.random.RandomState(1).randn(n)
        x, conv, err, nit = rmg.amg_2_v(A, P, b2, np.zeros(n), res_tol=1e-8, pre_smoothing_steps=2,
                                        post_smoothing_steps=2)
        out["gs2_x"], out["gs2_conv"], out["gs2_err"], out["gs2_nit"] = x, conv, err, nit
        xj = rmg.jacobi(A, b2, x0.copy(), omega=0.666, nu=3)
        out["jacobi_x"] = xj
        np.savez_compressed(os.path.join(HERE, f"ref_{name}.npz"), **out)
        print(name, "N", n, "k", Agg.shape[1], "nnzP", P.nnz, "gs_nit", out["gs_nit"], "conv", float(out["gs_conv"]))

    # ---- reference-owned push Bellman-Ford + Agg assembly (pure torch/python, small) ----
    rs = np.random.RandomState(11)
    Ag = oml.poisson((9, 8)).tocoo()
    w = (rs.rand(Ag.nnz).astype(np.float32) + 0.01)
    w[rs.rand(Ag.nnz) < 0.1] = 0.0                     # ReLU outputs may be exactly 0
    S_T = torch.sparse_coo_tensor(np.vstack([Ag.row, Ag.col]), w, Ag.shape).coalesce()
    centers = torch.tensor(np.sort(rs.permutation(72)[:8]))
    dist, near = rgraph.modified_bellman_ford(S_T, centers)
    agg_T = rgraph.nearest_center_to_agg(centers, near)
    agg = agg_T.coalesce()
    np.savez_compressed(os.path.join(HERE, "ref_modified_bf.npz"),
                        row=S_T.indices()[0].numpy(), col=S_T.indices()[1].numpy(), w=S_T.values().numpy(),
                        n=72, centers=centers.numpy(), dist=dist.numpy(), nearest=near.numpy(),
                        agg_row=agg.indices()[0].numpy(), agg_col=agg.indices()[1].numpy(),
                        agg_val=agg.values().numpy())
    print("modified_bf ok", dist.numpy()[:5], near.numpy()[:5])


if __name__ == "__main__":
    main()
