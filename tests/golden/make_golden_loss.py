"""Generate tests/golden/ref_amg_loss_*.npz by running the UNMODIFIED reference module
/root/reference/ns/model/loss.py (forward value AND the gradient `loss.backward()` leaves on the values of P)
in the build container.

Run:  python tests/golden/make_golden_loss.py        (needs /root/reference; CPU only)

`loss.py` imports two third-party extension packages that are not installable here; they are shimmed in
`sys.modules` by plain differentiable torch expressions of their documented contracts:
  * torch_sparse.spspmm(indexA, valueA, indexB, valueB, m, k, n) -> (index, value): the sparse product on its
    structural pattern, row-major sorted; torch_sparse.spmm(index, value, m, n, X) = scatter-add of
    value[:, None] * X[col] over the rows; torch_sparse.transpose(index, value, m, n) -> swapped + sorted;
  * torch_sparse_solve.solve(A[1,n,n] sparse, b[1,n,m]) -> x: exact solve (KLU there, torch.linalg.solve of the
    densified block here), differentiable in A's values and in b.
Everything else that executes is the reference's own code: `amg_loss` (:32-96), `add_lagrange_rowcols` / `add_lagrange_vec`
(:11-30) and the reference's ns/lib/sparse_tensor.py wrappers.  Each case is run twice, with two different fp32 summation orders
inside the shims (`grad`, `grad_alt`): their distance is the fp32 noise floor any re-ordered implementation can be
held to, and it is what the tolerance of the GPU test is derived from.
"""
import os
import sys
import types

import numpy as np
import scipy.sparse as sp
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

VARIANT = {"alt": False}


def _spmm(index, value, m, n, matrix):
    row, col = index[0].long(), index[1].long()
    if VARIANT["alt"]:                       # dense product: another summation order
        S = torch.zeros(m, n, dtype=value.dtype).index_put((row, col), value, accumulate=True)
        return S @ matrix
    out = torch.zeros(m, matrix.shape[1], dtype=matrix.dtype)
    return out.index_add(0, row, value[:, None] * matrix[col])


def _spspmm(indexA, valueA, indexB, valueB, m, k, n, coalesced=False):
    Ad = torch.zeros(m, k, dtype=valueA.dtype).index_put((indexA[0].long(), indexA[1].long()), valueA, accumulate=True)
    Bd = torch.zeros(k, n, dtype=valueB.dtype).index_put((indexB[0].long(), indexB[1].long()), valueB, accumulate=True)
    pat = sp.csr_matrix((np.ones(indexA.shape[1]), (indexA[0].numpy(), indexA[1].numpy())), shape=(m, k)) @ \
        sp.csr_matrix((np.ones(indexB.shape[1]), (indexB[0].numpy(), indexB[1].numpy())), shape=(k, n))
    pat = sp.csr_matrix(pat)
    pat.sort_indices()
    coo = pat.tocoo()
    r, c = torch.from_numpy(coo.row.astype(np.int64)), torch.from_numpy(coo.col.astype(np.int64))
    C = (Ad.double() @ Bd.double()).to(valueA.dtype) if VARIANT["alt"] else Ad @ Bd
    return torch.stack([r, c]), C[r, c]


def _transpose(index, value, m, n, coalesced=True):
    r, c = index[1].long(), index[0].long()
    order = torch.argsort(r * m + c, stable=True)
    return torch.stack([r[order], c[order]]), value[order]


def _solve(A, b):
    A = A.coalesce()
    dense = torch.zeros(tuple(A.shape), dtype=A.dtype).index_put(tuple(A.indices()), A.values(), accumulate=True)
    return torch.linalg.solve(dense, b)


_orig_unsqueeze = torch.Tensor.unsqueeze


def _unsqueeze_compat(self, dim):
    """torch >= 2 cannot back-propagate through `unsqueeze` of a sparse COO tensor (loss.py:79 relies on it: the
    backward pass ends in 'aten::as_strided ... SparseCPU').  Same result, built from the differentiable
    `values()`: a compatibility patch for the torch version of this container, not a change of the reference."""
    if self.layout == torch.sparse_coo and dim == 0:
        S = self.coalesce()
        idx = torch.cat([torch.zeros((1, S.indices().shape[1]), dtype=torch.long), S.indices()])
        return torch.sparse_coo_tensor(idx, S.values(), (1,) + tuple(S.shape))
    return _orig_unsqueeze(self, dim)


def install_shims():
    ts = types.ModuleType("torch_sparse")
    ts.spspmm, ts.spmm, ts.transpose = _spspmm, _spmm, _transpose
    tss = types.ModuleType("torch_sparse_solve")
    tss.solve = _solve
    sys.modules["torch_sparse"] = ts
    sys.modules["torch_sparse_solve"] = tss
    torch.Tensor.unsqueeze = _unsqueeze_compat


def coo_f32(M):
    M = sp.csr_matrix(M)
    M.sort_indices()
    c = M.tocoo()
    return torch.sparse_coo_tensor(np.vstack([c.row, c.col]).astype(np.int64), torch.from_numpy(c.data.astype(np.float32)),
                                   M.shape).coalesce()


def cases():
    from oracle import multilevel as oml
    from oracle import reference_path as rp
    out = {}
    # 1. the 9-node 1-D Poisson of demos/1d_poisson.py:32-60: 3 aggregates of 3 nodes, P = P_SA, 50 test vectors
    n_aggs, n_per = 3, 3
    n = n_aggs * n_per
    A = (sp.eye(n) * 2 - sp.eye(n, k=-1) - sp.eye(n, k=1)).tocsr()
    Agg = sp.csr_matrix(np.kron(np.eye(n_aggs), np.ones((n_per, 1))))
    P = (sp.eye(n) - (2.0 / 3.0) * sp.diags(1.0 / A.diagonal()) @ A) @ Agg
    tv = np.random.RandomState(0).normal(0, 1, (n, 50)).astype(np.float32)
    tv /= np.linalg.norm(tv, 2, axis=0)
    out["poisson1d_9"] = (A, sp.csr_matrix(P), tv, dict(tot_num_loop=20), False)
    # 2. its Neumann twin (demos/1d_poisson.py:37-42) with the Lagrange-bordered coarse solve
    An = A.tolil()
    An[0, 0], An[0, 1], An[-1, -1], An[-1, -2] = 1, -1, 1, -1
    An = An.tocsr()
    Pn = (sp.eye(n) - (2.0 / 3.0) * sp.diags(1.0 / An.diagonal()) @ An) @ Agg
    out["neumann1d_9"] = (An, sp.csr_matrix(Pn), tv, dict(tot_num_loop=20), True)
    # 3. 2-D Poisson 14 x 14, Lloyd aggregates, a perturbed smoothed-aggregation P, 8 test vectors, 5 loops
    A2 = oml.poisson((14, 14))
    Agg2, _, _ = rp.lloyd_aggregation(A2, ratio=0.1, distance="unit", rand=0)
    P2 = sp.csr_matrix((sp.eye(A2.shape[0]) - (2.0 / 3.0) * sp.diags(1.0 / A2.diagonal()) @ A2) @ Agg2)
    P2.sort_indices()
    P2.data = P2.data * (1.0 + 0.2 * np.random.RandomState(3).randn(P2.nnz))
    tv2 = np.random.RandomState(1).normal(0, 1, (A2.shape[0], 8)).astype(np.float32)
    tv2 /= np.linalg.norm(tv2, 2, axis=0)
    out["poisson2d_14"] = (A2, P2, tv2, dict(tot_num_loop=5), False)
    # 4. two pre / two post sweeps on a 3-D problem
    A3 = oml.poisson((6, 6, 5))
    Agg3, _, _ = rp.lloyd_aggregation(A3, ratio=0.08, distance="unit", rand=1)
    P3 = sp.csr_matrix((sp.eye(A3.shape[0]) - (2.0 / 3.0) * sp.diags(1.0 / A3.diagonal()) @ A3) @ Agg3)
    P3.sort_indices()
    P3.data = P3.data * (1.0 + 0.1 * np.random.RandomState(4).randn(P3.nnz))
    tv3 = np.random.RandomState(2).normal(0, 1, (A3.shape[0], 5)).astype(np.float32)
    tv3 /= np.linalg.norm(tv3, 2, axis=0)
    out["poisson3d_6x6x5_nu2"] = (A3, P3, tv3, dict(tot_num_loop=6, no_prerelax=2, no_postrelax=2), False)
    return out


def run(rloss, A, P, tv, kw, neumann):
    A_T = coo_f32(A)
    P_T = coo_f32(P)
    vals = P_T.values().clone().requires_grad_(True)
    P_leaf = torch.sparse_coo_tensor(P_T.indices(), vals, P_T.shape).coalesce()
    loss = rloss.amg_loss(P_leaf, A_T, torch.from_numpy(tv.copy()), device="cpu", neumann_solve_fix=neumann, **kw)
    loss.backward()
    return float(loss), vals.grad.numpy().copy(), P_T


def main():
    install_shims()
    sys.path.insert(0, REF)
    import ns.model.loss as rloss
    assert rloss.__file__.startswith(REF)
    for name, (A, P, tv, kw, neumann) in cases().items():
        try:
            VARIANT["alt"] = False
            loss, grad, P_T = run(rloss, A, P, tv, kw, neumann)
            VARIANT["alt"] = True
            loss_alt, grad_alt, _ = run(rloss, A, P, tv, kw, neumann)
        except Exception as e:                      # e.g. float-typed indices of add_lagrange_rowcols under a modern torch
            print(name, "NOT GENERATED:", type(e).__name__, e)
            continue
        A = sp.csr_matrix(A)
        A.sort_indices()
        idx = P_T.indices().numpy()
        np.savez_compressed(os.path.join(HERE, f"ref_amg_loss_{name}.npz"),
                            A_indptr=A.indptr.astype(np.int32), A_indices=A.indices.astype(np.int32),
                            A_data=A.data.astype(np.float32), n=A.shape[0], k=P_T.shape[1],
                            P_row=idx[0].astype(np.int32), P_col=idx[1].astype(np.int32), P_val=P_T.values().numpy(),
                            test_vecs=tv, neumann=neumann, loss=loss, grad=grad, loss_alt=loss_alt, grad_alt=grad_alt,
                            **{f"kw_{k}": v for k, v in kw.items()})
        print(name, "loss", loss, "alt", loss_alt, "|grad|max", np.abs(grad).max(),
              "noise", np.abs(grad - grad_alt).max())


if __name__ == "__main__":
    main()
