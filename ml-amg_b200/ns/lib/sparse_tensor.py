"""Mirror of the reference's `ns.lib.sparse_tensor` (/root/reference/ns/lib/sparse_tensor.py) without
torch_sparse: spspmm -> hash SpGEMM kernel, spmm -> SpMM kernel, spT -> transpose kernel."""
import scipy.sparse as sp
import torch

from mlamg import core


def spspmm(A, B):
    '''Sparse * Sparse mat-mat (:9-20) -> coalesced torch COO on A.device'''
    assert (A.shape[1] == B.shape[0])
    C = core.spgemm(core.DeviceCSR.from_torch(A), core.DeviceCSR.from_torch(B, A.dtype))
    return C.to_torch_coo().to(A.device)


def spmm(A, B):
    '''Sparse * Dense mat-mat (:22-29)'''
    assert (A.shape[1] == B.shape[0])
    Ad = core.DeviceCSR.from_torch(A)
    Bd = B.to(device="cuda", dtype=Ad.dtype)
    if Bd.dim() == 1:
        return core.spmv(Ad, Bd.contiguous()).to(B.device)
    return core.spmm(Ad, Bd.contiguous()).to(B.device)


def spT(A):
    '''Sparse transpose (:31-38)'''
    return core.transpose(core.DeviceCSR.from_torch(A)).to_torch_coo().to(A.device)


def diag(A):
    '''Diagonal of a sparse tensor (:40-52); entries without a stored diagonal stay 1 as in the reference'''
    A = A.coalesce()
    n = min(A.shape[0], A.shape[1])
    d = torch.ones(n, dtype=A.dtype, device=A.device)
    idx, val = A.indices(), A.values()
    on = idx[0] == idx[1]
    d[idx[0][on]] = val[on]
    return d


def to_scipy(T):
    '''torch COO -> scipy CSR (:54-59)'''
    T = T.coalesce()
    indices = T.indices().cpu().numpy()
    coo = sp.coo_matrix((T.values().cpu().numpy(), (indices[0], indices[1])), shape=tuple(T.shape))
    return coo.tocsr()
