"""Mirror of the reference's `ns.lib.sparse` format seam (/root/reference/ns/lib/sparse.py).
Indices are integer tensors: the reference builds them through `torch.Tensor(...)` (float32, :26-30),
which corrupts indices above 2^24 and is deliberately not reproduced."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

import ns.lib.sparse_tensor


def col_normalize_csr(A_sp, ord=1):
    if not sp.isspmatrix_csr(A_sp):
        A_sp = A_sp.tocsr()
    norms = spla.norm(A_sp, axis=0, ord=ord)
    return sp.csr_matrix((A_sp.data / norms[A_sp.indices], A_sp.indices, A_sp.indptr), A_sp.shape)


def to_torch_sparse(A):
    '''scipy sparse -> coalesced torch COO, float32 values (:20-32)'''
    A = A.tocoo()
    idx = torch.from_numpy(np.vstack([A.row, A.col]).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(A.data.astype(np.float32)), A.shape).coalesce()


def get_diagonal(A_T, as_vector=True):
    '''(:35-48)'''
    values = A_T.values()
    indices = A_T.indices()
    diag_entries = (indices[0] == indices[1])
    if as_vector:
        return values[diag_entries]
    return torch.sparse_coo_tensor(indices[:, diag_entries], values[diag_entries], size=A_T.shape)


def triu(A_T, diag=0):
    '''(:51-75)'''
    values, indices = A_T.values(), A_T.indices()
    mask = (indices[1] - indices[0]) >= diag
    return torch.sparse_coo_tensor(indices[:, mask], values[mask], size=A_T.shape)


def tril(A_T, diag=0):
    '''(:78-102)'''
    values, indices = A_T.values(), A_T.indices()
    mask = (indices[0] - indices[1]) >= diag
    return torch.sparse_coo_tensor(indices[:, mask], values[mask], size=A_T.shape)


scipy_to_torch = to_torch_sparse
torch_to_scipy = ns.lib.sparse_tensor.to_scipy
