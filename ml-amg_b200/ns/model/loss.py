"""Forward value of the reference's two-grid loss (/root/reference/ns/model/loss.py:32-96) on the
multi-vector kernels: SpMM for A X / P^T (A X) / P e_H, hash SpGEMM for P^T A P, dense coarse solve
(`neumann_solve_fix`: the coarse operator bordered with the Lagrange row / column of :11-30).
Autograd through the cycle is out of scope (GA training needs no gradients)."""
import numpy as np
import torch

import mlamg
from mlamg import core
from mlamg._lib import lib, check, F64


def _lagrange_inverse(A_H):
    """dense inverse of [[A_H, 1], [1^T, 0]] (add_lagrange_rowcols, :11-27) in fp64"""
    k = A_H.shape[0]
    dense = torch.zeros(k, k, dtype=torch.float64, device="cuda")
    check(lib.mlamg_csr_to_dense(F64, k, core.ptr(A_H.rowptr), core.ptr(A_H.col), core.ptr(A_H.val), core.ptr(dense), core.stream()))
    aug = torch.zeros(k + 1, k + 1, dtype=torch.float64, device="cuda")
    aug[:k, :k] = dense
    aug[:k, k] = 1.0
    aug[k, :k] = 1.0
    work = torch.empty_like(aug)
    check(lib.mlamg_dense_inverse_f64(k + 1, core.ptr(aug), core.ptr(work), core.stream()))
    return aug


def amg_loss(P, A, test_vecs, tot_num_loop=5, no_prerelax=1, no_postrelax=1, device='cuda',
             neumann_solve_fix=False):
    omega = 2. / 3.
    Ad = core.DeviceCSR.wrap(A, torch.float32)
    Pd = core.DeviceCSR.wrap(P, torch.float32)
    Rd = core.transpose(Pd)
    Dinv_v = core.smoother_diag(Ad, 'jacobi', omega)                 # (1/D) * omega, fp32  (:49-50)
    A_H = mlamg.galerkin(Ad, Pd, Rd, drop=False).astype(torch.float64)   # .double()        (:53-54)
    AH_inv = _lagrange_inverse(A_H) if neumann_solve_fix else core.dense_inverse(A_H)      # fp64 coarse solve (:66-67, :79)
    N = Ad.shape[0]
    if not isinstance(test_vecs, torch.Tensor):
        np.random.seed(0)
        x = torch.tensor(np.random.normal(0, 1, (N, test_vecs))).float()
        x = x / torch.linalg.norm(x, 2, dim=0)
        x = x.to("cuda")
    else:
        x = test_vecs.to(device="cuda", dtype=torch.float32)
    x = x.contiguous()
    errs = torch.zeros((tot_num_loop + 1, x.shape[1]), device="cuda")
    for no_loop in range(tot_num_loop + 1):
        for _ in range(no_prerelax):
            x = x - Dinv_v[:, None] * core.spmm(Ad, x)
        r_H = core.spmm(Rd, core.spmm(Ad, x))
        if neumann_solve_fix:                                         # add_lagrange_vec (:29-30), then drop the multiplier (:81-82)
            r_H = torch.cat([r_H, torch.zeros(1, r_H.shape[1], dtype=r_H.dtype, device=r_H.device)], dim=0)
            e_H = (AH_inv @ (-r_H).double())[:-1].float().contiguous()
        else:
            e_H = (AH_inv @ (-r_H).double()).float().contiguous()
        x = x + core.spmm(Pd, e_H)
        for _ in range(no_postrelax):
            x = x - Dinv_v[:, None] * core.spmm(Ad, x)
        x = (x - x.mean(0)).contiguous()
        errs[no_loop] = torch.linalg.vector_norm(x, ord=2, dim=0)
    n_err = 3
    convs = (errs[-1] / errs[-n_err]) ** (1 / (n_err - 1))
    loss = torch.softmax(convs, dim=0) @ convs
    return loss
