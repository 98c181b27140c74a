"""Forward value of the reference's two-grid loss (/root/reference/ns/model/loss.py:32-96) on the
multi-vector kernels: SpMM for A X / P^T (A X) / P e_H, hash SpGEMM for P^T A P, dense coarse solve.
Autograd through the cycle is out of scope (GA training needs no gradients)."""
import numpy as np
import torch

import mlamg
from mlamg import core


def amg_loss(P, A, test_vecs, tot_num_loop=5, no_prerelax=1, no_postrelax=1, device='cuda',
             neumann_solve_fix=False):
    if neumann_solve_fix:
        raise NotImplementedError("neumann_solve_fix (Lagrange-augmented coarse solve) is not on the built path")
    omega = 2. / 3.
    Ad = core.DeviceCSR.wrap(A, torch.float32)
    Pd = core.DeviceCSR.wrap(P, torch.float32)
    Rd = core.transpose(Pd)
    Dinv_v = core.smoother_diag(Ad, 'jacobi', omega)                 # (1/D) * omega, fp32  (:49-50)
    A_H = mlamg.galerkin(Ad, Pd, Rd, drop=False).astype(torch.float64)   # .double()        (:53-54)
    AH_inv = core.dense_inverse(A_H)                                  # fp64 coarse solve    (:79)
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
