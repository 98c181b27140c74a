"""The reference's two-grid loss (/root/reference/ns/model/loss.py:32-96) on the multi-vector kernels, forward AND
backward: SpMM for A X / P^T (A X) / P e_H, hash SpGEMM for P^T A P, dense fp64 coarse solve (`neumann_solve_fix`: the
coarse operator bordered with the Lagrange row / column of :11-30), SDDMM / SpMM-transpose / pattern sampling for the
gradient of the stored values of P (`mlamg.autograd`).  `loss.backward()` works as in demos/1d_poisson.py:91-95 when P
is a torch sparse COO tensor that requires grad (or a DeviceCSR whose `val` does); A is a constant of the loss.

Held to the unmodified reference (value and gradient) by tests/golden/ref_amg_loss_*.npz."""
import numpy as np
import torch
import torch.nn.functional as nnF

from mlamg import core, MlamgError
from mlamg import autograd as ag

# The coarse solve is an explicit dense fp64 inverse (k^2 doubles, twice) and the gradient of P^T A P goes through a dense
# N x k block: fine for the training grids of the reference (N of a few thousand, utils/create_data.py), refused beyond these
MAX_COARSE = 8192
MAX_BLOCK_ELEMENTS = 1 << 31


def add_lagrange_rowcols(A, device='cpu'):
    """:11-27  sparse COO [[A, 1], [1^T, 0]] (integer index tensors; the reference's are float-typed)"""
    A = A.coalesce()
    n, m = A.size()
    dev = A.device
    extra = torch.cat([torch.stack([torch.arange(n, device=dev), torch.full((n,), m, device=dev)]),
                       torch.stack([torch.full((m,), n, device=dev), torch.arange(m, device=dev)])], dim=1)
    return torch.sparse_coo_tensor(torch.cat((A.indices(), extra), dim=1),
                                   torch.cat((A.values(), torch.ones(n + m, dtype=A.dtype, device=dev))),
                                   (n + 1, m + 1)).coalesce()


def add_lagrange_vec(x, device='cpu'):
    """:29-30  one zero row under the block of right-hand sides"""
    return torch.cat((x, torch.zeros(1, x.shape[1], dtype=x.dtype, device=x.device)), dim=0)


def _lagrange_border(A_H):
    """[[A_H, 1], [1^T, 0]] (add_lagrange_rowcols, :11-27), dense fp64"""
    k = A_H.shape[0]
    border = torch.zeros(k + 1, k + 1, dtype=A_H.dtype, device=A_H.device)
    border[:k, k] = 1.0
    border[k, :k] = 1.0
    return nnF.pad(A_H, (0, 1, 0, 1)) + border


def amg_loss(P, A, test_vecs, tot_num_loop=5, no_prerelax=1, no_postrelax=1, device='cuda',
             neumann_solve_fix=False):
    omega = 2. / 3.
    Ad = core.DeviceCSR.wrap(A, torch.float32)
    Ad = Ad.with_values(Ad.val.detach())
    Pd = core.DeviceCSR.wrap(P, torch.float32)
    pv = Pd.val                                                       # may carry the autograd link to the caller's P
    if Pd.shape[1] > MAX_COARSE or Pd.shape[0] * Pd.shape[1] >= MAX_BLOCK_ELEMENTS:
        raise MlamgError(3, f"amg_loss: {Pd.shape[0]} x {Pd.shape[1]} interpolation is beyond the dense coarse solve of the loss "
                            f"(coarse size <= {MAX_COARSE}, N * k < 2^31)")
    dev = pv.device
    Pop = ag.SparseOperand(Pd)
    Aop = ag.SparseOperand(Ad)
    At = Aop.csr_t(Ad.val)
    Dinv_v = core.smoother_diag(Ad, 'jacobi', omega)                 # (1/D) * omega, fp32  (:49-50)
    A_H = ag.galerkin_dense(Pop, pv, Ad, At)                         # spspmm, .double()    (:53-54)
    if neumann_solve_fix:
        A_H = _lagrange_border(A_H)                                  # (:66-67)
    AH_inv = ag.dense_inverse(A_H)                                   # fp64 coarse solve    (:79)
    N = Ad.shape[0]
    if not isinstance(test_vecs, torch.Tensor):
        np.random.seed(0)
        x = torch.tensor(np.random.normal(0, 1, (N, test_vecs))).float()
        x = x / torch.linalg.norm(x, 2, dim=0)
        x = x.to(dev)
    else:
        x = test_vecs.to(device=dev, dtype=torch.float32)
    x = x.contiguous()

    def relax(x):
        return x - Dinv_v[:, None] * ag.spmm(Aop, Ad.val, x)

    errs = []
    for no_loop in range(tot_num_loop + 1):
        for _ in range(no_prerelax):
            x = relax(x)
        r_H = ag.spmm(Pop, pv, ag.spmm(Aop, Ad.val, x), transposed=True)
        if neumann_solve_fix:                                         # add_lagrange_vec (:29-30), then drop the multiplier (:81-82)
            r_H = nnF.pad(r_H, (0, 0, 0, 1))
            e_H = (AH_inv @ (-r_H).double())[:-1].float()
        else:
            e_H = (AH_inv @ (-r_H).double()).float()
        x = x + ag.spmm(Pop, pv, e_H)
        for _ in range(no_postrelax):
            x = relax(x)
        x = x - x.mean(0)
        errs.append(torch.linalg.vector_norm(x, ord=2, dim=0))
    n_err = 3
    convs = (errs[-1] / errs[-n_err]) ** (1 / (n_err - 1))
    loss = torch.softmax(convs, dim=0) @ convs
    return loss
