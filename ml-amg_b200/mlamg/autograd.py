"""Differentiable forms of the multi-vector operators of the reference's two-grid loss
(/root/reference/ns/model/loss.py:32-96) and of the learned prolongator P = P_hat Agg
(/root/reference/ns/model/agg_interp.py:481-484).

The reference back-propagates through torch_sparse (`spmm`, `spspmm`, `transpose`) and torch_sparse_solve
(KLU) — demos/1d_poisson.py:91-95 trains the PNet that way.  Here every forward AND backward product is a
kernel of libmlamg_b200.so; torch only chains them (autograd.Function) and differentiates the elementwise
glue of the cycle.  With S a sparse operand on a fixed pattern, X a dense row-major block:

    Y = S X      dL/dS_ij = <dL/dY_i, X_j>   (SDDMM)      dL/dX = S^T dL/dY   (SpMM)
    Y = S^T X    dL/dS_ij = <X_i, dL/dY_j>   (SDDMM)      dL/dX = S dL/dY     (SpMM)
    A_H = P^T A P (dense k x k)    dL/dP = (A P G^T + A^T P G) sampled on P's pattern, G = dL/dA_H
    M = A_H^-1                     dL/dA_H = -M^T (dL/dM) M^T
    P = P_hat Agg                  dL/dP_hat_ij = dL/dP_{i, agg(j)}

A itself is a constant of the loss (the reference never asks for its gradient).
"""
import torch

from . import core
from . import hierarchy as _hier


class SparseOperand:
    """A fixed CSR pattern, its transposed pattern and the permutation carrying the stored values into the
    transposed order (obtained once by transposing the pattern with its entry numbers as values)."""

    def __init__(self, S):
        self.rowptr, self.col, self.shape = S.rowptr, S.col, S.shape
        nnz = S.nnz
        tags = torch.arange(nnz, dtype=torch.float64, device=S.col.device)
        T = core.transpose(core.DeviceCSR(S.rowptr, S.col, tags, S.shape))
        self.t_rowptr, self.t_col = T.rowptr, T.col
        self.perm = T.val.round().long()

    def pattern(self):
        return core.DeviceCSR(self.rowptr, self.col, None, self.shape)

    def csr(self, vals):
        return core.DeviceCSR(self.rowptr, self.col, vals.contiguous(), self.shape)

    def csr_t(self, vals):
        return core.DeviceCSR(self.t_rowptr, self.t_col, vals[self.perm].contiguous(), (self.shape[1], self.shape[0]))


class _SpMM(torch.autograd.Function):
    """Y = S X (transposed=False) or Y = S^T X (transposed=True) for the operand `op` with stored values `vals`."""

    @staticmethod
    def forward(ctx, vals, X, op, transposed):
        ctx.op, ctx.transposed = op, transposed
        ctx.save_for_backward(vals, X)
        M = op.csr_t(vals) if transposed else op.csr(vals)
        return core.spmm(M, X.contiguous())

    @staticmethod
    def backward(ctx, G):
        vals, X = ctx.saved_tensors
        op = ctx.op
        G = G.contiguous()
        g_vals = g_X = None
        if ctx.needs_input_grad[0]:
            X = X.contiguous()
            g_vals = core.sddmm(op.pattern(), X, G) if ctx.transposed else core.sddmm(op.pattern(), G, X)
        if ctx.needs_input_grad[1]:
            M = op.csr(vals) if ctx.transposed else op.csr_t(vals)
            g_X = core.spmm(M, G)
        return g_vals, g_X, None, None


def spmm(op, vals, X, transposed=False):
    return _SpMM.apply(vals, X, op, transposed)


class _GalerkinDense(torch.autograd.Function):
    """dense fp64 copy of A_H = P^T A P, the product itself evaluated by the hash SpGEMM in P's dtype
    (loss.py:53-54: spspmm in fp32, then `.double()`)."""

    @staticmethod
    def forward(ctx, pvals, Pop, A, At):
        ctx.Pop, ctx.A, ctx.At = Pop, A, At
        ctx.save_for_backward(pvals)
        A_H = _hier.galerkin(A, Pop.csr(pvals), Pop.csr_t(pvals), drop=False)
        return core.csr_to_dense(A_H.astype(torch.float64))

    @staticmethod
    def backward(ctx, G):
        (pvals,) = ctx.saved_tensors
        Pd = ctx.Pop.csr(pvals)
        Gf = G.to(pvals.dtype)
        T = core.spmm(ctx.A, core.spmm(Pd, Gf.t().contiguous()))
        T += core.spmm(ctx.At, core.spmm(Pd, Gf.contiguous()))
        return core.sample_dense(ctx.Pop.pattern(), T), None, None, None


def galerkin_dense(Pop, pvals, A, At):
    return _GalerkinDense.apply(pvals, Pop, A, At)


class _DenseInverse(torch.autograd.Function):
    """M = A_H^-1 in fp64 (LU once per loss evaluation: the coarse solve of loss.py:79 for every test vector and
    every iteration is then one dense product)."""

    @staticmethod
    def forward(ctx, dense):
        inv = core.dense_inverse_f64(dense)
        ctx.save_for_backward(inv)
        return inv

    @staticmethod
    def backward(ctx, G):
        (inv,) = ctx.saved_tensors
        it = inv.t()
        return -(it @ G @ it)


def dense_inverse(dense):
    return _DenseInverse.apply(dense)


class _AggProductValues(torch.autograd.Function):
    """stored values of P = P_hat Agg as a function of P_hat's values (the product itself is computed by the
    caller with the ordered SpGEMM; this node only carries the gradient back onto A's pattern)."""

    @staticmethod
    def forward(ctx, phat_vals, pvals, A, labels, P):
        ctx.A, ctx.labels, ctx.P = A, labels, P
        return pvals.clone()

    @staticmethod
    def backward(ctx, G):
        return core.agg_product_backward(ctx.A, ctx.labels, ctx.P, G.contiguous()), None, None, None, None


def agg_product_values(phat_vals, P, A, labels):
    """P.val with its autograd link to phat_vals (A: pattern carrier of P_hat, labels int32[n])"""
    return _AggProductValues.apply(phat_vals, P.val, A, labels, P)
