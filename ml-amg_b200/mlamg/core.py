"""Device CSR container and thin wrappers over the C ABI (one wrapper per entry point family).

PyTorch is plumbing here: it owns device memory and the CUDA stream; all arithmetic happens in
libmlamg_b200.so.  No wrapper has a CPU path — a missing CUDA device raises RuntimeError.
"""
import ctypes

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib
from ._lib import lib, check

_TORCH2DT = {torch.float32: _lib.F32, torch.float64: _lib.F64}
_NP2TORCH = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("mlamg_b200 needs a CUDA device: the hot path has no CPU fallback")


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def dt(t):
    try:
        return _TORCH2DT[t.dtype if isinstance(t, torch.Tensor) else t]
    except KeyError:
        raise TypeError(f"mlamg supports float32/float64 values, got {t.dtype if isinstance(t, torch.Tensor) else t}")


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def as_vec(x, dtype, device="cuda"):
    """numpy / torch vector -> contiguous CUDA tensor of `dtype` (copy only when needed)."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.to(device=device, dtype=dtype).contiguous()


def as_i32(x, device="cuda"):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.to(device=device, dtype=torch.int32).contiguous()


class DeviceCSR:
    """CSR matrix resident in HBM: rowptr int32[n+1], col int32[nnz], val f32|f64[nnz]."""

    __slots__ = ("rowptr", "col", "val", "shape")

    def __init__(self, rowptr, col, val, shape):
        self.rowptr, self.col, self.val = rowptr, col, val
        self.shape = (int(shape[0]), int(shape[1]))

    # -- construction ---------------------------------------------------------------------
    @classmethod
    def from_scipy(cls, A, dtype=None, device="cuda"):
        require_cuda()
        A = sp.csr_matrix(A)
        if not A.has_canonical_format:
            A = A.copy()
            A.sum_duplicates()          # sorted, unique columns: the ordered SpGEMM relies on it
        if A.indptr[-1] >= 2 ** 31 or max(A.shape) >= 2 ** 31:
            raise OverflowError("matrix exceeds int32 indexing")
        if dtype is None:
            dtype = _NP2TORCH.get(A.dtype, torch.float64)
        return cls(as_i32(A.indptr, device), as_i32(A.indices, device),
                   torch.from_numpy(np.ascontiguousarray(A.data)).to(device=device, dtype=dtype).contiguous(), A.shape)

    @classmethod
    def from_arrays(cls, indptr, indices, data, shape, dtype=None, device="cuda"):
        """Raw CSR arrays, taken AS STORED (no sorting): the aggregation tie-breaking of the
        reference depends on the stored neighbour order."""
        require_cuda()
        data = np.ascontiguousarray(data)
        if dtype is None:
            dtype = _NP2TORCH.get(data.dtype, torch.float64)
        return cls(as_i32(np.asarray(indptr), device), as_i32(np.asarray(indices), device),
                   torch.from_numpy(data).to(device=device, dtype=dtype).contiguous(), shape)

    @classmethod
    def from_torch(cls, A, dtype=None, device="cuda"):
        """torch sparse COO/CSR (any device) -> DeviceCSR.  Indices are integer tensors (the
        reference's float-typed indices, ns/lib/sparse.py:26-30, are not reproduced).  The stored values keep
        their autograd link to A (a coalesced COO tensor that requires grad: mlamg.autograd, ns/model/loss.py)."""
        require_cuda()
        if A.layout == torch.sparse_coo:
            A = A.coalesce().to(device)
            idx = A.indices()
            n, m = A.shape
            counts = torch.bincount(idx[0], minlength=n)
            rowptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
            rowptr[1:] = torch.cumsum(counts, 0)
            val = A.values()
            return cls(rowptr.to(torch.int32), idx[1].to(torch.int32).contiguous(),
                       val.to(dtype or val.dtype).contiguous(), (n, m))
        if A.layout == torch.sparse_csr:
            A = A.to(device)
            val = A.values()
            return cls(A.crow_indices().to(torch.int32).contiguous(), A.col_indices().to(torch.int32).contiguous(),
                       val.to(dtype or val.dtype).contiguous(), A.shape)
        raise TypeError("expected a torch sparse COO or CSR tensor")

    @classmethod
    def wrap(cls, A, dtype=None):
        if isinstance(A, DeviceCSR):
            return A if dtype is None or A.val.dtype == dtype else A.astype(dtype)
        if isinstance(A, torch.Tensor):
            return cls.from_torch(A, dtype)
        return cls.from_scipy(A, dtype)

    # -- views ----------------------------------------------------------------------------
    @property
    def nnz(self):
        return int(self.col.numel())

    @property
    def dtype(self):
        return self.val.dtype

    def astype(self, dtype):
        return DeviceCSR(self.rowptr, self.col, self.val.to(dtype), self.shape)

    def with_values(self, val):
        return DeviceCSR(self.rowptr, self.col, val.contiguous(), self.shape)

    def to_scipy(self):
        return sp.csr_matrix((self.val.cpu().numpy(), self.col.cpu().numpy(), self.rowptr.cpu().numpy()),
                             shape=self.shape)

    def to_torch_coo(self):
        n = self.shape[0]
        rows = torch.repeat_interleave(torch.arange(n, device=self.col.device),
                                       (self.rowptr[1:] - self.rowptr[:-1]).long())
        return torch.sparse_coo_tensor(torch.stack([rows, self.col.long()]), self.val, self.shape).coalesce()

    def nbytes(self):
        return self.rowptr.numel() * 4 + self.col.numel() * 4 + self.val.numel() * self.val.element_size()

    # -- kernels --------------------------------------------------------------------------
    def matvec(self, x, out=None):
        return spmv(self, x, out)

    def __matmul__(self, other):
        if isinstance(other, DeviceCSR):
            return spgemm(self, other)
        return spmv(self, other)

    @property
    def T(self):
        return transpose(self)


class DeviceSELL:
    """SELL-32 copy of a DeviceCSR for the apply kernels (slice_ptr int32[ceil(n/32)+1], col/val padded)."""

    __slots__ = ("slice_ptr", "col", "val", "shape", "nnz")

    def __init__(self, A):
        n = A.shape[0]
        dev = A.val.device
        self.shape, self.nnz = A.shape, A.nnz
        self.slice_ptr = torch.empty((n + 31) // 32 + 1, dtype=torch.int32, device=dev)
        padded = ctypes.c_longlong(0)
        check(lib.mlamg_sell_slice_ptr(n, ptr(A.rowptr), ptr(self.slice_ptr), ctypes.byref(padded), stream()))
        self.col = torch.empty(max(padded.value, 1), dtype=torch.int32, device=dev)
        self.val = torch.empty(max(padded.value, 1), dtype=A.dtype, device=dev)
        check(lib.mlamg_sell_fill(dt(A.val), n, ptr(A.rowptr), ptr(A.col), ptr(A.val), ptr(self.slice_ptr), ptr(self.col),
                                  ptr(self.val), stream()))

    @property
    def padding(self):
        return self.col.numel() / max(self.nnz, 1)

    def rowop(self, op, x, b=None, dw=None, out=None, norm=False):
        n = self.shape[0]
        if out is None:
            out = torch.empty(n, dtype=self.val.dtype, device=x.device)
        nrm = torch.zeros(1, dtype=torch.float64, device=x.device) if norm else None
        check(lib.mlamg_sell_rowop(dt(self.val), op, n, ptr(self.slice_ptr), ptr(self.col), ptr(self.val), ptr(x), ptr(b),
                                   ptr(dw), ptr(out), ptr(nrm), stream()))
        return (out, float(nrm.sqrt().item())) if norm else out

    def spmv(self, x, out=None):
        return self.rowop(0, x, out=out)

    def residual(self, x, b, out=None, norm=False):
        return self.rowop(2, x, b=b, out=out, norm=norm)

    def jacobi_sweep(self, dw, b, x_in, x_out=None):
        return self.rowop(3, x_in, b=b, dw=dw, out=x_out)


def set_csr_lanes(lanes=-1):
    check(lib.mlamg_set_csr_lanes(int(lanes)))


def set_csr_batch(nb=0):
    check(lib.mlamg_set_csr_batch(int(nb)))


# ------------------------------------------------------------------ V-cycle apply kernels
def spmv(A, x, out=None):
    n, m = A.shape
    assert x.numel() == m and x.dtype == A.dtype
    if out is None:
        out = torch.empty(n, dtype=A.dtype, device=x.device)
    check(lib.mlamg_spmv_csr(dt(A.val), n, A.nnz, ptr(A.rowptr), ptr(A.col), ptr(A.val), ptr(x), ptr(out), stream()))
    return out


def spmv_perm(A, x, row_order, out=None):
    """y = A x, rows visited in `row_order` (int32 permutation) — same result, different cache behaviour."""
    n, m = A.shape
    assert x.numel() == m and row_order.numel() == n and row_order.dtype == torch.int32
    if out is None:
        out = torch.empty(n, dtype=A.dtype, device=x.device)
    check(lib.mlamg_spmv_csr_perm(dt(A.val), n, A.nnz, ptr(A.rowptr), ptr(A.col), ptr(A.val), ptr(x), ptr(out),
                                  ptr(row_order), stream()))
    return out


def rowop(A, op, x, y, b=None, dw=None, rows=None, row_range=None, aux=None):
    """generic row-op (0 y=Ax | 1 y+=Ax | 2 y=b-Ax | 3 y=x+dw.*(b-Ax) | 4 aux=dw.*b, y=b-A(dw.*b) | 5 y=aux+dw.*b+Ax |
    6 = op 4 with A holding the column-scaled values a_ij*dw_j | 7 y=dw.*(aux+b)+Ax) over all rows, the int32 list `rows`, or the contiguous range row_range=(begin, end)"""
    begin = 0
    if rows is not None:
        n = rows.numel()
    elif row_range is not None:
        begin, n = int(row_range[0]), int(row_range[1] - row_range[0])
    else:
        n = A.shape[0]
    if n <= 0:
        return y
    check(lib.mlamg_rowop_csr(dt(A.val), op, n, max(1, int(A.nnz * n / max(A.shape[0], 1))), ptr(A.rowptr), ptr(A.col),
                              ptr(A.val), ptr(x), ptr(b), ptr(dw), ptr(y), ptr(aux), ptr(rows), begin, None, stream()))
    return y


def prolong_smooth_zero(Q, e, rhs, r, dw, x_out=None):
    """x_out = dw .* (rhs + r) + Q e — prolong_smooth when x_in is the zero-guess sweep dw.*rhs (never materialised)."""
    n = Q.shape[0]
    if x_out is None:
        x_out = torch.empty_like(rhs)
    check(lib.mlamg_prolong_smooth_zero_csr(dt(Q.val), n, Q.nnz, ptr(Q.rowptr), ptr(Q.col), ptr(Q.val), ptr(e), ptr(rhs), ptr(r),
                                            ptr(dw), ptr(x_out), stream()))
    return x_out


def csr_to_w32(A):
    """W32 copies of A's col / val (slot-major inside every window of 32 rows; rowptr is shared).  -> (col_w32, val_w32)"""
    col = torch.empty_like(A.col)
    val = torch.empty_like(A.val)
    if A.nnz == 0:
        return col, val
    check(lib.mlamg_csr_to_w32(dt(A.val), A.shape[0], ptr(A.rowptr), ptr(A.col), ptr(A.val), ptr(col), ptr(val), stream()))
    return col, val


def rowop_w32(A, w32, op, x, y, b=None, dw=None, row_range=None, aux=None):
    """rowop() on the W32 copies of A (ops 0, 2, 5, 7) over all rows or the contiguous range row_range=(begin, end)"""
    n_total = A.shape[0]
    begin, end = (0, n_total) if row_range is None else (int(row_range[0]), int(row_range[1]))
    if end <= begin:
        return y
    check(lib.mlamg_rowop_w32(dt(A.val), op, end - begin, begin, n_total, ptr(A.rowptr), ptr(w32[0]), ptr(w32[1]), ptr(x), ptr(b),
                              ptr(dw), ptr(y), ptr(aux), stream()))
    return y


def residual_w32(A, w32, x, b, out=None):
    """r = b - A x on the W32 copies of A"""
    if out is None:
        out = torch.empty(A.shape[0], dtype=A.dtype, device=x.device)
    check(lib.mlamg_residual_w32(dt(A.val), A.shape[0], ptr(A.rowptr), ptr(w32[0]), ptr(w32[1]), ptr(x), ptr(b), ptr(out), stream()))
    return out


def prolong_smooth_zero_w32(Q, w32, e, rhs, r, dw, x_out=None):
    """prolong_smooth_zero on the W32 copies of Q (w32 = csr_to_w32(Q)): coalesced operator stream, bit-identical result"""
    if x_out is None:
        x_out = torch.empty_like(rhs)
    check(lib.mlamg_prolong_smooth_zero_w32(dt(Q.val), Q.shape[0], ptr(Q.rowptr), ptr(w32[0]), ptr(w32[1]), ptr(e), ptr(rhs), ptr(r),
                                            ptr(dw), ptr(x_out), stream()))
    return x_out


def prolong_smooth(Q, e, x_in, r, dw, x_out=None):
    """x_out = x_in + dw .* r + Q e (prolongation fused with the first post-smoothing sweep; x_out may be x_in)."""
    n = Q.shape[0]
    if x_out is None:
        x_out = torch.empty_like(x_in)
    check(lib.mlamg_prolong_smooth_csr(dt(Q.val), n, Q.nnz, ptr(Q.rowptr), ptr(Q.col), ptr(Q.val), ptr(e), ptr(x_in), ptr(r),
                                       ptr(dw), ptr(x_out), stream()))
    return x_out


def spmv_add(A, x, y):
    """y += A x (prolongation-and-correct)."""
    n, m = A.shape
    assert x.numel() == m and y.numel() == n
    check(lib.mlamg_spmv_add_csr(dt(A.val), n, A.nnz, ptr(A.rowptr), ptr(A.col), ptr(A.val), ptr(x), ptr(y), stream()))
    return y


def residual(A, x, b, out=None, norm=False):
    """r = b - A x; with norm=True also returns ||r||_2 (python float, one host sync)."""
    n = A.shape[0]
    if out is None:
        out = torch.empty(n, dtype=A.dtype, device=x.device)
    nrm = torch.zeros(1, dtype=torch.float64, device=x.device) if norm else None
    check(lib.mlamg_residual_csr(dt(A.val), n, A.nnz, ptr(A.rowptr), ptr(A.col), ptr(A.val), ptr(x), ptr(b), ptr(out),
                                 ptr(nrm), stream()))
    if norm:
        return out, float(nrm.sqrt().item())
    return out


def jacobi_sweep(A, dw, b, x_in, x_out=None):
    """x_out = x_in + dw .* (b - A x_in) — one fused pass over A."""
    n = A.shape[0]
    if x_out is None:
        x_out = torch.empty_like(x_in)
    check(lib.mlamg_jacobi_csr(dt(A.val), n, A.nnz, ptr(A.rowptr), ptr(A.col), ptr(A.val), ptr(dw), ptr(b), ptr(x_in),
                               ptr(x_out), stream()))
    return x_out


def jacobi_zero_residual(A, dw, b, x_out=None, r_out=None, norm=False):
    """x = dw .* b and r = b - A x in one pass over A (first zero-guess sweep fused with the residual)."""
    n = A.shape[0]
    if x_out is None:
        x_out = torch.empty_like(b)
    if r_out is None:
        r_out = torch.empty_like(b)
    nrm = torch.zeros(1, dtype=torch.float64, device=b.device) if norm else None
    check(lib.mlamg_jacobi_zero_residual_csr(dt(A.val), n, A.nnz, ptr(A.rowptr), ptr(A.col), ptr(A.val), ptr(dw), ptr(b),
                                             ptr(x_out), ptr(r_out), ptr(nrm), stream()))
    if norm:
        return x_out, r_out, float(nrm.sqrt().item())
    return x_out, r_out


def scaled_values(A, dw_cols):
    """values of A D_w on A's pattern: a_ij * dw_cols[j] (dw_cols indexed by COLUMN; pass the ext vector for the
    row-partitioned levels)"""
    return (A.val * dw_cols[A.col.long()]).contiguous()


def jacobi_zero_residual_scaled(A, val_scaled, dw, b, x_out=None, r_out=None):
    """jacobi_zero_residual on the column-scaled values (gathers b alone)"""
    n = A.shape[0]
    if x_out is None:
        x_out = torch.empty_like(b)
    if r_out is None:
        r_out = torch.empty_like(b)
    check(lib.mlamg_jacobi_zero_residual_scaled_csr(dt(A.val), n, A.nnz, ptr(A.rowptr), ptr(A.col), ptr(val_scaled), ptr(dw),
                                                    ptr(b), ptr(x_out), ptr(r_out), None, stream()))
    return x_out, r_out


def jacobi_zero(dw, b, out=None):
    if out is None:
        out = torch.empty_like(b)
    check(lib.mlamg_jacobi_zero(dt(b), b.numel(), ptr(dw), ptr(b), ptr(out), stream()))
    return out


def smoother_diag(A, mode="jacobi", omega=2.0 / 3.0):
    """dw = omega/diag(A) ('jacobi') or 1/rowsum|A| ('l1_jacobi')."""
    m = {"jacobi": 0, "l1_jacobi": 1}[mode]
    dw = torch.empty(A.shape[0], dtype=A.dtype, device=A.val.device)
    check(lib.mlamg_smoother_diag(dt(A.val), m, float(omega), A.shape[0], ptr(A.rowptr), ptr(A.col), ptr(A.val), ptr(dw),
                                  stream()))
    return dw


def spmm(A, X, alpha=1.0, beta=0.0, out=None):
    """Y = alpha A X + beta Y for a row-major N x k block (ns/model/loss.py multi-vector cycle)."""
    n, m = A.shape
    assert X.dim() == 2 and X.shape[0] == m and X.is_contiguous()
    k = X.shape[1]
    if out is None:
        out = torch.zeros(n, k, dtype=A.dtype, device=X.device)
    check(lib.mlamg_spmm_csr(dt(A.val), n, k, ptr(A.rowptr), ptr(A.col), ptr(A.val), ptr(X), ptr(out), float(alpha),
                             float(beta), stream()))
    return out


def sddmm(S, U, V):
    """out[j] = <U[row(j), :], V[col(j), :]> on the pattern of S: the gradient of S's stored values in S X / S^T X
    (backward pass of ns/model/loss.py's multi-vector cycle)."""
    n, m = S.shape
    assert U.dim() == 2 and V.dim() == 2 and U.shape == (n, V.shape[1]) and V.shape[0] == m
    assert U.is_contiguous() and V.is_contiguous() and U.dtype == V.dtype
    out = torch.empty(S.nnz, dtype=U.dtype, device=U.device)
    check(lib.mlamg_sddmm_csr(dt(U), n, U.shape[1], ptr(S.rowptr), ptr(S.col), ptr(U), ptr(V), ptr(out), stream()))
    return out


def sample_dense(S, Dm):
    """out[j] = Dm[row(j), col(j)] on the pattern of S (Dm: dense S.shape block, row-major)."""
    n, m = S.shape
    assert Dm.shape == (n, m) and Dm.is_contiguous()
    out = torch.empty(S.nnz, dtype=Dm.dtype, device=Dm.device)
    check(lib.mlamg_csr_sample_dense(dt(Dm), n, m, ptr(S.rowptr), ptr(S.col), ptr(Dm), ptr(out), stream()))
    return out


def agg_product_backward(A, labels, P, g_p):
    """gradient of P_hat's values (on A's pattern) in P = P_hat Agg, given the gradient of P's stored values"""
    assert g_p.is_contiguous() and g_p.numel() == P.nnz
    out = torch.empty(A.nnz, dtype=g_p.dtype, device=g_p.device)
    check(lib.mlamg_agg_product_backward(dt(g_p), A.shape[0], ptr(A.rowptr), ptr(A.col), ptr(labels), ptr(P.rowptr),
                                         ptr(P.col), ptr(g_p), ptr(out), stream()))
    return out


def csr_to_dense(A):
    """square DeviceCSR -> dense row-major block in A's dtype"""
    n = A.shape[0]
    assert A.shape[0] == A.shape[1]
    dense = torch.empty(n, n, dtype=A.dtype, device=A.val.device)
    check(lib.mlamg_csr_to_dense(dt(A.val), n, ptr(A.rowptr), ptr(A.col), ptr(A.val), ptr(dense), stream()))
    return dense


def dense_inverse_f64(dense):
    """inverse of a dense fp64 block (LU with partial pivoting, cuSOLVER); the argument is left untouched"""
    n = dense.shape[0]
    out = dense.detach().clone().contiguous()
    work = torch.empty_like(out)
    check(lib.mlamg_dense_inverse_f64(n, ptr(out), ptr(work), stream()))
    return out


def dot(x, y):
    res = torch.zeros(1, dtype=torch.float64, device=x.device)
    check(lib.mlamg_dot(dt(x), x.numel(), ptr(x), ptr(y), ptr(res), stream()))
    return float(res.item())


def axpby(alpha, x, beta, y):
    check(lib.mlamg_axpby(dt(x), x.numel(), float(alpha), ptr(x), float(beta), ptr(y), stream()))
    return y


def gemv(M, x, out=None):
    n = M.shape[0]
    if out is None:
        out = torch.empty(n, dtype=M.dtype, device=M.device)
    check(lib.mlamg_gemv(dt(M), n, ptr(M), ptr(x), ptr(out), stream()))
    return out


class GaussSeidelSchedule:
    """Dependency levels of the forward sweep (built once per operator)."""

    def __init__(self, A):
        n = A.shape[0]
        self.A = A
        self.level = torch.empty(n, dtype=torch.int32, device=A.val.device)
        self.order = torch.empty(n, dtype=torch.int32, device=A.val.device)
        self.level_ptr = (ctypes.c_int * (n + 1))()
        nlev = ctypes.c_int(0)
        check(lib.mlamg_gs_schedule(n, ptr(A.rowptr), ptr(A.col), ptr(self.level), ptr(self.order), self.level_ptr,
                                    ctypes.byref(nlev), stream()))
        self.nlevels = nlev.value

    def sweep(self, b, x, iterations=1):
        A = self.A
        for _ in range(iterations):
            check(lib.mlamg_gauss_seidel(dt(A.val), A.shape[0], ptr(A.rowptr), ptr(A.col), ptr(A.val), ptr(b), ptr(x),
                                         ptr(self.order), self.level_ptr, self.nlevels, stream()))
        return x


# ------------------------------------------------------------------ setup kernels
def scan_i32(counts):
    n = counts.numel()
    out = torch.empty(n + 1, dtype=torch.int32, device=counts.device)
    check(lib.mlamg_scan_i32(ptr(counts), ptr(out), n, stream()))
    return out


def agg_from_labels(labels, ncoarse, dtype=torch.float64):
    """labels int32[n] (-1 = unaggregated) -> Agg DeviceCSR n x ncoarse (ns/lib/graph.py:234-238)."""
    n = labels.numel()
    rowptr = torch.empty(n + 1, dtype=torch.int32, device=labels.device)
    col = torch.empty(max(n, 1), dtype=torch.int32, device=labels.device)
    val = torch.empty(max(n, 1), dtype=dtype, device=labels.device)
    nnz = ctypes.c_int(0)
    check(lib.mlamg_agg_from_labels(dt(val), n, ptr(labels), ptr(rowptr), ptr(col), ptr(val), ctypes.byref(nnz), stream()))
    return DeviceCSR(rowptr, col[:nnz.value], val[:nnz.value], (n, ncoarse))


def center_rank_labels(centers, nearest):
    """nearest centre node id -> column rank (ns/lib/graph.py:76-84); KeyError like the reference."""
    n = nearest.numel()
    scratch = torch.empty(n, dtype=torch.int32, device=nearest.device)
    labels = torch.empty(n, dtype=torch.int32, device=nearest.device)
    check(lib.mlamg_center_rank_labels(n, centers.numel(), ptr(centers), ptr(nearest), ptr(scratch), ptr(labels), stream()))
    return labels


def sa_smoother(A, omega):
    """S = I - omega D^-1 A, rows stored in scipy's order for `eye - omega*Dinv@A` (diagonal last)."""
    sval = torch.empty_like(A.val)
    scol = torch.empty_like(A.col)
    check(lib.mlamg_sa_smoother(dt(A.val), A.shape[0], ptr(A.rowptr), ptr(A.col), ptr(A.val), float(omega), ptr(scol),
                                ptr(sval), stream()))
    return DeviceCSR(A.rowptr, scol, sval, A.shape)


def spgemm(A, B):
    m, k = A.shape
    k2, n = B.shape
    assert k == k2, "dimension mismatch"
    if A.dtype != B.dtype:
        raise TypeError("spgemm operands must share a dtype")
    dev = A.val.device
    c_rowptr = torch.empty(m + 1, dtype=torch.int32, device=dev)
    nnz = ctypes.c_longlong(0)
    check(lib.mlamg_spgemm_symbolic(m, k, n, ptr(A.rowptr), ptr(A.col), ptr(B.rowptr), ptr(B.col), ptr(c_rowptr),
                                    ctypes.byref(nnz), stream()))
    c_col = torch.empty(max(nnz.value, 1), dtype=torch.int32, device=dev)
    c_val = torch.empty(max(nnz.value, 1), dtype=A.dtype, device=dev)
    check(lib.mlamg_spgemm_numeric(dt(A.val), m, k, n, ptr(A.rowptr), ptr(A.col), ptr(A.val), ptr(B.rowptr), ptr(B.col),
                                   ptr(B.val), ptr(c_rowptr), ptr(c_col), ptr(c_val), stream()))
    return DeviceCSR(c_rowptr, c_col[:nnz.value], c_val[:nnz.value], (m, n))


def transpose(A):
    m, n = A.shape
    dev = A.val.device
    t_rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    t_col = torch.empty(max(A.nnz, 1), dtype=torch.int32, device=dev)
    t_val = torch.empty(max(A.nnz, 1), dtype=A.dtype, device=dev)
    check(lib.mlamg_csr_transpose(dt(A.val), m, n, A.nnz, ptr(A.rowptr), ptr(A.col), ptr(A.val), ptr(t_rowptr),
                                  ptr(t_col), ptr(t_val), stream()))
    return DeviceCSR(t_rowptr, t_col[:A.nnz], t_val[:A.nnz], (n, m))


def drop_zeros(A):
    """Remove stored entries equal to 0.0 (scipy's SpGEMM never stores them)."""
    m = A.shape[0]
    dev = A.val.device
    new_rowptr = torch.empty(m + 1, dtype=torch.int32, device=dev)
    nnz = ctypes.c_longlong(0)
    check(lib.mlamg_csr_nonzero_count(dt(A.val), m, ptr(A.rowptr), ptr(A.val), ptr(new_rowptr), ctypes.byref(nnz), stream()))
    if nnz.value == A.nnz:
        return A
    new_col = torch.empty(max(nnz.value, 1), dtype=torch.int32, device=dev)
    new_val = torch.empty(max(nnz.value, 1), dtype=A.dtype, device=dev)
    check(lib.mlamg_csr_nonzero_fill(dt(A.val), m, ptr(A.rowptr), ptr(A.col), ptr(A.val), ptr(new_rowptr), ptr(new_col),
                                     ptr(new_val), stream()))
    return DeviceCSR(new_rowptr, new_col[:nnz.value], new_val[:nnz.value], A.shape)


def sort_rows(A):
    check(lib.mlamg_csr_sort_rows(dt(A.val), A.shape[0], ptr(A.rowptr), ptr(A.col), ptr(A.val), stream()))
    return A


def lambda_max(A, tol=1e-13, maxiter=6000, symmetric=None, info=None):
    """|lambda_max(D^-1 A)| on the device to a reported accuracy (the reference calls ARPACK, multigrid.py:105).
    Symmetric A: Lanczos on D^-1/2 A D^-1/2 until the eigenvalue error estimate is <= tol*lambda; otherwise a
    power iteration with a Rayleigh-residual stopping rule.  symmetric=None tests symmetry with two SpMVs.
    info (optional dict) receives {'residual', 'steps', 'method'}."""
    lam, res = ctypes.c_double(0.0), ctypes.c_double(0.0)
    steps, method = ctypes.c_int(0), ctypes.c_int(0)
    sym = -1 if symmetric is None else (1 if symmetric else 0)
    check(lib.mlamg_lambda_max(dt(A.val), A.shape[0], A.nnz, ptr(A.rowptr), ptr(A.col), ptr(A.val), float(tol), int(maxiter),
                               sym, ctypes.byref(lam), ctypes.byref(res), ctypes.byref(steps), ctypes.byref(method),
                               stream()))
    if info is not None:
        info.update(residual=res.value, steps=steps.value, method="lanczos" if method.value == 1 else "power")
    return lam.value


def dense_inverse(A):
    """Dense inverse (in A's dtype) of a small DeviceCSR operator; SingularCoarseError if singular."""
    n = A.shape[0]
    dev = A.val.device
    A64 = A if A.dtype == torch.float64 else A.astype(torch.float64)
    dense = torch.empty(n, n, dtype=torch.float64, device=dev)
    check(lib.mlamg_csr_to_dense(_lib.F64, n, ptr(A64.rowptr), ptr(A64.col), ptr(A64.val), ptr(dense), stream()))
    work = torch.empty(n, n, dtype=torch.float64, device=dev)
    check(lib.mlamg_dense_inverse_f64(n, ptr(dense), ptr(work), stream()))
    return dense if A.dtype == torch.float64 else dense.to(A.dtype)


def poisson(shape, dtype=torch.float64, device="cuda"):
    """Dirichlet 5/7-point Laplacian generated on device, x fastest."""
    require_cuda()
    shape = tuple(int(s) for s in shape) + (1,) * (3 - len(shape))
    nx, ny, nz = shape
    N = nx * ny * nz
    nnz = int(lib.mlamg_poisson_nnz(nx, ny, nz))
    rowptr = torch.empty(N + 1, dtype=torch.int32, device=device)
    col = torch.empty(nnz, dtype=torch.int32, device=device)
    val = torch.empty(nnz, dtype=dtype, device=device)
    check(lib.mlamg_poisson_csr(dt(val), nx, ny, nz, ptr(rowptr), ptr(col), ptr(val), stream()))
    return DeviceCSR(rowptr, col, val, (N, N))


# ------------------------------------------------------------------ aggregation
def bellman_ford(G, seeds):
    """pyamg.graph.bellman_ford twin: -> (distances[N] in G.dtype, nearest[N] int32 node ids / -1)."""
    n = G.shape[0]
    seeds = as_i32(seeds, G.val.device)
    dist = torch.empty(n, dtype=G.dtype, device=G.val.device)
    near = torch.empty(n, dtype=torch.int32, device=G.val.device)
    sweeps = ctypes.c_int(0)
    check(lib.mlamg_bellman_ford(dt(G.val), n, ptr(G.rowptr), ptr(G.col), ptr(G.val), seeds.numel(), ptr(seeds),
                                 ptr(dist), ptr(near), ctypes.byref(sweeps), stream()))
    return dist, near, sweeps.value


def lloyd_cluster(G, seeds, maxiter=10):
    """pyamg.graph.lloyd_cluster twin: -> (distances, clusters (seed index / -1), seeds, iterations)."""
    n = G.shape[0]
    seeds = as_i32(seeds, G.val.device).clone()
    dist = torch.empty(n, dtype=G.dtype, device=G.val.device)
    clusters = torch.empty(n, dtype=torch.int32, device=G.val.device)
    iters = ctypes.c_int(0)
    check(lib.mlamg_lloyd_cluster(dt(G.val), n, ptr(G.rowptr), ptr(G.col), ptr(G.val), seeds.numel(), ptr(seeds),
                                  int(maxiter), ptr(dist), ptr(clusters), ctypes.byref(iters), stream()))
    return dist, clusters, seeds, iters.value


def modified_bellman_ford(S, centers):
    """ns.lib.graph.modified_bellman_ford twin on a DeviceCSR (= coalesced COO) with float32 weights."""
    n = S.shape[0]
    if S.dtype != torch.float32:
        S = S.astype(torch.float32)
    centers = as_i32(centers, S.val.device)
    dist = torch.empty(n, dtype=torch.float32, device=S.val.device)
    near = torch.empty(n, dtype=torch.int64, device=S.val.device)
    passes = ctypes.c_int(0)
    check(lib.mlamg_modified_bellman_ford(n, ptr(S.rowptr), ptr(S.col), ptr(S.val), centers.numel(), ptr(centers),
                                          ptr(dist), ptr(near), ctypes.byref(passes), stream()))
    return dist, near, passes.value
