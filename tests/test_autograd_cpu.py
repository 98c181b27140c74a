"""CPU: the autograd wiring of the two-grid loss and of the learned prolongator (mlamg/autograd.py, ns/model/loss.py,
ns/model/agg_interp.py) held to the value AND gradient the unmodified reference produced
(tests/golden/ref_amg_loss_*.npz, written by tests/golden/make_golden_loss.py).

There is no GPU here and the product has no CPU path, so the C-ABI wrappers the autograd nodes call
(`core.spmm`, `core.sddmm`, `core.sample_dense`, `core.transpose`, ...) are replaced — in this test only — by
scipy / numpy stand-ins of the same contracts.  What is exercised is the product's own differentiation code: which
kernel is called with which operands, transposes, permutations, dtype casts and the torch glue between them.  The
kernels themselves are held to the same goldens on the GPU (tests/test_gpu_zz_autograd.py)."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from helpers import GOLDEN

CASES = sorted(os.path.basename(f)[len("ref_amg_loss_"):-4] for f in glob.glob(os.path.join(GOLDEN, "ref_amg_loss_*.npz")))


def _sp(S, vals=None):
    v = (S.val if vals is None else vals).detach().numpy()
    return sp.csr_matrix((v, S.col.numpy(), S.rowptr.numpy()), shape=S.shape)


def _dev(M, dtype):
    from mlamg.core import DeviceCSR
    M = sp.csr_matrix(M)
    M.sort_indices()
    return DeviceCSR(torch.from_numpy(M.indptr.astype(np.int32)), torch.from_numpy(M.indices.astype(np.int32)),
                     torch.from_numpy(M.data.copy()).to(dtype), M.shape)


@pytest.fixture
def cpu_kernels(monkeypatch):
    """scipy / numpy stand-ins of the C-ABI wrappers used by the autograd nodes (test-only)"""
    from mlamg import core, hierarchy

    def transpose(A):
        coo = _sp(A).tocoo()
        order = np.lexsort((coo.row, coo.col))
        T = sp.csr_matrix((np.ones(coo.nnz), (coo.col, coo.row)), shape=(A.shape[1], A.shape[0]))
        T.sort_indices()
        out = _dev(T, A.val.dtype)
        out.val = A.val.detach()[torch.from_numpy(order)].clone()
        return out

    def spmm(A, X, alpha=1.0, beta=0.0, out=None):
        assert X.is_contiguous() and X.dtype == A.val.dtype
        return torch.from_numpy(np.ascontiguousarray(alpha * (_sp(A) @ X.detach().numpy()))).to(X.dtype)

    def rows_of(S):
        return np.repeat(np.arange(S.shape[0]), np.diff(S.rowptr.numpy()))

    def sddmm(S, U, V):
        assert U.is_contiguous() and V.is_contiguous()
        r, c = rows_of(S), S.col.numpy()
        return torch.from_numpy(np.einsum("jk,jk->j", U.numpy()[r], V.numpy()[c])).to(U.dtype)

    def sample_dense(S, Dm):
        assert Dm.shape == S.shape and Dm.is_contiguous()
        return torch.from_numpy(Dm.numpy()[rows_of(S), S.col.numpy()].copy())

    def agg_product_backward(A, labels, P, g_p):
        G = sp.csr_matrix((g_p.numpy(), P.col.numpy(), P.rowptr.numpy()), shape=P.shape).toarray()
        r, lab = rows_of(A), labels.numpy()[A.col.numpy()]
        return torch.from_numpy(np.where(lab >= 0, G[r, np.maximum(lab, 0)], 0).astype(g_p.numpy().dtype))

    def smoother_diag(A, mode="jacobi", omega=2.0 / 3.0):
        assert mode == "jacobi"
        return torch.from_numpy((np.float32(omega) / _sp(A).diagonal()).astype(np.float32))

    def galerkin(A, P, R=None, drop=True, symmetric=False):
        Ps, As = _sp(P), _sp(A)
        if R is not None:
            assert abs(_sp(R) - Ps.T).max() == 0
        return _dev(Ps.T @ As @ Ps, P.val.dtype)

    def csr_to_dense(A):
        return torch.from_numpy(_sp(A).toarray())

    def dense_inverse_f64(dense):
        return torch.linalg.inv(dense.detach())

    def agg_from_labels(labels, ncoarse, dtype=torch.float64):
        lab = labels.numpy()
        keep = lab >= 0
        return _dev(sp.csr_matrix((np.ones(keep.sum()), (np.nonzero(keep)[0], lab[keep])), shape=(lab.size, ncoarse)), dtype)

    def learned_prolongator(P_hat, Agg, drop=False):
        # keep explicit zeros: structural product
        pat = sp.csr_matrix((np.ones(P_hat.nnz), P_hat.col.numpy(), P_hat.rowptr.numpy()), shape=P_hat.shape) @ \
            sp.csr_matrix((np.ones(Agg.nnz), Agg.col.numpy(), Agg.rowptr.numpy()), shape=Agg.shape)
        pat = sp.csr_matrix(pat)
        pat.sort_indices()
        val = (_sp(P_hat) @ _sp(Agg)).toarray()
        coo = pat.tocoo()
        pat.data = val[coo.row, coo.col].astype(P_hat.val.numpy().dtype)
        return _dev(pat, P_hat.val.dtype)

    for name, fn in dict(transpose=transpose, spmm=spmm, sddmm=sddmm, sample_dense=sample_dense,
                         agg_product_backward=agg_product_backward, smoother_diag=smoother_diag,
                         csr_to_dense=csr_to_dense, dense_inverse_f64=dense_inverse_f64,
                         agg_from_labels=agg_from_labels, require_cuda=lambda: None).items():
        monkeypatch.setattr(core, name, fn)
    monkeypatch.setattr(hierarchy, "galerkin", galerkin)
    import mlamg
    monkeypatch.setattr(mlamg, "learned_prolongator", learned_prolongator)
    monkeypatch.setattr(core, "as_vec", lambda x, dtype, device="cpu": (torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x).to(dtype).contiguous())
    monkeypatch.setattr(core, "as_i32", lambda x, device="cpu": (torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x).to(torch.int32).contiguous())
    return core


def load(name):
    z = np.load(os.path.join(GOLDEN, f"ref_amg_loss_{name}.npz"))
    n, k = int(z["n"]), int(z["k"])
    A = sp.csr_matrix((z["A_data"], z["A_indices"], z["A_indptr"]), shape=(n, n))
    P = sp.csr_matrix((z["P_val"], (z["P_row"], z["P_col"])), shape=(n, k))
    P.sort_indices()
    kw = {key[3:]: int(z[key]) for key in z.files if key.startswith("kw_")}
    return z, A, P, kw


def grad_tolerance(z):
    """what two fp32 evaluations of the same formula can be held to: well above the distance between the reference's own
    two summation orders, far below any wiring error (a wrong operand or transpose changes the gradient by O(1))"""
    g = z["grad"]
    return max(50 * np.abs(g - z["grad_alt"]).max(), 2e-4 * np.abs(g).max())


def test_goldens_exist():
    assert len(CASES) >= 4 and any("neumann" in c for c in CASES)


@pytest.mark.parametrize("name", CASES)
def test_amg_loss_value_and_gradient_vs_reference(cpu_kernels, name):
    import ns.model.loss as loss
    z, A, P, kw = load(name)
    assert np.array_equal(P.tocoo().row, z["P_row"]) and np.array_equal(P.indices, z["P_col"])   # CSR order = coalesced COO order
    Ad, Pd = _dev(A, torch.float32), _dev(P, torch.float32)
    Pd.val.requires_grad_(True)
    val = loss.amg_loss(Pd, Ad, torch.from_numpy(z["test_vecs"].copy()), neumann_solve_fix=bool(z["neumann"]), **kw)
    assert abs(float(val.detach()) - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    val.backward()
    g = Pd.val.grad.numpy()
    print(f"{name}: loss rel err {abs(float(val.detach()) - float(z['loss'])) / float(z['loss']):.1e}, gradient max err "
          f"{np.abs(g - z['grad']).max():.2e} (tolerance {grad_tolerance(z):.2e}, |grad|max {np.abs(z['grad']).max():.2e})")
    assert np.abs(g - z["grad"]).max() <= grad_tolerance(z), (np.abs(g - z["grad"]).max(), grad_tolerance(z))


def test_amg_loss_accepts_a_sparse_coo_P_that_requires_grad(cpu_kernels):
    """the call of demos/1d_poisson.py:91-95: P is a torch sparse tensor produced by the model"""
    import ns.model.loss as loss
    from mlamg.core import DeviceCSR
    z, A, P, kw = load("poisson2d_14")
    vals = torch.from_numpy(z["P_val"].copy()).requires_grad_(True)
    idx = torch.from_numpy(np.vstack([z["P_row"], z["P_col"]]).astype(np.int64))
    P_T = torch.sparse_coo_tensor(idx, vals, P.shape)
    Pd = DeviceCSR.from_torch(P_T, torch.float32, device="cpu")
    val = loss.amg_loss(Pd, _dev(A, torch.float32), torch.from_numpy(z["test_vecs"].copy()), **kw)
    val.backward()
    assert np.abs(vals.grad.numpy() - z["grad"]).max() <= grad_tolerance(z)


def test_gradient_reaches_the_edge_outputs_through_P_hat_times_Agg(cpu_kernels):
    """agg_interp.py:481-484 + loss: d loss / d P_hat = (d loss / d P)[i, agg(j)] on A's pattern"""
    import ns.model.loss as loss
    import ns.model.agg_interp as ai
    z, A, P, kw = load("poisson2d_14")
    n, k = P.shape
    rs = np.random.RandomState(5)
    labels = rs.randint(0, k, n).astype(np.int32)
    ph = torch.from_numpy((rs.rand(A.nnz).astype(np.float32) + 0.1)).requires_grad_(True)
    Ad = _dev(A, torch.float32)
    P_T, Pd = ai.learned_prolongator(Ad, ph, torch.from_numpy(labels), k)
    tv = torch.from_numpy(z["test_vecs"].copy())
    val = loss.amg_loss(Pd, Ad, tv, **kw)
    val.backward()
    # the same chain with torch's own dense autograd
    ph2 = ph.detach().clone().requires_grad_(True)
    r = torch.from_numpy(np.repeat(np.arange(n), np.diff(A.indptr)).astype(np.int64))
    c = torch.from_numpy(A.indices.astype(np.int64))
    Agg = torch.zeros(n, k).index_put((torch.arange(n), torch.from_numpy(labels).long()), torch.ones(n))
    Pdense = torch.zeros(n, n).index_put((r, c), ph2, accumulate=True) @ Agg
    Adense = torch.from_numpy(A.toarray().astype(np.float32))
    Dinv = (2.0 / 3.0) / torch.diagonal(Adense)
    AH_inv = torch.linalg.inv((Pdense.T @ Adense @ Pdense).double())
    x, errs = tv.clone(), []
    for _ in range(kw["tot_num_loop"] + 1):
        x = x - Dinv[:, None] * (Adense @ x)
        e = (AH_inv @ (-(Pdense.T @ (Adense @ x))).double()).float()
        x = x + Pdense @ e
        x = x - Dinv[:, None] * (Adense @ x)
        x = x - x.mean(0)
        errs.append(torch.linalg.vector_norm(x, dim=0))
    convs = (errs[-1] / errs[-3]) ** 0.5
    ref = torch.softmax(convs, 0) @ convs
    ref.backward()
    assert abs(float(val) - float(ref)) <= 1e-5 * abs(float(ref))
    scale = ph2.grad.abs().max().item()
    print(f"P_hat gradient: max err {(ph.grad - ph2.grad).abs().max().item():.2e} of scale {scale:.2e}")
    assert (ph.grad - ph2.grad).abs().max().item() <= 1e-3 * scale
    assert P_T.shape == (n, k)


def test_descent_on_the_edge_weights_lowers_the_loss(cpu_kernels):
    """demos/1d_poisson.py:84-101 without the PNet (torch_geometric is absent): Adam on the P_hat edge values of the
    9-node problem, starting from the smoothed-aggregation weights perturbed — the loss the optimiser sees must go down"""
    import ns.model.loss as loss
    import ns.model.agg_interp as ai
    z, A, P, kw = load("poisson1d_9")
    n, k = P.shape
    labels = torch.from_numpy(np.repeat(np.arange(k), n // k).astype(np.int32))
    S = sp.csr_matrix(sp.eye(n) - (2.0 / 3.0) * sp.diags(1.0 / A.diagonal()) @ A)        # I - omega D^-1 A on A's pattern
    S.sort_indices()
    assert np.array_equal(S.indices, A.indices)
    ph = torch.from_numpy((S.data * (1.0 + 0.5 * np.random.RandomState(0).randn(S.nnz))).astype(np.float32)).requires_grad_(True)
    Ad = _dev(A, torch.float32)
    tv = torch.from_numpy(z["test_vecs"].copy())
    opt = torch.optim.Adam([ph], lr=0.02)
    hist = []
    for _ in range(25):
        opt.zero_grad()
        _, Pd = ai.learned_prolongator(Ad, ph, labels, k)
        val = loss.amg_loss(Pd, Ad, tv, tot_num_loop=10)
        val.backward()
        opt.step()
        hist.append(float(val.detach()))
    assert hist[-1] < 0.8 * hist[0], hist


def test_amg_loss_refuses_sizes_beyond_its_dense_coarse_solve(cpu_kernels):
    import mlamg
    import ns.model.loss as loss
    n, k = 20000, loss.MAX_COARSE + 1
    P = sp.csr_matrix((np.ones(n, dtype=np.float32), (np.arange(n), np.arange(n) % k)), shape=(n, k))
    A = sp.eye(n, format="csr", dtype=np.float32)
    with pytest.raises(mlamg.MlamgError):
        loss.amg_loss(_dev(P, torch.float32), _dev(A, torch.float32), 2)
