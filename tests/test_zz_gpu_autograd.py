"""GPU: the differentiable two-grid loss (ns/model/loss.py) and learned prolongator (ns/model/agg_interp.py) through the
C ABI — value and gradient against what the UNMODIFIED reference produced (tests/golden/ref_amg_loss_*.npz, written by
tests/golden/make_golden_loss.py from /root/reference/ns/model/loss.py), plus the three backward kernels one by one.
Tolerances: fp32 arithmetic on both sides in different summation orders — `grad_tolerance` below."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from helpers import GOLDEN

pytestmark = pytest.mark.gpu

CASES = sorted(os.path.basename(f)[len("ref_amg_loss_"):-4] for f in glob.glob(os.path.join(GOLDEN, "ref_amg_loss_*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN, f"ref_amg_loss_{name}.npz"))
    n, k = int(z["n"]), int(z["k"])
    A = sp.csr_matrix((z["A_data"], z["A_indices"], z["A_indptr"]), shape=(n, n))
    P = sp.csr_matrix((z["P_val"], (z["P_row"], z["P_col"])), shape=(n, k))
    P.sort_indices()
    kw = {key[3:]: int(z[key]) for key in z.files if key.startswith("kw_")}
    return z, A, P, kw


def grad_tolerance(z):
    g = z["grad"]
    return max(50 * np.abs(g - z["grad_alt"]).max(), 2e-4 * np.abs(g).max())


def test_backward_kernels_vs_numpy():
    import mlamg
    from mlamg import core
    rs = np.random.RandomState(0)
    for dtype, tol in ((torch.float32, 1e-5), (torch.float64, 1e-13)):
        npdt = np.float32 if dtype == torch.float32 else np.float64
        S = sp.random(57, 23, density=0.2, random_state=rs, format="csr", dtype=np.float64)
        S.sort_indices()
        Sd = mlamg.DeviceCSR.from_scipy(S, dtype)
        rows = np.repeat(np.arange(57), np.diff(S.indptr))
        for k in (1, 5, 32, 70):
            U, V = rs.randn(57, k).astype(npdt), rs.randn(23, k).astype(npdt)
            out = core.sddmm(Sd, torch.from_numpy(U).cuda(), torch.from_numpy(V).cuda()).cpu().numpy()
            ref = np.einsum("jk,jk->j", U[rows].astype(np.float64), V[S.indices].astype(np.float64))
            assert np.abs(out - ref).max() <= tol * max(1.0, np.abs(ref).max()), (dtype, k)
        Dm = rs.randn(57, 23).astype(npdt)
        out = core.sample_dense(Sd, torch.from_numpy(Dm).cuda()).cpu().numpy()
        assert np.array_equal(out, Dm[rows, S.indices])
        # P = P_hat Agg: the gradient of an entry of P_hat is the gradient of the entry of P it was summed into
        A = sp.random(40, 40, density=0.15, random_state=rs, format="csr") + sp.eye(40)
        A = sp.csr_matrix(A)
        A.sort_indices()
        labels = rs.randint(-1, 6, 40).astype(np.int32)
        labels[:6] = np.arange(6)
        Ad = mlamg.DeviceCSR.from_scipy(A, dtype)
        lab_d = torch.from_numpy(labels).cuda()
        Agg = core.agg_from_labels(lab_d, 6, dtype)
        Pd = mlamg.learned_prolongator(Ad, Agg)
        gp = rs.randn(Pd.nnz).astype(npdt)
        out = core.agg_product_backward(Ad, lab_d, Pd, torch.from_numpy(gp).cuda()).cpu().numpy()
        G = sp.csr_matrix((gp, Pd.col.cpu().numpy(), Pd.rowptr.cpu().numpy()), shape=Pd.shape).toarray()
        ar = np.repeat(np.arange(40), np.diff(A.indptr))
        lab = labels[A.indices]
        assert np.array_equal(out, np.where(lab >= 0, G[ar, np.maximum(lab, 0)], 0).astype(npdt))
    # empty operands
    E = mlamg.DeviceCSR.from_scipy(sp.csr_matrix((3, 4)), torch.float32)
    assert core.sddmm(E, torch.zeros(3, 2, device="cuda"), torch.zeros(4, 2, device="cuda")).numel() == 0


@pytest.mark.parametrize("name", CASES)
def test_amg_loss_value_and_gradient_vs_reference(name):
    import mlamg
    import ns.model.loss as loss
    z, A, P, kw = load(name)
    before = mlamg.launch_count()
    vals = torch.from_numpy(z["P_val"].copy()).cuda().requires_grad_(True)
    idx = torch.from_numpy(np.vstack([z["P_row"], z["P_col"]]).astype(np.int64)).cuda()
    P_T = torch.sparse_coo_tensor(idx, vals, P.shape)                  # the reference's callers hand over a sparse COO tensor
    val = loss.amg_loss(P_T, mlamg.DeviceCSR.from_scipy(A, torch.float32), torch.from_numpy(z["test_vecs"].copy()),
                        neumann_solve_fix=bool(z["neumann"]), **kw)
    assert abs(float(val.detach()) - float(z["loss"])) <= 2e-5 * abs(float(z["loss"])), (float(val), float(z["loss"]))
    val.backward()
    g = vals.grad.cpu().numpy()
    err = np.abs(g - z["grad"]).max()
    print(f"{name}: loss {float(val.detach()):.8f} vs {float(z['loss']):.8f}, gradient max err {err:.2e} "
          f"(tolerance {grad_tolerance(z):.2e})")
    assert err <= grad_tolerance(z), (err, grad_tolerance(z))
    assert mlamg.launch_count() - before > 8 * (kw["tot_num_loop"] + 1)      # forward and backward ran on the library's kernels


def test_gradient_through_P_hat_times_Agg_and_DeviceCSR_input():
    """agg_interp.py:481-484 followed by the loss: d loss / d P_hat against torch's own dense autograd on the device"""
    import ns.model.loss as loss
    import ns.model.agg_interp as ai
    import mlamg
    z, A, P, kw = load("poisson2d_14")
    n, k = P.shape
    rs = np.random.RandomState(5)
    labels = rs.randint(0, k, n).astype(np.int32)
    ph = torch.from_numpy((rs.rand(A.nnz).astype(np.float32) + 0.1)).cuda().requires_grad_(True)
    Ad = mlamg.DeviceCSR.from_scipy(A, torch.float32)
    tv = torch.from_numpy(z["test_vecs"].copy()).cuda()
    P_T, Pd = ai.learned_prolongator(Ad, ph, torch.from_numpy(labels).cuda(), k)
    val = loss.amg_loss(Pd, Ad, tv, **kw)                              # DeviceCSR whose values carry the autograd link
    val.backward()
    g_direct = ph.grad.clone()
    ph.grad = None
    val2 = loss.amg_loss(P_T, Ad, tv, **kw)                            # the same through the sparse COO tensor
    val2.backward()
    assert (g_direct - ph.grad).abs().max().item() <= 1e-6 * g_direct.abs().max().item()
    assert abs(float(val.detach()) - float(val2.detach())) <= 1e-6 * abs(float(val.detach()))
    ph2 = ph.detach().clone().requires_grad_(True)
    r = torch.from_numpy(np.repeat(np.arange(n), np.diff(A.indptr)).astype(np.int64)).cuda()
    c = torch.from_numpy(A.indices.astype(np.int64)).cuda()
    Agg = torch.zeros(n, k, device="cuda").index_put((torch.arange(n, device="cuda"), torch.from_numpy(labels).long().cuda()),
                                                     torch.ones(n, device="cuda"))
    Pdense = torch.zeros(n, n, device="cuda").index_put((r, c), ph2, accumulate=True) @ Agg
    Adense = torch.from_numpy(A.toarray().astype(np.float32)).cuda()
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
    assert abs(float(val.detach()) - float(ref.detach())) <= 2e-5 * abs(float(ref.detach()))
    scale = ph2.grad.abs().max().item()
    assert (g_direct - ph2.grad).abs().max().item() <= 1e-3 * scale


def test_forward_only_inputs_still_work_and_A_stays_constant():
    import mlamg
    import ns.model.loss as loss
    z, A, P, kw = load("poisson3d_6x6x5_nu2")
    val = loss.amg_loss(P.astype(np.float32), A, torch.from_numpy(z["test_vecs"].copy()), **kw)        # scipy in, no gradient asked
    assert not val.requires_grad
    assert abs(float(val) - float(z["loss"])) <= 2e-5 * abs(float(z["loss"]))


def test_descent_on_the_edge_weights_lowers_the_loss():
    """demos/1d_poisson.py:84-101 without the PNet (torch_geometric is absent): Adam on the P_hat edge values of the
    9-node problem, starting from the smoothed-aggregation weights perturbed — the loss the optimiser sees must go down"""
    import ns.model.loss as loss
    import ns.model.agg_interp as ai
    import mlamg
    z, A, P, kw = load("poisson1d_9")
    n, k = P.shape
    labels = torch.from_numpy(np.repeat(np.arange(k), n // k).astype(np.int32)).cuda()
    S = sp.csr_matrix(sp.eye(n) - (2.0 / 3.0) * sp.diags(1.0 / A.diagonal()) @ A)        # I - omega D^-1 A on A's pattern
    S.sort_indices()
    assert np.array_equal(S.indices, A.indices)
    ph = torch.from_numpy((S.data * (1.0 + 0.5 * np.random.RandomState(0).randn(S.nnz))).astype(np.float32)).cuda().requires_grad_(True)
    Ad = mlamg.DeviceCSR.from_scipy(A, torch.float32)
    tv = torch.from_numpy(z["test_vecs"].copy()).cuda()
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
