"""GPU: hierarchy setup (T, P, RAP) and the cycle drivers against the oracle.
Parity bar (BASELINE.json north_star): aggregates and sparsity patterns bit-exact (canonical sorted
CSR, exact zeros dropped), P / RAP values and residual histories within 1e-12 relative in fp64
(1e-5 in fp32); iteration counts identical."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from helpers import GOLDEN_CASES, load_golden, csr_from, assert_csr_close, assert_csr_bitwise, rel_hist_err, hist_err0
from oracle import reference_path as rp, multilevel as oml, pyamg_restated as pr

pytestmark = pytest.mark.gpu
RTOL64 = 1e-12
RTOL32 = 1e-5


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_sa_prolongator_and_galerkin_vs_reference_golden(name):
    import mlamg
    import ns.lib.multigrid as mg
    z = load_golden(name)
    A, Agg = csr_from(z, "A"), csr_from(z, "Agg")
    omega = (4.0 / 3.0) / float(z["lam_max"])
    P = mg.smoothed_aggregation_jacobi(A, Agg, omega=omega)
    assert sp.isspmatrix_csr(P)
    assert_csr_bitwise(P, csr_from(z, "P"))             # stronger than the 1e-12 bar: same bits as scipy
    Pd = mlamg.DeviceCSR.from_scipy(csr_from(z, "P"))
    AH = mlamg.galerkin(mlamg.DeviceCSR.from_scipy(A), Pd).to_scipy()
    assert_csr_bitwise(AH, csr_from(z, "AH"))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_amg_2_v_mirror_vs_reference_golden(name):
    """ns.lib.multigrid.amg_2_v with the reference's smoother (exact Gauss-Seidel): same iteration
    count, residual history and convergence factor as the unmodified reference run."""
    import ns.lib.multigrid as mg
    z = load_golden(name)
    A, P = csr_from(z, "A"), csr_from(z, "P")
    n = A.shape[0]
    x0 = np.random.RandomState(0).randn(n)
    x0 /= np.linalg.norm(x0, 2)
    x, conv, err, nit = mg.amg_2_v(A, P, np.zeros(n), x0, res_tol=1e-10)
    assert nit == int(z["gs_nit"])
    assert hist_err0(err, z["gs_err"]) < RTOL64          # whole history, relative to the initial residual
    assert rel_hist_err(err[:10], z["gs_err"][:10]) < RTOL64 * 10
    assert abs(conv - float(z["gs_conv"])) < 1e-9
    assert np.abs(x - z["gs_x"]).max() < 1e-12
    b2 = np.random.RandomState(1).randn(n)
    x, conv, err, nit = mg.amg_2_v(A, P, b2, np.zeros(n), res_tol=1e-8, pre_smoothing_steps=2, post_smoothing_steps=2)
    assert nit == int(z["gs2_nit"]) and hist_err0(err, z["gs2_err"]) < RTOL64
    xj = mg.jacobi(A, b2, x0.copy(), omega=0.666, nu=3)
    assert np.abs(xj - z["jacobi_x"]).max() <= 1e-13 * np.abs(z["jacobi_x"]).max()


def test_amg_2_v_jacobi_and_error_tol_and_singular():
    import ns.lib.multigrid as mg
    z = load_golden("poisson2d_24_unit")
    A, P = csr_from(z, "A"), csr_from(z, "P")
    n = A.shape[0]
    x0 = np.random.RandomState(0).randn(n)
    x0 /= np.linalg.norm(x0, 2)
    for kw in (dict(res_tol=1e-9), dict(error_tol=1e-9)):
        ref = rp.amg_2_v(A, P, np.zeros(n), x0, smoother="jacobi", jacobi_weight=2 / 3, **kw)
        got = mg.amg_2_v(A, P, np.zeros(n), x0, smoother="jacobi", jacobi_weight=2 / 3, **kw)
        assert got[3] == ref[3]
        assert rel_hist_err(got[2][:12], ref[2][:12]) < RTOL64 * 10
        assert hist_err0(got[2], ref[2]) < RTOL64
    with pytest.raises(RuntimeError):
        mg.amg_2_v(A, P, np.zeros(n), x0)
    # singular coarse operator: (x, 1.0, err, 0) without raising (multigrid.py:166-170)
    Pz = sp.csr_matrix(np.hstack([P.toarray()[:, :1], P.toarray()[:, :1]]))
    out = mg.amg_2_v(A, Pz, np.zeros(n), x0, res_tol=1e-9)
    assert out[1] == 1.0 and out[3] == 0


@pytest.mark.parametrize("shape,ratio,smoother", [((48, 40), 0.1, "jacobi"), ((14, 12, 10), 0.05, "jacobi"),
                                                  ((40, 40), 0.1, "l1_jacobi")])
def test_multilevel_setup_and_cycles_fp64(shape, ratio, smoother):
    import mlamg
    A = oml.poisson(shape)
    lam = [2.0, 1.9, 1.8, 1.7, 1.6, 1.5]
    ref = oml.build_hierarchy(A, ratio=ratio, distance="unit", rand=0, lam_max=lam, max_coarse=30, smoother=smoother)
    H = mlamg.build_hierarchy(A, aggregates="lloyd", ratio=ratio, distance="unit", rand=0, lam_max=lam, max_coarse=30,
                              smoother=smoother)
    assert len(H.levels) == len(ref) >= 3
    for Lg, Lr in zip(H.levels, ref):
        assert_csr_bitwise(Lg.A.to_scipy(), Lr.A)        # pattern and values identical at every level
        if Lr.P is not None:
            assert np.array_equal(Lg.labels.cpu().numpy(), Lr.labels), "aggregate labels must be bit-exact"
            assert_csr_bitwise(Lg.P.to_scipy(), Lr.P)
            assert_csr_bitwise(Lg.R.to_scipy(), Lr.R)
            assert np.allclose(Lg.dw.cpu().numpy(), Lr.dw, rtol=1e-15)
    n = A.shape[0]
    b = np.random.RandomState(0).randn(n)
    bd = torch.from_numpy(b).cuda()
    # one cycle from zero (preconditioner apply) and from a nonzero guess, V(1,1) and V(2,3)
    for nu1, nu2 in ((1, 1), (2, 3), (0, 1), (1, 0)):
        xr = oml.vcycle(ref, b.copy(), None, nu1, nu2)
        xg = H.vcycle(bd, None, nu1, nu2).cpu().numpy()
        assert np.abs(xg - xr).max() <= RTOL64 * np.abs(xr).max(), (nu1, nu2)
        x0 = np.random.RandomState(1).randn(n)
        xr = oml.vcycle(ref, b.copy(), x0.copy(), nu1, nu2)
        xg = H.vcycle(bd, torch.from_numpy(x0).cuda(), nu1, nu2).cpu().numpy()
        assert np.abs(xg - xr).max() <= RTOL64 * np.abs(xr).max(), (nu1, nu2)
    # graph replay gives the same bits as plain launches
    x_plain = H.vcycle(bd, None, 1, 1)
    H.use_graph(True)
    xbuf = torch.empty_like(bd)
    from mlamg import core
    core.check(core.lib.mlamg_vcycle(H._h, core.ptr(bd), core.ptr(xbuf), 1, 1, 1, core.stream()))
    core.check(core.lib.mlamg_vcycle(H._h, core.ptr(bd), core.ptr(xbuf), 1, 1, 1, core.stream()))
    assert torch.equal(xbuf, x_plain)
    H.use_graph(False)
    # stationary iteration and PCG: residual histories and iteration counts
    xr, res_r = oml.solve(ref, b, tol=1e-8, maxiter=60)
    xg, res_g = H.solve(b, tol=1e-8, maxiter=60, return_residuals=True)
    assert len(res_g) == len(res_r)
    assert hist_err0(res_g, res_r) < RTOL64 and rel_hist_err(res_g[:8], res_r[:8]) < RTOL64 * 10
    xr, res_r, it_r = oml.pcg(ref, b, tol=1e-8, maxiter=100)
    xg, res_g = H.solve(b, tol=1e-8, maxiter=100, accel="cg", return_residuals=True)
    assert len(res_g) - 1 == it_r
    assert hist_err0(res_g, res_r) < RTOL64 * 10 and rel_hist_err(res_g[:6], res_r[:6]) < RTOL64 * 100
    assert np.linalg.norm(b - A @ xg) <= 2e-8 * np.linalg.norm(b)
    # preconditioner through host buffers and scipy CG
    import scipy.sparse.linalg as spla
    M = H.aspreconditioner()
    z_host = M @ b
    assert np.abs(z_host - oml.vcycle(ref, b.copy(), None, 1, 1)).max() <= RTOL64 * np.abs(z_host).max()
    xs, info = spla.cg(A, b, M=M, rtol=1e-8, maxiter=100)
    assert info == 0
    assert H.cycle_bytes() > 0 and "levels" in repr(H)


def test_fused_post_operator_cycle_matches_plain_cycle():
    """Hierarchy(fuse_post=True) (default: one pass over Q = (I - D_w A) P on the way up) against the plain
    prolongation + sweep, for several sweep counts, zero and non-zero guesses"""
    import mlamg
    A = oml.poisson((20, 18, 11))
    lam = [2.0, 1.9, 1.8, 1.7]
    kw = dict(aggregates="lloyd", ratio=0.05, distance="unit", rand=0, lam_max=lam, max_coarse=30)
    Hf = mlamg.build_hierarchy(A, fuse_post=True, fuse_pre=True, **kw)
    Hp = mlamg.build_hierarchy(A, fuse_post=False, fuse_pre=False, **kw)
    assert len(Hf._Q) == len(Hf.levels) - 1 and not Hp._Q and len(Hf._scaled) == len(Hf.levels) - 1 and not Hp._scaled
    assert Hf.cycle_bytes(1, 1) < Hp.cycle_bytes(1, 1)
    n = A.shape[0]
    b = torch.from_numpy(np.random.RandomState(0).randn(n)).cuda()
    x0 = torch.from_numpy(np.random.RandomState(1).randn(n)).cuda()
    for nu1, nu2 in ((1, 1), (2, 2), (0, 1), (1, 0), (3, 1)):
        for guess in (None, x0):
            xf = Hf.vcycle(b, None if guess is None else guess.clone(), nu1, nu2)
            xp = Hp.vcycle(b, None if guess is None else guess.clone(), nu1, nu2)
            assert float((xf - xp).abs().max()) <= 1e-13 * float(xp.abs().max()), (nu1, nu2)


def test_gmres_acceleration_matches_the_restated_algorithm():
    """Hierarchy.solve(accel='gmres') (PyAMG.py:119) against oracle.multilevel.gmres: same iteration count, same
    preconditioned-residual history, with restarts and with an iteration cap inside a restart cycle"""
    import mlamg
    A = oml.poisson((26, 22))
    lam = [2.0, 1.9, 1.8, 1.7]
    ref = oml.build_hierarchy(A, ratio=0.1, distance="unit", rand=0, lam_max=lam, max_coarse=20)
    H = mlamg.build_hierarchy(A, aggregates="lloyd", ratio=0.1, distance="unit", rand=0, lam_max=lam, max_coarse=20)
    b = np.random.RandomState(3).randn(A.shape[0])
    for restart, maxiter in ((30, 100), (6, 100), (5, 7)):
        xr, res_r, it_r = oml.gmres(ref, b, tol=1e-8, maxiter=maxiter, restart=restart)
        xg, res_g = H.solve(b, tol=1e-8, maxiter=maxiter, accel="gmres", restart=restart, return_residuals=True)
        assert len(res_g) - 1 == it_r, (restart, maxiter, len(res_g) - 1, it_r)
        assert hist_err0(res_g, res_r) < 1e-10
        assert np.abs(xg - xr).max() <= 1e-9 * np.abs(xr).max()
    xg = H.solve(b, tol=1e-8, maxiter=100, accel="gmres")
    assert np.linalg.norm(b - A @ xg) <= 1e-6 * np.linalg.norm(b)
    x0 = np.random.RandomState(4).randn(A.shape[0])
    xr, res_r, it_r = oml.gmres(ref, b, x0=x0, tol=1e-8, maxiter=100, restart=8)
    xg, res_g = H.solve(b, x0=x0, tol=1e-8, maxiter=100, accel="gmres", restart=8, return_residuals=True)
    assert len(res_g) - 1 == it_r and hist_err0(res_g, res_r) < 1e-10


def test_multilevel_fp32():
    import mlamg
    A = oml.poisson((40, 36))
    lam = [2.0] * 6
    ref = oml.build_hierarchy(A, ratio=0.1, rand=0, lam_max=lam, max_coarse=30)
    H = mlamg.build_hierarchy(A.astype(np.float32), ratio=0.1, rand=0, lam_max=lam, max_coarse=30)
    assert H.dtype == torch.float32 and len(H.levels) == len(ref)
    for Lg, Lr in zip(H.levels, ref):
        assert_csr_close(Lg.A.to_scipy().astype(np.float64), Lr.A, RTOL32)
        if Lr.P is not None:
            assert np.array_equal(Lg.labels.cpu().numpy(), Lr.labels)
    b = np.random.RandomState(0).randn(A.shape[0])
    xr = oml.vcycle(ref, b.copy(), None, 1, 1)
    xg = H.vcycle(torch.from_numpy(b.astype(np.float32)).cuda(), None, 1, 1).cpu().numpy()
    assert np.abs(xg - xr).max() <= RTOL32 * np.abs(xr).max()


def test_external_aggregates_and_learned_P():
    """GNN-style inputs (agg_interp.py:469-484): centres -> Bellman-Ford -> Agg -> P = P_hat Agg -> RAP."""
    import mlamg
    import ns.lib.graph as g
    torch.manual_seed(0)
    A = oml.poisson((30, 30))
    n = A.shape[0]
    rs = np.random.RandomState(0)
    k = int(np.ceil(0.1 * n))
    top_k = np.sort(rs.permutation(n)[:k])
    Cw = np.maximum(rs.randn(A.nnz), 0).astype(np.float32) + np.float32(1e-3)
    C = sp.csr_matrix((Cw, A.indices, A.indptr), shape=A.shape)
    d_ref, near_ref = pr.bellman_ford(C, top_k)
    dist, near, _ = mlamg.bellman_ford(mlamg.DeviceCSR.from_arrays(C.indptr, C.indices, C.data, C.shape), top_k)
    assert np.array_equal(near.cpu().numpy(), near_ref) and np.array_equal(dist.cpu().numpy(), d_ref)
    agg_T = g.nearest_center_to_agg(torch.from_numpy(top_k), near.cpu())
    Agg_ref = rp.nearest_center_to_agg(top_k, near_ref)
    agg_sp = sp.coo_matrix((agg_T.values().numpy(), agg_T.indices().numpy()), shape=tuple(agg_T.shape)).tocsr()
    assert (agg_sp != Agg_ref).nnz == 0
    phat = np.maximum(rs.randn(A.nnz), 0).astype(np.float32)
    P_hat = sp.csr_matrix((phat, A.indices, A.indptr), shape=A.shape)
    P_ref = rp.learned_prolongator(P_hat, Agg_ref)
    labels = mlamg.center_rank_labels(torch.from_numpy(top_k.astype(np.int32)).cuda(), near)
    Aggd = mlamg.agg_from_labels(labels, k, torch.float32)
    Pd = mlamg.learned_prolongator(mlamg.DeviceCSR.from_scipy(P_hat), Aggd)
    assert_csr_close(mlamg.drop_zeros(Pd).to_scipy().astype(np.float64), P_ref.astype(np.float64), RTOL32)
    # full hierarchy from external labels + learned weights (fp64 copy of the fp32 producers)
    H = mlamg.build_hierarchy(A, aggregates=[(labels, k)], P_hat=[phat.astype(np.float64)], max_coarse=k + 1)
    ref_P = sp.csr_matrix(P_hat.astype(np.float64) @ Agg_ref.astype(np.float64))
    assert_csr_close(mlamg.drop_zeros(H.levels[0].P).to_scipy(), ref_P, RTOL64)
    assert_csr_close(H.levels[1].A.to_scipy(), ref_P.T @ A @ ref_P, RTOL64)


def test_amg_loss_forward_and_torch_twins():
    import ns.model.loss as loss
    import ns.lib.multigrid as mg
    import ns.lib.sparse as nsp
    z = load_golden("poisson2d_24_unit")
    A, P = csr_from(z, "A"), csr_from(z, "P")
    A_T, P_T = nsp.to_torch_sparse(A), nsp.to_torch_sparse(P)
    val = float(loss.amg_loss(P_T, A_T, 8, tot_num_loop=5))
    ref, _ = rp.amg_loss_forward(P, A, 8, tot_num_loop=5)
    assert abs(val - ref) <= 2e-4 * abs(ref)
    n = A.shape[0]
    x0 = np.random.RandomState(0).randn(n).astype(np.float32)
    x0 /= np.linalg.norm(x0)
    xt = torch.from_numpy(x0.copy())
    cf = mg.amg_2_v_torch(A_T, P_T, torch.zeros(n), xt, jacobi_weight=2 / 3)
    assert not np.array_equal(xt.numpy(), x0)            # the caller's iterate is updated in place, as in the reference
    ref = rp.amg_2_v_torch_like(A.astype(np.float32), P.astype(np.float32), np.zeros(n, dtype=np.float32), x0, jacobi_weight=2 / 3)
    assert abs(float(cf) - float(ref)) <= 1e-3 * abs(float(ref))
    assert nsp.torch_to_scipy(A_T).shape == A.shape
    # neumann_solve_fix: singular operator (graph Laplacian, constant null space), Lagrange-bordered coarse solve (:11-30, :66-82)
    G = sp.csr_matrix(oml.poisson((12, 11)))
    G = G - sp.diags(G.diagonal())
    L = (sp.diags(-np.asarray(G.sum(axis=1)).ravel()) + G).tocsr()
    Agg, _, _ = rp.lloyd_aggregation(L, ratio=0.15, distance="unit", rand=0)
    Pn = sp.csr_matrix(rp.smoothed_aggregation_jacobi(L, Agg, omega=2.0 / 3.0))
    val_n = float(loss.amg_loss(nsp.to_torch_sparse(Pn), nsp.to_torch_sparse(L), 6, tot_num_loop=5, neumann_solve_fix=True))
    ref_n, _ = rp.amg_loss_forward(Pn, L, 6, tot_num_loop=5, neumann_solve_fix=True)
    assert abs(val_n - ref_n) <= 5e-4 * abs(ref_n), (val_n, ref_n)


class _FakeVec:
    def __init__(self, a):
        self.array_r = a
        self.out = None

    def setArray(self, a):
        self.out = np.array(a)


class _FakeMat:
    def __init__(self, A):
        self.A = sp.csr_matrix(A)

    def getValuesCSR(self):
        return self.A.indptr, self.A.indices, self.A.data


class _FakePC:
    def __init__(self, A):
        self.m = _FakeMat(A)
        self.appctx = {}

    def getType(self):
        return "python"

    def getOptionsPrefix(self):
        return ""

    def getOperators(self):
        return self.m, self.m


def test_pc_plugins_vs_oracle():
    """MLAMG.apply against the restated MLAMG.py:148-212 (same seeded random guess, same P): identical iterate to
    rounding; PyAMG.apply against the oracle's multilevel hierarchy + preconditioned GMRES."""
    import scipy.sparse.linalg as spla
    from ns.preconditioner.MLAMG import MLAMG
    from ns.preconditioner.PyAMG import PyAMG
    A = oml.poisson((20, 20))
    n = A.shape[0]
    b = np.random.RandomState(0).randn(n)
    # --- MLAMG: P is built as the plugin builds it (Lloyd ratio 0.1 unit rand 0, SA with the ARPACK omega) and injected,
    # so that both sides iterate on identical operators
    Agg, _, _ = rp.lloyd_aggregation(A, ratio=0.1, distance="unit", rand=0)
    P = sp.csr_matrix(rp.smoothed_aggregation_jacobi(A, Agg, omega=(4.0 / 3.0) / rp.lambda_max_dinv_a(A)))
    pc = _FakePC(A)
    pc.appctx["mlamg_P"] = P
    p = MLAMG()
    p.initialize(pc)
    X, Y = _FakeVec(b), _FakeVec(None)
    np.random.seed(0)
    p.apply(pc, X, Y)
    np.random.seed(0)
    x0 = np.random.normal(size=n)                                   # the plugin's random guess (MLAMG.py:209)
    lu = spla.splu(sp.csc_matrix(P.T @ A @ P), permc_spec="COLAMD")
    x_ref, it_ref = rp.mlamg_amg_2_v(A, P, lu.solve, sp.diags((2.0 / 3.0) / A.diagonal()), b, x0.copy(), amg_rtol=1e-8)
    assert np.linalg.norm(b - A @ Y.out) <= 1e-8 * (1 + 1e-6)
    assert np.abs(Y.out - x_ref).max() <= 1e-11 * np.abs(x_ref).max(), np.abs(Y.out - x_ref).max()
    # the same loop through the stationary solver of the two-level hierarchy: iteration count + history
    H = p.two.hierarchy("jacobi", 2.0 / 3.0)
    xd = torch.from_numpy(x0).cuda()
    _, hist = H.solve_abs(torch.from_numpy(b).cuda(), xd, 1e-8, 500, 1, 1, H.SOLVE_NO_INITIAL_CHECK)
    assert len(hist) - 1 == it_ref, (len(hist) - 1, it_ref)
    assert H.loop_mode in ("while-node", "host")
    # default construction (no injected P): the plugin's own Lloyd + Lanczos-omega P gives the same solve to 1e-8
    pc2 = _FakePC(A)
    p2 = MLAMG()
    p2.initialize(pc2)
    Y2 = _FakeVec(None)
    np.random.seed(0)
    p2.apply(pc2, X, Y2)
    assert np.abs(Y2.out - x_ref).max() <= 1e-7 * np.abs(x_ref).max()
    # --- PyAMG: multilevel hierarchy + left-preconditioned GMRES (PyAMG.py:94,119)
    pc3 = _FakePC(A)
    q = PyAMG()
    q.initialize(pc3)
    Y3 = _FakeVec(None)
    q.apply(pc3, X, Y3)
    lams = [(4.0 / 3.0) / L.omega_sa for L in q.Amg.levels[:-1]]
    ref = oml.build_hierarchy(A, ratio=0.1, distance="unit", rand=0, lam_max=lams, max_levels=10, max_coarse=500)
    xr, res_r, it_r = oml.gmres(ref, b, tol=1e-8, maxiter=200)
    assert np.abs(Y3.out - xr).max() <= 1e-10 * np.abs(xr).max()


def test_solver_loops_while_node_equals_host_fallback(monkeypatch):
    """the device-resident PCG / stationary loops: the CUDA-graph WHILE node and the host-driven fallback run the same
    kernels in the same order -> identical histories and iterates; both against the oracle"""
    import mlamg
    A = oml.poisson((40, 36))
    n = A.shape[0]
    b = np.random.RandomState(3).randn(n)
    lam = [2.0, 1.9, 1.8]
    ref = oml.build_hierarchy(A, ratio=0.1, distance="unit", rand=0, lam_max=lam, max_coarse=30)
    out = {}
    for mode in ("while", "host"):
        if mode == "host":
            monkeypatch.setenv("MLAMG_SOLVER_HOST_LOOP", "1")
        H = mlamg.build_hierarchy(A, aggregates="lloyd", ratio=0.1, distance="unit", rand=0, lam_max=lam, max_coarse=30)
        xg, res_g = H.solve(b, tol=1e-9, maxiter=100, accel="cg", return_residuals=True)
        xs, res_s = H.solve(b, tol=1e-6, maxiter=100, return_residuals=True)
        x2, res_2 = H.solve(b, tol=1e-9, maxiter=7, accel="cg", return_residuals=True)       # maxiter reached
        x3, res_3 = H.solve(np.zeros(n), tol=1e-9, maxiter=7, accel="cg", return_residuals=True)   # converged at entry
        assert len(res_2) == 8 and len(res_3) == 1
        out[mode] = (xg, res_g, xs, res_s, H.loop_mode)
    assert out["host"][4] == "host"
    if out["while"][4] == "while-node":
        for i in range(4):
            assert np.array_equal(out["while"][i], out["host"][i])
    xr, res_r, it_r = oml.pcg(ref, b, tol=1e-9, maxiter=100)
    assert len(out["while"][1]) - 1 == it_r and hist_err0(out["while"][1], res_r) < RTOL64 * 10
    xr, res_r = oml.solve(ref, b, tol=1e-6, maxiter=100)
    assert len(out["while"][3]) == len(res_r) and hist_err0(out["while"][3], res_r) < RTOL64


def test_amg_2_v_modes_on_device_loop():
    """amg_2_v with Jacobi smoothing runs behind mlamg_solve_ex: error_tol mode (||x||), singular mode (pseudo-inverse +
    mean removal), max_iter exhaustion — all against the restated reference loop"""
    import ns.lib.multigrid as mg
    z = load_golden("poisson2d_24_unit")
    A, P = csr_from(z, "A"), csr_from(z, "P")
    n = A.shape[0]
    x0 = np.random.RandomState(0).randn(n)
    x0 /= np.linalg.norm(x0, 2)
    kw = dict(smoother="jacobi", jacobi_weight=2 / 3)
    ref = rp.amg_2_v(A, P, np.zeros(n), x0, error_tol=1e-9, pre_smoothing_steps=2, post_smoothing_steps=1, **kw)
    got = mg.amg_2_v(A, P, np.zeros(n), x0, error_tol=1e-9, pre_smoothing_steps=2, post_smoothing_steps=1, **kw)
    assert got[3] == ref[3] and hist_err0(got[2], ref[2]) < RTOL64 and abs(got[1] - ref[1]) < 1e-9
    ref = rp.amg_2_v(A, P, np.zeros(n), x0, res_tol=1e-30, max_iter=12, **kw)
    got = mg.amg_2_v(A, P, np.zeros(n), x0, res_tol=1e-30, max_iter=12, **kw)
    assert got[3] == ref[3] == 12 and len(got[2]) == 12 and hist_err0(got[2], ref[2]) < RTOL64
    # singular (pure Neumann) problem: graph Laplacian of the grid, constant null space
    G = sp.csr_matrix(oml.poisson((12, 11)))
    G = G - sp.diags(G.diagonal())
    L = (sp.diags(-np.asarray(G.sum(axis=1)).ravel()) + G).tocsr()
    Agg, _, _ = rp.lloyd_aggregation(L, ratio=0.15, distance="unit", rand=0)
    Pn = sp.csr_matrix(rp.smoothed_aggregation_jacobi(L, Agg, omega=2.0 / 3.0))
    m = L.shape[0]
    b = np.random.RandomState(1).randn(m)
    b -= b.mean()
    ref = rp.amg_2_v(L, Pn, b, np.zeros(m), res_tol=1e-8, singular=True, **kw)
    got = mg.amg_2_v(L, Pn, b, np.zeros(m), res_tol=1e-8, singular=True, **kw)
    # the reference solves the singular coarse problem with lsqr at its default 1e-6 tolerances; the exact pseudo-inverse
    # used here agrees with it to that accuracy only
    assert abs(got[3] - ref[3]) <= 1 and hist_err0(got[2][:8], ref[2][:8]) < 1e-4


def test_two_level_inner_hierarchy_coarse_solve():
    """coarse operators above the dense limit are solved by an inner AMG-preconditioned CG (rtol 1e-14): same two-level
    histories as the dense coarse inverse"""
    import ns.lib.multigrid as mg
    z = load_golden("poisson2d_24_unit")
    A, P = csr_from(z, "A"), csr_from(z, "P")
    n = A.shape[0]
    x0 = np.random.RandomState(0).randn(n)
    x0 /= np.linalg.norm(x0, 2)
    ref = rp.amg_2_v(A, P, np.zeros(n), x0, res_tol=1e-10)
    old = mg.MAX_DENSE_COARSE
    try:
        mg.MAX_DENSE_COARSE = 16
        got = mg.amg_2_v(A, P, np.zeros(n), x0, res_tol=1e-10)
        gotj = mg.amg_2_v(A, P, np.zeros(n), x0, res_tol=1e-10, smoother="jacobi", jacobi_weight=2 / 3)
    finally:
        mg.MAX_DENSE_COARSE = old
    assert got[3] == ref[3] and hist_err0(got[2], ref[2]) < RTOL64 * 10
    refj = rp.amg_2_v(A, P, np.zeros(n), x0, res_tol=1e-10, smoother="jacobi", jacobi_weight=2 / 3)
    assert gotj[3] == refj[3] and hist_err0(gotj[2], refj[2]) < RTOL64 * 10
