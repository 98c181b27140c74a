"""CPU: the oracle against (a) the golden vectors produced by the UNMODIFIED reference modules,
(b) its own independent pure-Python twin, (c) known-answer facts."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import pyamg_restated as pr, reference_path as rp, multilevel as oml
from helpers import GOLDEN_CASES, GOLDEN_CASES_CONFIG_SHAPES, load_golden, csr_from, assert_csr_close, rel_hist_err, grid_graph


@pytest.mark.parametrize("name", GOLDEN_CASES + GOLDEN_CASES_CONFIG_SHAPES)
def test_restated_path_matches_reference_golden(name):
    z = load_golden(name)
    A, C = csr_from(z, "A"), csr_from(z, "C")
    Agg, roots, seeds = rp.lloyd_aggregation(C, ratio=float(z["ratio"]), distance=str(z["distance"]), rand=int(z["rand"]))
    Agg_ref = csr_from(z, "Agg")
    assert np.array_equal(seeds, z["seeds"]) and np.array_equal(roots, z["roots"])
    assert (Agg != Agg_ref).nnz == 0 and Agg.dtype == np.int8
    omega = (4.0 / 3.0) / float(z["lam_max"])
    P = rp.smoothed_aggregation_jacobi(A, Agg, omega=omega)
    assert_csr_close(P, csr_from(z, "P"), 1e-14)
    assert_csr_close(rp.galerkin(A, sp.csr_matrix(P)), csr_from(z, "AH"), 1e-13)
    # the two-level driver, with the reference's own P
    Pref = csr_from(z, "P")
    n = A.shape[0]
    x0 = np.random.RandomState(0).randn(n)
    x0 /= np.linalg.norm(x0, 2)
    x, conv, err, nit = rp.amg_2_v(A, Pref, np.zeros(n), x0, res_tol=1e-10)
    assert nit == int(z["gs_nit"])
    assert rel_hist_err(err, z["gs_err"]) < 1e-12
    assert abs(conv - float(z["gs_conv"])) < 1e-12
    b2 = np.random.RandomState(1).randn(n)
    x, conv, err, nit = rp.amg_2_v(A, Pref, b2, np.zeros(n), res_tol=1e-8, pre_smoothing_steps=2, post_smoothing_steps=2)
    assert nit == int(z["gs2_nit"]) and rel_hist_err(err, z["gs2_err"]) < 1e-10
    xj = rp.jacobi(A, b2, x0.copy(), omega=0.666, nu=3)
    assert np.allclose(xj, z["jacobi_x"], rtol=1e-14, atol=1e-15)


def test_lambda_max_known_answer():
    # rho(D^-1 A) of the n x n 5-point Dirichlet Laplacian is 1 + cos(pi/(n+1))
    n = 24
    lam = float(load_golden("poisson2d_24_unit")["lam_max"])
    assert abs(lam - (1 + np.cos(np.pi / (n + 1)))) < 1e-10


def test_modified_bellman_ford_matches_reference_golden():
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "ref_modified_bf.npz"))
    n = int(z["n"])
    S = sp.coo_matrix((z["w"], (z["row"], z["col"])), shape=(n, n))
    dist, near = rp.modified_bellman_ford(S, z["centers"])
    assert np.array_equal(dist, z["dist"]) and np.array_equal(near, z["nearest"])
    Agg = rp.nearest_center_to_agg(z["centers"], near).tocoo()
    assert np.array_equal(Agg.row, z["agg_row"]) and np.array_equal(Agg.col, z["agg_col"])


@pytest.mark.parametrize("weights,dtype", [("unit", np.float64), ("random", np.float64), ("relu", np.float32)])
def test_c_restatement_matches_python_twin(weights, dtype):
    G = grid_graph((9, 7), weights, seed=3, dtype=dtype)
    seeds = np.random.RandomState(5).permutation(G.shape[0])[:6].astype(np.int32)
    d1, z1 = pr.bellman_ford(G, seeds)
    d2, z2 = pr.bellman_ford_py(G, seeds)
    assert np.array_equal(d1, d2) and np.array_equal(z1, z2) and d1.dtype == dtype
    a = pr.lloyd_cluster(G, seeds.copy(), 10)
    b = pr.lloyd_cluster_py(G, seeds.copy(), 10)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)


def test_gauss_seidel_c_matches_python_twin_and_triangular_solve():
    A = oml.poisson((7, 6))
    rs = np.random.RandomState(0)
    b, x0 = rs.randn(42), rs.randn(42)
    x1, x2 = x0.copy(), x0.copy()
    pr.gauss_seidel(A, x1, b, iterations=2)
    pr.gauss_seidel_py(A, x2, b, iterations=2)
    assert np.array_equal(x1, x2)
    # reference's own scipy form (multigrid.py:58-90) agrees to rounding
    import scipy.sparse.linalg as spla
    L, U = sp.tril(A).tocsr(), sp.triu(A, k=1).tocsr()
    x3 = x0.copy()
    for _ in range(2):
        x3 = spla.spsolve_triangular(L, b - U @ x3)
    assert np.allclose(x1, x3, rtol=1e-13)


def test_unreachable_nodes_and_keyerror():
    # two disconnected components, seeds only in the first: labels -1, Agg rows empty, KeyError in the BF path
    A = sp.block_diag([oml.poisson((4,)), oml.poisson((3,))]).tocsr()
    G = sp.csr_matrix((np.ones(A.nnz), A.indices, A.indptr), shape=A.shape)
    d, z = pr.bellman_ford(G, np.array([1], dtype=np.int32))
    assert (z[4:] == -1).all() and (z[:4] == 1).all() and (d[4:] == np.finfo(float).max).all()
    with pytest.raises(KeyError):
        rp.nearest_center_to_agg(np.array([1]), z)
    _, w, _ = pr.lloyd_cluster(G, np.array([1], dtype=np.int32), 5)
    assert (w[4:] == -1).all()


def test_multilevel_oracle_converges():
    A = oml.poisson((32, 32))
    lv = oml.build_hierarchy(A, ratio=0.1, lam_max=lambda M: 2.0, max_coarse=40)
    assert len(lv) >= 3
    b = np.random.RandomState(0).randn(A.shape[0])
    x, res, it = oml.pcg(lv, b, tol=1e-8)
    assert res[-1] <= 1e-8 * np.linalg.norm(b) and it < 60
    assert np.linalg.norm(b - A @ x) <= 2e-8 * np.linalg.norm(b)


def test_error_conventions():
    A = oml.poisson((6, 6))
    with pytest.raises(ValueError):
        rp.lloyd_aggregation(A, ratio=0.0)
    with pytest.raises(ValueError):
        rp.lloyd_aggregation(A, ratio=0.5, distance="bogus")
    with pytest.raises(TypeError):
        rp.lloyd_aggregation(A.tocoo(), ratio=0.5)
    with pytest.raises(TypeError):
        rp.lloyd_aggregation(A, ratio=0.5, rand="x")
    with pytest.raises(RuntimeError):
        rp.amg_2_v(A, A, np.zeros(36), np.zeros(36))


def test_gmres_restatement_converges_and_matches_a_direct_solve():
    """oracle.multilevel.gmres (the algorithm `accel='gmres'` is held to, parity unpinned against pyamg): preconditioned
    residual norms never increase inside a restart cycle, the iterate converges to the direct solution, restarts and the
    iteration cap behave, and an exact preconditioner converges in one step"""
    import scipy.sparse.linalg as spla
    A = oml.poisson((17, 15))
    lv = oml.build_hierarchy(A, ratio=0.1, distance="unit", rand=0, lam_max=[2.0, 1.9, 1.8], max_coarse=20)
    b = np.random.RandomState(5).randn(A.shape[0])
    xd = spla.spsolve(sp.csc_matrix(A), b)
    for restart in (30, 5):
        x, res, it = oml.gmres(lv, b, tol=1e-10, maxiter=200, restart=restart)
        assert it == len(res) - 1 and res[-1] <= 1e-10 * res[0] * 10
        assert np.abs(x - xd).max() <= 1e-7 * np.abs(xd).max()
        if restart == 30:
            assert np.all(np.diff(res) <= 1e-12 * res[0])          # one cycle: monotone
    x, res, it = oml.gmres(lv, b, tol=1e-10, maxiter=7, restart=5)
    assert it == 7 and len(res) == 8
    # GMRES needs no more iterations than the stationary V-cycle iteration it accelerates
    _, res_s = oml.solve(lv, b, tol=1e-8, maxiter=200)
    _, res_g, it_g = oml.gmres(lv, b, tol=1e-8, maxiter=200, restart=50)
    assert it_g <= len(res_s) - 1
    one = [lv[0]]                                                  # single level = exact coarse solve as the preconditioner
    x, res, it = oml.gmres(one, b, tol=1e-10, maxiter=10)
    assert it <= 1 and np.abs(x - xd).max() <= 1e-9 * np.abs(xd).max()


def test_amg_loss_oracle_neumann_fix_handles_the_constant_null_space():
    """ns/model/loss.py:11-30,66-82 restated: on a singular operator (graph Laplacian) the bordered coarse solve gives a
    finite convergence estimate below 1, and on a regular operator it agrees with the plain solve up to the mean removal"""
    from oracle import multilevel as oml
    G = sp.csr_matrix(oml.poisson((12, 11)))
    G = G - sp.diags(G.diagonal())
    L = (sp.diags(-np.asarray(G.sum(axis=1)).ravel()) + G).tocsr()
    Agg, _, _ = rp.lloyd_aggregation(L, ratio=0.15, distance="unit", rand=0)
    P = sp.csr_matrix(rp.smoothed_aggregation_jacobi(L, Agg, omega=2.0 / 3.0))
    assert np.abs(P @ np.ones(P.shape[1]) - 1.0).max() < 1e-12          # the coarse space reproduces the null space
    val, errs = rp.amg_loss_forward(P, L, 6, tot_num_loop=5, neumann_solve_fix=True)
    assert np.isfinite(val) and 0.0 < val < 1.0 and np.all(np.diff(errs, axis=0) < 0)


@pytest.mark.parametrize("name", ["poisson1d_9", "neumann1d_9", "poisson2d_14", "poisson3d_6x6x5_nu2"])
def test_amg_loss_oracle_matches_the_unmodified_reference(name):
    """ns/model/loss.py:32-96 restated vs the loss value the reference's own `amg_loss` returned
    (tests/golden/make_golden_loss.py: the 9-node 1-D problem of demos/1d_poisson.py, its Neumann twin with the
    Lagrange-bordered coarse solve, 2-D / 3-D Poisson with Lloyd aggregates, nu = 1 and 2)"""
    import os
    from helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, f"ref_amg_loss_{name}.npz"))
    n, k = int(z["n"]), int(z["k"])
    A = sp.csr_matrix((z["A_data"], z["A_indices"], z["A_indptr"]), shape=(n, n))
    P = sp.csr_matrix((z["P_val"], (z["P_row"], z["P_col"])), shape=(n, k))
    kw = {key[3:]: int(z[key]) for key in z.files if key.startswith("kw_")}
    val, _ = rp.amg_loss_forward(P, A, z["test_vecs"], neumann_solve_fix=bool(z["neumann"]), **kw)
    assert abs(val - float(z["loss"])) <= 1e-6 * float(z["loss"])
    assert abs(float(z["loss_alt"]) - float(z["loss"])) <= 1e-6 * float(z["loss"])      # the reference's own fp32 noise floor


def _modes():
    import os
    from helpers import GOLDEN
    return np.load(os.path.join(GOLDEN, "ref_two_level_modes.npz"))


def test_two_level_driver_modes_match_the_unmodified_reference():
    """ns/lib/multigrid.py in the modes the per-problem goldens do not cover (tests/golden/make_golden_modes.py):
    error_tol measure, singular=True (lsqr + mean removal), missing tolerance, singular coarse operator, the scipy
    Gauss-Seidel form and the fp32 torch twins"""
    z = _modes()
    A, P = csr_from(z, "A"), csr_from(z, "P")
    n = A.shape[0]
    x0, b2 = z["x0"], z["b2"]
    x, conv, err, nit = rp.amg_2_v(A, P, np.zeros(n), x0, error_tol=1e-9)
    assert nit == int(z["errtol_nit"]) and rel_hist_err(err, z["errtol_err"]) < 1e-12
    assert abs(conv - float(z["errtol_conv"])) < 1e-10 and np.abs(x - z["errtol_x"]).max() < 1e-12
    with pytest.raises(RuntimeError) as ei:
        rp.amg_2_v(A, P, np.zeros(n), x0)
    assert str(ei.value) == str(z["no_tol_message"]) == "One of res_tol or error_tol must be set!"
    Pz = csr_from(z, "Pz")
    xs, convs, errs, nits = rp.amg_2_v(A, Pz, np.zeros(n), x0, res_tol=1e-10)
    assert convs == float(z["singcoarse_conv"]) == 1.0 and nits == int(z["singcoarse_nit"]) == 0 and bool(z["singcoarse_x_is_x0"])
    assert np.array_equal(xs, x0)
    xg = x0.copy()
    pr.gauss_seidel(A, xg, b2, iterations=3)                      # pyamg's sweep == the reference's own triangular-solve form
    assert np.abs(xg - z["gs_scipy_x"]).max() < 1e-13
    # fp32 twins
    A32, P32 = A.astype(np.float32), P.astype(np.float32)
    xj = rp.jacobi_torch_like(A32, z["jacobi_torch_b"], x0.astype(np.float32), (1.0 / A32.diagonal()).astype(np.float32), omega=0.666, nu=3)
    assert np.abs(xj - z["jacobi_torch_x"]).max() <= 2e-6 * np.abs(z["jacobi_torch_x"]).max()
    cf = rp.amg_2_v_torch_like(A32, P32, np.zeros(n, dtype=np.float32), x0.astype(np.float32), jacobi_weight=2.0 / 3.0)
    assert abs(float(cf) - float(z["amg_2_v_torch_conv"])) <= 1e-3 * float(z["amg_2_v_torch_conv"])
    cf2 = rp.amg_2_v_torch_like(A32, P32, np.zeros(n, dtype=np.float32), x0.astype(np.float32), pre_smoothing_steps=2,
                                post_smoothing_steps=2, jacobi_weight=0.5, max_iter=12)
    assert abs(float(cf2) - float(z["amg_2_v_torch_conv_nu2"])) <= 1e-3 * float(z["amg_2_v_torch_conv_nu2"])
    # singular=True on the Neumann problem
    L, PN = csr_from(z, "L"), csr_from(z, "PN")
    x, conv, err, nit = rp.amg_2_v(L, PN, np.zeros(L.shape[0]), z["xn0"], res_tol=1e-8, singular=True)
    # lsqr stops at its default 1e-6 tolerances: the history is reproducible to about that accuracy only (stored-order of P^T A P changes
    # the last bits of its matvecs), measured 1e-7
    assert nit == int(z["sing_nit"]) and rel_hist_err(err, z["sing_err"]) < 2e-6 and abs(conv - float(z["sing_conv"])) < 1e-5


@pytest.mark.parametrize("name", ["poisson2d_20", "poisson3d_8"])
def test_mlamg_pc_loop_matches_the_unmodified_reference(name):
    """ns/preconditioner/MLAMG.py:143-212 restated (`rp.mlamg_jacobi`, `rp.mlamg_amg_2_v`) vs what the reference class itself
    returned from `apply` (tests/golden/make_golden_pc.py: seeded random guess, splu COLAMD coarse solve)"""
    import os
    import scipy.sparse.linalg as spla
    from helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "ref_mlamg_pc.npz"))
    A = sp.csr_matrix((z[f"{name}_A_data"], z[f"{name}_A_indices"], z[f"{name}_A_indptr"]))
    P = sp.csr_matrix((z[f"{name}_P_data"], z[f"{name}_P_indices"], z[f"{name}_P_indptr"]), shape=tuple(z[f"{name}_P_shape"]))
    n = A.shape[0]
    w, rtol, b = float(z[f"{name}_jacobi_weight"]), float(z[f"{name}_amg_rtol"]), z[f"{name}_b"]
    Dw = sp.diags(1.0 / A.diagonal()) * w
    lu = spla.splu(sp.csc_matrix(P.T @ A @ P), permc_spec="COLAMD")
    np.random.seed(0)
    x0 = np.random.normal(size=n)                                       # MLAMG.py:209
    x, it = rp.mlamg_amg_2_v(A, P, lu.solve, Dw, b, x0, amg_rtol=rtol)
    assert np.abs(x - z[f"{name}_x"]).max() <= 1e-12 * np.abs(z[f"{name}_x"]).max() and 1 < it < 500
    xj = rp.mlamg_jacobi(A, Dw, b, np.random.RandomState(4).randn(n), nu=3)
    assert np.abs(xj - z[f"{name}_jacobi_x"]).max() <= 1e-13 * np.abs(z[f"{name}_jacobi_x"]).max()


def _oracle_measure(name, A):
    A = sp.csr_matrix(A)
    if name == "abs":
        return abs(A)
    if name == "invabs":
        return sp.csr_matrix((1.0 / np.abs(A.data), A.indices, A.indptr), A.shape)
    if name == "unit":
        return sp.csr_matrix((np.ones_like(A.data), A.indices, A.indptr), A.shape)
    E = pr.evolution_strength_of_connection(A)
    if name == "evolution":
        return E + sp.csr_matrix((np.ones_like(A.data), A.indices, A.indptr), A.shape) * 0.1
    return E + sp.csr_matrix((1.0 / np.abs(A.data), A.indices, A.indptr), A.shape)


def test_evaluation_sequences_match_the_unmodified_reference_drivers():
    """The oracle's statement of utils/common.py:40-111 (seed -> strength measure -> Lloyd 'same' -> SA -> amg_2_v) against the
    convergence factors the reference's own evaluate_ref_conv / evaluate_dataset returned (tests/golden/make_golden_eval.py);
    ARPACK's lambda_max as recorded there is injected."""
    from helpers import load_eval_golden
    z, grids = load_eval_golden()
    for measure in ("abs", "evolution", "invabs", "unit", "olson"):
        for g, (name, A) in enumerate(grids.items()):
            np.random.seed(0)
            C = sp.csr_matrix(_oracle_measure(measure, A))
            Agg, _, _ = rp.lloyd_aggregation(C, ratio=0.2, distance="same")             # rand=None: the global stream (common.py:91)
            P = rp.smoothed_aggregation_jacobi(A, Agg, omega=(4.0 / 3.0) / float(z[f"ref_conv_{measure}_lam"][g]))
            np.random.seed(0)
            x = np.random.randn(A.shape[1])
            x /= np.linalg.norm(x, 2)
            conv = rp.amg_2_v(A, P, np.zeros(A.shape[1]), x, res_tol=1e-10, jacobi_weight=2. / 3.)[1]
            assert abs(conv - float(z[f"ref_conv_{measure}"][g])) < 1e-10, (measure, name)
    for key, S, alpha in (("dataset_conv_default", "olson", 0.2), ("dataset_conv_invabs", "invabs", 0.3)):
        for g, (name, A) in enumerate(grids.items()):
            np.random.seed(0)
            C = sp.csr_matrix(_oracle_measure(S, A))
            Agg, _, _ = rp.lloyd_aggregation(C, ratio=alpha, distance="same", rand=0)   # common.py:58
            P = rp.smoothed_aggregation_jacobi(A, Agg, omega=(4.0 / 3.0) / float(z[f"{key}_lam"][g]))
            x = np.random.RandomState(0).randn(A.shape[1])
            x /= np.linalg.norm(x, 2)
            conv = rp.amg_2_v(A, P, np.zeros(A.shape[1]), x, res_tol=1e-10)[1]
            assert abs(conv - float(z[key][g])) < 1e-10, (key, name)
