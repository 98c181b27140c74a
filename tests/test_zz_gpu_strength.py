"""GPU: evolution strength of connection (mlamg/strength.py, csrc/strength.cu) through the C ABI.
  * every kernel against the numpy statement of its contract (tests/strength_ref.py): bit for bit;
  * the whole measure against the oracle's restatement of pyamg (oracle.pyamg_restated) for an injected rho: pattern
    and values bit for bit (k = 1, 2, 4; with / without symmetrisation; another drop tolerance), incl. the reference's
    two measures 'evolution' and 'olson' (utils/common.py:27,30);
  * pyamg's Arnoldi estimate of rho from the same seeded numpy stream: 1e-10;
  * the reference's default evaluation sequence (utils/common.py:84-111 with the 'olson' measure) against the same
    sequence on the oracle."""
import importlib.util
import os

import numpy as np
import numpy.linalg as la
import pytest
import scipy.sparse as sp
import torch

import strength_ref as sr
from helpers import ROOT
from oracle import pyamg_restated as pr, reference_path as rp, multilevel as oml

pytestmark = pytest.mark.gpu


def anisotropic(n, eps):
    Ax = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n, n))
    eye = sp.eye(n)
    return (sp.kron(eye, Ax) + eps * sp.kron(Ax, eye)).tocsr()


def problems():
    from mlamg import problems as pb
    out = {"poisson2d": sp.csr_matrix(oml.poisson((19, 17))), "poisson3d": sp.csr_matrix(oml.poisson((9, 8, 7))),
           "anisotropic": anisotropic(16, 1e-3), "voronoi_jump": pb.voronoi_jump_problem(14, seed=5)[0],
           "delaunay": pb.delaunay_laplacian(400, seed=1)[0]}
    rs = np.random.RandomState(3)
    W = sp.csr_matrix(oml.poisson((12, 11))).astype(np.float64)
    W.data = W.data * (1.0 + 0.3 * rs.rand(W.nnz))              # symmetric pattern, non-symmetric values
    out["nonsym_values"] = W
    res = {}
    for k, v in out.items():
        v = sp.csr_matrix(v).astype(np.float64)
        v.sort_indices()
        res[k] = v
    return res


def _sp(S):
    return sp.csr_matrix((S.val.cpu().numpy(), S.col.cpu().numpy(), S.rowptr.cpu().numpy()), shape=S.shape)


def assert_bitwise(X, Y, what=""):
    X, Y = sp.csr_matrix(X), sp.csr_matrix(Y)
    X.sort_indices()
    Y.sort_indices()
    assert X.shape == Y.shape and np.array_equal(X.indptr, Y.indptr) and np.array_equal(X.indices, Y.indices), what
    bad = np.nonzero(X.data != Y.data)[0]
    assert bad.size == 0, f"{what}: {bad.size} of {X.nnz} values differ, max {np.abs(X.data - Y.data).max():.3e}"


@pytest.mark.parametrize("name", ["poisson2d", "delaunay", "nonsym_values"])
def test_every_kernel_against_its_contract(name):
    import mlamg
    from mlamg import core, strength
    A = problems()[name]
    Ad = mlamg.DeviceCSR.from_scipy(A)
    assert np.array_equal(Ad.col.cpu().numpy(), A.indices)
    rho = 1.93
    S, DA = strength.evolution_step(Ad, rho, want_dinv_a=True)
    s_ref, t_ref, flag = sr.evolution_step(A.indptr, A.indices, A.data, 1.0 / rho)
    assert flag == 0 and np.array_equal(S.val.cpu().numpy(), s_ref) and np.array_equal(DA.val.cpu().numpy(), t_ref)
    T = core.transpose(S)
    Ts = sp.csr_matrix((s_ref, A.indices, A.indptr), shape=A.shape).T.tocsr()
    Ts.sort_indices()
    assert np.array_equal(T.col.cpu().numpy(), Ts.indices) and np.array_equal(T.val.cpu().numpy(), Ts.data)
    z = strength.incomplete_matmul(T, S, Ad).cpu().numpy()
    z_ref = sr.incomplete_matmul(Ts.indptr, Ts.indices, Ts.data, A.indptr, A.indices, s_ref, A.indptr, A.indices)
    assert np.array_equal(z, z_ref)
    Z = core.DeviceCSR(Ad.rowptr, Ad.col, torch.from_numpy(z_ref.copy()).cuda(), A.shape)
    m = strength.evolution_measure_(Z).val.cpu().numpy()
    m_ref = sr.evolution_measure(A.indptr, A.indices, z_ref)
    assert np.array_equal(m, m_ref) and (m_ref == 0).sum() >= A.shape[0]
    # rules of the measure on hand-made values: stored zero, weak ratio, obtuse angle, near-perfect connection
    hp = np.array([0, 5], dtype=np.int32)
    hj = np.array([0, 1, 2, 3, 4], dtype=np.int32)
    hx = np.array([2.0, 0.0, 1e6, -1.0, 2.0 * (1 + 1e-10)])
    H = core.DeviceCSR(torch.from_numpy(np.concatenate([hp, np.full(4, 5, np.int32)])).cuda(), torch.from_numpy(hj).cuda(),
                       torch.from_numpy(hx.copy()).cuda(), (5, 5))
    hm = strength.evolution_measure_(H).val.cpu().numpy()
    assert np.array_equal(hm, sr.evolution_measure(np.concatenate([hp, np.full(4, 5, np.int32)]), hj, hx))
    assert hm[0] == 0 and hm[1] == 0 and hm[2] == 0 and hm[3] == 0 and hm[4] == 1e-4
    M = sp.csr_matrix((m_ref, A.indices.copy(), A.indptr.copy()), shape=A.shape)      # copies: eliminate_zeros works in place
    M.eliminate_zeros()
    Md = mlamg.DeviceCSR.from_scipy(M)
    f = strength.distance_filter_(Md, 4.0).val.cpu().numpy()
    f_ref = sr.distance_filter(M.indptr, M.indices, M.data, 4.0)
    assert np.array_equal(f, f_ref)
    M.data = f_ref
    M.eliminate_zeros()
    Md = mlamg.DeviceCSR.from_scipy(M)
    for symm in (True, False):
        o = strength.symmetrize_on(Ad, Md, symm).val.cpu().numpy()
        assert np.array_equal(o, sr.evolution_symmetrize(A.indptr, A.indices, M.indptr, M.indices, M.data, symm))
    O = sp.csr_matrix((sr.evolution_symmetrize(A.indptr, A.indices, M.indptr, M.indices, M.data, True), A.indices.copy(), A.indptr.copy()),
                      shape=A.shape)
    O.eliminate_zeros()
    Od = mlamg.DeviceCSR.from_scipy(O)
    inv_ref = sr.invert_scale_rows(O.indptr, O.data)
    assert np.array_equal(strength.invert_scale_rows_(Od).val.cpu().numpy(), inv_ref)
    O.data = inv_ref
    w = 1.0 / np.abs(A.data)
    c = strength.pattern_add(Ad, torch.from_numpy(w).cuda(), mlamg.DeviceCSR.from_scipy(O)).val.cpu().numpy()
    assert np.array_equal(c, sr.pattern_add(A.indptr, A.indices, w, O.indptr, O.indices, O.data))


@pytest.mark.parametrize("name", ["poisson2d", "poisson3d", "anisotropic", "voronoi_jump", "delaunay", "nonsym_values"])
def test_measure_is_bit_identical_to_the_oracle_for_an_injected_rho(name):
    import mlamg
    from mlamg import strength
    A = problems()[name]
    rho = 1.9 if name != "anisotropic" else 1.97
    before = mlamg.launch_count()
    for kw in (dict(), dict(k=4), dict(k=1), dict(symmetrize_measure=False), dict(epsilon=2.0)):
        ref = pr.evolution_strength_of_connection(A, rho=rho, **kw)
        got = _sp(strength.evolution_strength_of_connection(A, rho=rho, **kw))
        assert_bitwise(got, ref, f"{name} {kw}")
    assert mlamg.launch_count() - before >= 5 * 8
    E = pr.evolution_strength_of_connection(A, rho=rho)
    olson = E + sp.csr_matrix((1.0 / np.abs(A.data), A.indices, A.indptr), A.shape)
    assert_bitwise(_sp(strength.olson_measure(A, rho=rho)), olson, "olson")
    evo = E + sp.csr_matrix((np.ones_like(A.data), A.indices, A.indptr), A.shape) * 0.1
    assert_bitwise(_sp(strength.evolution_measure_plus_pattern(A, rho=rho)), evo, "evolution")


def test_arnoldi_estimate_and_argument_errors():
    import mlamg
    from mlamg import strength
    for name in ("poisson2d", "voronoi_jump", "delaunay"):
        A = problems()[name]
        DinvA = sp.csr_matrix(sp.diags(1.0 / A.diagonal()) @ A)
        np.random.seed(0)
        ref, trace = pr.approximate_spectral_radius(DinvA, return_trace=True)
        after_ref = np.random.rand()
        np.random.seed(0)
        got, trace_g = strength.approximate_spectral_radius(mlamg.DeviceCSR.from_scipy(DinvA), return_trace=True)
        assert np.random.rand() == after_ref                                  # same number of draws from the global stream
        assert all(abs(e / r - 0.01) > 2e-4 for r, e in trace), "test problem sits on the restart threshold"
        assert len(trace) == len(trace_g) and abs(got - ref) <= 1e-10 * ref, (name, got, ref)
        np.random.seed(0)
        C_ref = pr.evolution_strength_of_connection(A)
        np.random.seed(0)
        C_got = _sp(strength.evolution_strength_of_connection(A))
        assert np.array_equal(C_ref.indices, C_got.indices) and np.abs(C_ref.data - C_got.data).max() <= 1e-9
    A = problems()["poisson2d"]
    with pytest.raises(ValueError):
        strength.evolution_strength_of_connection(A, epsilon=0.5)
    with pytest.raises(NotImplementedError):
        strength.evolution_strength_of_connection(A, k=3)
    with pytest.raises(NotImplementedError):
        strength.evolution_strength_of_connection(sp.csr_matrix(sp.triu(A)), rho=1.9)
    Cz = (A - sp.diags(A.diagonal())).tocsr()
    Cz.eliminate_zeros()
    with pytest.raises(ValueError):
        strength.evolution_strength_of_connection(Cz, rho=1.9)


class _G:
    def __init__(self, A):
        self.A = A


def test_default_evaluation_sequence_with_the_olson_measure():
    """utils/common.py:84-111 / :40-82 with the reference's default strength measure, against the same sequence on the
    oracle (pyamg restated): seeded global stream -> Arnoldi rho -> evolution measure + 1/|A| -> Lloyd ('same' distances)
    -> smoothed aggregation -> two-grid convergence factor.  Unstructured meshes: no exact distance ties that the last
    bits of rho could flip."""
    from mlamg import problems as pb
    spec = importlib.util.spec_from_file_location("mlamg_utils_common", os.path.join(ROOT, "ml-amg_b200", "utils", "common.py"))
    common = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(common)
    grids = [_G(pb.delaunay_laplacian(500, seed=2)[0]), _G(pb.voronoi_jump_problem(300, seed=4, mesh="delaunay", npts=300)[0])]
    lam = 2.0

    def oracle_olson(A):
        A = sp.csr_matrix(A)
        return pr.evolution_strength_of_connection(A) + sp.csr_matrix((1.0 / np.abs(A.data), A.indices, A.indptr), A.shape)

    def oracle_conv(A, seeded_rand):
        np.random.seed(0)
        C = sp.csr_matrix(oracle_olson(A))
        Agg, _, _ = rp.lloyd_aggregation(C, ratio=0.2, distance="same", rand=seeded_rand)
        P = rp.smoothed_aggregation_jacobi(A, Agg, omega=(4.0 / 3.0) / lam)
        if seeded_rand is None:
            np.random.seed(0)
            x = np.random.randn(A.shape[1])
        else:
            x = np.random.RandomState(0).randn(A.shape[1])
        x /= la.norm(x, 2)
        return rp.amg_2_v(A, P, np.zeros(A.shape[1]), x, res_tol=1e-10, jacobi_weight=2. / 3.)[1]

    got = common.evaluate_ref_conv(grids, common.strength_measure_funcs["olson"], alpha=0.2, lam_max=lambda A: lam)
    ref = [oracle_conv(sp.csr_matrix(g.A), None) for g in grids]
    assert np.allclose(got, ref, rtol=0, atol=1e-8), (got, ref)
    got = common.evaluate_dataset(None, grids, alpha=0.2, lam_max=lambda A: lam)          # S=None: the 'olson' default
    ref = [oracle_conv(sp.csr_matrix(g.A), 0) for g in grids]
    assert np.allclose(got, ref, rtol=0, atol=1e-8), (got, ref)
    assert all(0.0 < c < 1.0 for c in got)
    evo = common.strength_measure_funcs["evolution"](grids[0].A)
    assert evo.shape == grids[0].A.shape and evo.nnz == sp.csr_matrix(grids[0].A).nnz
