"""CPU: (a) the oracle's restatement of pyamg's evolution strength of connection (known-answer properties),
(b) the product's orchestration of it (mlamg/strength.py) with the C-ABI wrappers replaced — in this test only —
by the numpy contract statements of tests/strength_ref.py: the product code that decides which kernel sees which
operand must reproduce the oracle BIT FOR BIT for an injected rho, and pyamg's Arnoldi estimate of rho to 1e-10
from the same numpy random stream.  The kernels themselves are held to the same references on the GPU
(tests/test_zz_gpu_strength.py)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import strength_ref as sr
from oracle import pyamg_restated as pr, multilevel as oml


def anisotropic(n, eps):
    Ax = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n, n))
    eye = sp.eye(n)
    return (sp.kron(eye, Ax) + eps * sp.kron(Ax, eye)).tocsr()


def problems():
    from mlamg import problems as pb
    out = {"poisson2d": sp.csr_matrix(oml.poisson((9, 8))), "poisson3d": sp.csr_matrix(oml.poisson((5, 4, 4))),
           "anisotropic": anisotropic(8, 1e-3), "voronoi_jump": pb.voronoi_jump_problem(10, seed=5)[0],
           "delaunay": pb.delaunay_laplacian(150, seed=1)[0]}
    rs = np.random.RandomState(3)
    W = sp.csr_matrix(oml.poisson((7, 6))).astype(np.float64)
    W.data = W.data * (1.0 + 0.3 * rs.rand(W.nnz))              # symmetric pattern, non-symmetric values
    out["nonsym_values"] = W
    return {k: sp.csr_matrix(v).astype(np.float64) for k, v in out.items()}


# ---------------------------------------------------------------- (a) the oracle itself
def test_oracle_evolution_strength_known_answers():
    A = sp.csr_matrix(oml.poisson((10, 10)))
    np.random.seed(0)
    C = pr.evolution_strength_of_connection(A)
    assert C.nnz == A.nnz and abs(C - C.T).max() < 1e-15          # isotropic: every connection is strong, measure symmetric
    assert np.all(C.diagonal() > 0) and C.data.max() == 1.0 and np.all(C.data > 0)
    assert np.allclose(np.asarray(abs(C).max(axis=1).todense()).ravel(), 1.0)      # rows scaled by their largest entry
    An = anisotropic(10, 1e-3)
    np.random.seed(0)
    Cn = pr.evolution_strength_of_connection(An)
    coo = Cn.tocoo()
    off = coo.row != coo.col
    assert np.all(np.abs(coo.row[off] - coo.col[off]) == 1), "only the strong (x) direction survives the drop tolerance"
    # rho: pyamg's estimate is a 1 %-class lower bound of the spectral radius, reproducible from the seeded global stream
    DinvA = sp.diags(1.0 / A.diagonal()) @ A
    np.random.seed(0)
    rho1, trace = pr.approximate_spectral_radius(DinvA, return_trace=True)
    np.random.seed(0)
    rho2 = pr.approximate_spectral_radius(DinvA)
    exact = 1.0 + np.cos(np.pi / 11)
    assert rho1 == rho2 and 0.97 * exact < rho1 <= exact * (1 + 1e-12)
    assert trace[-1][1] / trace[-1][0] < 0.01
    # the injected-rho form and the powers of two
    C4 = pr.evolution_strength_of_connection(A, k=4, rho=exact)
    C1 = pr.evolution_strength_of_connection(A, k=1, rho=exact)
    assert C4.nnz == A.nnz and C1.nnz == A.nnz


# ---------------------------------------------------------------- (b) the product's orchestration
def _dev(M):
    from mlamg.core import DeviceCSR
    M = sp.csr_matrix(M)
    M.sort_indices()
    return DeviceCSR(torch.from_numpy(M.indptr.astype(np.int32)), torch.from_numpy(M.indices.astype(np.int32)),
                     torch.from_numpy(M.data.astype(np.float64)), M.shape)


def _sp(S):
    return sp.csr_matrix((S.val.numpy().copy(), S.col.numpy().copy(), S.rowptr.numpy().copy()), shape=S.shape)


def _arrs(S):
    return S.rowptr.numpy(), S.col.numpy(), S.val.numpy()


@pytest.fixture
def cpu_kernels(monkeypatch):
    from mlamg import core, strength

    def transpose(A):
        T = _sp(A).T.tocsr()
        T.sort_indices()
        # keep explicit zeros (scipy's transpose does)
        return _dev_keep(T)

    def _dev_keep(M):
        from mlamg.core import DeviceCSR
        return DeviceCSR(torch.from_numpy(M.indptr.astype(np.int32)), torch.from_numpy(M.indices.astype(np.int32)),
                         torch.from_numpy(M.data.astype(np.float64)), M.shape)

    def drop_zeros(A):
        M = _sp(A)
        if (M.data != 0).all():
            return A
        M.eliminate_zeros()
        return _dev_keep(M)

    def spgemm(A, B):
        M = (_sp(A) @ _sp(B)).tocsr()
        M.sort_indices()
        return _dev_keep(M)

    def wrap(A, dtype=None):
        return A if isinstance(A, core.DeviceCSR) else _dev(A)

    monkeypatch.setattr(core, "require_cuda", lambda: None)
    monkeypatch.setattr(core, "transpose", transpose)
    monkeypatch.setattr(core, "drop_zeros", drop_zeros)
    monkeypatch.setattr(core, "spgemm", spgemm)
    monkeypatch.setattr(core.DeviceCSR, "wrap", staticmethod(wrap))
    monkeypatch.setattr(core, "spmv", lambda A, x, out=None: torch.from_numpy(_sp(A) @ x.numpy()))
    monkeypatch.setattr(core, "dot", lambda x, y: float(np.dot(x.numpy(), y.numpy())))

    def axpby(alpha, x, beta, y):
        y.mul_(beta).add_(x, alpha=alpha)
        return y
    monkeypatch.setattr(core, "axpby", axpby)

    def evolution_step(A, rho, want_dinv_a=False):
        s, t, flag = sr.evolution_step(*_arrs(A), 1.0 / float(rho))
        if flag:
            raise ValueError("evolution strength needs a matrix that stores its diagonal")
        return A.with_values(torch.from_numpy(s)), (A.with_values(torch.from_numpy(t)) if want_dinv_a else None)

    def incomplete_matmul(T, Bt, S):
        return torch.from_numpy(sr.incomplete_matmul(*_arrs(T), *_arrs(Bt), S.rowptr.numpy(), S.col.numpy()))

    def evolution_measure_(Z):
        Z.val = torch.from_numpy(sr.evolution_measure(*_arrs(Z)))
        return Z

    def distance_filter_(M, epsilon):
        M.val = torch.from_numpy(sr.distance_filter(*_arrs(M), epsilon))
        return M

    def symmetrize_on(A, M, symmetrize=True):
        out = sr.evolution_symmetrize(A.rowptr.numpy(), A.col.numpy(), *_arrs(M), symmetrize)
        return core.DeviceCSR(A.rowptr, A.col, torch.from_numpy(out), A.shape)

    def invert_scale_rows_(M):
        M.val = torch.from_numpy(sr.invert_scale_rows(M.rowptr.numpy(), M.val.numpy()))
        return M

    def pattern_add(A, w, E):
        out = sr.pattern_add(A.rowptr.numpy(), A.col.numpy(), w.numpy(), *_arrs(E))
        return core.DeviceCSR(A.rowptr, A.col, torch.from_numpy(out), A.shape)

    for name, fn in dict(evolution_step=evolution_step, incomplete_matmul=incomplete_matmul, evolution_measure_=evolution_measure_,
                         distance_filter_=distance_filter_, symmetrize_on=symmetrize_on, invert_scale_rows_=invert_scale_rows_,
                         pattern_add=pattern_add).items():
        monkeypatch.setattr(strength, name, fn)
    return strength


def assert_bitwise(X, Y):
    X, Y = sp.csr_matrix(X), sp.csr_matrix(Y)
    X.sort_indices()
    Y.sort_indices()
    assert X.shape == Y.shape and np.array_equal(X.indptr, Y.indptr) and np.array_equal(X.indices, Y.indices)
    assert np.array_equal(X.data, Y.data), np.abs(X.data - Y.data).max()


@pytest.mark.parametrize("name", ["poisson2d", "poisson3d", "anisotropic", "voronoi_jump", "delaunay", "nonsym_values"])
def test_orchestration_reproduces_the_oracle_bit_for_bit(cpu_kernels, name):
    strength = cpu_kernels
    A = problems()[name]
    rho = 1.9 if name != "anisotropic" else 1.97
    for kw in (dict(), dict(k=4), dict(k=1), dict(symmetrize_measure=False), dict(epsilon=2.0)):
        ref = pr.evolution_strength_of_connection(A, rho=rho, **kw)
        got = _sp(strength.evolution_strength_of_connection(_dev(A), rho=rho, **kw))
        assert_bitwise(got, ref)
    # the reference's two measures (utils/common.py:27,30)
    E = pr.evolution_strength_of_connection(A, rho=rho)
    olson = E + sp.csr_matrix((1.0 / np.abs(A.data), A.indices, A.indptr), A.shape)
    assert_bitwise(_sp(strength.olson_measure(_dev(A), rho=rho)), olson)
    evo = E + sp.csr_matrix((np.ones_like(A.data), A.indices, A.indptr), A.shape) * 0.1
    assert_bitwise(_sp(strength.evolution_measure_plus_pattern(_dev(A), rho=rho)), evo)


@pytest.mark.parametrize("name", ["poisson2d", "voronoi_jump", "delaunay"])
def test_arnoldi_estimate_of_rho_follows_pyamg_from_the_same_stream(cpu_kernels, name):
    strength = cpu_kernels
    A = problems()[name]
    DinvA = sp.csr_matrix(sp.diags(1.0 / A.diagonal()) @ A)
    np.random.seed(0)
    ref, trace = pr.approximate_spectral_radius(DinvA, return_trace=True)
    after_ref = np.random.rand()
    np.random.seed(0)
    got, trace_g = strength.approximate_spectral_radius(_dev(DinvA), return_trace=True)
    after_got = np.random.rand()
    assert after_ref == after_got, "the device path must consume the global stream exactly like pyamg (n draws)"
    assert len(trace) == len(trace_g) and abs(got - ref) <= 1e-10 * ref
    assert all(abs(e / r - 0.01) > 2e-4 for r, e in trace), "test problem sits on the restart threshold: choose another"
    # default call = estimate + measure: same aggregation input as the oracle's default call, up to the estimate's last bits
    np.random.seed(0)
    C_ref = pr.evolution_strength_of_connection(A)
    np.random.seed(0)
    C_got = _sp(strength.evolution_strength_of_connection(_dev(A)))
    assert np.array_equal(C_ref.indices, C_got.indices) and np.abs(C_ref.data - C_got.data).max() <= 1e-9


def test_argument_errors(cpu_kernels):
    strength = cpu_kernels
    A = problems()["poisson2d"]
    with pytest.raises(ValueError):
        strength.evolution_strength_of_connection(_dev(A), epsilon=0.5)
    with pytest.raises(ValueError):
        strength.evolution_strength_of_connection(_dev(A), k=0)
    with pytest.raises(NotImplementedError):
        strength.evolution_strength_of_connection(_dev(A), k=3)
    B = sp.csr_matrix(sp.triu(A))
    with pytest.raises(NotImplementedError):
        strength.evolution_strength_of_connection(_dev(B), rho=1.9)
    Cz = (A - sp.diags(A.diagonal())).tocsr()
    Cz.eliminate_zeros()
    with pytest.raises(ValueError):
        strength.evolution_strength_of_connection(_dev(Cz), rho=1.9)
