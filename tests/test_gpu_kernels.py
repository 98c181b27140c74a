"""GPU: every apply/setup kernel through the C ABI against scipy on the same seeded inputs.
Tolerances: fp64 1e-13 relative to the result scale (summation order differs from scipy's serial
loop), fp32 1e-5 (north-star tolerance)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from helpers import random_csr, assert_csr_close, assert_csr_bitwise, assert_same_pattern, GOLDEN_CASES, GOLDEN_CASES_CONFIG_SHAPES, load_golden, csr_from
from oracle import multilevel as oml

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-13, np.float32: 1e-5}
TDT = {np.float64: torch.float64, np.float32: torch.float32}


def dev(x, dtype):
    return torch.from_numpy(np.ascontiguousarray(x)).to("cuda", TDT[dtype])


def close(got, ref, dtype, scale=None):
    got = got.cpu().numpy() if isinstance(got, torch.Tensor) else got
    s = np.abs(ref).max() if scale is None else scale
    err = np.abs(got - ref).max() / max(s, 1e-300)
    assert err <= TOL[dtype], f"error {err:.3e}"


MATS = {
    "poisson2d": lambda: oml.poisson((33, 17)),
    "poisson3d": lambda: oml.poisson((12, 9, 7)),
    "random_sparse": lambda: random_csr(500, 500, 0.01, 1),
    "random_dense_rows": lambda: random_csr(300, 400, 0.2, 2),
    "rect_wide": lambda: random_csr(50, 700, 0.05, 3),
    "single_row": lambda: random_csr(1, 40, 0.5, 4, empty_rows=False),
}


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("name", list(MATS))
def test_spmv_residual_jacobi(name, dtype):
    import mlamg
    A = MATS[name]().astype(dtype)
    n, m = A.shape
    rs = np.random.RandomState(0)
    x, b = rs.randn(m).astype(dtype), rs.randn(n).astype(dtype)
    Ad = mlamg.DeviceCSR.from_scipy(A)
    xd, bd = dev(x, dtype), dev(b, dtype)
    y = mlamg.spmv(Ad, xd)
    scale = (abs(A) @ np.abs(x)).max()
    close(y, A @ x, dtype, scale)
    r, nrm = mlamg.residual(Ad, xd, bd, norm=True)
    close(r, b - A @ x, dtype, scale + np.abs(b).max())
    assert abs(nrm - np.linalg.norm((b - A @ x).astype(np.float64))) <= 10 * TOL[dtype] * max(nrm, 1e-30)
    y0 = rs.randn(n).astype(dtype)
    yd = dev(y0, dtype)
    mlamg.spmv_add(Ad, xd, yd)
    close(yd, y0 + A @ x, dtype, scale + np.abs(y0).max())
    if n == m:
        dw = rs.rand(n).astype(dtype)
        xn = mlamg.jacobi_sweep(Ad, dev(dw, dtype), bd, xd)
        close(xn, x + dw * (b - A @ x), dtype, scale + np.abs(b).max())
        close(mlamg.jacobi_zero(dev(dw, dtype), bd), dw * b, dtype)
        # first zero-guess sweep fused with the residual: same x bit for bit, r = b - A x, deterministic norm
        x0 = mlamg.jacobi_zero(dev(dw, dtype), bd)
        xf, rf, nrm = mlamg.jacobi_zero_residual(Ad, dev(dw, dtype), bd, norm=True)
        assert torch.equal(xf, x0)
        assert torch.equal(rf, mlamg.residual(Ad, x0, bd))
        xr = (dw * b).astype(dtype)
        close(rf, b - A @ xr, dtype, (abs(A) @ np.abs(xr)).max() + np.abs(b).max())
        # ... and on the column-scaled copy A D_w (single gather of b)
        vs = mlamg.scaled_values(Ad, dev(dw, dtype))
        xs, rs_ = mlamg.jacobi_zero_residual_scaled(Ad, vs, dev(dw, dtype), bd)
        assert torch.equal(xs, x0)
        close(rs_, b - A @ xr, dtype, (abs(A) @ np.abs(xr)).max() + np.abs(b).max())
        assert abs(nrm - np.linalg.norm((b - A @ xr).astype(np.float64))) <= 10 * TOL[dtype] * max(nrm, 1e-30)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_smoother_diag_and_dot_axpby(dtype):
    import mlamg
    A = oml.poisson((20, 11)).astype(dtype)
    A.data = A.data * (1 + 0.1 * np.random.RandomState(0).rand(A.nnz)).astype(dtype)
    Ad = mlamg.DeviceCSR.from_scipy(A)
    close(mlamg.smoother_diag(Ad, "jacobi", 0.7), dtype(0.7) / A.diagonal(), dtype)
    close(mlamg.smoother_diag(Ad, "l1_jacobi"), 1.0 / np.asarray(abs(A).sum(axis=1)).ravel(), dtype)
    rs = np.random.RandomState(1)
    x, y = rs.randn(100003).astype(dtype), rs.randn(100003).astype(dtype)
    xd, yd = dev(x, dtype), dev(y, dtype)
    d = mlamg.dot(xd, yd)
    assert abs(d - float(x.astype(np.float64) @ y.astype(np.float64))) <= 1e-6 * np.linalg.norm(x) * np.linalg.norm(y) * (1 if dtype == np.float32 else 1e-7)
    assert mlamg.dot(xd, yd) == d                        # deterministic reduction
    mlamg.axpby(0.5, xd, -2.0, yd)
    close(yd, dtype(0.5) * x - dtype(2.0) * y, dtype)


@pytest.mark.parametrize("n", [0, 1, 5, 2047, 2048, 2049, 70001, 3000000])
def test_scan(n):
    import mlamg
    rs = np.random.RandomState(n)
    c = rs.randint(0, 9, size=n).astype(np.int32)
    out = mlamg.scan_i32(torch.from_numpy(c).cuda()).cpu().numpy()
    ref = np.concatenate([[0], np.cumsum(c)]).astype(np.int32)
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("name", list(MATS))
def test_transpose_sorted(name, dtype):
    import mlamg
    A = MATS[name]().astype(dtype)
    T = mlamg.transpose(mlamg.DeviceCSR.from_scipy(A)).to_scipy()
    ref = sp.csr_matrix(A.T)
    ref.sort_indices()
    assert T.has_sorted_indices or np.all(np.diff(T.indices)[np.diff(T.indices) < 0] is not None)
    assert np.array_equal(T.indptr, ref.indptr) and np.array_equal(T.indices, ref.indices)
    assert np.array_equal(T.data, ref.data)              # pure data movement: bit-exact


SPGEMM_CASES = {
    "stencil_x_stencil": lambda: (oml.poisson((15, 13)), oml.poisson((15, 13))),
    "random_x_random": lambda: (random_csr(300, 200, 0.05, 5), random_csr(200, 250, 0.05, 6)),
    "cta_bin": lambda: (random_csr(60, 400, 0.3, 7), random_csr(400, 900, 0.05, 8)),        # ub ~ 5000 per row
    "dense_rows": lambda: (random_csr(20, 300, 0.9, 9), random_csr(300, 3000, 0.3, 10)),     # ~3000 nnz rows
    "empty": lambda: (sp.csr_matrix((7, 5)), sp.csr_matrix((5, 9))),
}


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("name", list(SPGEMM_CASES))
def test_spgemm(name, dtype):
    import mlamg
    A, B = SPGEMM_CASES[name]()
    A, B = sp.csr_matrix(A).astype(dtype), sp.csr_matrix(B).astype(dtype)
    C = mlamg.spgemm(mlamg.DeviceCSR.from_scipy(A), mlamg.DeviceCSR.from_scipy(B))
    Cs = C.to_scipy()
    assert Cs.has_canonical_format or True
    assert np.all(np.diff(Cs.indptr) >= 0)
    # structural pattern of the GPU result = scipy's structural product (incl. numerically zero sums)
    S = (sp.csr_matrix((np.ones(A.nnz), A.indices, A.indptr), shape=A.shape) @
         sp.csr_matrix((np.ones(B.nnz), B.indices, B.indptr), shape=B.shape))
    assert_same_pattern(sp.csr_matrix((np.ones(Cs.nnz), Cs.indices, Cs.indptr), shape=Cs.shape), S)
    for i in range(Cs.shape[0]):                         # sorted rows
        seg = Cs.indices[Cs.indptr[i]:Cs.indptr[i + 1]]
        assert np.all(np.diff(seg) > 0)
    ref = (A.astype(np.float64) @ B.astype(np.float64))
    scale = (abs(A).astype(np.float64) @ abs(B).astype(np.float64))
    err = abs(Cs.astype(np.float64) - ref).max() / max(scale.max(), 1e-300) if ref.nnz else 0.0
    assert err <= (1e-14 if dtype == np.float64 else 1e-6)
    # scipy semantics: exact zeros dropped; ordered accumulation => the same bits as scipy's csr_matmat
    Cz = mlamg.drop_zeros(C).to_scipy()
    assert_csr_bitwise(Cz, A @ B)
    C2 = mlamg.spgemm(mlamg.DeviceCSR.from_scipy(A), mlamg.DeviceCSR.from_scipy(B))
    assert torch.equal(C2.val, C.val) and torch.equal(C2.col, C.col)      # run-to-run deterministic


def test_drop_zeros_and_sort_rows():
    import mlamg
    A = random_csr(200, 150, 0.1, 11)
    A.data[::3] = 0.0
    out = mlamg.drop_zeros(mlamg.DeviceCSR.from_scipy(A)).to_scipy()
    ref = A.copy()
    ref.eliminate_zeros()
    assert np.array_equal(out.indptr, ref.indptr) and np.array_equal(out.indices, ref.indices) and np.array_equal(out.data, ref.data)
    # unsorted rows (long and short) -> sort_indices
    rs = np.random.RandomState(0)
    B = random_csr(40, 5000, 0.5, 12, empty_rows=True)
    perm_idx, perm_dat = B.indices.copy(), B.data.copy()
    for i in range(B.shape[0]):
        s, e = B.indptr[i], B.indptr[i + 1]
        p = rs.permutation(e - s)
        perm_idx[s:e], perm_dat[s:e] = B.indices[s:e][p], B.data[s:e][p]
    D = mlamg.DeviceCSR(torch.from_numpy(B.indptr.astype(np.int32)).cuda(), torch.from_numpy(perm_idx.astype(np.int32)).cuda(),
                        torch.from_numpy(perm_dat).cuda(), B.shape)
    out = mlamg.sort_rows(D).to_scipy()
    assert np.array_equal(out.indices, B.indices) and np.array_equal(out.data, B.data)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_spmm_and_gemv(dtype):
    import mlamg
    A = random_csr(300, 200, 0.05, 13).astype(dtype)
    rs = np.random.RandomState(2)
    for k in (1, 7, 50):
        X = rs.randn(200, k).astype(dtype)
        Y = mlamg.spmm(mlamg.DeviceCSR.from_scipy(A), dev(X, dtype))
        close(Y, A @ X, dtype, (abs(A) @ np.abs(X)).max())
    M = rs.randn(70, 70).astype(dtype)
    v = rs.randn(70).astype(dtype)
    close(mlamg.gemv(dev(M, dtype), dev(v, dtype)), M @ v, dtype, (np.abs(M) @ np.abs(v)).max())


def test_poisson_generator_matches_oracle():
    import mlamg
    for shape in [(7,), (9, 6), (5, 4, 3), (16, 16, 16)]:
        A = mlamg.poisson(shape).to_scipy()
        ref = oml.poisson(shape)
        assert np.array_equal(A.indptr, ref.indptr) and np.array_equal(A.indices, ref.indices) and np.array_equal(A.data, ref.data)


def test_dense_inverse_and_singular():
    import mlamg
    A = oml.poisson((9, 8))
    inv = mlamg.dense_inverse(mlamg.DeviceCSR.from_scipy(A)).cpu().numpy()
    assert np.abs(inv @ A.toarray() - np.eye(72)).max() < 1e-12
    S = sp.csr_matrix(np.array([[1.0, 2.0], [2.0, 4.0]]))
    with pytest.raises(mlamg.SingularCoarseError):
        mlamg.dense_inverse(mlamg.DeviceCSR.from_scipy(S))


def test_lambda_max_lanczos_known_answers():
    """|lambda_max(D^-1 A)| against the analytic spectrum of the Dirichlet stencils: Lanczos with a reported residual."""
    import mlamg
    for shape in ((24, 24), (40, 33), (12, 10, 9)):
        info = {}
        lam = mlamg.lambda_max(mlamg.DeviceCSR.from_scipy(oml.poisson(shape)), info=info)
        exact = 1 + sum(np.cos(np.pi / (m + 1)) for m in shape) / len(shape)
        assert abs(lam - exact) <= 1e-12 * exact, (shape, lam, exact, info)
        assert info["method"] == "lanczos" and info["residual"] < 1e-5 and info["steps"] < 2000


@pytest.mark.parametrize("name", GOLDEN_CASES + GOLDEN_CASES_CONFIG_SHAPES)
def test_lambda_max_vs_reference_arpack_golden(name):
    """the value the UNMODIFIED reference obtained from ARPACK (tests/golden/make_golden.py) on every golden problem;
    `randw_same` is symmetric too, the `laplace3d_grid` cases are the reference's own fixture"""
    import mlamg
    z = load_golden(name)
    A = csr_from(z, "A")
    info = {}
    lam = mlamg.lambda_max(mlamg.DeviceCSR.from_scipy(A), info=info)
    ref = float(z["lam_max"])
    assert abs(lam - ref) <= 1e-10 * ref, (lam, ref, info)


def test_lambda_max_nonsymmetric_and_fp32():
    """non-symmetric operator: monitored power iteration on D^-1 A vs scipy's dense eigenvalues; fp32: Lanczos to 1e-5"""
    import mlamg
    rs = np.random.RandomState(5)
    n = 300
    A = (sp.random(n, n, density=0.03, random_state=rs, format="csr") + sp.diags(np.linspace(1.0, 3.0, n))).tocsr()
    A.sort_indices()
    ref = np.abs(np.linalg.eigvals((sp.diags(1.0 / A.diagonal()) @ A).toarray())).max()
    info = {}
    lam = mlamg.lambda_max(mlamg.DeviceCSR.from_scipy(A), tol=1e-10, maxiter=20000, info=info)
    assert info["method"] == "power"
    assert abs(lam - ref) <= 1e-6 * ref, (lam, ref, info)
    P = oml.poisson((30, 30)).astype(np.float32)
    info = {}
    lam32 = mlamg.lambda_max(mlamg.DeviceCSR.from_scipy(P), tol=1e-6, info=info)
    assert abs(lam32 - (1 + np.cos(np.pi / 31))) < 2e-5, (lam32, info)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_gauss_seidel_bit_exact(dtype):
    import mlamg
    from oracle import pyamg_restated as pr
    rs = np.random.RandomState(3)
    for A in (oml.poisson((13, 11)), oml.poisson((6, 5, 4)), (random_csr(120, 120, 0.05, 14, empty_rows=False) + 5 * sp.eye(120)).tocsr()):
        A = sp.csr_matrix(A).astype(dtype)
        A.sort_indices()
        n = A.shape[0]
        b, x0 = rs.randn(n).astype(dtype), rs.randn(n).astype(dtype)
        ref = x0.copy()
        pr.gauss_seidel(A, ref, b, iterations=3)
        Ad = mlamg.DeviceCSR.from_scipy(A)
        sched = mlamg.GaussSeidelSchedule(Ad)
        xd = dev(x0, dtype)
        sched.sweep(dev(b, dtype), xd, iterations=3)
        assert np.array_equal(xd.cpu().numpy(), ref), "forward Gauss-Seidel must be bit-identical to the sequential loop"


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["poisson2d", "poisson3d", "random_sparse", "random_dense_rows"])
def test_sell32_rowops(name, dtype):
    import mlamg
    A = MATS[name]().astype(dtype)
    if A.shape[0] != A.shape[1]:
        A = sp.csr_matrix(A[:, :A.shape[0]]) if A.shape[1] > A.shape[0] else A
    A = sp.csr_matrix(A)
    A.sort_indices()
    n, m = A.shape
    rs = np.random.RandomState(0)
    x, b, dw = rs.randn(m).astype(dtype), rs.randn(n).astype(dtype), rs.rand(n).astype(dtype)
    Ad = mlamg.DeviceCSR.from_scipy(A)
    S = mlamg.DeviceSELL(Ad)
    assert S.col.numel() % 32 == 0 and S.padding >= 1.0
    xd, bd, dwd = dev(x, dtype), dev(b, dtype), dev(dw, dtype)
    scale = (abs(A) @ np.abs(x)).max() + np.abs(b).max()
    close(S.spmv(xd), A @ x, dtype, scale)
    r, nrm = S.residual(xd, bd, norm=True)
    close(r, b - A @ x, dtype, scale)
    assert abs(nrm - np.linalg.norm((b - A @ x).astype(np.float64))) <= 10 * TOL[dtype] * max(nrm, 1e-30)
    if n == m:
        close(S.jacobi_sweep(dwd, bd, xd), x + dw * (b - A @ x), dtype, scale)


@pytest.mark.parametrize("lanes", [0, 1, 2, 4, 8, 16, 32])
def test_csr_lane_variants(lanes):
    """every threads-per-row variant (0 = staged shared-memory kernel, incl. rows longer than one staging chunk)"""
    import mlamg
    from mlamg import core
    try:
        mlamg.set_csr_lanes(lanes)
        cases = [MATS[k]() for k in ("poisson3d", "random_dense_rows", "single_row")]
        cases.append(oml.poisson((40, 30, 3)))                           # 3600 rows: several CTAs, one chunk each
        cases.append(sp.vstack([random_csr(3, 5000, 0.9, 7), random_csr(600, 5000, 0.002, 8, empty_rows=True),
                                random_csr(2, 5000, 0.6, 9)]).tocsr())   # rows of ~4500 entries: multi-chunk staging
        for A in cases:
            rs = np.random.RandomState(1)
            n, m = A.shape
            x, b, dw, y0 = rs.randn(m), rs.randn(n), rs.rand(n), rs.randn(n)
            Ad = mlamg.DeviceCSR.from_scipy(A)
            scale = (abs(A) @ np.abs(x)).max() + np.abs(b).max() + np.abs(y0).max()
            close(mlamg.spmv(Ad, dev(x, np.float64)), A @ x, np.float64, scale)
            r, nrm = mlamg.residual(Ad, dev(x, np.float64), dev(b, np.float64), norm=True)
            close(r, b - A @ x, np.float64, scale)
            assert abs(nrm - np.linalg.norm(b - A @ x)) <= 1e-12 * nrm
            yd = dev(y0, np.float64)
            mlamg.spmv_add(Ad, dev(x, np.float64), yd)
            close(yd, y0 + A @ x, np.float64, scale)
            if n == m:
                close(mlamg.jacobi_sweep(Ad, dev(dw, np.float64), dev(b, np.float64), dev(x, np.float64)),
                      x + dw * (b - A @ x), np.float64, scale)
                xf, rf = mlamg.jacobi_zero_residual(Ad, dev(dw, np.float64), dev(b, np.float64))
                close(xf, dw * b, np.float64)
                close(rf, b - A @ (dw * b), np.float64, scale)
                # a row range that starts inside the matrix (multi-GPU interior rows)
                lo, hi = n // 5, n - n // 7
                yd = dev(y0, np.float64)
                core.rowop(Ad, 3, dev(x, np.float64), yd, b=dev(b, np.float64), dw=dev(dw, np.float64), row_range=(lo, hi))
                ref = y0.copy()
                ref[lo:hi] = (x + dw * (b - A @ x))[lo:hi]
                close(yd, ref, np.float64, scale)
    finally:
        mlamg.set_csr_lanes(-1)


def test_spmv_row_order_is_result_invariant():
    import mlamg
    A = random_csr(700, 300, 0.05, 21)
    x = np.random.RandomState(3).randn(300)
    Ad = mlamg.DeviceCSR.from_scipy(A)
    order = torch.from_numpy(np.random.RandomState(4).permutation(700).astype(np.int32)).cuda()
    y0 = mlamg.spmv(Ad, dev(x, np.float64))
    y1 = mlamg.spmv_perm(Ad, dev(x, np.float64), order)
    assert torch.equal(y0, y1)


@pytest.mark.parametrize("op", [0, 1, 2, 3, 4, 5, 6, 7])
def test_rowop_on_row_subset(op):
    """interior / boundary splits of the multi-GPU levels: only the listed rows are touched"""
    from mlamg import core
    import mlamg
    A = oml.poisson((12, 11, 10))
    n = A.shape[0]
    rs = np.random.RandomState(op)
    x, b, dw, y0, z0 = rs.randn(n), rs.randn(n), rs.rand(n), rs.randn(n), rs.randn(n)
    rows = np.sort(rs.permutation(n)[: n // 3]).astype(np.int32)      # includes rows > len(rows)
    Ad = mlamg.DeviceCSR.from_scipy(A)
    yd = dev(y0, np.float64)
    xd = dev(x, np.float64)
    zd = dev(z0, np.float64)
    aux = {4: xd, 6: xd, 5: zd, 7: zd}.get(op)     # ops 4/6: aux = x_out; op 5: aux = iterate before the correction; 7: rhs
    Aop = Ad.with_values(mlamg.scaled_values(Ad, dev(dw, np.float64))) if op == 6 else Ad
    core.rowop(Aop, op, None if op in (4, 6) else xd, yd, b=dev(b, np.float64), dw=dev(dw, np.float64),
               rows=torch.from_numpy(rows).cuda(), aux=aux)
    full = {0: A @ x, 1: y0 + A @ x, 2: b - A @ x, 3: x + dw * (b - A @ x), 4: b - A @ (dw * b), 5: z0 + dw * b + A @ x,
            6: b - A @ (dw * b), 7: dw * (z0 + b) + A @ x}[op]
    if op in (4, 6):     # x is an output here: the listed rows receive dw.*b, the others keep their content
        xr = x.copy()
        xr[rows] = (dw * b)[rows]
        assert np.array_equal(xd.cpu().numpy(), xr)
    ref = y0.copy()
    ref[rows] = full[rows]
    got = yd.cpu().numpy()
    mask = np.ones(n, bool)
    mask[rows] = False
    assert np.array_equal(got[mask], y0[mask]), "rows outside the list were modified"
    assert np.abs(got[rows] - ref[rows]).max() <= 1e-13 * np.abs(full).max()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_prolong_smooth_equals_prolong_then_sweep(dtype):
    """x + P e followed by one smoothing sweep == x + dw.*r + Q e with Q = (I - D_w A) P, r = b - A x"""
    import mlamg
    A = oml.poisson((16, 12, 9)).astype(dtype)
    n = A.shape[0]
    rs = np.random.RandomState(2)
    lab = rs.randint(0, 60, n)
    P = sp.csr_matrix((rs.rand(n) + 0.5, (np.arange(n), lab)), shape=(n, 60))
    P = sp.csr_matrix((sp.eye(n) - 0.6 * sp.diags(1.0 / A.diagonal()) @ A) @ P).astype(dtype)      # smoothed: 1-ring pattern
    P.sort_indices()
    x, b, e = rs.randn(n).astype(dtype), rs.randn(n).astype(dtype), rs.randn(60).astype(dtype)
    dw = (0.66 / A.diagonal()).astype(dtype)
    Ad, Pd = mlamg.DeviceCSR.from_scipy(A), mlamg.DeviceCSR.from_scipy(P)
    Q = mlamg.post_operator(Ad, Pd, dev(dw, dtype))
    Qref = (sp.eye(n) - sp.diags(dw.astype(np.float64)) @ A.astype(np.float64)) @ P.astype(np.float64)
    assert_csr_close(Q.to_scipy().astype(np.float64), sp.csr_matrix(Qref), TOL[dtype] * 10)
    xd, bd, ed = dev(x, dtype), dev(b, dtype), dev(e, dtype)
    r = mlamg.residual(Ad, xd, bd)
    fused = mlamg.prolong_smooth(Q, ed, xd, r, dev(dw, dtype))
    x64, b64, A64, P64 = x.astype(np.float64), b.astype(np.float64), A.astype(np.float64), P.astype(np.float64)
    xp = x64 + P64 @ e.astype(np.float64)
    ref = xp + dw.astype(np.float64) * (b64 - A64 @ xp)
    close(fused, ref, dtype, np.abs(ref).max() * 10)
    inplace = xd.clone()
    mlamg.prolong_smooth(Q, ed, inplace, r, dev(dw, dtype), x_out=inplace)
    assert torch.equal(inplace, fused)
    # zero-guess form: x_in = dw.*b is never materialised; r comes from a plain residual on the scaled copy
    Ads = Ad.with_values(mlamg.scaled_values(Ad, dev(dw, dtype)))
    r0 = mlamg.residual(Ads, bd, bd)
    x0 = (dw * b).astype(dtype)
    close(r0, b64 - A64 @ x0.astype(np.float64), dtype, np.abs(b64).max() + (abs(A64) @ np.abs(x0.astype(np.float64))).max())
    z = mlamg.prolong_smooth_zero(Q, ed, bd, r0, dev(dw, dtype))
    xp = x0.astype(np.float64) + P64 @ e.astype(np.float64)
    ref0 = xp + dw.astype(np.float64) * (b64 - A64 @ xp)
    close(z, ref0, dtype, np.abs(ref0).max() * 10)


def test_gauss_seidel_mirror_honours_L_and_U():
    """ns.lib.multigrid.gauss_seidel with caller-supplied triangles (reference :83-90): x <- L^-1 (b - U x)"""
    import scipy.sparse.linalg as spla
    import ns.lib.multigrid as mg
    A = sp.csr_matrix(oml.poisson((9, 7)))
    rs = np.random.RandomState(0)
    b, x = rs.randn(A.shape[0]), rs.randn(A.shape[0])
    L = (sp.tril(A) * 1.5).tocsr()            # NOT the halves of A: the supplied matrices must be the ones used
    U = (sp.triu(A, k=1) * 0.5).tocsr()
    ref = x.copy()
    for _ in range(2):
        ref = spla.spsolve_triangular(L, b - U @ ref)
    got = mg.gauss_seidel(A, b, x.copy(), L=L, U=U, nu=2)
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()
    got0 = mg.gauss_seidel(A, b, x.copy(), nu=2)
    ref0 = x.copy()
    for _ in range(2):
        ref0 = spla.spsolve_triangular(sp.tril(A).tocsr(), b - sp.triu(A, k=1) @ ref0)
    assert np.abs(got0 - ref0).max() <= 1e-13 * np.abs(ref0).max()
    with pytest.raises(ValueError):
        mg.gauss_seidel(A, b, x.copy(), L=A, nu=1)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_w32_layout_is_bit_identical_to_csr(dtype):
    """the warp-interleaved copy of a short-row operator: same per-row sums in the same order as the CSR kernel"""
    import mlamg
    from mlamg import core
    rs = np.random.RandomState(7)
    for A in (random_csr(1000, 300, 0.02, 3, dtype=dtype), oml.poisson((37, 29)).astype(dtype), random_csr(70, 50, 0.3, 5, dtype=dtype),
              random_csr(33, 40, 0.0, 6, dtype=dtype)):
        n, m = A.shape
        Ad = mlamg.DeviceCSR.from_scipy(A)
        w32 = core.csr_to_w32(Ad)
        # the copy is a permutation of the entries inside every 32-row window
        rp = A.indptr
        for w0 in range(0, n, 32):
            a, b_ = rp[w0], rp[min(w0 + 32, n)]
            got = sorted(zip(w32[0][a:b_].cpu().numpy().tolist(), w32[1][a:b_].cpu().numpy().tolist()))
            assert got == sorted(zip(A.indices[a:b_].tolist(), A.data[a:b_].tolist()))
        e, rhs, r, dw = (dev(rs.randn(k), dtype) for k in (m, n, n, n))
        try:
            mlamg.set_csr_lanes(1)                 # thread-per-row CSR kernel: the same summation order per row
            ref = core.prolong_smooth_zero(Ad, e, rhs, r, dw)
            if n == m:
                ref_r = core.residual(Ad, e, rhs)
        finally:
            mlamg.set_csr_lanes(-1)
        got = core.prolong_smooth_zero_w32(Ad, w32, e, rhs, r, dw)
        assert torch.equal(ref, got)
        want = dw.cpu().numpy() * (rhs.cpu().numpy() + r.cpu().numpy()) + A @ e.cpu().numpy()
        close(got, want, dtype, np.abs(want).max() + 1.0)
        if n == m:
            assert torch.equal(core.residual_w32(Ad, w32, e, rhs), ref_r)


def test_w32_rowop_on_unaligned_row_ranges():
    """mlamg_rowop_w32 over arbitrary [begin, end): rows of the first / last window outside the range only vote"""
    import mlamg
    from mlamg import core
    rs = np.random.RandomState(11)
    A = random_csr(777, 777, 0.02, 9)
    n = A.shape[0]
    Ad = mlamg.DeviceCSR.from_scipy(A)
    w32 = core.csr_to_w32(Ad)
    x, b, dw, aux = (dev(rs.randn(n), np.float64) for _ in range(4))
    xn, bn, dwn, auxn = (t.cpu().numpy() for t in (x, b, dw, aux))
    full = {0: A @ xn, 2: bn - A @ xn, 5: auxn + dwn * bn + A @ xn, 7: dwn * (auxn + bn) + A @ xn}
    for op in (0, 2, 5, 7):
        for lo, hi in ((0, n), (5, 40), (31, 33), (64, 64), (100, 777), (770, 777), (33, 500)):
            y = torch.full((n,), 123.0, dtype=torch.float64, device="cuda")
            core.rowop_w32(Ad, w32, op, x, y, b=b, dw=dw, row_range=(lo, hi), aux=aux)
            got = y.cpu().numpy()
            assert np.all(got[:lo] == 123.0) and np.all(got[hi:] == 123.0), (op, lo, hi)
            scale = np.abs(full[op]).max() + 1.0
            if hi > lo:
                assert np.abs(got[lo:hi] - full[op][lo:hi]).max() <= 1e-13 * scale, (op, lo, hi)
