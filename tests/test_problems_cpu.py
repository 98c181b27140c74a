"""CPU: the synthetic problem generators (host side) and the `.grid` container — no kernel runs here."""
import os

import numpy as np
import scipy.sparse as sp

from oracle import multilevel as oml


def test_structured_p1_laplacian_is_the_5_point_stencil():
    from mlamg import problems
    pts, tris, bnd = problems.structured_triangles(9, 9)           # square cells: h cancels, diagonal edges decouple
    A, p = problems.remove_dirichlet(problems.p1_stiffness(pts, tris), pts, bnd)
    ref = oml.poisson((8, 8))
    assert A.shape == ref.shape and abs(A - ref).max() < 1e-12
    assert p.shape == (64, 2) and p.min() > 0 and p.max() < 1
    pts, tris, bnd = problems.structured_triangles(9, 7)           # anisotropic cells: still symmetric, zero row sums inside
    A = problems.p1_stiffness(pts, tris)
    assert abs(A - A.T).max() < 1e-13 and abs(np.asarray(A.sum(axis=1))).max() < 1e-12


def test_voronoi_jump_and_delaunay_shapes():
    from mlamg import problems
    for mesh, size in (("structured", 20), ("delaunay", 1500)):
        A, pts, jumps = problems.voronoi_jump_problem(size, seed=2, mesh=mesh, npts=size)
        assert jumps.shape[1] == 3 and jumps.shape[0] in (2, 3) and np.ptp(jumps[:, 2]) > 1e3
        assert abs(A - A.T).max() <= 1e-12 * abs(A).max()
        assert np.linalg.eigvalsh(A.toarray()).min() > 0                       # SPD after Dirichlet elimination
        assert A.has_sorted_indices and pts.shape == (A.shape[0], 2)
    A, pts = problems.delaunay_laplacian(4000, seed=0)
    assert 6.5 < A.nnz / A.shape[0] < 7.2                                       # ~7 entries per row (config 4)
    assert abs(A - A.T).max() <= 1e-12 * abs(A).max()
    assert (np.asarray(A.sum(axis=1)).ravel() > -1e-12).all()                   # weakly diagonally dominant rows sum >= 0
    # Morton ordering: contiguous row blocks are spatially compact (half of the rows ~ half of the domain)
    half = pts[: len(pts) // 2]
    assert (half[:, 1].max() - half[:, 1].min()) < 0.75


def test_rotated_anisotropic_and_1d_shapes():
    from mlamg import problems
    # tensor form reduces to the scalar form for K = kappa I, and to the axis-aligned anisotropic stencil for theta = 0
    pts, tris, bnd = problems.structured_triangles(6, 6)
    A_iso = problems.p1_stiffness(pts, tris, None)
    assert abs(problems.p1_stiffness(pts, tris, None, tensor=np.eye(2)) - A_iso).max() < 1e-13
    A, p = problems.rotated_anisotropic_problem(8, epsilon=0.01, theta=0.0)
    d = np.sort(np.unique(np.round(A.data, 10)))
    assert np.allclose(d, [-1.0, -0.01, 2.02])                       # -1 along x, -eps along y, 2(1+eps) on the diagonal
    for theta in (0.3, np.pi / 4):
        A, p = problems.rotated_anisotropic_problem(10, epsilon=1e-3, theta=theta)
        assert abs(A - A.T).max() < 1e-12 and np.linalg.eigvalsh(A.toarray()).min() > 0
    A, p = problems.rotated_anisotropic_problem(0, epsilon=0.1, theta=1.0, mesh="delaunay", npts=400)
    assert abs(A - A.T).max() < 1e-12 and np.linalg.eigvalsh(A.toarray()).min() > 0
    # 1-D generators (data.py:244-297)
    A, x = problems.poisson_1d(9)
    h = x[1] - x[0]
    assert np.allclose(A.toarray()[4, 3:6] * h * h, [-1, 2, -1]) and A.shape == (9, 9)
    A, x = problems.poisson_1d(9, neumann=True)
    assert abs(np.asarray(A.sum(axis=1))).max() < 1e-9                # constants in the null space


def test_morton_order_is_a_permutation_and_local():
    from mlamg import problems
    rs = np.random.RandomState(1)
    pts = rs.rand(5000, 2)
    order = problems.morton_order(pts)
    assert np.array_equal(np.sort(order), np.arange(5000))
    q = pts[order]
    step = np.linalg.norm(np.diff(q, axis=0), axis=1)
    assert np.median(step) < 0.05


def test_random_gnn_outputs_are_reproducible_and_typed():
    from mlamg import problems
    A = oml.poisson((12, 12))
    a = problems.random_gnn_outputs(A, 0.1, seed=0)
    b = problems.random_gnn_outputs(A, 0.1, seed=0)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    top_k, bf, ph = a
    assert len(top_k) == int(np.ceil(0.1 * 144)) and np.all(np.diff(top_k) > 0)
    assert bf.dtype == np.float32 and ph.dtype == np.float32 and len(bf) == A.nnz and bf.min() >= 0 and (bf == 0).any()


def test_grid_container_roundtrip_and_reference_fixture(tmp_path):
    from ns.model.data import Grid
    A = sp.csr_matrix(oml.poisson((5, 4)))
    x = np.random.RandomState(0).rand(20, 2)
    g = Grid(A, x, {"dim": 2})
    f = str(tmp_path / "t")
    g.save(f)
    assert os.path.exists(f + ".grid")
    h = Grid.load(f)
    assert (h.A != A).nnz == 0 and np.array_equal(h.x, x) and h.extra["dim"] == 2 and h.extra["filename"].endswith(".grid")
    g.save(str(tmp_path / "a_second"))
    both = Grid.load_dir(str(tmp_path))
    assert [os.path.basename(x.extra["filename"]) for x in both] == ["a_second.grid", "t.grid"]
    ref = "/root/reference/demos/laplace_3d.grid"             # the reference's only bundled data file (build container only)
    if os.path.exists(ref):
        r = Grid.load(ref)
        assert r.A.shape == (1331, 1331) and r.A.nnz == 17191 and r.x.shape == (1331, 3) and r.extra["dim"] == 3


def test_distributed_delaunay_rows_equal_the_global_mesh():
    """config 4's distributed generator: every rank triangulates its own strip plus a certified halo band (and the thin
    layers along the left/right edges) and must produce exactly its rows of the global P1 Laplacian"""
    import scipy.sparse as sp
    from mlamg import problems
    for seed, npts, world in [(0, 3000, 3), (1, 20000, 4), (2, 60000, 8)]:
        A, offs = problems.delaunay_laplacian_strips(npts, seed, world)
        assert abs(A - A.T).max() <= 1e-12 * abs(A).max()
        for r in range(world):
            info = {}
            rp_, col, val, offs2 = problems.delaunay_laplacian_distributed(npts, seed, world, r, info=info)
            assert np.array_equal(offs, offs2)
            Al = sp.csr_matrix((val, col, rp_), shape=(len(rp_) - 1, A.shape[1]))
            Ar = sp.csr_matrix(A[offs[r]:offs[r + 1]])
            assert np.array_equal(Al.indptr, Ar.indptr) and np.array_equal(Al.indices, Ar.indices), (seed, npts, world, r)
            assert np.abs(Al.data - Ar.data).max() <= 1e-13 * np.abs(Ar.data).max()
            if world >= 4 and npts >= 20000:
                assert info["local_points"] < 0.7 * npts            # nowhere near the whole point set
