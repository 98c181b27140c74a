"""Synthetic problem generators for the named benchmark shapes (host side, numpy/scipy; NOT on the hot path).

The reference builds its datasets with pygmsh meshes and `pyamg.gallery.fem.gradgradform`
(ns/model/data.py:349-394, utils/create_data.py:69-78, demos/voronoi_jump_disc.py:10-22); neither pygmsh nor
pyamg exists here, so the same *shapes* are generated directly:

  * P1 (linear triangle) stiffness matrix  a(u,v) = int kappa grad u . grad v  with one kappa per element
    (evaluated at the centroid, as gradgradform does), Dirichlet boundary rows/columns removed;
  * kappa piecewise constant on the Voronoi cells of Ns in {2,3} seeds, 10^U(-4,4) re-drawn until the
    spread exceeds 1e3 (the reference's 'jump' dataset);
  * rotated-anisotropic diffusion K = Q diag(1, eps) Q^T (the 'anisotropic' dataset, data.py:301-347) and the 1-D
    finite-difference matrices (data.py:244-297);
  * structured-triangle and Delaunay meshes, Morton (Z-curve) ordering so that contiguous row blocks are
    spatially compact (the row partition of the multi-GPU levels);
  * stand-ins for the GNN outputs of FullAggNet.forward (agg_interp.py:459-481): centres = top-k of a random
    score, Bellman-Ford weights and P_hat weights = relu(N(0,1)) in float32 on A's pattern, torch.manual_seed.
"""
import numpy as np
import scipy.sparse as sp


# ------------------------------------------------------------------------------------------ meshes
def structured_triangles(nx, ny):
    """(nx+1) x (ny+1) points on [0,1]^2, two triangles per cell.  -> (pts float64[n,2], tris int64[m,3], boundary mask)"""
    xs, ys = np.linspace(0.0, 1.0, nx + 1), np.linspace(0.0, 1.0, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    pts = np.column_stack([X.ravel(), Y.ravel()])
    idx = np.arange((nx + 1) * (ny + 1), dtype=np.int64).reshape(ny + 1, nx + 1)
    a, b_, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    tris = np.concatenate([np.column_stack([a, b_, d]), np.column_stack([a, d, c])])
    bnd = np.zeros(pts.shape[0], dtype=bool)
    bnd[idx[0, :]] = bnd[idx[-1, :]] = bnd[idx[:, 0]] = bnd[idx[:, -1]] = True
    return pts, tris, bnd


def delaunay_triangles(npts, seed=0):
    """Delaunay triangulation of npts uniform random points in [0,1]^2; the convex-hull vertices are the
    Dirichlet boundary.  -> (pts, tris, boundary mask)"""
    from scipy.spatial import Delaunay
    rs = np.random.RandomState(seed)
    pts = rs.uniform(0.0, 1.0, size=(int(npts), 2))
    tri = Delaunay(pts)
    bnd = np.zeros(pts.shape[0], dtype=bool)
    bnd[np.unique(tri.convex_hull)] = True
    return pts, tri.simplices.astype(np.int64), bnd


def morton_order(pts, bits=16):
    """permutation sorting 2-D points along the Z-curve"""
    q = np.clip((pts * (1 << bits)).astype(np.uint64), 0, (1 << bits) - 1)

    def spread(v):
        v = (v | (v << 16)) & np.uint64(0x0000FFFF0000FFFF)
        v = (v | (v << 8)) & np.uint64(0x00FF00FF00FF00FF)
        v = (v | (v << 4)) & np.uint64(0x0F0F0F0F0F0F0F0F)
        v = (v | (v << 2)) & np.uint64(0x3333333333333333)
        v = (v | (v << 1)) & np.uint64(0x5555555555555555)
        return v
    key = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1))
    return np.argsort(key, kind="stable")


# ------------------------------------------------------------------------------------------ coefficients
def voronoi_jump_kappa(centroids, rng, n_seeds=None):
    """piecewise-constant diffusion on the Voronoi cells of 2 or 3 random seeds (utils/create_data.py:69-78).
    -> (kappa per centroid, jumps array [x, y, d])"""
    ns = int(rng.choice([2, 3])) if n_seeds is None else int(n_seeds)
    S = rng.uniform(0.0, 1.0, (ns, 2))
    while True:
        D = 10.0 ** rng.uniform(-4.0, 4.0, ns)
        if np.ptp(D) > 1e3:
            break
    d2 = ((centroids[:, None, :] - S[None, :, :]) ** 2).sum(axis=2)
    return D[np.argmin(d2, axis=1)], np.column_stack([S, D])


# ------------------------------------------------------------------------------------------ assembly
def p1_stiffness(pts, tris, kappa=None, tensor=None):
    """P1 stiffness matrix (all nodes).  kappa: None (Laplacian) or one scalar per triangle; tensor: optional constant
    symmetric 2x2 diffusion tensor K (a(u,v) = int grad u . K grad v, the rotated-anisotropic shape of
    ns/model/data.py:301-347)."""
    p0, p1, p2 = pts[tris[:, 0]], pts[tris[:, 1]], pts[tris[:, 2]]
    # gradients of the barycentric functions: g_i = rot90(edge opposite to i) / (2 area)
    e0, e1, e2 = p2 - p1, p0 - p2, p1 - p0
    area2 = e2[:, 0] * (-e1[:, 1]) - e2[:, 1] * (-e1[:, 0])          # 2 * signed area = cross(p1-p0, p2-p0)
    scale = (1.0 if kappa is None else np.asarray(kappa, dtype=np.float64)) / (2.0 * np.abs(area2))
    E = [e0, e1, e2]
    if tensor is not None:
        # grad phi_i = rot90(E_i) / (2 area): with G = rot90, g_i . K g_j = E_i . (G^T K G) E_j
        K = np.asarray(tensor, dtype=np.float64)
        G = np.array([[0.0, -1.0], [1.0, 0.0]])
        M = G.T @ K @ G
    rows, cols, vals = [], [], []
    for i in range(3):
        for j in range(3):
            rows.append(tris[:, i])
            cols.append(tris[:, j])
            if tensor is None:
                vals.append(scale * (E[i][:, 0] * E[j][:, 0] + E[i][:, 1] * E[j][:, 1]))
            else:
                Ej = E[j] @ M.T
                vals.append(scale * (E[i][:, 0] * Ej[:, 0] + E[i][:, 1] * Ej[:, 1]))
    n = pts.shape[0]
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def remove_dirichlet(A, pts, bnd):
    """drop boundary rows/columns (R A R^T of data.py:384-390) and exact zeros"""
    keep = np.nonzero(~bnd)[0]
    Ad = sp.csr_matrix(A[keep][:, keep])
    Ad.eliminate_zeros()
    Ad.sort_indices()
    return Ad, pts[keep]


def voronoi_jump_problem(n_side=64, seed=0, mesh="structured", npts=None, n_seeds=None):
    """2-D jump-coefficient diffusion (BASELINE config 3 shape).  mesh: 'structured' (n_side x n_side cells) or
    'delaunay' (npts random points).  Rows in Morton order.  -> (A csr float64, pts, jumps)"""
    rng = np.random.RandomState(seed)
    if mesh == "structured":
        pts, tris, bnd = structured_triangles(n_side, n_side)
    elif mesh == "delaunay":
        pts, tris, bnd = delaunay_triangles(npts or n_side * n_side, seed)
    else:
        raise ValueError(f"unknown mesh {mesh!r}")
    cent = (pts[tris[:, 0]] + pts[tris[:, 1]] + pts[tris[:, 2]]) / 3.0
    kappa, jumps = voronoi_jump_kappa(cent, rng, n_seeds)
    A, p = remove_dirichlet(p1_stiffness(pts, tris, kappa), pts, bnd)
    order = morton_order(p)
    A = sp.csr_matrix(A[order][:, order])
    A.sort_indices()
    return A, p[order], jumps


def rotated_anisotropic_problem(n_side=64, epsilon=1e-2, theta=np.pi / 6, mesh="structured", npts=None, seed=0):
    """2-D rotated-anisotropic diffusion K = Q diag(1, eps) Q^T (the reference's 'anisotropic' dataset,
    ns/model/data.py:301-347, utils/create_data.py:62-68), P1, Dirichlet, Morton-ordered rows.  -> (A csr, pts)"""
    if mesh == "structured":
        pts, tris, bnd = structured_triangles(n_side, n_side)
    else:
        pts, tris, bnd = delaunay_triangles(npts or n_side * n_side, seed)
    c, s_ = np.cos(theta), np.sin(theta)
    Q = np.array([[c, -s_], [s_, c]])
    K = Q @ np.diag([1.0, float(epsilon)]) @ Q.T
    A, p = remove_dirichlet(p1_stiffness(pts, tris, None, tensor=K), pts, bnd)
    order = morton_order(p)
    A = sp.csr_matrix(A[order][:, order])
    A.sort_indices()
    return A, p[order]


def poisson_1d(n, neumann=False, xdim=(0.0, 1.0)):
    """1-D finite-difference Poisson matrices of ns/model/data.py:244-297 (Dirichlet: n interior points; Neumann: n
    points with one-sided end rows), scaled by h^-2.  -> (A csr, x coordinates)"""
    if not neumann:
        x = np.linspace(xdim[0], xdim[1], n + 2)[1:-1]
        h = abs(x[1] - x[0])
        A = (sp.eye(n) * 2 - sp.eye(n, k=-1) - sp.eye(n, k=1)) * (h ** -2.)
        return A.tocsr(), x
    x = np.linspace(xdim[0], xdim[1], n)
    h = abs(x[1] - x[0])
    A = (sp.eye(n) * 2 - sp.eye(n, k=-1) - sp.eye(n, k=1)).tolil()
    A[0, 0] = 1; A[0, 1] = -1
    A[-1, -1] = 1; A[-1, -2] = -1
    return (A.tocsr() * (h ** -2.)).tocsr(), x


def delaunay_laplacian(npts, seed=0):
    """P1 Laplacian on a Delaunay mesh of npts random points, Dirichlet hull, Morton-ordered rows
    (BASELINE config 4 shape: ~7 entries per row).  -> (A csr float64, pts)"""
    pts, tris, bnd = delaunay_triangles(npts, seed)
    A, p = remove_dirichlet(p1_stiffness(pts, tris, None), pts, bnd)
    order = morton_order(p)
    A = sp.csr_matrix(A[order][:, order])
    A.sort_indices()
    return A, p[order]


# ------------------------------------------------------------------------------------------ GNN stand-ins
def random_gnn_outputs(A, alpha=0.1, seed=0):
    """Stand-ins for the three network outputs of FullAggNet.forward with the dtypes and supports the real nets
    produce (agg_interp.py:459-481): top_k = sorted ids of the k = ceil(alpha m) best random scores (int64),
    BF edge weights and P_hat edge weights = relu(N(0,1)) float32 on A's stored pattern (CSR order incl. the
    diagonal, data.py:39-46).  torch.manual_seed(seed) makes them reproducible."""
    import torch
    g = torch.Generator().manual_seed(int(seed))
    m = A.shape[0]
    k = int(np.ceil(alpha * m))
    scores = torch.rand(m, generator=g)
    top_k = torch.sort(torch.topk(scores, k).indices).values
    bf = torch.relu(torch.randn(A.nnz, generator=g)).to(torch.float32)
    ph = torch.relu(torch.randn(A.nnz, generator=g)).to(torch.float32)
    return top_k.numpy().astype(np.int64), bf.numpy(), ph.numpy()


# ------------------------------------------------------------------------------------------ distributed Delaunay (config 4)
def _morton_key(pts, bits=16):
    q = np.clip((pts * (1 << bits)).astype(np.uint64), 0, (1 << bits) - 1)

    def spread(v):
        v = (v | (v << 16)) & np.uint64(0x0000FFFF0000FFFF)
        v = (v | (v << 8)) & np.uint64(0x00FF00FF00FF00FF)
        v = (v | (v << 4)) & np.uint64(0x0F0F0F0F0F0F0F0F)
        v = (v | (v << 2)) & np.uint64(0x3333333333333333)
        v = (v | (v << 1)) & np.uint64(0x5555555555555555)
        return v
    return spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1))


def hull_vertices(pts, directions=32):
    """indices of the convex-hull vertices of a large planar point set: points strictly inside the polygon of the
    extreme points in `directions` directions cannot be hull vertices (Akl-Toussaint), Qhull runs on the rest."""
    from scipy.spatial import ConvexHull
    th = np.linspace(0.0, 2.0 * np.pi, directions, endpoint=False)
    ext = np.unique([int(np.argmax(pts[:, 0] * np.cos(t) + pts[:, 1] * np.sin(t))) for t in th])
    poly = pts[ext]
    c = poly.mean(axis=0)
    poly = poly[np.argsort(np.arctan2(poly[:, 1] - c[1], poly[:, 0] - c[0]))]       # counter-clockwise
    inside = np.ones(pts.shape[0], dtype=bool)
    for a, b in zip(poly, np.roll(poly, -1, axis=0)):
        cross = (b[0] - a[0]) * (pts[:, 1] - a[1]) - (b[1] - a[1]) * (pts[:, 0] - a[0])
        inside &= cross > 1e-14
    cand = np.nonzero(~inside)[0]
    return cand[ConvexHull(pts[cand]).vertices]


def strip_order(pts, world, keep=None):
    """Global numbering of the row partition of config 4: strip r owns y in [r/world, (r+1)/world); inside a strip rows
    follow the Z-curve.  keep: boolean mask of the points that become rows (the others, Dirichlet nodes, get -1).
    -> (new_index int64[npts], offsets int64[world+1])"""
    n = pts.shape[0]
    strip = np.minimum((pts[:, 1] * world).astype(np.int64), world - 1)
    key = _morton_key(pts)
    sel = np.nonzero(keep)[0] if keep is not None else np.arange(n)
    order = sel[np.lexsort((key[sel], strip[sel]))]
    new = np.full(n, -1, dtype=np.int64)
    new[order] = np.arange(order.size)
    offsets = np.concatenate([[0], np.cumsum(np.bincount(strip[sel], minlength=world))]).astype(np.int64)
    return new, offsets


def delaunay_laplacian_strips(npts, seed=0, world=1):
    """GLOBAL reference of the distributed generator below: P1 Laplacian on the Delaunay mesh of npts uniform random
    points, Dirichlet hull removed, rows in strip_order.  -> (A csr, offsets)"""
    pts, tris, bnd = delaunay_triangles(npts, seed)
    new, offsets = strip_order(pts, world, ~bnd)
    A = p1_stiffness(pts, tris, None)
    keep = np.nonzero(~bnd)[0]
    perm = keep[np.argsort(new[keep])]
    Ad = sp.csr_matrix(A[perm][:, perm])
    Ad.eliminate_zeros()
    Ad.sort_indices()
    return Ad, offsets


def delaunay_laplacian_distributed(npts, seed, world, rank, halo_factor=10.0, info=None):
    """Rows of rank `rank` of delaunay_laplacian_strips(npts, seed, world) WITHOUT triangulating the whole point set:
    every rank draws the same points (the RNG stream is cheap), triangulates only its own strip plus a halo band of
    width delta, and keeps the triangles it can certify: a local Delaunay triangle is a triangle of the global mesh
    when the part of its circumdisc that lies inside the domain is inside the band the rank holds all points of.  If a
    triangle touching an owned node cannot be certified the band is doubled and the step repeated.
    -> (rowptr int32, global col int64, val float64, offsets)   rows = owned non-Dirichlet nodes in strip_order"""
    from scipy.spatial import Delaunay
    rs = np.random.RandomState(seed)
    pts = rs.uniform(0.0, 1.0, size=(int(npts), 2))
    bnd = np.zeros(pts.shape[0], dtype=bool)
    bnd[hull_vertices(pts)] = True
    new, offsets = strip_order(pts, world, ~bnd)
    lo, hi = rank / world, (rank + 1) / world
    strip = np.minimum((pts[:, 1] * world).astype(np.int64), world - 1)
    own_mask = strip == rank
    delta = halo_factor / np.sqrt(float(npts))
    while True:
        blo = lo - delta if lo - delta > 0.0 else -np.inf
        bhi = hi + delta if hi + delta < 1.0 else np.inf
        eps = delta
        # the band of the strip plus thin layers along the left / right edges over the full height: the hull has only
        # O(log n) vertices, so the triangles along those edges are long slivers between far-apart hull vertices; their
        # (empty) circumdiscs meet the domain in caps a few 1/(n L) wide that hug the edge
        loc = np.nonzero(((pts[:, 1] >= blo) & (pts[:, 1] <= bhi)) | (pts[:, 0] <= eps) | (pts[:, 0] >= 1.0 - eps))[0]
        p = pts[loc]
        tris = Delaunay(p).simplices.astype(np.int64)
        own_l = own_mask[loc]
        touch = own_l[tris].any(axis=1)
        t = tris[touch]
        a, b, c = p[t[:, 0]], p[t[:, 1]], p[t[:, 2]]
        # circumcentre / radius
        d = 2.0 * (a[:, 0] * (b[:, 1] - c[:, 1]) + b[:, 0] * (c[:, 1] - a[:, 1]) + c[:, 0] * (a[:, 1] - b[:, 1]))
        a2, b2, c2 = (a ** 2).sum(1), (b ** 2).sum(1), (c ** 2).sum(1)
        ux = (a2 * (b[:, 1] - c[:, 1]) + b2 * (c[:, 1] - a[:, 1]) + c2 * (a[:, 1] - b[:, 1])) / d
        uy = (a2 * (c[:, 0] - b[:, 0]) + b2 * (a[:, 0] - c[:, 0]) + c2 * (b[:, 0] - a[:, 0])) / d
        rho = np.sqrt((a[:, 0] - ux) ** 2 + (a[:, 1] - uy) ** 2) * (1.0 + 1e-12)

        def part_covered(limit, above):
            """the part of the disc beyond y = limit (above / below) lies outside 0 <= x <= 1 or inside an edge layer"""
            if not np.isfinite(limit):
                return np.ones(ux.shape, dtype=bool)
            dist = (limit - uy) if above else (uy - limit)             # signed distance from the centre to the line
            empty = dist >= rho                                        # the disc does not reach beyond the line
            wx = np.sqrt(np.maximum(rho ** 2 - np.maximum(dist, 0.0) ** 2, 0.0))
            return empty | (ux + wx <= eps) | (ux - wx >= 1.0 - eps)
        ok = part_covered(bhi, True) & part_covered(blo, False)
        if ok.all() or (blo == -np.inf and bhi == np.inf):
            break
        delta *= 2.0
    if info is not None:
        info.update(local_points=int(loc.size), owned=int(own_mask.sum()), delta=float(delta), triangles=int(t.shape[0]))
    # P1 stiffness of the kept triangles, rows of owned non-Dirichlet nodes, global columns
    A = p1_stiffness(p, t, None)
    g = new[loc]                                                        # local point -> global row id (-1 = Dirichlet)
    rows_l = np.nonzero(own_l & (g >= 0))[0]
    rows_l = rows_l[np.argsort(g[rows_l])]
    Ar = sp.csr_matrix(A[rows_l])
    coo = Ar.tocoo()
    keepc = g[coo.col] >= 0
    n_glob = int(offsets[-1])
    Ag = sp.coo_matrix((coo.data[keepc], (coo.row[keepc], g[coo.col[keepc]])), shape=(rows_l.size, n_glob)).tocsr()
    Ag.eliminate_zeros()
    Ag.sort_indices()
    assert rows_l.size == int(offsets[rank + 1] - offsets[rank])
    return Ag.indptr.astype(np.int32), Ag.indices.astype(np.int64), Ag.data, offsets
