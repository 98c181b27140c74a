"""GPU: Bellman-Ford / Lloyd aggregation must reproduce the sequential reference loops BIT-EXACTLY
(distances, labels incl. tie-breaking, moved seeds), on unit-weight grids (every node is a tie),
random weights, GNN-like fp32 weights with exact zeros, disconnected graphs and the golden inputs."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from helpers import GOLDEN_CASES, GOLDEN, load_golden, csr_from, grid_graph
from oracle import pyamg_restated as pr, reference_path as rp, multilevel as oml

pytestmark = pytest.mark.gpu

GRAPHS = [
    ("grid2d_unit", lambda: grid_graph((31, 23), "unit"), 0.1),
    ("grid3d_unit", lambda: grid_graph((11, 9, 8), "unit"), 0.03),
    ("grid2d_random", lambda: grid_graph((25, 25), "random", seed=1), 0.08),
    ("grid2d_random_nonsym", lambda: grid_graph((25, 20), "random", seed=2, symmetric=False), 0.08),
    ("grid2d_relu_f32", lambda: grid_graph((24, 24), "relu", seed=3, dtype=np.float32), 0.1),
    ("grid3d_relu_f32", lambda: grid_graph((9, 9, 9), "relu", seed=4, dtype=np.float32), 0.05),
    ("chain_unit", lambda: grid_graph((400,), "unit"), 0.02),
]


def to_dev(G):
    import mlamg
    return mlamg.DeviceCSR.from_arrays(G.indptr, G.indices, G.data, G.shape)


@pytest.mark.parametrize("name,make,ratio", GRAPHS, ids=[g[0] for g in GRAPHS])
def test_bellman_ford_bit_exact(name, make, ratio):
    import mlamg
    G = make()
    N = G.shape[0]
    for seed in (0, 1):
        seeds = np.sort(np.random.RandomState(seed).permutation(N)[:max(1, int(np.ceil(ratio * N)))]).astype(np.int32)
        d_ref, z_ref = pr.bellman_ford(G, seeds)
        d, z, sweeps = mlamg.bellman_ford(to_dev(G), seeds)
        assert np.array_equal(d.cpu().numpy(), d_ref), "distances differ"
        assert np.array_equal(z.cpu().numpy(), z_ref), "nearest-seed labels (tie-breaking) differ"
        assert d.cpu().numpy().dtype == G.dtype
        assert sweeps >= 1


@pytest.mark.parametrize("name,make,ratio", GRAPHS, ids=[g[0] for g in GRAPHS])
def test_lloyd_cluster_bit_exact(name, make, ratio):
    import mlamg
    G = make()
    N = G.shape[0]
    seeds = np.random.RandomState(0).permutation(N)[:max(1, int(np.ceil(ratio * N)))].astype(np.int32)
    for maxiter in (1, 10):
        d_ref, w_ref, s_ref = pr.lloyd_cluster(G, seeds.copy(), maxiter)
        d, w, s, iters = mlamg.lloyd_cluster(to_dev(G), seeds.copy(), maxiter)
        assert np.array_equal(w.cpu().numpy(), w_ref), "cluster labels differ"
        assert np.array_equal(s.cpu().numpy(), s_ref), "moved seeds differ"
        assert np.array_equal(d.cpu().numpy(), d_ref), "inward distances differ"
    # determinism: run twice, bitwise equal
    a = mlamg.lloyd_cluster(to_dev(G), seeds.copy(), 10)
    b = mlamg.lloyd_cluster(to_dev(G), seeds.copy(), 10)
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[0], b[0])


def test_unreachable_component():
    import mlamg
    A = sp.block_diag([oml.poisson((6, 5)), oml.poisson((4, 4))]).tocsr()
    G = sp.csr_matrix((np.ones(A.nnz), A.indices, A.indptr), shape=A.shape)
    seeds = np.array([3, 17], dtype=np.int32)
    d_ref, z_ref = pr.bellman_ford(G, seeds)
    d, z, _ = mlamg.bellman_ford(to_dev(G), seeds)
    assert np.array_equal(z.cpu().numpy(), z_ref) and np.array_equal(d.cpu().numpy(), d_ref)
    assert (z_ref[30:] == -1).all()
    _, w_ref, s_ref = pr.lloyd_cluster(G, seeds.copy(), 10)
    _, w, s, _ = mlamg.lloyd_cluster(to_dev(G), seeds.copy(), 10)
    assert np.array_equal(w.cpu().numpy(), w_ref) and np.array_equal(s.cpu().numpy(), s_ref)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_lloyd_aggregation_mirror_matches_reference_golden(name):
    """ns.lib.graph.lloyd_aggregation (GPU) == the unmodified reference run (golden): AggOp, roots, seeds."""
    import ns.lib.graph as g
    z = load_golden(name)
    C = csr_from(z, "C")
    Agg, roots, seeds = g.lloyd_aggregation(C, ratio=float(z["ratio"]), distance=str(z["distance"]), rand=int(z["rand"]))
    ref = csr_from(z, "Agg")
    assert Agg.dtype == np.int8 and sp.isspmatrix_csr(Agg) and Agg.shape == ref.shape
    assert np.array_equal(Agg.indptr, ref.indptr) and np.array_equal(Agg.indices, ref.indices)
    assert np.array_equal(roots, z["roots"]) and np.array_equal(seeds, z["seeds"])


def test_lloyd_aggregation_csc_input_and_modes():
    import ns.lib.graph as g
    A = oml.poisson((14, 12))
    C = sp.csr_matrix((np.random.RandomState(0).rand(A.nnz) + 0.1, A.indices, A.indptr), shape=A.shape)
    for mat in (C, C.tocsc()):
        for mode in ("unit", "abs", "inv", "same", "min"):
            ref = rp.lloyd_aggregation(mat, ratio=0.1, distance=mode, rand=2)
            got = g.lloyd_aggregation(mat, ratio=0.1, distance=mode, rand=2)
            assert (ref[0] != got[0]).nnz == 0 and np.array_equal(ref[1], got[1]) and np.array_equal(ref[2], got[2])


def test_modified_bellman_ford_and_agg_golden():
    """reference-owned push Bellman-Ford (graph.py:7-53) and nearest_center_to_agg (graph.py:56-86)"""
    import ns.lib.graph as g
    z = np.load(GOLDEN + "/ref_modified_bf.npz")
    n = int(z["n"])
    S_T = torch.sparse_coo_tensor(np.vstack([z["row"], z["col"]]), torch.from_numpy(z["w"]), (n, n)).coalesce()
    centers = torch.from_numpy(z["centers"])
    dist, near = g.modified_bellman_ford(S_T, centers)
    assert dist.dtype == torch.float32 and near.dtype == torch.int64
    assert np.array_equal(dist.numpy(), z["dist"]) and np.array_equal(near.numpy(), z["nearest"])
    agg = g.nearest_center_to_agg(centers, near).coalesce()
    assert np.array_equal(agg.indices()[0].numpy(), z["agg_row"]) and np.array_equal(agg.indices()[1].numpy(), z["agg_col"])
    assert np.array_equal(agg.values().numpy(), z["agg_val"])
    with pytest.raises(KeyError):
        g.nearest_center_to_agg(centers, torch.full((n,), -1, dtype=torch.int64))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_modified_bellman_ford_random(seed):
    import mlamg
    rs = np.random.RandomState(seed)
    A = oml.poisson((17, 13)).tocoo()
    w = np.maximum(rs.randn(A.nnz), 0).astype(np.float32) + (rs.rand(A.nnz) < 0.7) * 0.01
    w = w.astype(np.float32)
    S = sp.coo_matrix((w, (A.row, A.col)), shape=A.shape)
    centers = np.sort(rs.permutation(A.shape[0])[:9])
    d_ref, z_ref = rp.modified_bellman_ford(S, centers)
    Sd = mlamg.DeviceCSR.from_scipy(S.tocsr().astype(np.float32))
    d, zl, passes = mlamg.modified_bellman_ford(Sd, centers)
    assert np.array_equal(d.cpu().numpy(), d_ref) and np.array_equal(zl.cpu().numpy(), z_ref)
