"""CPU: the host-side mirrors of /root/reference/ns/model/data.py — `.grid` format, structured generators, the
`graph_from_matrix*` constructors as tensor ops — against files and matrices produced by the UNMODIFIED reference
(tests/golden/make_golden_data.py), against networkx's own edge order, and against known answers."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from helpers import GOLDEN


def test_structured_1d_generators_equal_the_reference_bit_for_bit():
    import ns.model.data as data
    z = np.load(os.path.join(GOLDEN, "ref_grids_1d.npz"))
    for key, fn in (("dirichlet_7", data.Grid.structured_1d_poisson_dirichlet), ("dirichlet_12_shifted", data.Grid.structured_1d_poisson_dirichlet),
                    ("neumann_6", data.Grid.structured_1d_poisson_neumann), ("neumann_9_shifted", data.Grid.structured_1d_poisson_neumann)):
        g = fn(int(z[f"{key}_n"]), tuple(z[f"{key}_xdim"]))
        A = sp.csr_matrix(g.A)
        A.sort_indices()
        assert np.array_equal(A.indptr, z[f"{key}_indptr"]) and np.array_equal(A.indices, z[f"{key}_indices"])
        assert np.array_equal(A.data, z[f"{key}_data"]) and np.array_equal(g.x, z[f"{key}_x"])


def test_grid_file_written_by_the_reference_loads_and_round_trips(tmp_path):
    import ns.model.data as data
    g = data.Grid.load(os.path.join(GOLDEN, "ref_written_by_reference.grid"))
    ref = data.Grid.structured_1d_poisson_neumann(5)
    assert (g.A != ref.A).nnz == 0 and np.array_equal(g.x, ref.x) and g.extra["note"] == "written by the reference"
    g2 = data.Grid.structured_2d_poisson_dirichlet(6, 5, epsilon=0.1, theta=0.4)
    path = str(tmp_path / "mirror_written")
    g2.save(path)
    back = data.Grid.load(path)
    assert (back.A != g2.A).nnz == 0 and np.array_equal(back.x, g2.x) and back.extra["epsilon"] == 0.1
    assert [os.path.basename(x.extra["filename"]) for x in data.Grid.load_dir(str(tmp_path))] == ["mirror_written.grid"]
    if os.path.isdir("/root/reference"):              # build container only: the reference's OWN loader reads the mirror's file
        out = subprocess.run([sys.executable, os.path.join(GOLDEN, "make_golden_data.py"), "--check-load", path + ".grid"],
                             capture_output=True, text=True, timeout=300)
        line = [ln for ln in out.stdout.splitlines() if ln.startswith("LOADED")]
        assert line, out.stderr[-2000:]
        _, n, nnz, total, npts, extra = line[0].split(" ", 5)
        assert int(n) == g2.A.shape[0] and int(nnz) == g2.A.nnz and float(total) == float(abs(g2.A).sum()) and int(npts) == g2.x.shape[0]
        assert "epsilon" in extra and "theta" in extra


def test_structured_2d_generators_known_answers():
    import ns.model.data as data
    g = data.Grid.structured_2d_poisson_dirichlet(7, 7)
    A = g.A.toarray()
    assert A.shape == (49, 49) and np.allclose(A, A.T) and np.allclose(np.diag(A), 4.0)     # right triangles, hx = hy: the 5-point stencil
    off = A - np.diag(np.diag(A))
    assert set(np.round(np.unique(off), 12)) == {-1.0, 0.0} and np.linalg.eigvalsh(A).min() > 0
    assert np.allclose(g.x[0], [1 / 8, 1 / 8]) and np.allclose(g.x[1], [2 / 8, 1 / 8])        # lexicographic, x fastest
    # anisotropy along the axes: eps scales the y part only
    gx = data.Grid.structured_2d_poisson_dirichlet(5, 4, epsilon=0.0).A
    gy = data.Grid.structured_2d_poisson_dirichlet(5, 4, epsilon=1.0).A - gx
    ge = data.Grid.structured_2d_poisson_dirichlet(5, 4, epsilon=0.01).A
    assert abs(ge - (gx + 0.01 * gy)).max() < 1e-13
    # rotation by 90 degrees swaps the roles of x and y
    gr = data.Grid.structured_2d_poisson_dirichlet(5, 4, epsilon=0.01, theta=np.pi / 2).A
    assert abs(gr - (0.01 * gx + gy)).max() < 1e-12
    # the reference maps coordinates by (v + lo) * (hi - lo)
    gs = data.Grid.structured_2d_poisson_dirichlet(3, 3, xdim=(0, 2), ydim=(0, 3))
    assert np.allclose(gs.x.max(axis=0), [1.5, 2.25]) and gs.extra == {'epsilon': 1.0, 'theta': 0.0}
    gn = data.Grid.structured_2d_poisson_neumann(6, 5, epsilon=0.3, theta=0.2)
    assert gn.A.shape == (30, 30) and abs(gn.A @ np.ones(30)).max() < 1e-13 and abs(gn.A - gn.A.T).max() < 1e-13
    assert np.linalg.eigvalsh(gn.A.toarray())[0] > -1e-12


def test_graph_constructors_follow_networkx_edge_order():
    nx = pytest.importorskip("networkx")
    import ns.model.data as data
    rs = np.random.RandomState(0)
    A = sp.random(9, 9, density=0.3, random_state=rs, format="csr") + sp.diags(rs.rand(9) + 1.0)
    A = sp.csr_matrix(A)
    A.data[3] = 0.0                                   # an explicit zero stays an edge
    # the reference: nx.from_scipy_sparse_matrix(A, edge_attribute='weight', create_using=nx.DiGraph) then G.edges
    G = nx.from_scipy_sparse_array(A, edge_attribute="weight", parallel_edges=False, create_using=nx.DiGraph)
    edges = list(G.edges(data="weight"))
    g = data.graph_from_matrix_basic(A, device="cpu")
    assert g.edge_index.shape == (2, len(edges)) and g.edge_attr.shape == (len(edges), 1) and g.edge_attr.dtype == torch.float32
    assert [(int(u), int(v)) for u, v in g.edge_index.t()] == [(u, v) for u, v, _ in edges]
    assert np.array_equal(g.edge_attr[:, 0].numpy(), np.abs(np.array([w for _, _, w in edges], dtype=np.float32)))
    assert torch.equal(g.x, torch.ones(9) / 9) and g.num_nodes == 9
    Agg = sp.csr_matrix((np.ones(9), (np.arange(9), np.arange(9) // 3)), shape=(9, 3))
    g2 = data.graph_from_matrix(A, Agg, device="cpu")
    same = (g2.edge_index[0] // 3) == (g2.edge_index[1] // 3)
    assert g2.edge_attr.shape == (len(edges), 2) and torch.equal(g2.edge_attr[:, 1], (~same).float())
    assert torch.equal(g2.edge_attr[:, 0], g.edge_attr[:, 0])
    xv = torch.arange(9.0)
    g3 = data.graph_from_matrix_node_vals(A, xv, device="cpu")
    assert torch.equal(g3.x, xv) and np.array_equal(g3.edge_attr[:, 0].numpy(), np.array([w for _, _, w in edges], dtype=np.float32))
    g4 = data.graph_from_matrix_node_vals_with_inv(A, xv, device="cpu")
    assert g4.edge_attr.shape == (len(edges), 2)
    assert torch.all(g4.edge_attr[:, 1] == np.float32(1.0 / edges[-1][2]))      # the reference's scalar-overwrite quirk
    assert g.to("cpu").edge_attr.shape == g.edge_attr.shape


def test_lagrange_helpers_and_their_reference_shapes():
    import ns.model.loss as loss
    A = torch.sparse_coo_tensor(torch.tensor([[0, 1, 1], [0, 0, 1]]), torch.tensor([2.0, -1.0, 3.0]), (2, 2))
    B = loss.add_lagrange_rowcols(A).to_dense()
    assert torch.equal(B, torch.tensor([[2.0, 0.0, 1.0], [-1.0, 3.0, 1.0], [1.0, 1.0, 0.0]]))
    v = loss.add_lagrange_vec(torch.ones(2, 3))
    assert v.shape == (3, 3) and torch.equal(v[-1], torch.zeros(3))


def test_topk_vec():
    import ns.model.agg_interp as ai
    x = torch.tensor([[0.3], [0.9], [0.1], [0.5]])
    v = ai.topk_vec(x, 2)
    assert torch.equal(v, torch.tensor([0.0, 1.0, 0.0, 1.0])) and torch.equal(torch.where(v == 1)[0], torch.tensor([1, 3]))
