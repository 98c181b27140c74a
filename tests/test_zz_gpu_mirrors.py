"""GPU: small mirrors added for API completeness — `ns.lib.multigrid.gauss_seidel_torch` (reference :93-99) against the
dense triangular-solve formula it states, and the `graph_from_matrix*` constructors producing device tensors."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import multilevel as oml

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-5)])
def test_gauss_seidel_torch_vs_dense_triangular_solves(dtype, tol):
    import ns.lib.multigrid as mg
    rs = np.random.RandomState(0)
    A = sp.csr_matrix(oml.poisson((9, 8))).astype(np.float64)
    A.data = A.data * (1.0 + 0.2 * rs.rand(A.nnz))                  # non-symmetric values
    Ad = torch.from_numpy(A.toarray()).to(dtype)
    b, x = torch.from_numpy(rs.randn(72)).to(dtype), torch.from_numpy(rs.randn(72)).to(dtype)
    U = torch.triu(Ad, 1)
    ref = x.clone()
    for _ in range(3):
        ref = torch.linalg.solve_triangular(Ad, (b - U @ ref).unsqueeze(1), upper=False).squeeze(1)
    x0 = x.clone()
    for kw in (dict(), dict(U=U), dict(U=U.to_sparse())):
        got = mg.gauss_seidel_torch(Ad, b, x, nu=3, **kw)
        assert got.dtype == dtype and got.device == x.device and torch.equal(x, x0)        # new tensor, caller's x untouched
        assert (got - ref).abs().max().item() <= tol * ref.abs().max().item()
    got = mg.gauss_seidel_torch(Ad.to_sparse(), b.cuda(), x.cuda(), nu=3)
    assert got.is_cuda and (got.cpu() - ref).abs().max().item() <= tol * ref.abs().max().item()
    with pytest.raises(ValueError):
        mg.gauss_seidel_torch(Ad, b, x, U=Ad)


def test_graph_constructors_on_the_device():
    import ns.model.data as data
    A = sp.csr_matrix(oml.poisson((6, 5)))
    g = data.graph_from_matrix_basic(A)
    ei, ea = data.edge_list(A)
    assert g.edge_index.is_cuda and torch.equal(g.edge_index, ei) and torch.equal(g.edge_attr[:, 0], ea)
    Agg = sp.csr_matrix((np.ones(30), (np.arange(30), np.arange(30) // 6)), shape=(30, 5))
    g2 = data.graph_from_matrix(A, Agg)
    assert g2.edge_attr.shape == (A.nnz, 2) and g2.edge_attr.is_cuda
    assert int(g2.edge_attr[:, 1].sum().item()) == int(((ei[0] // 6) != (ei[1] // 6)).sum().item())
