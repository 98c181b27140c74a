"""`.grid` files of the reference (bz2-pickled dict, /root/reference/ns/model/data.py:208-235) straight onto
the device, and the matrix -> edge-list convention its networks use (:22-46).

Mirrored: the container and its file format, the structured generators (`Grid.structured_1d_poisson_*`,
`Grid.structured_2d_poisson_*`: same names and arguments; the 2-D ones assemble P1 elements with `mlamg.problems`
instead of `pyamg.gallery.fem`, which is absent) and the `graph_from_matrix*` constructors as pure tensor ops
(the reference goes through networkx and torch_geometric, both out of reach of the device).
Out of scope: the pygmsh-based unstructured generators (`mlamg.problems` generates those shapes from Delaunay
meshes) and the plotting helpers.
"""
import bz2
import os
import pickle

import numpy as np
import scipy.sparse as sp
import torch

from mlamg import core


class Grid:
    """A, node coordinates and free-form extras — the reference's dataset item (data.py:65-84)."""

    def __init__(self, A_csr, x=None, extra=None):
        self.A = A_csr.tocsr()
        self.x = x
        self.extra = extra if extra is not None else {}
        self._device = None

    @property
    def networkx(self):
        raise NotImplementedError("networkx view is not needed by the device path")

    def device_csr(self, dtype=None):
        """A resident in HBM (cached)"""
        if self._device is None or (dtype is not None and self._device.dtype != dtype):
            self._device = core.DeviceCSR.from_scipy(self.A, dtype)
        return self._device

    def save(self, fname):
        """same on-disk layout as data.py:207-219: A as the (data, indices, indptr) tuple"""
        if '.grid' not in fname:
            fname = fname + '.grid'
        A = self.A.tocsr()
        with bz2.BZ2File(fname, "wb") as f:
            pickle.dump({'A': (A.data, A.indices, A.indptr), 'x': self.x, 'extra': self.extra}, f)

    @staticmethod
    def load(fname):
        """data.py:221-234"""
        if '.grid' not in fname:
            fname = fname + '.grid'
        with bz2.BZ2File(fname, "rb") as f:
            loaded = pickle.load(f)
        extra = loaded['extra'] if 'extra' in loaded else {}
        extra['filename'] = fname
        A = loaded['A']
        if isinstance(A, tuple):
            A = sp.csr_matrix(A)
        return Grid(A, loaded['x'], extra)


def load_dir(directory):
    """every `.grid` file of a directory (data.py:236-242), sorted by name so that dataset order is reproducible"""
    return [Grid.load(os.path.join(directory, f)) for f in sorted(os.listdir(directory)) if '.grid' in f.lower()]


Grid.load_dir = staticmethod(load_dir)


def edge_list(A, with_values=True):
    """edge_index (2 x nnz, CSR order incl. the diagonal) and edge_attr = |A_ij| float32 — the seam between A and
    the networks' per-edge outputs (graph_from_matrix_basic, data.py:39-46), as device tensors."""
    Ad = core.DeviceCSR.wrap(A)
    n = Ad.shape[0]
    rows = torch.repeat_interleave(torch.arange(n, device=Ad.col.device), (Ad.rowptr[1:] - Ad.rowptr[:-1]).long())
    edge_index = torch.stack([rows, Ad.col.long()])
    if not with_values:
        return edge_index
    return edge_index, Ad.val.abs().to(torch.float32)


# ------------------------------------------------------------------------------------------ generators (data.py:244-297, 436-560)
def _structured_1d_poisson_dirichlet(n, xdim=(0, 1)):
    """data.py:244-267: n interior points, finite differences, scaled by h^-2"""
    from mlamg import problems
    A, x = problems.poisson_1d(n, neumann=False, xdim=xdim)
    return Grid(A, np.column_stack((x, np.zeros_like(x))))


def _structured_1d_poisson_neumann(n, xdim=(0, 1)):
    """data.py:269-297: n points, one-sided end rows"""
    from mlamg import problems
    A, x = problems.poisson_1d(n, neumann=True, xdim=xdim)
    return Grid(A, np.column_stack((x, np.zeros_like(x))))


def _rotated_tensor(epsilon, theta):
    c, s_ = np.cos(theta), np.sin(theta)
    Q = np.array([[c, -s_], [s_, c]])
    return Q @ np.diag([1., epsilon]) @ Q.T


def _structured_2d_poisson_dirichlet(n_pts_x, n_pts_y, xdim=(0, 1), ydim=(0, 1), epsilon=1.0, theta=0.0):
    """data.py:436-497: n_pts_x x n_pts_y interior points of a structured triangle mesh, P1, diffusion tensor
    Q diag(1, eps) Q^T, Dirichlet rows / columns removed; rows in lexicographic vertex order (x fastest) as there.
    The coordinate map is the reference's own `(v + lo) * (hi - lo)`."""
    from mlamg import problems
    pts, tris, bnd = problems.structured_triangles(n_pts_x + 1, n_pts_y + 1)
    v = pts.copy()
    v[:, 0] = (v[:, 0] + xdim[0]) * (xdim[1] - xdim[0])
    v[:, 1] = (v[:, 1] + ydim[0]) * (ydim[1] - ydim[0])
    A = problems.p1_stiffness(v, tris, None, tensor=_rotated_tensor(epsilon, theta))
    A_d, x = problems.remove_dirichlet(A, v, bnd)
    return Grid(A_d, x, {'epsilon': epsilon, 'theta': theta})


def _structured_2d_poisson_neumann(n_pts_x, n_pts_y, xdim=(0, 1), ydim=(0, 1), epsilon=1.0, theta=0.0):
    """data.py:499-543: all n_pts_x x n_pts_y points kept (homogeneous Neumann: singular operator)"""
    from mlamg import problems
    pts, tris, _ = problems.structured_triangles(n_pts_x - 1, n_pts_y - 1)
    v = pts.copy()
    v[:, 0] = (v[:, 0] + xdim[0]) * (xdim[1] - xdim[0])
    v[:, 1] = (v[:, 1] + ydim[0]) * (ydim[1] - ydim[0])
    return Grid(problems.p1_stiffness(v, tris, None, tensor=_rotated_tensor(epsilon, theta)), v)


Grid.structured_1d_poisson_dirichlet = staticmethod(_structured_1d_poisson_dirichlet)
Grid.structured_1d_poisson_neumann = staticmethod(_structured_1d_poisson_neumann)
Grid.structured_2d_poisson_dirichlet = staticmethod(_structured_2d_poisson_dirichlet)
Grid.structured_2d_poisson_neumann = staticmethod(_structured_2d_poisson_neumann)


# ------------------------------------------------------------------------------------------ graph constructors (data.py:22-63)
class GraphData:
    """The three tensors the reference's networks read from a `torch_geometric.data.Data` (x, edge_index, edge_attr);
    torch_geometric is absent, attribute names are kept."""

    def __init__(self, x, edge_index, edge_attr):
        self.x, self.edge_index, self.edge_attr = x, edge_index, edge_attr

    def to(self, device):
        return GraphData(self.x.to(device), self.edge_index.to(device), self.edge_attr.to(device))

    @property
    def num_nodes(self):
        return int(self.x.shape[0])


def _default_device(device):
    return torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")


def _edges(A, device):
    """edges in the order networkx + `from_networkx` produce them: CSR storage order, stored diagonal (self loops)
    and explicit zeros included"""
    A = sp.csr_matrix(A)
    rows = np.repeat(np.arange(A.shape[0], dtype=np.int64), np.diff(A.indptr))
    edge_index = torch.from_numpy(np.vstack([rows, A.indices.astype(np.int64)])).to(device)
    weight = torch.from_numpy(np.asarray(A.data, dtype=np.float64)).to(device)
    return edge_index, weight


def graph_from_matrix_basic(A, device=None):
    """data.py:39-46: x = 1/n per node, edge_attr = |a_ij| float32 of shape (E, 1)"""
    device = _default_device(device)
    n = A.shape[0]
    edge_index, w = _edges(A, device)
    return GraphData(torch.ones(n, device=device) / n, edge_index, w.float().abs().view(-1, 1))


def graph_from_matrix(A, agg_op, device=None):
    """data.py:22-37: edge_attr = |[a_ij, cluster_adj_ij]| float32 (E, 2), cluster_adj = 0 inside an aggregate, 1 across"""
    device = _default_device(device)
    n = A.shape[0]
    edge_index, w = _edges(A, device)
    clusters = torch.from_numpy(np.array(sp.csr_matrix(agg_op).argmax(axis=1)).flatten().astype(np.int64)).to(device)
    adj = (clusters[edge_index[0]] != clusters[edge_index[1]]).to(torch.float64)
    return GraphData(torch.ones(n, device=device) / n, edge_index, torch.stack([w, adj], dim=1).float().abs())


def graph_from_matrix_node_vals(A, x, device=None):
    """data.py:48-51: caller-supplied node features, signed weights (E, 1)"""
    device = _default_device(device)
    edge_index, w = _edges(A, device)
    return GraphData(x.to(device) if isinstance(x, torch.Tensor) else torch.as_tensor(x, device=device), edge_index,
                     w.float().view(-1, 1))


def graph_from_matrix_node_vals_with_inv(A, x, device=None):
    """data.py:53-63.  The reference's loop overwrites its dict with a scalar, so EVERY edge receives 1 / (weight of the
    last edge); that is what its networks were trained on and what is reproduced here."""
    device = _default_device(device)
    edge_index, w = _edges(A, device)
    inv = torch.full_like(w, 1.0 / float(w[-1])) if w.numel() else w
    return GraphData(x.to(device) if isinstance(x, torch.Tensor) else torch.as_tensor(x, device=device), edge_index,
                     torch.stack([w, inv], dim=1).float())
