"""Drop-in mirror of the reference's `ns.lib.graph` (same names, argument order, defaults, return
tuples and errors) with the loops executed by the sm_100a kernels in libmlamg_b200.so.

Reference: /root/reference/ns/lib/graph.py
  modified_bellman_ford   :7-53    -> mlamg_modified_bellman_ford
  nearest_center_to_agg   :56-86   -> mlamg_center_rank_labels (+ COO assembly)
  lloyd_aggregation       :156-239 -> mlamg_lloyd_cluster (pyamg.graph.lloyd_cluster at :232)
`num_connected_components` / `check_aggregates_connected` (:89-153) are debug helpers that nothing
calls; they are out of scope.
"""
import numpy as np
import scipy.sparse as sp
import torch

import mlamg
from mlamg import core


def modified_bellman_ford(S_T, centers):
    """S_T: torch sparse COO strength matrix; centers: 1-D integer tensor.
    -> (distance float32[n], nearest_center int64[n]) on centers.device."""
    core.require_cuda()
    S = core.DeviceCSR.from_torch(S_T, torch.float32)
    dist, near, _ = core.modified_bellman_ford(S, centers.to("cuda"))
    return dist.to(centers.device), near.to(centers.device)


def nearest_center_to_agg(top_k, nearest_center):
    """-> torch sparse COO (n x m) float32 aggregate assignment, coalesced, on top_k.device.
    Raises KeyError when a node's nearest centre is not in top_k (e.g. -1: unreachable)."""
    core.require_cuda()
    n = len(nearest_center)
    m = len(top_k)
    labels = core.center_rank_labels(core.as_i32(top_k), core.as_i32(nearest_center))
    idx = torch.stack([torch.arange(n, device=labels.device), labels.long()])
    return torch.sparse_coo_tensor(idx, torch.ones(n, device=labels.device), (n, m)).coalesce().to(top_k.device)


def lloyd_aggregation(C, ratio=0.03, distance='unit', maxiter=10, rand=None):
    """Aggregate nodes using Lloyd clustering (reference docstring: graph.py:157-193).

    Returns (AggOp csr int8 N x num_seeds, roots, seeds) exactly like the reference."""
    if ratio <= 0 or ratio > 1:
        raise ValueError('ratio must be > 0.0 and <= 1.0')
    if not (sp.isspmatrix_csr(C) or sp.isspmatrix_csc(C)):
        raise TypeError('expected csr_matrix or csc_matrix')
    if distance not in ('unit', 'abs', 'inv', 'same', 'min'):
        raise ValueError(f'Unrecognized value distance={distance}')
    if rand is not None and not isinstance(rand, (int, np.integer, np.random.RandomState)):
        raise TypeError('rand should be an integer seed value or a random state')
    core.require_cuda()
    data = C.data
    if distance == 'unit':
        data = np.ones_like(np.real(data)).astype(float)     # reference: float64 unit lengths
        distance_dev = 'same'
    elif np.iscomplexobj(data):
        # reference order (:201-223): the transform acts on the complex entries (abs = modulus), the real part is
        # taken afterwards
        data = {'abs': lambda d: abs(d), 'inv': lambda d: 1.0 / abs(d), 'same': lambda d: d,
                'min': lambda d: d - d.min()}[distance](data)
        data = np.ascontiguousarray(np.real(data))
        distance_dev = 'same'
    else:
        distance_dev = distance
    # CSC arrays are used as stored (pyamg asgraph keeps CSC: the transposed graph); no sorting
    G = core.DeviceCSR.from_arrays(C.indptr, C.indices, data, C.shape)
    labels, num_seeds, roots, seeds = mlamg.lloyd_labels(G, ratio=ratio, distance=distance_dev, maxiter=maxiter,
                                                         rand=rand)
    clusters = labels.cpu().numpy()
    row = (clusters >= 0).nonzero()[0]
    col = clusters[row]
    ones = np.ones(len(row), dtype='int8')
    AggOp = sp.coo_matrix((ones, (row, col)), shape=(C.shape[0], num_seeds)).tocsr()
    return AggOp, roots.cpu().numpy(), seeds
