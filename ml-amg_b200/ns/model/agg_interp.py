"""Device-resident post-GNN tail of the reference's `FullAggNet.forward`
(/root/reference/ns/model/agg_interp.py:469-484): network outputs -> Bellman-Ford aggregates -> Agg ->
P = P_hat Agg.  The GNN layers themselves (TAGConv / NNConv message passing, torch_geometric) are out of scope
(SURVEY.md §2.1 row 7); this module is what they feed.

The reference moves the edge weights to the host (`.cpu().numpy()`, :469-470), runs
`pyamg.graph.bellman_ford` (:475), builds Agg with a Python loop (ns/lib/graph.py:56-86) and multiplies with
`torch.sparse.mm` (:484).  Here everything stays in HBM: `mlamg_bellman_ford` (bit-exact emulation of the
sequential sweeps incl. tie-breaking), `mlamg_center_rank_labels`, `mlamg_agg_from_labels`, ordered SpGEMM.

Edge order convention (ns/model/data.py:39-46): the networks emit one value per stored entry of A in CSR
order, diagonal included.
"""
import numpy as np
import torch

import mlamg
from mlamg import core


def topk_vec(x, k):
    """:14-22  indicator vector of the k largest scores (1.0 at the aggregate centres), on x's device"""
    if len(x.shape) != 1:
        x = x.squeeze()
    assert len(x.shape) == 1
    top_k = torch.argsort(x, descending=True)[:k]
    top_k_vec = torch.zeros(x.shape, device=x.device)
    top_k_vec[top_k] = 1.0
    return top_k_vec


def _pattern(A):
    Ad = core.DeviceCSR.wrap(A)
    return Ad


def bellman_ford_aggregates(A, top_k, BF_edges):
    """:469-476.  A: scipy / torch sparse / DeviceCSR (pattern carrier); top_k: sorted centre node ids;
    BF_edges: one non-negative weight per stored entry of A (float32 from the CNet, kept as float32).
    -> (agg_T torch sparse COO n x k float32 on the device, labels int32[n], distance, nearest_center int32[n]).
    Raises KeyError when a node cannot reach any centre (the reference's dict lookup fails on -1,
    graph.py:83; its callers score that as convergence 1.0, utils/common.py:67-70)."""
    core.require_cuda()
    Ad = _pattern(A)
    w = core.as_vec(BF_edges, BF_edges.dtype if isinstance(BF_edges, torch.Tensor) else
                    {np.dtype(np.float32): torch.float32}.get(np.asarray(BF_edges).dtype, torch.float64))
    C = Ad.with_values(w)
    centers = core.as_i32(top_k)
    dist, nearest, _ = core.bellman_ford(C, centers)
    labels = core.center_rank_labels(centers, nearest)
    n, k = Ad.shape[0], int(centers.numel())
    idx = torch.stack([torch.arange(n, device=labels.device), labels.long()])
    agg_T = torch.sparse_coo_tensor(idx, torch.ones(n, device=labels.device), (n, k)).coalesce()
    return agg_T, labels, dist, nearest


def learned_prolongator(A, P_hat_edges, labels, k):
    """:481-484.  P_hat = PNet edge weights on A's pattern; P = P_hat Agg (explicit zeros kept, as
    `torch.sparse.mm(...).coalesce()` keeps them).  -> (P_T torch sparse COO on the device, P DeviceCSR).
    Differentiable in P_hat_edges (a tensor that requires grad), like the reference's `torch.sparse.mm`."""
    core.require_cuda()
    Ad = _pattern(A)
    ph = core.as_vec(P_hat_edges, P_hat_edges.dtype if isinstance(P_hat_edges, torch.Tensor) else
                     {np.dtype(np.float32): torch.float32}.get(np.asarray(P_hat_edges).dtype, torch.float64))
    labels = core.as_i32(labels, ph.device)
    Agg = core.agg_from_labels(labels, int(k), ph.dtype)
    P = mlamg.learned_prolongator(Ad.with_values(ph.detach()), Agg)
    if ph.requires_grad:                       # training: the gradient of P's values flows back onto the PNet's edge outputs
        from mlamg import autograd as ag
        P = P.with_values(ag.agg_product_values(ph, P, Ad, labels))
    return P.to_torch_coo(), P


def forward_tail(A, top_k, BF_edges, P_hat_edges):
    """The whole tail with both network outputs given up front (benchmarks with random-init stand-ins;
    the real PNet sees Agg before it emits P_hat).  -> (agg_T, P_T, labels, P DeviceCSR)"""
    agg_T, labels, _, _ = bellman_ford_aggregates(A, top_k, BF_edges)
    P_T, P = learned_prolongator(A, P_hat_edges, labels, len(top_k))
    return agg_T, P_T, labels, P
