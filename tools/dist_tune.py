"""Timing of distributed-cycle variants (development aid).

    torchrun --nproc-per-node 2 tools/dist_tune.py [n]

Prints one JSON line (rank 0): ms per V(1,1) cycle for the peer-window and NCCL halo transports, eager and
replayed from a CUDA graph, with and without the interior/boundary row split, plus the cost of single pieces
(one exchange, one fine sweep with and without its exchange, the replicated tail)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ml-amg_b200")]
import numpy as np
import torch
import torch.distributed as dist


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import mlamg
    from mlamg import distributed as md
    comm = md.Comm()
    rowptr, col, val = md.poisson_slab(n, world, rank)
    b = torch.from_numpy(np.random.RandomState(rank).randn(n ** 3)).cuda()
    x = torch.empty_like(b)

    def timed(fn, k=20):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record(); torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / k], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return round(float(t.item()), 4)

    lam0 = 1.0 + (2.0 * np.cos(np.pi / (n + 1)) + np.cos(np.pi / (n * world + 1))) / 3.0
    H = md.DistHierarchy(rowptr, col, val, comm, ratio=0.027, distance="unit", maxiter=10, rand=0, lam_max=[lam0],
                         max_levels=8, max_coarse=1000, replicate_below=500000)
    res = {"n": n, "world": world, "dist_levels": len(H.levels), "tail": [l.A.shape[0] for l in H.tail.levels]}
    for _ in range(100):
        H.vcycle(b, x, 1, 1)
    def graph_timed(fn, reps=10):
        """GPU-side cost of fn: `reps` back-to-back calls captured in one CUDA graph (no CPU launch overhead)"""
        fn(); torch.cuda.synchronize(); dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            for _ in range(reps):
                fn()
        return round(timed(g.replay, k=10) / reps, 4)

    res["peer_eager"] = timed(lambda: H.vcycle(b, x, 1, 1))
    res["peer_graph"] = timed(H.capture(b, x, 1, 1))
    md.PEER_INPLACE = False
    res["peer_graph_unpack"] = timed(H.capture(b, x, 1, 1))
    md.PEER_INPLACE = True
    H.overlap = False
    res["peer_graph_nosplit"] = timed(H.capture(b, x, 1, 1))
    H.overlap = True
    H._graph = None
    L0 = H.levels[0]
    cs = H._channels(1, 1)
    ch = cs[(0, "post", 0)]
    halo = L0.x[0][L0.n:]

    def exch():
        ch.push(L0.x[0]); ch.unpack(halo)
    res["g_peer_exchange_L0"] = graph_timed(exch)
    res["g_jacobi_L0_all_rows"] = graph_timed(lambda: L0.A.rowop(3, L0.x[0], L0.x[1], b=b, dw=L0.dw))
    res["g_jacobi_L0_peer_apply"] = graph_timed(lambda: L0.A.apply(3, L0.x[0], L0.x[1], b=b, dw=L0.dw, chan=ch))
    res["g_jacobi_L0_interior_only"] = graph_timed(lambda: L0.A.rowop(3, L0.x[0], L0.x[1], b=b, dw=L0.dw, row_range=L0.A.interior_range))
    res["g_jacobi_L0_boundary_list"] = graph_timed(lambda: L0.A.rowop(3, L0.x[0], L0.x[1], b=b, dw=L0.dw, rows=L0.A.boundary))
    tail_ch = cs["tail"]
    res["g_tail_only_peer"] = graph_timed(lambda: H._tail_solve(H._tail_local_b(), 1, 1, tail_ch))
    if len(H.levels) > 1:
        L1 = H.levels[1]
        ch1 = cs[(1, "post", 0)]
        res["g_jacobi_L1_all_rows"] = graph_timed(lambda: L1.A.rowop(3, L1.x[0], L1.x[1], b=L1.b, dw=L1.dw))
        res["g_jacobi_L1_peer_apply"] = graph_timed(lambda: L1.A.apply(3, L1.x[0], L1.x[1], b=L1.b, dw=L1.dw, chan=ch1))
        H.overlap = False
        res["g_jacobi_L1_peer_apply_nosplit"] = graph_timed(lambda: L1.A.apply(3, L1.x[0], L1.x[1], b=L1.b, dw=L1.dw, overlap=False, chan=ch1))
        H.overlap = True
        res["L1"] = {"rows": L1.n, "nnz": L1.A.csr.nnz, "halo": L1.A.plan.n_halo, "boundary_rows": int(L1.A.boundary.numel()),
                     "P_nnz": L0.P.csr.nnz, "P_halo": L0.P.plan.n_halo, "R_halo": L0.R.plan.n_halo}
    H.halo = "nccl"
    res["nccl_eager"] = timed(lambda: H.vcycle(b, x, 1, 1))
    H.halo = "peer"
    H.check_exchange()
    if rank == 0:
        print(json.dumps(res), flush=True)
    H.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
