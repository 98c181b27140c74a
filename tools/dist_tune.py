"""Timing of distributed-cycle variants (development aid). torchrun --nproc-per-node 2 tools/dist_tune.py [n]"""
import os, sys, time, json
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
        return float(t.item())

    for rb in (500000, 1200000 * world):
        H = md.DistHierarchy(rowptr, col, val, comm, ratio=0.027, distance="unit", maxiter=10, rand=0, lam_max=[2.0],
                             max_levels=8, max_coarse=1000, replicate_below=rb)
        for _ in range(100):
            H.vcycle(b, x, 1, 1)
        res = {"replicate_below": rb, "dist_levels": len(H.levels), "tail": [l.A.shape[0] for l in H.tail.levels]}
        H.overlap = True
        res["overlap"] = timed(lambda: H.vcycle(b, x, 1, 1))
        H.overlap = False
        res["no_overlap"] = timed(lambda: H.vcycle(b, x, 1, 1))
        H.overlap = True
        H.tail.use_graph(True)
        res["overlap_tailgraph"] = timed(lambda: H.vcycle(b, x, 1, 1))
        # pieces: exchange alone, level-0 jacobi alone
        L0 = H.levels[0]
        res["exchange_L0"] = timed(lambda: L0.A.plan.exchange(L0.x[0], L0.n))
        res["jacobi_L0_all_rows"] = timed(lambda: L0.A.rowop(3, L0.x[0], L0.x[1], b=b, dw=L0.dw))
        res["jacobi_L0_apply_overlap"] = timed(lambda: L0.A.apply(3, L0.x[0], L0.x[1], b=b, dw=L0.dw, overlap=True, comm_stream=H.comm_stream))
        res["tail_only"] = timed(lambda: H._tail_solve(H._tail_local_b(), 1, 1))
        if len(H.levels) > 1:
            L1 = H.levels[1]
            res["jacobi_L1_apply_overlap"] = timed(lambda: L1.A.apply(3, L1.x[0], L1.x[1], b=L1.b, dw=L1.dw, overlap=True, comm_stream=H.comm_stream))
            res["jacobi_L1_all_rows"] = timed(lambda: L1.A.rowop(3, L1.x[0], L1.x[1], b=L1.b, dw=L1.dw))
        if rank == 0:
            print(json.dumps(res), flush=True)
        del H
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
