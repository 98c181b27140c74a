"""A/B timing of the distributed cycle's layout choices on N ranks (development aid).

    torchrun --nproc-per-node N tools/dist_scale_tune.py [n]

One JSON line per variant (rank 0): geometry (slab / cube), boundary rows on the side stream or behind the interior rows,
the replication threshold of the coarse tail; setup seconds and ms per V(1,1) cycle replayed from a CUDA graph."""
import json
import os
import sys
import time

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
    from mlamg import distributed as md
    comm = md.Comm()

    def timed(fn, k=30):
        for _ in range(10):
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

    variants = [("cube", 500000), ("slab", 500000), ("cube", 20000), ("cube", 3000000)]
    if len(sys.argv) > 2:
        variants = [(g, int(r)) for g, r in (v.split(":") for v in sys.argv[2:])]
    for geometry, rb in variants:
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.time()
        rowptr, col, val = md.poisson_slab(n, world, rank, geometry=geometry)
        H = md.DistHierarchy(rowptr, col, val, comm, ratio=0.027, distance="unit", maxiter=10, rand=0,
                             lam_max=[md.slab_lambda_max(n, world, geometry)], max_levels=8, max_coarse=1000, replicate_below=rb)
        torch.cuda.synchronize(); dist.barrier()
        setup_s = time.time() - t0
        b = torch.from_numpy(np.random.RandomState(rank).randn(n ** 3)).cuda()
        x = torch.empty_like(b)
        res = {"n": n, "world": world, "geometry": geometry, "replicate_below": rb, "setup_s": round(setup_s, 2),
               "dist_levels": [int(o[-1]) for o in H.offsets[:-1]], "tail": [l.A.shape[0] for l in H.tail.levels],
               "halo_fine": H.levels[0].A.plan.n_halo,
               "setup_stages_s": {k: round(v, 3) for k, v in H.setup_profile.acc.items()}}
        for side, key in ((1, "ms_boundary_side"), (0, "ms_boundary_after"), (2, "ms_boundary_tail_stream")):
            md.PEER_BOUNDARY_SIDE = side
            replay = H.capture(b, x, 1, 1)
            res[key] = timed(replay)
            H._graph = None
        md.PEER_BOUNDARY_SIDE = 1
        md.EARLY_PUSH = False
        replay = H.capture(b, x, 1, 1)
        res["ms_boundary_side_no_early_push"] = timed(replay)
        H._graph = None
        md.EARLY_PUSH = True
        res["w32_levels"] = [sorted(getattr(L.Q, "w32", {})) + sorted("A:" + k for k in getattr(L.A, "w32", {})) for L in H.levels]
        res["early_push_flags"] = [{k: bool(getattr(L, k, False)) for k in ("early_R", "early_next_res", "early_up_Q", "early_up_A")}
                                   for L in H.levels]
        H.check_exchange()
        if rank == 0:
            print(json.dumps(res), flush=True)
        H.close()
        del H, rowptr, col, val
        torch.cuda.empty_cache()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
