"""Per-kernel timeline of ONE distributed V-cycle on rank 0 (development aid; kineto trace, nothing is replayed).

    torchrun --nproc-per-node 2 tools/dist_trace.py [n] [cube|slab] [replicate_below]

Prints, in launch order, every kernel of one CUDA-graph replay of the cycle with its duration — the multi-GPU
counterpart of the ncu launch list of the single-GPU cycle (ncu must not wrap multi-rank runs with spinning kernels)."""
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
    from mlamg import distributed as md
    comm = md.Comm()
    geometry = sys.argv[2] if len(sys.argv) > 2 else "cube"
    rb = int(sys.argv[3]) if len(sys.argv) > 3 else 500000
    rowptr, col, val = md.poisson_slab(n, world, rank, geometry=geometry)
    b = torch.from_numpy(np.random.RandomState(rank).randn(n ** 3)).cuda()
    x = torch.empty_like(b)
    lam0 = md.slab_lambda_max(n, world, geometry)
    H = md.DistHierarchy(rowptr, col, val, comm, ratio=0.027, distance="unit", maxiter=10, rand=0, lam_max=[lam0],
                         max_levels=8, max_coarse=1000, replicate_below=rb)
    replay = H.capture(b, x, 1, 1)
    for _ in range(50):
        replay()
    torch.cuda.synchronize(); dist.barrier()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            replay()
        torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        evs = [e for e in prof.events() if str(e.device_type).endswith("CUDA") and e.name and "Memcpy" not in e.name[:0]]
        evs.sort(key=lambda e: e.time_range.start)
        per = len(evs) // 3
        last = evs[-per:] if per else evs
        t0 = last[0].time_range.start
        rows = []
        for e in last:
            rows.append({"t_us": round(e.time_range.start - t0, 1), "dur_us": round(e.time_range.end - e.time_range.start, 1),
                         "kernel": e.name[:110]})
        span = last[-1].time_range.end - t0
        print(json.dumps({"n": n, "world": world, "geometry": geometry, "replicate_below": rb, "kernels_in_cycle": len(last), "cycle_span_us": round(span, 1),
                          "sum_kernel_us": round(sum(r["dur_us"] for r in rows), 1)}))
        for r in rows:
            print(f'{r["t_us"]:9.1f} {r["dur_us"]:8.1f}  {r["kernel"]}')
    H._graph = None
    H.check_exchange()
    H.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
