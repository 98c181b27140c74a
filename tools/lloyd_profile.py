"""Lloyd aggregation at a given grid size, repeated: seconds, passes, sweeps and rows evaluated per call (development aid).

    python tools/lloyd_profile.py [n] [reps]        MLAMG_AGG_FRONTIER=0 for the full-pass variant"""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ml-amg_b200")]
import numpy as np
import torch


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    import mlamg
    from mlamg import core, hierarchy as hm
    from mlamg._lib import lib
    A = mlamg.poisson((n, n, n), torch.float64)
    G = hm.distance_transform(A, "unit")
    seeds = hm.lloyd_seeds(A.shape[0], 0.027, 0).astype(np.int32)
    for rep in range(reps):
        lib.mlamg_agg_stats(None, 1)
        torch.cuda.synchronize()
        t = time.perf_counter()
        d, c, roots, it = core.lloyd_cluster(G, seeds, maxiter=10)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        st = (ctypes.c_longlong * 8)()
        lib.mlamg_agg_stats(st, 0)
        names = ["ordered_passes", "free_passes", "sweeps", "chain_passes", "rows_ordered", "rows_free", "lloyd_iterations"]
        print(json.dumps({"n": n, "rep": rep, "seconds": round(dt, 4), "frontier": os.environ.get("MLAMG_AGG_FRONTIER", "1"),
                          **{k: int(st[i]) for i, k in enumerate(names)},
                          "host_wait_in_flag_fetch_s": round(st[7] / 1e9, 4),
                          "rows_ordered_per_n": round(st[4] / A.shape[0], 2), "rows_free_per_n": round(st[5] / A.shape[0], 2)}), flush=True)


if __name__ == "__main__":
    main()
