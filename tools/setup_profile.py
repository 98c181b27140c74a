"""Where the hierarchy setup time goes (development aid, one GPU):  python tools/setup_profile.py [n]

Wraps the setup wrappers of mlamg.core with synchronising timers and prints seconds per stage for the
n^3 Poisson hierarchy (Lloyd aggregation, SA prolongator, Galerkin product, apply copies)."""
import collections
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
    import mlamg
    from mlamg import core, hierarchy as hm
    acc = collections.OrderedDict()

    def wrap(mod, name):
        fn = getattr(mod, name)

        def timed(*a, **k):
            torch.cuda.synchronize()
            t = time.perf_counter()
            out = fn(*a, **k)
            torch.cuda.synchronize()
            acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
            return out
        setattr(mod, name, timed)
    for name in ("lloyd_cluster", "spgemm", "transpose", "drop_zeros", "sa_smoother", "agg_from_labels", "dense_inverse",
                 "smoother_diag", "lambda_max", "sort_rows"):
        wrap(core, name)
    for name in ("_permuted", "lloyd_seeds", "distance_transform"):
        wrap(hm, name)
    exact = 1.0 + np.cos(np.pi / (n + 1))
    for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
        acc.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        A = mlamg.poisson((n, n, n), torch.float64)
        H = mlamg.build_hierarchy(A, aggregates="lloyd", ratio=0.027, distance="unit", maxiter=10, rand=0,
                                  lam_max=lambda M: exact if M.shape[0] == n ** 3 else mlamg.lambda_max(M),
                                  max_coarse=1000, max_levels=8)
        torch.cuda.synchronize()
        total = time.perf_counter() - t0
        print(json.dumps({"rep": rep, "n": n, "total_s": round(total, 3), "levels": [l.A.shape[0] for l in H.levels],
                          "stages_s": {k: round(v, 3) for k, v in acc.items()},
                          "unaccounted_s": round(total - sum(v for k, v in acc.items() if k not in ("sort_rows",)), 3)}), flush=True)
        del H, A


if __name__ == "__main__":
    main()
