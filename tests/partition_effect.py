"""Test infrastructure (CPU, oracle only): what per-rank aggregation costs against single-domain aggregation.

    python tests/partition_effect.py > profiles/r02_partitioned_vs_single_domain_oracle.jsonl

The multi-GPU hierarchy aggregates every rank's diagonal block on its own (aggregates never straddle ranks), so it equals the
PARTITIONED oracle's hierarchy, not the single-domain one.  This script builds both with the oracle on z-slab problems of
48 x 48 x 24 DOF per rank and reports PCG iteration counts (rtol 1e-8) and operator complexities side by side."""
import sys, time, json
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np
from oracle import multilevel as oml
nz_per = 24; nx = 48
rows = []
for world in (1, 2, 4, 8):
    shape = (nx, nx, nz_per * world)              # z-slabs, weak scaling like the bench's slab geometry (small)
    A = oml.poisson(shape)
    n = A.shape[0]
    b = np.random.RandomState(0).randn(n)
    lam = [2.0] * 8
    t0 = time.time()
    single = oml.build_hierarchy(A, ratio=0.027, distance="unit", rand=0, lam_max=lam, max_coarse=500)
    _, res_s, it_s = oml.pcg(single, b, tol=1e-8)
    if world > 1:
        offs = np.arange(world + 1) * (n // world)
        part, _ = oml.build_hierarchy_partitioned(A, offs, ratio=0.027, distance="unit", rand=0, lam_max=lam, max_coarse=500, replicate_below=4000)
        _, res_p, it_p = oml.pcg(part, b, tol=1e-8)
        oc_p = sum(L.A.nnz for L in part) / part[0].A.nnz
    else:
        it_p, oc_p = it_s, sum(L.A.nnz for L in single) / single[0].A.nnz
    oc_s = sum(L.A.nnz for L in single) / single[0].A.nnz
    rows.append(dict(world=world, dof=n, pcg_iterations_single_domain=it_s, pcg_iterations_partitioned=it_p,
                     operator_complexity_single=round(oc_s, 4), operator_complexity_partitioned=round(oc_p, 4), seconds=round(time.time() - t0, 1)))
    print(json.dumps(rows[-1]), flush=True)
