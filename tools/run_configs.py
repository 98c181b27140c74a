"""Timed runs of the BASELINE.json configurations other than the bench workload (development aid).

    python tools/run_configs.py 1            # 2D 5-point 256^2, the reference's canonical two-level path, CPU oracle timed beside
    python tools/run_configs.py 2            # 3D 7-point 128^3: setup + AMG-preconditioned CG (1 GPU)
    python tools/run_configs.py 3            # 2D Voronoi jump diffusion, 4M DOF, GNN-style aggregates and P weights (1 GPU)
    torchrun --nproc-per-node N tools/run_configs.py 4 [npts]   # Delaunay P1 Laplacian, row-partitioned over N GPUs

One JSON line per configuration (rank 0).  Parity of the same paths is in tests/test_gpu_configs.py; this script only
measures."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ml-amg_b200")]
import numpy as np
import torch


def sync_time(fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t


def cycle_ms(H, b, reps=20):
    from mlamg import core
    x = torch.empty_like(b)
    H.use_graph(True)
    fn = lambda: core.check(core.lib.mlamg_vcycle(H._h, core.ptr(b), core.ptr(x), 1, 1, 1, core.stream()))
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    H.use_graph(False)
    return e0.elapsed_time(e1) / reps


def config1():
    """utils/evaluate_dataset.py:59-101 -> ns/lib/graph.py:156-239 -> ns/lib/multigrid.py:102-210 at 256^2, stage by
    stage, second (warm) call timed; the CPU oracle (scipy + restated pyamg C loops, 1 core) beside every stage."""
    import mlamg
    import ns.lib.graph as g
    import ns.lib.multigrid as mg
    from oracle import multilevel as oml, reference_path as rp
    A = oml.poisson((256, 256))
    n = A.shape[1]
    b = np.zeros(n)
    x = np.random.RandomState(0).randn(n)
    x /= np.linalg.norm(x, 2)

    def warm(fn, reps=2):
        out = None
        for _ in range(reps):
            out, t = sync_time(fn)
        return out, t

    def cpu(fn):
        t = time.perf_counter()
        out = fn()
        return out, time.perf_counter() - t
    (Agg, roots, seeds), t_agg = warm(lambda: g.lloyd_aggregation(A, ratio=0.1, distance="unit", rand=0))
    (Agg_c, roots_c, _), t_agg_c = cpu(lambda: rp.lloyd_aggregation(A, ratio=0.1, distance="unit", rand=0))
    info = {}
    Ad = mlamg.DeviceCSR.from_scipy(A)
    lam, t_lam = warm(lambda: mlamg.lambda_max(Ad, info=info))
    lam_c, t_lam_c = cpu(lambda: rp.lambda_max_dinv_a(A))
    P, t_p = warm(lambda: mg.smoothed_aggregation_jacobi(A, Agg, omega=(4.0 / 3.0) / lam))
    P_c, t_p_c = cpu(lambda: sp_csr(rp.smoothed_aggregation_jacobi(A, Agg_c, omega=(4.0 / 3.0) / lam_c)))
    out = {"config": 1, "workload": "poisson2d_5pt_256^2 lloyd 0.1 unit rand 0 + SA P + amg_2_v(res_tol 1e-10), fp64",
           "dof": n, "coarse": int(Agg.shape[1]), "labels_equal": bool(np.array_equal(Agg.indices, Agg_c.indices)),
           "lambda_max": {"gpu": lam, "cpu_arpack": lam_c, "lanczos_steps": info.get("steps"), "residual": info.get("residual")},
           "stages_ms": {"lloyd_aggregation": [round(t_agg * 1e3, 2), round(t_agg_c * 1e3, 2)],
                         "lambda_max": [round(t_lam * 1e3, 2), round(t_lam_c * 1e3, 2)],
                         "smoothed_aggregation_jacobi": [round(t_p * 1e3, 2), round(t_p_c * 1e3, 2)]},
           "stages_ms_columns": ["gpu (warm, incl. H2D/D2H of the scipy in/outputs)", "cpu oracle (1 core)"]}
    for sm in ("gauss_seidel", "jacobi"):
        kw = dict(res_tol=1e-10, jacobi_weight=2.0 / 3.0, smoother=sm)
        got, t_g = warm(lambda: mg.amg_2_v(A, P, b, x.copy(), **kw))
        ref, t_c = cpu(lambda: rp.amg_2_v(A, P_c, b, x.copy(), **kw))
        out["stages_ms"][f"amg_2_v[{sm}]"] = [round(t_g * 1e3, 2), round(t_c * 1e3, 2)]
        out[f"amg_2_v[{sm}]"] = {"iterations": [int(got[3]), int(ref[3])], "conv_factor": [float(got[1]), float(ref[1])],
                                 "history_err0": float(np.max(np.abs(got[2] - ref[2])) / ref[2][0])}
    # inside amg_2_v[jacobi] on the GPU: two-level setup vs the device-resident loop
    tl, t_setup = warm(lambda: mg._TwoLevel(A, P))
    H, t_h = sync_time(lambda: tl.hierarchy("jacobi", 2.0 / 3.0))
    bd = torch.zeros(n, dtype=torch.float64, device="cuda")
    xd = torch.from_numpy(x).cuda()
    H.solve_abs(bd, xd.clone(), 1e-10, 500, 1, 1, H.SOLVE_NO_INITIAL_CHECK)
    x2 = xd.clone()
    (_, hist), t_loop = sync_time(lambda: H.solve_abs(bd, x2, 1e-10, 500, 1, 1, H.SOLVE_NO_INITIAL_CHECK))
    out["amg_2_v[jacobi]_gpu_breakdown_ms"] = {"galerkin+dense_inverse": round(t_setup * 1e3, 2), "Q/scaled copies + handle": round(t_h * 1e3, 2),
                                               "loop": round(t_loop * 1e3, 2), "iterations": len(hist) - 1, "loop_mode": H.loop_mode}
    print(json.dumps(out), flush=True)


def sp_csr(M):
    import scipy.sparse as sp
    return sp.csr_matrix(M)


def config2():
    import mlamg
    n = 128
    exact = 1.0 + np.cos(np.pi / (n + 1))
    A = mlamg.poisson((n, n, n), torch.float64)
    H, t_setup = sync_time(lambda: mlamg.build_hierarchy(
        A, aggregates="lloyd", ratio=0.027, distance="unit", maxiter=10, rand=0,
        lam_max=lambda M: exact if M.shape[0] == n ** 3 else mlamg.lambda_max(M), max_coarse=1000, max_levels=8))
    b = mlamg.spmv(A, torch.ones(n ** 3, dtype=torch.float64, device="cuda"))
    (x, res), t_solve = sync_time(lambda: H.solve(b, tol=1e-8, maxiter=100, accel="cg", return_residuals=True))
    (x, res), t_solve = sync_time(lambda: H.solve(b, tol=1e-8, maxiter=100, accel="cg", return_residuals=True))
    ms = cycle_ms(H, b)
    print(json.dumps({"config": 2, "workload": "poisson3d_7pt_128^3 setup + PCG(V(1,1) Jacobi) rtol 1e-8", "dof": n ** 3,
                      "levels": [l.A.shape[0] for l in H.levels], "setup_s": round(t_setup, 3), "pcg_iterations": len(res) - 1,
                      "pcg_solve_ms": round(t_solve * 1e3, 2), "pcg_loop_mode": H.loop_mode, "final_rel_residual": float(res[-1] / np.linalg.norm(b.cpu().numpy())),
                      "vcycle_ms": round(ms, 4), "vcycle_gdof_per_s": round(n ** 3 / ms / 1e6, 2)}), flush=True)


def config3():
    import mlamg
    from mlamg import problems
    import ns.model.agg_interp as ai
    t0 = time.perf_counter()
    A, pts, jumps = problems.voronoi_jump_problem(2001, seed=0, mesh="structured")
    t_gen = time.perf_counter() - t0
    n = A.shape[0]
    top_k, bf, ph = problems.random_gnn_outputs(A, alpha=0.1, seed=0)
    Ad = mlamg.DeviceCSR.from_scipy(A)
    tk = torch.from_numpy(top_k).cuda()
    bfd, phd = torch.from_numpy(bf).cuda(), torch.from_numpy(ph + np.float32(0.1)).cuda()
    ai.bellman_ford_aggregates(Ad, tk, bfd)                      # warm-up (allocator, module load)
    (agg_T, labels, dist, near), t_bf = sync_time(lambda: ai.bellman_ford_aggregates(Ad, tk, bfd))
    (P_T, P), t_p = sync_time(lambda: ai.learned_prolongator(Ad, phd, labels, len(top_k)))
    H, t_h = sync_time(lambda: mlamg.build_hierarchy(Ad, aggregates=[(labels, len(top_k))], P_hat=[phd.double()], fallback="lloyd",
                                                     ratio=0.1, distance="unit", rand=0, max_coarse=1000, max_levels=8))
    b = torch.randn(n, dtype=torch.float64, device="cuda")
    ms = cycle_ms(H, b)
    print(json.dumps({"config": 3, "workload": "voronoi jump diffusion 2000^2 P1, random-init GNN outputs (alpha 0.1)", "dof": n,
                      "nnz": A.nnz, "kappa_jumps": jumps[:, 2].tolist(), "host_generation_s": round(t_gen, 2),
                      "bellman_ford_aggregates_ms": round(t_bf * 1e3, 2), "learned_P_ms": round(t_p * 1e3, 2),
                      "hierarchy_setup_s": round(t_h, 3), "levels": [l.A.shape[0] for l in H.levels],
                      "vcycle_ms": round(ms, 4), "vcycle_gdof_per_s": round(n / ms / 1e6, 2)}), flush=True)


def config4(npts):
    import torch.distributed as dist
    import mlamg
    from mlamg import problems, distributed as md
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = md.Comm()
    t0 = time.perf_counter()
    ginfo = {}
    # every rank triangulates ONLY its own strip plus a certified halo band (mlamg.problems.delaunay_laplacian_distributed)
    rp_l, col_l, val_l, offs = problems.delaunay_laplacian_distributed(npts, 0, world, rank, info=ginfo)
    t_gen = time.perf_counter() - t0
    N = int(offs[-1])
    offsets = [int(v) for v in offs]
    nnz_glob = int(comm.allreduce_sum(float(len(col_l))))
    rowptr = torch.from_numpy(rp_l).cuda()
    col = torch.from_numpy(col_l.astype(np.int32)).cuda()
    val = torch.from_numpy(val_l.copy()).cuda()
    torch.cuda.synchronize(); comm.barrier()
    t0 = time.perf_counter()
    H = md.DistHierarchy(rowptr, col, val, comm, ratio=0.1, distance="unit", maxiter=10, rand=0, lam_max=None, max_levels=8,
                         max_coarse=1000, replicate_below=200000)
    torch.cuda.synchronize(); comm.barrier()
    t_setup = time.perf_counter() - t0
    bg = np.random.RandomState(0).randn(N)
    b = torch.from_numpy(bg[offsets[rank]:offsets[rank + 1]]).cuda()
    x = torch.empty_like(b)
    H.pcg(b, tol=1e-8, maxiter=5)                               # warm-up
    torch.cuda.synchronize(); comm.barrier()
    t0 = time.perf_counter()
    xs, res, it = H.pcg(b, tol=1e-8, maxiter=200)
    torch.cuda.synchronize(); comm.barrier()
    t_solve = time.perf_counter() - t0
    replay = H.capture(b, x, 1, 1) if world > 1 and H.halo == "peer" else (lambda: H.vcycle(b, x, 1, 1))
    for _ in range(10):
        replay()
    torch.cuda.synchronize(); comm.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        replay()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 20], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    H.check_exchange()
    if rank == 0:
        print(json.dumps({"config": 4, "workload": f"P1 Laplacian on a Delaunay mesh of {npts} random points, y-strips with Z-curve order inside, "
                          f"row-partitioned over {world} GPU(s)", "dof": N, "nnz": nnz_glob, "host_generation_s": round(t_gen, 1),
                          "generator": {"per_rank_points_triangulated": ginfo.get("local_points"), "owned": ginfo.get("owned"),
                                        "halo_band": ginfo.get("delta")},
                          "distributed_levels": [int(o[-1]) for o in H.offsets[:-1]], "replicated_levels": [l.A.shape[0] for l in H.tail.levels],
                          "halo_entries_fine": H.levels[0].A.plan.n_halo if H.levels else 0,
                          "setup_s": round(t_setup, 2), "pcg_iterations": int(it), "pcg_solve_ms": round(t_solve * 1e3, 1),
                          "final_rel_residual": float(res[-1] / res[0]), "vcycle_ms": round(float(ms.item()), 4),
                          "vcycle_gdof_per_s": round(N / float(ms.item()) / 1e6, 2)}), flush=True)
    H._graph = None
    if world > 1:
        H.close()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "2"
    if which == "1":
        config1()
    elif which == "2":
        config2()
    elif which == "3":
        config3()
    else:
        config4(int(sys.argv[2]) if len(sys.argv) > 2 else 2000000)
