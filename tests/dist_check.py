"""Multi-rank parity check of the row-partitioned hierarchy against the partitioned CPU oracle.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py [n] [--cube] [--delaunay | --delaunay-strips]

Each rank owns a z-slab of the global n x n x (n_z*world) Poisson grid.  Checks (per rank, its rows):
aggregates bit-exact, P and the Galerkin operator of every distributed level bit-identical to the global
scipy computation, one V-cycle and the PCG residual history within 1e-12 of the oracle.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ml-amg_b200")]
import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist


def run_check(n, comm, geometry="slab", delaunay=False, verbose=True, extras=True, strips=False):
    """Build the row-partitioned hierarchy on this communicator and hold it to the partitioned CPU oracle.
    -> dict(ok, labels_equal, P_bitwise, A_bitwise, vcycle_rel_err, pcg_iters_equal, pcg_hist_err, ...) for THIS rank.
    extras: also the transport/graph self-consistency checks (skipped by bench.py's pre-timing parity leg)."""
    import mlamg
    from mlamg import distributed as md
    from oracle import multilevel as oml
    rank, world = comm.rank, comm.world
    saved = (md.OVERLAP_MIN_NNZ, md.PEER_SPLIT_MIN_NNZ)
    md.OVERLAP_MIN_NNZ = 0          # exercise the interior/boundary overlap path even on tiny levels
    md.PEER_SPLIT_MIN_NNZ = 0
    if delaunay:
        # BASELINE config 4 shape: P1 Laplacian on a Delaunay mesh of n random points (Morton-ordered rows, ~7 entries
        # per row), contiguous row blocks.  Every rank generates the same global matrix and keeps its rows.
        from mlamg import problems
        if strips:
            # the distributed generator of config 4 (every rank triangulates its own strip + certified halo only); the
            # global matrix built here is only the oracle's input
            A, offs_np = problems.delaunay_laplacian_strips(n, 0, world)
            N = A.shape[0]
            offsets = [int(v) for v in offs_np]
            rp_l, col_l, val_l, offs_l = problems.delaunay_laplacian_distributed(n, 0, world, rank)
            assert list(offs_l) == offsets
            rowptr = torch.from_numpy(rp_l).cuda()
            col = torch.from_numpy(col_l.astype(np.int32)).cuda()
            val = torch.from_numpy(val_l.copy()).cuda()
            # the oracle is given exactly the rows the ranks hold (entry sums may differ in the last bit between the two
            # assemblies), so that P / Galerkin can be compared bit for bit
            parts = comm.all_gather_obj((rp_l, col_l, val_l))
            A = sp.vstack([sp.csr_matrix((v, c, r), shape=(len(r) - 1, N)) for r, c, v in parts]).tocsr()
            A.sort_indices()
        else:
            A, _ = problems.delaunay_laplacian(n, seed=0)
            N = A.shape[0]
            offsets = [int(round(r * N / world)) for r in range(world + 1)]
            Al = A[offsets[rank]:offsets[rank + 1]]
            rowptr = torch.from_numpy(Al.indptr.astype(np.int32)).cuda()
            col = torch.from_numpy(Al.indices.astype(np.int32)).cuda()
            val = torch.from_numpy(Al.data.copy()).cuda()
        lam = [2.0, 1.9, 1.8, 1.7, 1.6]
        kw = dict(ratio=0.08, distance="unit", maxiter=10, rand=0, lam_max=lam, max_levels=6, max_coarse=30,
                  replicate_below=max(60, N // 20))
    else:
        rowptr, col, val = md.poisson_slab(n, world, rank, geometry=geometry)
        nx, ny, nzl = md.slab_geometry(n, world, geometry)
        N_loc = nx * ny * nzl
        offsets = [r * N_loc for r in range(world + 1)]
        lam = [2.0, 1.9, 1.8, 1.7, 1.6]
        kw = dict(ratio=0.06, distance="unit", maxiter=10, rand=0, lam_max=lam, max_levels=6, max_coarse=30,
                  replicate_below=max(60, N_loc * world // 40))
        A = oml.poisson((nx, ny, nzl * world))
    H = md.DistHierarchy(rowptr, col, val, comm, **kw)
    ref, offs = oml.build_hierarchy_partitioned(A, offsets, **kw)
    out = {"ok": True, "labels_equal": True, "P_bitwise": True, "A_bitwise": True, "failed": []}

    def check(name, cond, key=None):
        if not cond:
            out["ok"] = False
            out["failed"].append(name)
            if key:
                out[key] = False
            if verbose:
                print(f"[rank {rank}] FAIL {name}", flush=True)

    check("number of distributed levels", len(H.levels) == len(offs) - 1 and len(H.levels) >= 1)
    check("total levels", len(H.levels) + len(H.tail.levels) == len(ref))
    for l, L in enumerate(H.levels):
        lo, hi = int(H.offsets[l][rank]), int(H.offsets[l][rank + 1])
        check(f"L{l} offsets", list(H.offsets[l]) == list(offs[l]))
        clo = int(H.offsets[l + 1][rank])
        lab = L.labels.cpu().numpy().astype(np.int64)
        check(f"L{l} labels", np.array_equal(np.where(lab >= 0, lab + clo, -1), ref[l].labels[lo:hi]), "labels_equal")
        Pg = L.P_global.to_scipy()
        Pr = sp.csr_matrix(ref[l].P[lo:hi])
        check(f"L{l} P pattern", np.array_equal(Pg.indptr, Pr.indptr) and np.array_equal(Pg.indices, Pr.indices), "P_bitwise")
        check(f"L{l} P bits", Pg.nnz == Pr.nnz and np.array_equal(Pg.data, Pr.data), "P_bitwise")
    # Galerkin operators: level l+1 rows of this rank (distributed) / whole matrix (tail)
    for l in range(1, len(H.levels)):
        L = H.levels[l]
        lo, hi = int(H.offsets[l][rank]), int(H.offsets[l][rank + 1])
        Al = L.A.csr.to_scipy()
        order = getattr(L, "perm_new2old", None)            # block-local apply renumbering (new -> old), if any
        order = order.cpu().numpy() if order is not None else np.arange(hi - lo)
        ext = np.concatenate([lo + order, L.A.plan.halo_ids.cpu().numpy()])
        Ag = sp.csr_matrix((Al.data, ext[Al.indices], Al.indptr), shape=(hi - lo, ref[l].A.shape[1]))
        Ag.sort_indices()
        Ar = sp.csr_matrix(sp.csr_matrix(ref[l].A[lo:hi])[order])
        Ar.sort_indices()
        check(f"L{l} A bits", np.array_equal(Ag.indptr, Ar.indptr) and np.array_equal(Ag.indices, Ar.indices)
              and np.array_equal(Ag.data, Ar.data), "A_bitwise")
    nd = len(H.levels)
    for k, Lt in enumerate(H.tail.levels):
        At, Ar = Lt.A.to_scipy(), ref[nd + k].A
        check(f"tail{k} A bits", At.shape == Ar.shape and np.array_equal(At.indptr, Ar.indptr)
              and np.array_equal(At.indices, Ar.indices) and np.array_equal(At.data, Ar.data), "A_bitwise")
    # cycles
    lo, hi = offsets[rank], offsets[rank + 1]
    bg = np.random.RandomState(0).randn(A.shape[0])
    b = torch.from_numpy(bg[lo:hi]).cuda()
    x = torch.empty_like(b)
    out["vcycle_rel_err"] = 0.0
    for nu1, nu2 in ((1, 1), (2, 2)):
        H.vcycle(b, x, nu1, nu2)
        xr = oml.vcycle(ref, bg.copy(), None, nu1, nu2)
        err = float(np.abs(x.cpu().numpy() - xr[lo:hi]).max() / np.abs(xr).max())
        out["vcycle_rel_err"] = max(out["vcycle_rel_err"], err)
        check(f"vcycle({nu1},{nu2}) rel err {err:.2e}", err < 1e-12)
    if world > 1 and H.halo != "peer" and verbose:
        print(f"[rank {rank}] peer transport unavailable: cycles above ran on the NCCL transport", flush=True)
    if extras and world > 1 and H.halo == "peer":
        # the same cycle on the other halo transport (NCCL all-to-all vs peer windows) and without the
        # interior/boundary split agrees to rounding (the threads-per-row heuristic depends on the launch size);
        # repeated eager cycles and CUDA-graph replays of the peer cycle agree bit for bit
        def close(a, c):
            return float((a - c).abs().max() / c.abs().max()) < 1e-13
        H.vcycle(b, x, 1, 1)
        x_peer = x.clone()
        H.halo = "nccl"
        H.vcycle(b, x, 1, 1)
        check("peer and NCCL transports agree", close(x, x_peer))
        H.halo = "peer"
        H.overlap = False
        H.vcycle(b, x, 1, 1)
        check("peer transport without row split agrees", close(x, x_peer))
        H.overlap = True
        for _ in range(5):                       # parity double-buffering: repeated use of every channel
            x.zero_()
            H.vcycle(b, x, 1, 1)
        check("repeated peer cycles agree bitwise", torch.equal(x, x_peer))
        replay = H.capture(b, x, 1, 1)
        for _ in range(4):
            x.zero_()
            replay()
        torch.cuda.synchronize()
        check("CUDA-graph replay of the peer cycle agrees bitwise", torch.equal(x, x_peer))
        H.check_exchange()
    out["pcg_iters_equal"], out["pcg_hist_err"] = True, 0.0
    xr, res_r, it_r = oml.pcg(ref, bg, tol=1e-8, maxiter=100)
    for overlap in ((True, False) if extras else (True,)):
        H.overlap = overlap
        xs, res, it = H.pcg(b, tol=1e-8, maxiter=100)
        e = float(np.max(np.abs(res - res_r[:len(res)])) / res_r[0])
        out["pcg_hist_err"] = max(out["pcg_hist_err"], e)
        check(f"pcg overlap={overlap} iterations {it} vs {it_r}, history err {e:.2e}", it == it_r and e < 1e-11, "pcg_iters_equal")
        check("pcg solution", np.abs(xs.cpu().numpy() - xr[lo:hi]).max() <= 1e-9 * np.abs(xr).max())
    H.overlap = True
    out.update(dist_levels=len(H.levels), tail_levels=len(H.tail.levels), halo_entries=int(H.levels[0].A.plan.n_halo),
               dof_per_gpu=int(hi - lo), halo_transport=H.halo, pcg_iterations=int(it_r))
    H._graph = None
    if world > 1:
        H.close()
    md.OVERLAP_MIN_NNZ, md.PEER_SPLIT_MIN_NNZ = saved
    return out


def main():
    args = [a for a in sys.argv[1:]]
    strips = "--delaunay-strips" in args
    delaunay = "--delaunay" in args or strips
    geometry = "cube" if "--cube" in args else "slab"
    nums = [a for a in args if not a.startswith("--")]
    n = int(nums[0]) if nums else (4000 if delaunay else 12)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mlamg import distributed as md
    out = run_check(n, md.Comm(), geometry=geometry, delaunay=delaunay, strips=strips)
    print(f"[rank {rank}] {'PASS' if out['ok'] else 'FAIL'}: {out['dist_levels']} distributed + {out['tail_levels']} replicated "
          f"levels, halo {out['halo_entries']} entries, {geometry if not delaunay else 'delaunay'}", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
