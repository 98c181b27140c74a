"""CPU, world_size 2, gloo: the host-side plumbing of the row-partitioned multi-GPU levels
(partition, halo plan + exchange, row fetch, distributed transpose, gather) against scipy, and the
partitioned oracle itself.  No kernel runs here (there is no GPU in this container)."""
import os
import socket
import sys
import traceback

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, random_csr
from oracle import multilevel as oml


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _global_matrix(n=97, seed=5):
    A = random_csr(n, n, 0.08, seed, empty_rows=True) + sp.diags(np.arange(1.0, n + 1))
    A = sp.csr_matrix(A)
    A.sort_indices()
    return A


def _worker(rank, world, port, q):
    try:
        for p in (ROOT, os.path.join(ROOT, "ml-amg_b200")):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from mlamg import distributed as md
        comm = md.Comm()
        world = comm.world
        A = _global_matrix()
        n = A.shape[0]
        cut = [0, 41, n]
        lo, hi = cut[rank], cut[rank + 1]
        Al = sp.csr_matrix(A[lo:hi])
        rowptr = torch.from_numpy(Al.indptr.astype(np.int32))
        col = torch.from_numpy(Al.indices.astype(np.int32))
        val = torch.from_numpy(Al.data.copy())
        offs = md.partition_offsets(hi - lo, comm)
        assert list(offs) == cut
        # localize + halo exchange of a vector
        col_loc, halo = md.localize(col, hi - lo, lo)
        plan = md.HaloPlan(halo, offs, comm)
        xg = np.random.RandomState(0).randn(n)
        x_ext = torch.zeros(hi - lo + plan.n_halo, dtype=torch.float64)
        x_ext[:hi - lo] = torch.from_numpy(xg[lo:hi])
        plan.exchange(x_ext, hi - lo)
        assert np.array_equal(x_ext[hi - lo:].numpy(), xg[halo.numpy()])
        y = sp.csr_matrix((Al.data, col_loc.numpy(), Al.indptr), shape=(hi - lo, x_ext.numel())) @ x_ext.numpy()
        assert np.allclose(y, (A @ xg)[lo:hi], rtol=1e-14)
        lab = torch.zeros(x_ext.numel(), dtype=torch.int32)
        lab[:hi - lo] = torch.arange(lo, hi, dtype=torch.int32)
        plan.exchange(lab, hi - lo)
        assert np.array_equal(lab[hi - lo:].numpy(), halo.numpy())
        # distributed transpose
        t_rp, t_col, t_val = md.dist_transpose(rowptr, col, val, offs, offs, comm)
        At = sp.csr_matrix(A.T)
        At.sort_indices()
        ref = sp.csr_matrix(At[lo:hi])
        assert np.array_equal(t_rp.numpy(), ref.indptr) and np.array_equal(t_col.numpy(), ref.indices)
        assert np.array_equal(t_val.numpy(), ref.data)
        # rectangular transpose (row partition != column partition)
        B = random_csr(n, 31, 0.2, 9)
        ccut = [0, 12, 31]
        Bl = sp.csr_matrix(B[lo:hi])
        b_rp, b_col, b_val = md.dist_transpose(torch.from_numpy(Bl.indptr.astype(np.int32)),
                                               torch.from_numpy(Bl.indices.astype(np.int32)), torch.from_numpy(Bl.data.copy()),
                                               offs, np.array(ccut), comm)
        Bt = sp.csr_matrix(B.T)
        Bt.sort_indices()
        refb = sp.csr_matrix(Bt[ccut[rank]:ccut[rank + 1]])
        assert np.array_equal(b_rp.numpy(), refb.indptr) and np.array_equal(b_col.numpy(), refb.indices)
        assert np.array_equal(b_val.numpy(), refb.data)
        # fetch rows owned by the other rank
        f_rp, f_col, f_val = md.fetch_rows(rowptr, col, val, offs, halo, comm)
        reff = sp.csr_matrix(A[halo.numpy()])
        assert np.array_equal(f_rp.numpy(), reff.indptr) and np.array_equal(f_col.numpy(), reff.indices)
        assert np.array_equal(f_val.numpy(), reff.data)
        # diagonal block and gather
        d_rp, d_col, d_val = md.diag_block(rowptr, col, val, hi - lo, lo)
        refd = sp.csr_matrix(A[lo:hi, lo:hi])
        refd.sort_indices()
        assert np.array_equal(d_rp.numpy(), refd.indptr) and np.array_equal(d_col.numpy(), refd.indices)
        g_rp, g_col, g_val = md.gather_csr(rowptr, col, val, comm)
        assert np.array_equal(g_rp.numpy(), A.indptr) and np.array_equal(g_col.numpy(), A.indices) and np.array_equal(g_val.numpy(), A.data)
        assert abs(comm.allreduce_sum(rank + 1.0) - 3.0) < 1e-15
        # peer-window layout (host logic of csrc/peer.cu's channels): emulate every rank's pushes with the planned
        # remote offsets into byte arrays standing in for the windows, then unpack like the wait kernel
        specs = [(plan.send_counts, plan.recv_counts, 8),
                 ([3 + rank] * world, [3 + r for r in range(world)], 4)]       # halo channel, gather channel
        lay = md.plan_channels(specs, comm)
        assert lay["nbytes"] % 256 == 0 and lay["region"][0][0] == 0
        srcs = [x_ext[:hi - lo].numpy()[plan.send_idx.numpy()], np.arange(3 + rank, dtype=np.float32).repeat(1) + 10 * rank]
        writes = []                                   # (dest rank, byte offset, payload bytes) for parity 0 and 1
        for c, (sc, rc, esz) in enumerate(specs):
            start = 0
            for pdst in range(world):
                cnt = int(sc[pdst])
                seg = srcs[c][start:start + cnt] if c == 0 else srcs[c]
                start += cnt if c == 0 else 0
                for par in (0, 1):
                    writes.append((pdst, lay["remote"][pdst][c][par], np.ascontiguousarray(seg).tobytes()))
        allw = comm.all_gather_obj(writes)
        win = np.zeros(lay["nbytes"], dtype=np.uint8)
        for ws in allw:
            for pdst, off, payload in ws:
                if pdst == rank:
                    assert off + len(payload) <= lay["nbytes"]
                    win[off:off + len(payload)] = np.frombuffer(payload, dtype=np.uint8)
        for par in (0, 1):
            got = win[lay["region"][0][par]:lay["region"][0][par] + 8 * plan.n_halo].view(np.float64)
            assert np.array_equal(got, xg[halo.numpy()])
            n_all = sum(3 + r for r in range(world))
            got = win[lay["region"][1][par]:lay["region"][1][par] + 4 * n_all].view(np.float32)
            assert np.array_equal(got, np.concatenate([np.arange(3 + r, dtype=np.float32) + 10 * r for r in range(world)]))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:
        q.put((rank, traceback.format_exc()))


def _worker3(rank, world, port, q):
    """three ranks with unequal blocks: halo plans of a middle rank (two neighbours) and the window layout"""
    try:
        for p in (ROOT, os.path.join(ROOT, "ml-amg_b200")):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from mlamg import distributed as md
        comm = md.Comm()
        A = sp.csr_matrix(oml.poisson((9, 14)))              # 5-point grid, x fastest: slabs in y
        n = A.shape[0]
        cut = [0, 27, 90, n]
        lo, hi = cut[rank], cut[rank + 1]
        Al = sp.csr_matrix(A[lo:hi])
        col = torch.from_numpy(Al.indices.astype(np.int32))
        offs = md.partition_offsets(hi - lo, comm)
        assert list(offs) == cut
        col_loc, halo = md.localize(col, hi - lo, lo)
        plan = md.HaloPlan(halo, offs, comm)
        assert plan.n_halo == (9 if rank in (0, 2) else 18)          # one grid line per neighbour
        assert [c > 0 for c in plan.recv_counts] == [abs(r - rank) == 1 for r in range(world)]
        assert plan.send_counts == plan.recv_counts                  # symmetric stencil
        xg = np.random.RandomState(1).randn(n)
        x_ext = torch.zeros(hi - lo + plan.n_halo, dtype=torch.float64)
        x_ext[:hi - lo] = torch.from_numpy(xg[lo:hi])
        plan.exchange(x_ext, hi - lo)
        assert np.array_equal(x_ext[hi - lo:].numpy(), xg[halo.numpy()])
        # window layout: a halo channel (16-byte slots) and the padded coarse gather (every pair connected)
        sizes = [2, 0, 3]                                            # an EMPTY slice on rank 1
        specs = [(plan.send_counts, plan.recv_counts, 16), ([sizes[rank] + 1] * world, [s_ + 1 for s_ in sizes], 16)]
        lay = md.plan_channels(specs, comm)
        tables = comm.all_gather_obj((lay["region"], lay["nbytes"]))
        for p_ in range(world):
            region_p, nbytes_p = tables[p_]
            for c, (sc, rc, slot) in enumerate(specs):
                if int(sc[p_]) == 0:
                    continue
                off0, off1 = lay["remote"][p_][c]
                assert off0 % 16 == 0 and off1 - off0 == region_p[c][1] - region_p[c][0]
                assert region_p[c][0] <= off0 and off0 + int(sc[p_]) * slot <= region_p[c][1] <= nbytes_p
        # my own regions do not overlap and every sender's segment lies inside them, back to back in rank order
        rg = lay["region"]
        assert rg[0][0] == 0 and rg[0][1] <= rg[1][0] and rg[1][1] <= lay["nbytes"]
        starts = comm.all_gather_obj([lay["remote"][p_][1][0] for p_ in range(world)])   # where each rank writes into p_
        mine = [starts[s_][rank] for s_ in range(world)]
        expect = rg[1][0]
        for s_ in range(world):
            assert mine[s_] == expect
            expect += (sizes[s_] + 1) * 16
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:
        q.put((rank, traceback.format_exc()))


def test_halo_plans_and_window_layout_gloo_world3():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker3, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}:\n{msg}"


def test_distributed_plumbing_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}:\n{msg}"


def test_partitioned_oracle_reduces_to_single_domain_and_converges():
    A = oml.poisson((24, 20))
    n = A.shape[0]
    lam = [2.0, 1.9, 1.8, 1.7]
    one, offs = oml.build_hierarchy_partitioned(A, [0, n], ratio=0.1, lam_max=lam, max_coarse=20, replicate_below=100)
    ref = oml.build_hierarchy(A, ratio=0.1, lam_max=lam, max_coarse=20)
    assert len(one) == len(ref)
    for a, b in zip(one, ref):
        assert (a.A != b.A).nnz == 0
    two, offs2 = oml.build_hierarchy_partitioned(A, [0, 240, n], ratio=0.1, lam_max=lam, max_coarse=20, replicate_below=100)
    assert len(offs2[0]) == 3 and offs2[1][-1] == two[1].A.shape[0]
    # aggregates never straddle the partition
    lab = two[0].labels
    assert lab[:240].max() < offs2[1][1] <= lab[240:].min()
    b = np.random.RandomState(0).randn(n)
    x, res, it = oml.pcg(two, b, tol=1e-8)
    assert res[-1] <= 1e-8 * np.linalg.norm(b)
