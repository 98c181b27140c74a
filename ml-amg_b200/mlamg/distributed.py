"""Row-partitioned multi-GPU levels: one process per GPU, contiguous row blocks,
halo exchange over NCCL (NVLink 5 / NVSwitch) overlapped with the interior rows, coarse levels
agglomerated (replicated on every rank) below a size threshold.  The reference has no distributed
solve (its only parallelism is a task farm of independent CPU solves, ns/parallel/), so the distributed
algorithm is DEFINED here and mirrored on the CPU by `oracle.multilevel.build_hierarchy(partition=…)`:

  * aggregation: Lloyd on each rank's diagonal block (aggregates never straddle ranks; coarse dofs are
    numbered rank by rank, so every coarse level is again a contiguous row partition);
  * P = (I - w D^-1 A) Agg and A_H = P^T A P are the GLOBAL matrices, each rank computing its rows with
    the same ordered SpGEMM (rows keep their ascending-global stored order, so values are bit-identical
    to the single-domain scipy computation on the same aggregates);
  * cycle: per operator application one halo exchange of vector entries (`all_to_all_single` =
    grouped ncclSend/ncclRecv) issued on a side stream while the interior rows run; boundary rows
    follow.  Norms/dots: one 8-byte allreduce.  Below `replicate_below` global rows the level is
    all-gathered and every rank runs the single-GPU hierarchy for the levels beneath it (one
    all-gather of the restricted residual per cycle, no further communication).

The plumbing below (partition, halo plans, row fetch, distributed transpose) is plain torch tensor
code and works on CPU tensors with the gloo backend — that is how `tests/test_dist_cpu.py` covers it
with world_size 2.  Everything numerical runs in libmlamg_b200.so.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import core
from . import hierarchy as hmod
from . import _lib
from ._lib import lib, check


# =====================================================================================================
# communicator helpers
# =====================================================================================================
import os as _os
import time as _time

# overlap the halo exchange only when the interior kernel is long compared with an exchange (~40 us)
OVERLAP_MIN_NNZ = int(_os.environ.get("MLAMG_OVERLAP_MIN_NNZ", 20_000_000))


class Comm:
    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def all_gather_int(self, v):
        """python int per rank -> list of ints"""
        if self.world == 1:
            return [int(v)]
        out = [None] * self.world
        dist.all_gather_object(out, int(v), group=self.group)
        return out

    def all_gather_obj(self, obj):
        """picklable object per rank -> list in rank order"""
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def a2a(self, send, send_splits, recv_splits, out=None):
        """variable all-to-all of a 1-D tensor; splits are python int lists (elements per peer)"""
        n_recv = int(sum(recv_splits))
        if out is None:
            out = torch.empty(n_recv, dtype=send.dtype, device=send.device)
        if self.world == 1:
            if n_recv:
                out.copy_(send)
            return out
        dist.all_to_all_single(out, send.contiguous(), list(recv_splits), list(send_splits), group=self.group)
        return out

    def a2a_counts(self, counts):
        """counts[p] = elements I send to p  ->  elements I receive from each p"""
        if self.world == 1:
            return list(counts)
        t = torch.tensor(counts, dtype=torch.int64)
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        t = t.to(dev)
        out = torch.empty_like(t)
        dist.all_to_all_single(out, t, group=self.group)
        return [int(v) for v in out.cpu()]

    def all_gather_cat(self, t):
        """concatenate 1-D tensors of different length in rank order"""
        if self.world == 1:
            return t
        sizes = self.all_gather_int(t.numel())
        pad = max(max(sizes), 1)                      # equal-size collective (gloo and NCCL both take it)
        mine = torch.zeros(pad, dtype=t.dtype, device=t.device)
        mine[:t.numel()] = t
        outs = [torch.empty(pad, dtype=t.dtype, device=t.device) for _ in sizes]
        dist.all_gather(outs, mine, group=self.group)
        return torch.cat([o[:s] for o, s in zip(outs, sizes)])

    def allreduce_sum(self, value):
        if self.world == 1:
            return float(value)
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, group=self.group)
        return float(t.item())

    def barrier(self):
        if self.world > 1:
            dist.barrier(group=self.group)


def bind_cpu_affinity(device_index):
    """Pin this process to the CPU cores NVML reports as local to its GPU (one process per GPU): pinned host
    buffers allocated afterwards land on the GPU's NUMA node, so the host<->device copies of several ranks do
    not cross the socket interconnect.  Returns the core list, or None when NVML / affinity is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:                                           # noqa: BLE001
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (_os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = _os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            _os.sched_setaffinity(0, cpus)
        return cpus or None
    except Exception:                                               # noqa: BLE001
        return None


def partition_offsets(n_local, comm):
    counts = comm.all_gather_int(n_local)
    return np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)


def owners_of(ids, offsets):
    """rank owning each global id (ids: int tensor)"""
    bounds = torch.as_tensor(offsets[1:], dtype=torch.int64, device=ids.device)
    return torch.searchsorted(bounds, ids.to(torch.int64), right=True)


# =====================================================================================================
# CSR plumbing on raw tensors (rowptr int32, col int32 GLOBAL ids, val)
# =====================================================================================================
def row_lengths(rowptr):
    return (rowptr[1:] - rowptr[:-1]).to(torch.int64)


def rowptr_from_lengths(lengths):
    rp = torch.zeros(lengths.numel() + 1, dtype=torch.int64, device=lengths.device)
    rp[1:] = torch.cumsum(lengths, 0)
    return rp.to(torch.int32)


def localize(col_global, n_own, lo, offsets=None):
    """global column ids -> local ext numbering [owned 0..n_own) | halo n_own..); stored order kept.
    Returns (col_local int32, halo_ids int64 sorted unique)."""
    c = col_global.to(torch.int64)
    own = (c >= lo) & (c < lo + n_own)
    halo_ids = torch.unique(c[~own])           # sorted
    loc = torch.where(own, c - lo, n_own + torch.searchsorted(halo_ids, c))
    return loc.to(torch.int32), halo_ids


def csr_vstack(parts):
    """stack row blocks [(rowptr, col, val), …] (same column space)"""
    lens = torch.cat([row_lengths(p[0]) for p in parts])
    return rowptr_from_lengths(lens), torch.cat([p[1] for p in parts]), torch.cat([p[2] for p in parts])


def select_rows(rowptr, col, val, rows):
    """rows (int64 local indices, any order) -> CSR of those rows"""
    lens = row_lengths(rowptr)[rows]
    starts = rowptr.to(torch.int64)[rows]
    total = int(lens.sum())
    excl = torch.cumsum(lens, 0) - lens
    ent = torch.repeat_interleave(starts - excl, lens) + torch.arange(total, device=col.device)
    return rowptr_from_lengths(lens), col[ent], val[ent], lens


class HaloPlan:
    """Who sends which owned entries to whom so that x_ext[n_own:] = x_global[halo_ids]."""

    def __init__(self, halo_ids, offsets, comm):
        self.comm = comm
        self.halo_ids = halo_ids
        self.offsets = offsets
        self.n_halo = int(halo_ids.numel())
        lo = int(offsets[comm.rank])
        own = owners_of(halo_ids, offsets)
        self.recv_counts = [int(v) for v in torch.bincount(own, minlength=comm.world).cpu()]
        self.send_counts = comm.a2a_counts(self.recv_counts)
        req = comm.a2a(halo_ids.to(torch.int64), self.recv_counts, self.send_counts)   # ids others need from me
        self.send_idx = (req - lo).to(torch.int32).contiguous()
        self.n_send = int(self.send_idx.numel())
        self._buf = {}

    def exchange(self, x_ext, n_own):
        """fill x_ext[n_own:] from the owners (x_ext[:n_own] must be final)"""
        if self.comm.world == 1:
            return          # (never skipped on world > 1: the all-to-all is a collective)
        key = (x_ext.dtype, x_ext.device)
        buf = self._buf.get(key)
        if buf is None:
            buf = self._buf[key] = torch.empty(max(self.n_send, 1), dtype=x_ext.dtype, device=x_ext.device)
        send = buf[:self.n_send]
        if x_ext.is_cuda and x_ext.dtype in (torch.float32, torch.float64):
            check(lib.mlamg_gather(core.dt(x_ext), self.n_send, core.ptr(self.send_idx), core.ptr(x_ext), core.ptr(send),
                                   core.stream()))
        else:
            torch.index_select(x_ext[:n_own], 0, self.send_idx.long(), out=send)
        self.comm.a2a(send, self.send_counts, self.recv_counts, out=x_ext[n_own:n_own + self.n_halo])


# =====================================================================================================
# peer-memory exchange (CUDA IPC windows + device-side flags): csrc/peer.cu
# =====================================================================================================
# split a row-op into interior / boundary rows around the wait only when the interior kernel is long
# compared with the exchange latency (a few microseconds over NVLink)
PEER_SPLIT_MIN_NNZ = int(_os.environ.get("MLAMG_PEER_SPLIT_MIN_NNZ", 2_000_000))
# 1: the push kernel runs on a side stream next to the interior rows (forked/joined inside the captured graph)
PEER_FORK_PUSH = _os.environ.get("MLAMG_PEER_FORK_PUSH", "1") == "1"
# 1: the consuming row-op reads halo values in place from the receive region; 0: unpack kernel + plain row-op
PEER_INPLACE = _os.environ.get("MLAMG_PEER_INPLACE", "1") == "1"
# boundary rows (the HALO kernel, which spins on values that have not landed) on the forked high-priority stream right
# behind the push, BESIDE the interior rows instead of after them: their few CTAs are scheduled ahead of the interior
# kernel's queued CTAs and wait for the neighbours' stores while the interior rows keep the SMs busy
# 2: boundary rows on a second, normal-priority side stream enqueued behind the interior launch (they then run in the
# interior kernel's tail wave, when the halo has long arrived, instead of spinning on SM slots at the start)
PEER_BOUNDARY_SIDE = int(_os.environ.get("MLAMG_PEER_BOUNDARY_SIDE", "1"))
# 1: when everything an operator sends is written by the boundary rows of the operator before it (residual -> restriction),
# its push is issued right behind that boundary kernel on the side stream instead of after the previous operator's join
EARLY_PUSH = _os.environ.get("MLAMG_EARLY_PUSH", "1") == "1"
# rows per rank from which the interior rows of the residual / prolongation+post-sweep run on W32 copies (0 = never)
W32_MIN_ROWS = int(_os.environ.get("MLAMG_W32_MIN_ROWS", "100000"))
_ALIGN = 256


def _align(v):
    return (int(v) + _ALIGN - 1) // _ALIGN * _ALIGN


def plan_channels(specs, comm):
    """Window layout of a list of channel specs and the remote offsets each rank writes to.

    specs: [(send_counts[world], recv_counts[world], slot_bytes)], identical in number and order on every
    rank.  Pure host logic (two all_gather_object); returns
      {'nbytes': window size,
       'region': [(off_parity0, off_parity1)] per channel (local byte offsets),
       'remote': remote[p][c] = (off_parity0, off_parity1) in rank p's window where MY segment starts}"""
    world, me = comm.world, comm.rank
    nch = len(specs)
    off = 0
    region, table = [], []
    for send_counts, recv_counts, slot in specs:
        assert len(send_counts) == world and len(recv_counts) == world
        size = _align(max(int(sum(recv_counts)), 1) * slot)
        region.append((off, off + size))
        starts = np.concatenate([[0], np.cumsum(recv_counts)])[:-1]
        table.append([(off + int(st) * slot, off + size + int(st) * slot) for st in starts])   # per source rank
        off += 2 * size
    tables = comm.all_gather_obj(table)           # tables[p][c][source] = offsets in p's window
    counts = comm.all_gather_obj([list(map(int, sp_[1])) for sp_ in specs])     # counts[p][c][source]
    for c, (send_counts, _, _) in enumerate(specs):
        for p in range(world):
            if int(send_counts[p]) != counts[p][c][me]:
                raise RuntimeError(f"channel {c}: rank {me} sends {send_counts[p]} entries to {p}, which expects "
                                   f"{counts[p][c][me]}")
    remote = [[tables[p][c][me] for c in range(nch)] for p in range(world)]
    return {"nbytes": max(off, _ALIGN), "region": region, "remote": remote}


class PeerUnavailable(RuntimeError):
    """CUDA IPC windows could not be set up on every rank (raised on ALL ranks together)"""


class PeerWindow:
    """One cudaMalloc block per rank, exported with CUDA IPC and mapped by every peer (collective).
    Failures are detected collectively: either every rank gets a usable window or every rank raises PeerUnavailable."""

    def __init__(self, comm, nbytes):
        self.comm = comm
        self.nbytes = int(nbytes)
        self.base, self.ptrs = None, []
        p = ctypes.c_void_p()
        h = ctypes.create_string_buffer(64)
        err = None
        try:
            if _os.environ.get("MLAMG_TEST_PEER_FAIL") == str(comm.rank):      # failure injection for the fallback test
                raise RuntimeError("injected failure (MLAMG_TEST_PEER_FAIL)")
            check(lib.mlamg_peer_alloc(self.nbytes, ctypes.byref(p), h))
            self.base = p.value
        except Exception as exc:                                   # noqa: BLE001
            err = f"rank {comm.rank}: {exc}"
        handles = comm.all_gather_obj(None if err else h.raw)      # every window is allocated and zeroed before anyone maps it
        if any(x is None for x in handles):
            self._release_local()
            raise PeerUnavailable(err or "a peer could not allocate its window")
        opened = []
        try:
            for r, raw in enumerate(handles):
                if r == comm.rank:
                    self.ptrs.append(self.base)
                else:
                    q = ctypes.c_void_p()
                    check(lib.mlamg_peer_open(ctypes.create_string_buffer(raw, 64), ctypes.byref(q)))
                    self.ptrs.append(q.value)
                    opened.append(q.value)
        except Exception as exc:                                   # noqa: BLE001
            err = f"rank {comm.rank}: {exc}"
        oks = comm.all_gather_obj(err is None)
        if not all(oks):
            for q in opened:
                lib.mlamg_peer_close(ctypes.c_void_p(q))
            comm.barrier()
            self._release_local()
            raise PeerUnavailable(err or "a peer could not map the windows")
        comm.barrier()

    def _release_local(self):
        if self.base is not None:
            lib.mlamg_peer_free(ctypes.c_void_p(self.base))
        self.base, self.ptrs = None, []

    def close(self):
        """collective: unmap the peers' windows, then free the local one"""
        if self.base is None:
            return
        torch.cuda.synchronize()
        for r, q in enumerate(self.ptrs):
            if r != self.comm.rank:
                check(lib.mlamg_peer_close(ctypes.c_void_p(q)))
        self.comm.barrier()
        check(lib.mlamg_peer_free(ctypes.c_void_p(self.base)))
        self.base, self.ptrs = None, []


class Channel:
    """One exchange step bound to its slots in the windows (see csrc/peer.cu)."""

    def __init__(self, handle, send_idx, state, n_recv, name):
        self._h, self.send_idx, self.state, self.n_recv, self.name = handle, send_idx, state, n_recv, name

    def push(self, src, scale=None):
        """pack src[send_idx] (.* scale[send_idx]) into the peers' regions (tagged values), advance the sequence number"""
        check(lib.mlamg_channel_push(self._h, core.dt(src), core.ptr(self.send_idx), core.ptr(src), core.ptr(scale),
                                     core.stream()))

    def unpack(self, dst, dst_idx=None):
        """after push: wait for every slot and copy it to dst (optionally through an index map)"""
        assert dst_idx is not None or dst.numel() == self.n_recv
        check(lib.mlamg_channel_unpack(self._h, core.dt(dst), core.ptr(dst_idx), ctypes.c_void_p(dst.data_ptr()),
                                       core.stream()))

    def rowop(self, A, op, x_ext, n_own, y, b=None, dw=None, rows=None, row_range=None, aux=None):
        """after push: core.rowop whose gathers of halo columns read the receive region in place"""
        begin = 0
        if rows is not None:
            n = rows.numel()
        elif row_range is not None:
            begin, n = int(row_range[0]), int(row_range[1] - row_range[0])
        else:
            n = A.shape[0]
        if n <= 0:
            return y
        check(lib.mlamg_channel_rowop(self._h, core.dt(A.val), op, n, max(1, int(A.nnz * n / max(A.shape[0], 1))),
                                      core.ptr(A.rowptr), core.ptr(A.col), core.ptr(A.val), core.ptr(x_ext), int(n_own),
                                      core.ptr(b), core.ptr(dw), core.ptr(y), core.ptr(aux), core.ptr(rows), begin,
                                      core.stream()))
        return y


class ChannelSet:
    """All channels of one cycle shape in one window (collective constructor).

    specs: [(name, send_idx int32 cuda tensor | None, send_counts[world], recv_counts[world], dtype)]"""

    def __init__(self, comm, specs):
        self.comm = comm
        world = comm.world
        if world > 16:
            raise ValueError("peer channels support at most 16 ranks per node")
        slot = {torch.float32: int(lib.mlamg_channel_slot_bytes(_lib.F32)),
                torch.float64: int(lib.mlamg_channel_slot_bytes(_lib.F64))}
        lay = plan_channels([(sc, rc, slot[dt_]) for _, _, sc, rc, dt_ in specs], comm)
        self.window = PeerWindow(comm, lay["nbytes"])
        W = self.window
        self.state = torch.zeros(len(specs), 4, dtype=torch.int64, device="cuda")
        self.channels = {}
        VP = ctypes.c_void_p
        for c, (name, send_idx, sc, rc, dt_) in enumerate(specs):
            sp_ = [p for p in range(world) if int(sc[p]) > 0]
            s_counts = (ctypes.c_int * max(len(sp_), 1))(*[int(sc[p]) for p in sp_])
            s_dst0 = (VP * max(len(sp_), 1))(*[W.ptrs[p] + lay["remote"][p][c][0] for p in sp_])
            s_dst1 = (VP * max(len(sp_), 1))(*[W.ptrs[p] + lay["remote"][p][c][1] for p in sp_])
            n_recv = int(sum(int(v) for v in rc))
            h = VP()
            check(lib.mlamg_channel_create(len(sp_), s_counts, s_dst0, s_dst1, n_recv, VP(W.base + lay["region"][c][0]),
                                           VP(W.base + lay["region"][c][1]), VP(self.state[c].data_ptr()), ctypes.byref(h)))
            n_send = int(sum(int(sc[p]) for p in sp_))
            assert send_idx is None or send_idx.numel() == n_send
            self.channels[name] = Channel(h, send_idx, self.state[c], n_recv, name)
        torch.cuda.synchronize()
        comm.barrier()

    def __getitem__(self, name):
        return self.channels[name]

    def check(self):
        """raise if a wait timed out (host sync)"""
        err = self.state[:, 3].cpu()
        bad = torch.nonzero(err).flatten().tolist()
        if bad:
            names = list(self.channels)
            raise RuntimeError("peer exchange timed out on channel(s) " + ", ".join(str(names[i]) for i in bad))

    def close(self):
        for ch in self.channels.values():
            lib.mlamg_channel_destroy(ch._h)
        self.channels = {}
        self.window.close()


def fetch_rows(rowptr, col, val, offsets, needed_ids, comm):
    """CSR rows `needed_ids` (sorted unique global row ids owned by other ranks), in that order."""
    plan = HaloPlan(needed_ids, offsets, comm)
    _, scol, sval, slens = select_rows(rowptr, col, val, plan.send_idx.long())
    # rows per peer are contiguous in send order; entries per peer = sum of their lengths
    lens_recv = comm.a2a(slens, plan.send_counts, plan.recv_counts)
    def per_peer(lens, counts):
        l = lens.cpu().numpy()
        out, k = [], 0
        for n in counts:
            out.append(int(l[k:k + n].sum()))
            k += n
        return out
    send_e = per_peer(slens, plan.send_counts)
    recv_e = per_peer(lens_recv, plan.recv_counts)
    rcol = comm.a2a(scol, send_e, recv_e)
    rval = comm.a2a(sval, send_e, recv_e)
    return rowptr_from_lengths(lens_recv), rcol, rval


def dist_transpose(rowptr, col, val, row_offsets, col_offsets, comm):
    """Row-partitioned M (global column ids) -> row-partitioned M^T (rows = my block of the column
    partition, global column ids = row ids of M), sorted rows."""
    rank = comm.rank
    n_own = rowptr.numel() - 1
    lo_r = int(row_offsets[rank])
    lo_c, hi_c = int(col_offsets[rank]), int(col_offsets[rank + 1])
    n_rows_global = int(row_offsets[-1])
    rows = torch.repeat_interleave(torch.arange(n_own, device=col.device, dtype=torch.int64) + lo_r, row_lengths(rowptr))
    c = col.to(torch.int64)
    dest = owners_of(c, col_offsets)
    order = torch.argsort(dest, stable=True)
    counts = [int(v) for v in torch.bincount(dest, minlength=comm.world).cpu()]
    rcounts = comm.a2a_counts(counts)
    r_rows = comm.a2a(rows[order], counts, rcounts)
    r_cols = comm.a2a(c[order], counts, rcounts)
    r_vals = comm.a2a(val[order], counts, rcounts)
    j = r_cols - lo_c
    key = j * n_rows_global + r_rows
    perm = torch.argsort(key, stable=True)
    lens = torch.bincount(j, minlength=hi_c - lo_c)
    return rowptr_from_lengths(lens), r_rows[perm].to(torch.int32), r_vals[perm]


def diag_block(rowptr, col, val, n_own, lo):
    """entries whose column is owned, as a local CSR (the graph each rank aggregates on)"""
    c = col.to(torch.int64)
    own = (c >= lo) & (c < lo + n_own)
    rows = torch.repeat_interleave(torch.arange(n_own, device=col.device), row_lengths(rowptr))
    lens = torch.bincount(rows[own], minlength=n_own)
    return rowptr_from_lengths(lens), (c[own] - lo).to(torch.int32), val[own]


def gather_csr(rowptr, col, val, comm):
    """all ranks receive the whole matrix (rows in rank order, global column ids)"""
    lens = comm.all_gather_cat(row_lengths(rowptr))
    return rowptr_from_lengths(lens), comm.all_gather_cat(col), comm.all_gather_cat(val)


# =====================================================================================================
# distributed level + hierarchy (GPU)
# =====================================================================================================
_LOW_STREAM = {}


def _low_stream():
    d = torch.cuda.current_device()
    if d not in _LOW_STREAM:
        _LOW_STREAM[d] = torch.cuda.Stream(priority=0)
    return _LOW_STREAM[d]


class DistOperator:
    """Local rows of a row-partitioned operator, columns in ext numbering, with its halo plan."""

    def __init__(self, rowptr, col_global, val, n_cols_own, col_lo, col_offsets, comm):
        col_loc, halo = localize(col_global, n_cols_own, col_lo)
        self.plan = HaloPlan(halo, col_offsets, comm)
        self.n_rows = rowptr.numel() - 1
        self.n_cols_own = n_cols_own
        self.n_ext = n_cols_own + self.plan.n_halo
        self.csr = core.DeviceCSR(rowptr.contiguous(), col_loc.contiguous(), val.contiguous(), (self.n_rows, self.n_ext))
        self._split_rows()

    def _split_rows(self):
        """interior rows (no halo column) / boundary rows, as lists and — when contiguous — as ranges"""
        A = self.csr
        dev = A.col.device
        rows = torch.repeat_interleave(torch.arange(self.n_rows, device=dev), row_lengths(A.rowptr))
        brows = torch.unique(rows[A.col.long() >= self.n_cols_own])
        mask = torch.ones(self.n_rows, dtype=torch.bool, device=dev)
        mask[brows] = False
        self.interior = torch.nonzero(mask).flatten().to(torch.int32).contiguous()
        self.boundary = brows.to(torch.int32).contiguous()
        # contiguous row partitions of banded operators have a contiguous interior: [i0, i1) can then be run
        # as a plain row RANGE (coalesced, no index list) and the boundary as the two ranges around it
        self.interior_range = None
        if self.interior.numel() > 0:
            i0, i1 = int(self.interior[0]), int(self.interior[-1]) + 1
            if i1 - i0 == self.interior.numel():
                self.interior_range = (i0, i1)
        # overlap only pays when the interior kernel is much longer than an exchange (~40 us)
        self.overlap_ok = self.interior.numel() > 0 and A.nnz >= OVERLAP_MIN_NNZ
        self.peer_split_ok = self.interior.numel() > 0 and A.nnz >= PEER_SPLIT_MIN_NNZ

    def renumber(self, row_new2old=None, col_old2new=None):
        """Block-local renumbering: rows gathered by row_new2old, OWNED columns mapped by col_old2new (halo slots
        keep their place in the ext vector; the entries this rank SENDS are re-addressed instead)."""
        if col_old2new is not None:
            n_own = self.n_cols_own
            cmap = torch.cat([col_old2new.to(torch.int64),
                              torch.arange(n_own, self.n_ext, device=col_old2new.device, dtype=torch.int64)])
            self.plan.send_idx = col_old2new[self.plan.send_idx.long()].to(torch.int32).contiguous()
        else:
            cmap = None
        self.csr = hmod._permuted(self.csr, row_new2old, cmap)
        self._split_rows()

    def rowop(self, op, x_ext, y, b=None, dw=None, rows=None, row_range=None, aux=None):
        core.rowop(self._csr_for(op), op, x_ext, y, b=b, dw=dw, rows=rows, row_range=row_range, aux=aux)

    def _csr_for(self, op):
        return self.csr_scaled if op in (6, 8) else self.csr

    def build_scaled(self, dw):
        """column-scaled copy A D_w (op 6): dw of the halo columns comes from their owners (one exchange at setup)"""
        dw_ext = torch.zeros(self.n_ext, dtype=dw.dtype, device=dw.device)
        dw_ext[:self.n_cols_own] = dw
        self.plan.exchange(dw_ext, self.n_cols_own)
        self.csr_scaled = self.csr.with_values(core.scaled_values(self.csr, dw_ext))

    def build_w32(self, scaled_only=False):
        """W32 copies (core.csr_to_w32) of the local rows for the interior-row kernels of the cycle"""
        self.w32 = getattr(self, "w32", {})
        if not scaled_only:
            self.w32["csr"] = core.csr_to_w32(self.csr)
        if hasattr(self, "csr_scaled"):
            self.w32["scaled"] = core.csr_to_w32(self.csr_scaled)

    def channel_spec(self, name, dtype):
        """(name, send list, per-rank counts, dtype) of the halo exchange of this operator's input"""
        return (name, self.plan.send_idx, self.plan.send_counts, self.plan.recv_counts, dtype)

    def apply(self, op, x_ext, y, b=None, dw=None, overlap=True, comm_stream=None, chan=None, aux=None, after_boundary=None,
              prepushed=False):
        """halo exchange of x_ext, then the row-op; with overlap the interior rows run during the exchange.
        chan: peer-memory channel (push -> interior rows -> boundary rows reading the halo in place from the
        receive region; one stream, no collective); None: NCCL all-to-all on a side stream.
        op 4 / 6 (x = dw.*b, y = b - A x; 6 on the column-scaled copy): x_ext is the OUTPUT x, the neighbours receive
        dw.*b (4) or b (6) directly.
        op 8 (y = b - (A D_w) b on the scaled copy, x never materialised): x_ext is ignored.
        op 5 (y = aux + dw.*b + A x_ext): aux = iterate before the correction, b = residual, x_ext = coarse correction.
        op 7 (y = dw.*(aux + b) + A x_ext): aux = right-hand side, b = residual of dw.*rhs.
        after_boundary: callable run on the side stream right behind the boundary-row kernel (the NEXT operator's push
        when everything it sends is produced by these boundary rows: its halo then travels while the interior rows
        of THIS operator are still running); returns True when it was run.  prepushed: this operator's push has
        already been issued that way."""
        plan = self.plan
        xin = None if op in (4, 6) else x_ext      # gather vector (ops 4/6 gather b instead)
        if op in (4, 6):
            aux = x_ext
        kop = op
        if op == 8:                                # plain residual kernel on the scaled values, gathering b
            kop, xin = 2, b
        if plan.comm.world == 1:
            core.rowop(self._csr_for(op), kop, xin, y, b=b, dw=dw, aux=aux)
            return
        if chan is not None:
            n_own = self.n_cols_own
            split = overlap and self.peer_split_ok
            fork = split and PEER_FORK_PUSH and comm_stream is not None
            main = torch.cuda.current_stream()
            if fork:                   # the 4 us push kernel runs beside the interior rows instead of in front of them
                comm_stream.wait_stream(main)
            if not prepushed:
                with torch.cuda.stream(comm_stream if fork else main):
                    if op == 4:
                        chan.push(b, scale=dw)
                    elif op in (6, 8):
                        chan.push(b)
                    else:
                        chan.push(x_ext)
            elif not fork:             # pushed early on the side stream, but this operator runs unsplit on the main stream
                main.wait_stream(comm_stream)
            # (prepushed and fork: the wait above orders this operator's boundary kernel behind the whole previous
            # operator — its rows gather values the previous interior kernel wrote — while the push itself left earlier)
            A = self._csr_for(op)
            xk = b if op == 8 else xin
            rows = self.boundary if split else None
            side = PEER_BOUNDARY_SIDE if (fork and (PEER_INPLACE or op in (4, 6, 8))) else 0
            ran_after = False
            if side == 1:              # boundary rows right behind the push on the side stream, beside the interior rows
                with torch.cuda.stream(comm_stream):
                    chan.rowop(A, kop, xk, n_own, y, b=b, dw=dw, rows=rows, aux=aux)
                    if after_boundary is not None:
                        after_boundary()
                        ran_after = True
            if split:
                w32 = getattr(self, "w32", {}).get("scaled" if op in (6, 8) else "csr")
                if w32 is not None and self.interior_range is not None and kop in (0, 2, 5, 7):
                    # thread-per-row kernel on the W32 (warp-interleaved) copy: coalesced operator stream
                    core.rowop_w32(A, w32, kop, xk, y, b=b, dw=dw, aux=aux, row_range=self.interior_range)
                else:
                    core.rowop(A, kop, xk, y, b=b, dw=dw, aux=aux, row_range=self.interior_range,
                               rows=None if self.interior_range is not None else self.interior)
            if side == 2:              # boundary rows on a normal-priority stream, enqueued behind the interior launch
                lo = _low_stream()
                lo.wait_stream(comm_stream)
                with torch.cuda.stream(lo):
                    chan.rowop(A, kop, xk, n_own, y, b=b, dw=dw, rows=rows, aux=aux)
                main.wait_stream(lo)
            if fork:
                main.wait_stream(comm_stream)
            if side:
                return ran_after
            if PEER_INPLACE or op in (4, 6, 8):
                chan.rowop(A, kop, xk, n_own, y, b=b, dw=dw, rows=rows, aux=aux)
            else:
                chan.unpack(x_ext[n_own:n_own + plan.n_halo])
                core.rowop(A, kop, xk, y, b=b, dw=dw, rows=rows, aux=aux)
            return
        if op in (4, 6, 8):
            raise ValueError("the fused zero-guess sweep + residual needs the peer transport (halo columns of b, dw)")
        if not overlap or comm_stream is None or not self.overlap_ok:
            plan.exchange(x_ext, self.n_cols_own)
            self.rowop(op, x_ext, y, b, dw, aux=aux)
            return
        main = torch.cuda.current_stream()
        comm_stream.wait_stream(main)
        with torch.cuda.stream(comm_stream):
            plan.exchange(x_ext, self.n_cols_own)
        if self.interior_range is not None:
            i0, i1 = self.interior_range
            self.rowop(op, x_ext, y, b, dw, row_range=(i0, i1), aux=aux)
            main.wait_stream(comm_stream)
            self.rowop(op, x_ext, y, b, dw, row_range=(0, i0), aux=aux)
            self.rowop(op, x_ext, y, b, dw, row_range=(i1, self.n_rows), aux=aux)
        else:
            self.rowop(op, x_ext, y, b, dw, rows=self.interior, aux=aux)
            main.wait_stream(comm_stream)
            self.rowop(op, x_ext, y, b, dw, rows=self.boundary, aux=aux)


class DistLevel:
    pass


class _StageTimer:
    """per-stage wall-clock of the distributed setup (MLAMG_SETUP_PROFILE=1: device-synchronised; off: no-op)"""

    def __init__(self):
        self.on = _os.environ.get("MLAMG_SETUP_PROFILE", "0") == "1"
        self.acc = {}
        self._t = None

    def start(self):
        if self.on:
            torch.cuda.synchronize()
            self._t = _time.perf_counter()

    def lap(self, name):
        if self.on:
            torch.cuda.synchronize()
            now = _time.perf_counter()
            self.acc[name] = self.acc.get(name, 0.0) + now - self._t
            self._t = now


class DistHierarchy:
    """Distributed levels followed by a replicated single-GPU hierarchy."""

    def __init__(self, rowptr, col_global, val, comm=None, *, ratio=0.1, distance="unit", maxiter=10, rand=0,
                 lam_max=None, max_levels=10, max_coarse=500, replicate_below=200000, smoother="jacobi",
                 jacobi_weight=2.0 / 3.0, overlap=True, renumber=True, halo=None, fuse_post=True, fuse_pre=True):
        core.require_cuda()
        self.comm = comm or Comm()
        comm = self.comm
        self.dtype = val.dtype
        self.overlap = overlap
        # halo transport of the cycle: 'peer' = CUDA-IPC windows + device flags (csrc/peer.cu), 'nccl' = all-to-all
        self.halo = halo or _os.environ.get("MLAMG_HALO", "peer")
        if self.halo not in ("peer", "nccl"):
            raise ValueError("halo must be 'peer' or 'nccl'")
        self._chansets = {}
        # Q = (I - D_w A) P per level: prolongation + first post-smoothing sweep as one pass (see Hierarchy)
        self.fuse_post = bool(fuse_post)
        self.fuse_pre = bool(fuse_pre)
        # high priority: the tiny push kernels forked onto it are scheduled ahead of the queued CTAs of the interior kernel
        self.comm_stream = torch.cuda.Stream(priority=-1) if comm.world > 1 else None
        self.levels = []
        self.offsets = []
        self.setup_profile = _StageTimer()
        lvl = 0
        cur = (rowptr, col_global, val)
        offs = partition_offsets(rowptr.numel() - 1, comm)
        while True:
            n_glob = int(offs[-1])
            if n_glob <= replicate_below or lvl >= max_levels - 1:
                break
            L, nxt, coffs = self._build_level(cur, offs, lvl, ratio, distance, maxiter, rand, lam_max, smoother,
                                              jacobi_weight)
            self.levels.append(L)
            self.offsets.append(offs)
            cur, offs = nxt, coffs
            lvl += 1
        self.offsets.append(offs)
        # replicated tail
        T_ = self.setup_profile
        T_.start()
        g_rowptr, g_col, g_val = gather_csr(*cur, comm)
        T_.lap("tail_gather")
        n_glob = int(offs[-1])
        Ag = core.DeviceCSR(g_rowptr, g_col.to(torch.int32), g_val, (n_glob, n_glob))
        lam_tail = lam_max
        if isinstance(lam_max, (list, tuple)):
            rest = list(lam_max[lvl:])
            lam_tail = (lambda A, _r=rest, _k=[0]: (_r[_k[0]] if _k[0] < len(_r) and _r[_k[0]] is not None
                                                     else core.lambda_max(A), _k.__setitem__(0, _k[0] + 1))[0])
        self.tail = hmod.build_hierarchy(Ag, aggregates="lloyd", ratio=ratio, distance=distance, maxiter=maxiter, rand=rand,
                                         lam_max=lam_tail, max_levels=max(1, max_levels - lvl), max_coarse=max_coarse,
                                         smoother=smoother, jacobi_weight=jacobi_weight)
        T_.lap("tail_hierarchy")
        self.tail_offsets = offs
        if renumber:
            self._renumber()
        T_.lap("renumber")
        if self.fuse_pre:
            for L in self.levels:
                L.A.build_scaled(L.dw)
        T_.lap("scaled_copy")
        # W32 copies for the thread-per-row interior kernels (levels with enough rows per rank; the 7-entry fine operator is
        # DRAM-bound in plain CSR already): Q always, the scaled operator when its rows are long
        if W32_MIN_ROWS > 0:
            for L in self.levels:
                if L.n < W32_MIN_ROWS or not (self.fuse_pre and self.fuse_post and L.Q is not None):
                    continue
                L.Q.build_w32()
                if L.A.csr.nnz > 12 * L.n:
                    L.A.build_w32(scaled_only=True)

        def _subset(send_idx, rows):      # is everything that is sent written by these (boundary) rows?
            si, bd = send_idx.long(), rows.long()
            return bool(si.numel() == 0 or (bd.numel() > 0 and bool(torch.isin(si, bd).all())))
        for l, L in enumerate(self.levels):
            L.early_R = _subset(L.R.plan.send_idx, L.A.boundary)                  # residual -> restriction
            L.early_next_res = l + 1 < len(self.levels) and _subset(self.levels[l + 1].A.plan.send_idx, L.R.boundary)
            if l > 0:                                                            # last pass of level l -> prolongation of l-1
                Lf = self.levels[l - 1]
                up = (Lf.Q if Lf.Q is not None else Lf.P).plan.send_idx
                L.early_up_Q = L.Q is not None and _subset(up, L.Q.boundary)
                L.early_up_A = _subset(up, L.A.boundary)
        self._alloc()
        T_.lap("alloc")

    # ---------------------------------------------------------------------------------------------
    def _build_level(self, cur, offs, lvl, ratio, distance, maxiter, rand, lam_max, smoother, jacobi_weight):
        comm = self.comm
        rank = comm.rank
        rowptr, colg, val = cur
        n_own = rowptr.numel() - 1
        lo = int(offs[rank])
        L = DistLevel()
        L.n = n_own
        T_ = self.setup_profile
        T_.start()
        # 1. aggregates on the diagonal block
        b_rp, b_col, b_val = diag_block(rowptr, colg, val, n_own, lo)
        G = core.DeviceCSR(b_rp, b_col, b_val, (n_own, n_own))
        T_.lap("diag_block")
        labels, nc_local, roots, seeds = hmod.lloyd_labels(G, ratio=ratio, distance=distance, maxiter=maxiter, rand=rand)
        T_.lap("lloyd")
        coffs = partition_offsets(nc_local, comm)
        clo = int(coffs[rank])
        nc_glob = int(coffs[-1])
        L.labels = labels
        # 2. A in ext numbering, halo labels
        L.A = DistOperator(rowptr, colg, val, n_own, lo, offs, comm)
        T_.lap("dist_operator_A")
        lab_ext = torch.full((L.A.n_ext,), -1, dtype=torch.int32, device=val.device)
        lab_ext[:n_own] = torch.where(labels >= 0, labels + clo, labels)
        L.A.plan.exchange(lab_ext, n_own)
        Agg_ext = core.agg_from_labels(lab_ext, nc_glob, val.dtype)
        # 3. P = (I - w D^-1 A) Agg   (rows: owned fine, columns: GLOBAL coarse ids)
        if lam_max is None:
            lam = self._lambda_max(L)
        elif np.isscalar(lam_max):
            lam = float(lam_max)
        else:
            lam = lam_max[lvl] if lvl < len(lam_max) else None
            lam = self._lambda_max(L) if lam is None else float(lam)
        L.omega_sa = (4.0 / 3.0) / lam
        T_.lap("labels_exchange+lambda")
        Pg = core.drop_zeros(core.spgemm(core.sa_smoother(L.A.csr, L.omega_sa), Agg_ext))
        T_.lap("P_spgemm")
        # 4. Galerkin product in scipy's evaluation order: X^T = A^T P ; A_H^T = P^T X^T ; transpose
        t_rp, t_col, t_val = dist_transpose(rowptr, colg, val, offs, offs, comm)
        T_.lap("transpose_A")
        At = DistOperator(t_rp, t_col, t_val, n_own, lo, offs, comm)
        T_.lap("dist_operator_At")
        f_rp, f_col, f_val = fetch_rows(Pg.rowptr, Pg.col, Pg.val, offs, At.plan.halo_ids, comm)
        e_rp, e_col, e_val = csr_vstack([(Pg.rowptr, Pg.col, Pg.val), (f_rp, f_col, f_val)])
        T_.lap("fetch_P_rows")
        Xt = core.spgemm(At.csr, core.DeviceCSR(e_rp, e_col, e_val, (At.n_ext, nc_glob)))
        T_.lap("AtP_spgemm")
        r_rp, r_col, r_val = dist_transpose(Pg.rowptr, Pg.col, Pg.val, offs, coffs, comm)
        T_.lap("transpose_P")
        L.R = DistOperator(r_rp, r_col, r_val, n_own, lo, offs, comm)
        T_.lap("dist_operator_R")
        f_rp, f_col, f_val = fetch_rows(Xt.rowptr, Xt.col, Xt.val, offs, L.R.plan.halo_ids, comm)
        e_rp, e_col, e_val = csr_vstack([(Xt.rowptr, Xt.col, Xt.val), (f_rp, f_col, f_val)])
        T_.lap("fetch_X_rows")
        AHt = core.spgemm(L.R.csr, core.DeviceCSR(e_rp, e_col, e_val, (L.R.n_ext, nc_glob)))
        T_.lap("RX_spgemm")
        h_rp, h_col, h_val = dist_transpose(AHt.rowptr, AHt.col, AHt.val, coffs, coffs, comm)
        AH = core.drop_zeros(core.DeviceCSR(h_rp, h_col, h_val, (nc_local, nc_glob)))
        T_.lap("transpose_AH")
        # 5. apply structures
        L.P = DistOperator(Pg.rowptr, Pg.col, Pg.val, nc_local, clo, coffs, comm)
        L.P_global = Pg
        L.dw = core.smoother_diag(L.A.csr, smoother, jacobi_weight)
        L.nc = nc_local
        L.Q = None
        T_.lap("dist_operator_P")
        if self.fuse_post:
            # rows of P for every column of my rows of A (owned + halo fine nodes), then Q = (I - D_w A) P
            f_rp, f_col, f_val = fetch_rows(Pg.rowptr, Pg.col, Pg.val, offs, L.A.plan.halo_ids, comm)
            e_rp, e_col, e_val = csr_vstack([(Pg.rowptr, Pg.col, Pg.val), (f_rp, f_col, f_val)])
            Qg = hmod.post_operator(L.A.csr, core.DeviceCSR(e_rp, e_col, e_val, (L.A.n_ext, nc_glob)), L.dw)
            L.Q = DistOperator(Qg.rowptr, Qg.col, Qg.val, nc_local, clo, coffs, comm)
        T_.lap("Q_operator")
        return L, (AH.rowptr, AH.col, AH.val), coffs

    def _renumber(self):
        """Distributed levels l >= 1 get a block-local spatial numbering (dof -> rank of its first owned fine
        node), for the same reason as Hierarchy(renumber=True): the reference's seed numbering is random, which
        makes every gather of a coarse vector a separate L2 sector.  Only the apply operators change."""
        for l in range(1, len(self.levels)):
            Lf, Lc = self.levels[l - 1], self.levels[l]
            R = Lf.R.csr
            dev = R.col.device
            rows = torch.repeat_interleave(torch.arange(R.shape[0], device=dev), row_lengths(R.rowptr))
            cols = R.col.long()
            own = cols < Lf.R.n_cols_own
            key = torch.full((R.shape[0],), 2 ** 62, dtype=torch.int64, device=dev)
            key.scatter_reduce_(0, rows[own], cols[own], reduce="amin")
            order = torch.argsort(key, stable=True)            # new -> old
            inv = torch.empty_like(order)
            inv[order] = torch.arange(order.numel(), device=dev)
            Lf.P.renumber(None, inv)
            if Lf.Q is not None:
                Lf.Q.renumber(None, inv)
            Lf.R.renumber(order, None)
            Lc.A.renumber(order, inv)
            Lc.P.renumber(order, None)
            if Lc.Q is not None:
                Lc.Q.renumber(order, None)
            Lc.R.renumber(None, inv)
            Lc.dw = Lc.dw[order].contiguous()
            Lc.perm_new2old = order

    def _lambda_max(self, L, tol=1e-13, maxiter=6000, info=None):
        """|lambda_max(D^-1 A)| of a row-partitioned level: Lanczos on D^-1/2 A D^-1/2 with the distributed SpMV — the
        same recurrence, start vector (a hash of the GLOBAL row index, so the sequence does not depend on the
        partition), check schedule and stopping rule as csrc/eigen.cu (replaces ARPACK, multigrid.py:105).  Two
        8-byte allreduces per step.  The levels of this path are Galerkin products of a symmetric operator."""
        from scipy.linalg import eigh_tridiagonal
        comm, n = self.comm, L.n
        dev = L.A.csr.val.device
        s = torch.sqrt(core.smoother_diag(L.A.csr, "jacobi", 1.0))                     # 1/sqrt(a_ii)
        lo = int(L.A.plan.offsets[comm.rank]) if hasattr(L.A.plan, "offsets") else 0
        i = (torch.arange(n, device=dev, dtype=torch.int64) + lo) & 0xFFFFFFFF
        h = (i * 2654435761) & 0xFFFFFFFF
        h = h ^ (h >> 15)
        h = (h * 2246822519) & 0xFFFFFFFF
        h = h ^ (h >> 13)
        v = (0.5 + (h & 0xFFFF).to(torch.float64) / 65536.0) * torch.where((h & 0x10000) != 0, -1.0, 1.0)
        v = v.to(self.dtype)
        v /= np.sqrt(comm.allreduce_sum(core.dot(v, v)))
        vprev = torch.zeros_like(v)
        xe = torch.zeros(L.A.n_ext, dtype=self.dtype, device=dev)
        w = torch.empty(n, dtype=self.dtype, device=dev)
        al, be, nxt = [], [], 20
        theta, res_rel, steps = 0.0, 1.0, 0
        for j in range(maxiter):
            torch.mul(v, s, out=xe[:n])
            L.A.apply(0, xe, w, overlap=False)
            w *= s
            a = comm.allreduce_sum(core.dot(v, w))
            al.append(a)
            core.axpby(-a, v, 1.0, w)
            if be:
                core.axpby(-be[-1], vprev, 1.0, w)
            bnorm = np.sqrt(comm.allreduce_sum(core.dot(w, w)))
            be.append(bnorm)
            steps = j + 1
            if steps == nxt or steps == maxiter:
                if steps > 1:
                    ev, evec = eigh_tridiagonal(np.array(al), np.array(be[:-1]))
                else:
                    ev, evec = np.array(al), np.ones((1, 1))
                k = int(np.argmax(np.abs(ev)))
                theta = float(abs(ev[k]))
                gap = float(np.min(np.abs(np.delete(np.abs(ev), k) - theta))) if steps > 1 else 0.0
                res = bnorm * abs(float(evec[-1, k]))
                res_rel = res / theta if theta > 0 else 0.0
                if min(res, res * res / gap if gap > 0 else res) <= tol * theta or bnorm <= 1e-300:
                    break
                nxt = steps + max(10, steps // 4)
            vprev, v = v, vprev
            torch.mul(w, 1.0 / bnorm, out=v)
        if info is not None:
            info.update(residual=res_rel, steps=steps, method="lanczos")
        return theta

    # ---------------------------------------------------------------------------------------------
    def _alloc(self):
        dev, dt_ = "cuda", self.dtype
        for L in self.levels:
            L.x = [torch.zeros(L.A.n_ext, dtype=dt_, device=dev) for _ in range(2)]
            L.b = torch.zeros(L.n, dtype=dt_, device=dev)
            L.r = torch.zeros(L.R.n_ext, dtype=dt_, device=dev)
            L.e = torch.zeros(max(L.P.n_ext, L.Q.n_ext if L.Q is not None else 0), dtype=dt_, device=dev)
        n_tail = int(self.tail_offsets[-1])
        self.tail_b = torch.zeros(n_tail, dtype=dt_, device=dev)
        self.tail_x = torch.zeros(n_tail, dtype=dt_, device=dev)
        self._tail_sizes = [int(self.tail_offsets[i + 1] - self.tail_offsets[i]) for i in range(self.comm.world)]
        # equal-size all-gather of the restricted residual: padded per-rank slots + one index map back
        self._tail_pad = max(max(self._tail_sizes), 1)
        self._tail_in = torch.zeros(self._tail_pad, dtype=dt_, device=dev)
        self._tail_out = torch.zeros(self._tail_pad * self.comm.world, dtype=dt_, device=dev)
        self._tail_map = torch.cat([torch.arange(s, device=dev) + r * self._tail_pad
                                    for r, s in enumerate(self._tail_sizes)]).to(torch.int32).contiguous()
        self._graph = None

    @property
    def n_local(self):
        return self.levels[0].n if self.levels else int(self.tail_offsets[-1])

    def _channels(self, nu1, nu2):
        """peer channels of a V(nu1,nu2) cycle, built on first use (collective); None on the NCCL transport"""
        if self.halo != "peer" or self.comm.world == 1:
            return None
        cs = self._chansets.get((nu1, nu2))
        if cs is None:
            specs = []
            for l, L in enumerate(self.levels):
                for k in range(max(nu1 - 1, 0)):
                    specs.append(L.A.channel_spec((l, "pre", k), self.dtype))
                specs.append(L.A.channel_spec((l, "res"), self.dtype))
                specs.append(L.R.channel_spec((l, "R"), self.dtype))
                specs.append((L.Q if (L.Q is not None and nu2 > 0) else L.P).channel_spec((l, "P"), self.dtype))
                for k in range(nu2):
                    specs.append(L.A.channel_spec((l, "post", k), self.dtype))
            # coarse-level gather: every rank writes its slice of the restricted residual (+ one padding slot, so
            # that every pair of ranks is in step once per cycle even with an empty slice) into every window
            sizes = self._tail_sizes
            mine = sizes[self.comm.rank]
            one = torch.cat([torch.arange(mine, dtype=torch.int32), torch.tensor([-1], dtype=torch.int32)])
            idx = one.repeat(self.comm.world).cuda().contiguous()
            specs.append(("tail", idx, [mine + 1] * self.comm.world, [sz + 1 for sz in sizes], self.dtype))
            offs = self.tail_offsets
            self._tail_unpack_idx = torch.cat(
                [torch.cat([torch.arange(int(offs[r]), int(offs[r + 1]), dtype=torch.int32), torch.tensor([-1], dtype=torch.int32)])
                 for r in range(self.comm.world)]).cuda().contiguous()
            try:
                cs = self._chansets[(nu1, nu2)] = ChannelSet(self.comm, specs)
            except PeerUnavailable as exc:        # raised on every rank together: use the NCCL transport instead
                if self.comm.rank == 0:
                    print(f"mlamg: peer-memory halo transport unavailable ({exc}); using the NCCL all-to-all", flush=True)
                self.halo = "nccl"
                return None
        return cs

    def check_exchange(self):
        """host sync + raise if any peer exchange timed out.  `pcg` calls it before returning; `vcycle` and the replay
        callable of `capture` only ENQUEUE work, so a caller that consumes their result directly must call this after
        its own synchronisation (bench.py and tests/dist_check.py do)."""
        for cs in self._chansets.values():
            cs.check()

    def close(self):
        """collective: release the peer windows"""
        self._pcg_ws = None
        self._graph = None
        for cs in self._chansets.values():
            cs.close()
        self._chansets = {}

    def _first_op(self, l, nu1, nu2, chans):
        """(fused, scaled, lazy) of level l's first pass — does it push the right-hand side itself?"""
        L = self.levels[l]
        scaled = self.fuse_pre and hasattr(L.A, "csr_scaled")
        fused = nu1 == 1 and (chans is not None or self.comm.world == 1) and (scaled or L.A.csr.nnz <= 12 * L.n)
        lazy = fused and scaled and L.Q is not None and nu2 > 0
        return fused, scaled, lazy

    def vcycle(self, b, x_out, nu1=1, nu2=1):
        """x_out = V(nu1,nu2)(b) from a zero guess (preconditioner apply).  b, x_out: local slices.

        Early pushes (peer transport): when everything an operator sends to its neighbours is written by the BOUNDARY rows
        of the operator before it (checked once at setup: residual -> restriction, restriction -> the next level's first
        pass, a level's last pass -> the finer level's prolongation), its push is issued on the side stream right behind
        that boundary kernel; the halo then travels while the previous operator's interior rows are still running."""
        comm = self.comm
        cs = self.comm_stream if self.overlap else None
        chans = self._channels(nu1, nu2)
        ch = (lambda *key: chans[key]) if chans is not None else (lambda *key: None)
        early_ok = chans is not None and EARLY_PUSH
        rhs = b
        cur = []
        nl = len(self.levels)
        rhs_pushed = False
        for l, L in enumerate(self.levels):
            xa, xb = L.x
            n = L.n
            c = xa
            fused, scaled, L.lazy = self._first_op(l, nu1, nu2, chans)
            early = None
            if early_ok and getattr(L, "early_R", False):
                early = (lambda _c=ch(l, "R"), _r=L.r: _c.push(_r))
            r_pushed = False
            if L.lazy:     # x = dw.*b is never materialised: r = b - (A D_w) b, then x1 = dw.*(b + r) + Q e on the way up
                r_pushed = L.A.apply(8, None, L.r, b=rhs, overlap=self.overlap, comm_stream=cs, chan=ch(l, "res"),
                                     after_boundary=early, prepushed=rhs_pushed)
            elif fused:    # x = dw.*b and r = b - A x in one pass over A (x is never read back)
                r_pushed = L.A.apply(6 if scaled else 4, c, L.r, b=rhs, dw=L.dw, overlap=self.overlap, comm_stream=cs,
                                     chan=ch(l, "res"), after_boundary=early, prepushed=rhs_pushed)
            elif nu1 > 0:
                core.jacobi_zero(L.dw, rhs, c[:n])
            else:
                c[:n].zero_()
            for k in range(max(nu1 - 1, 0)):
                o = xb if c is xa else xa
                L.A.apply(3, c, o, b=rhs, dw=L.dw, overlap=self.overlap, comm_stream=cs, chan=ch(l, "pre", k))
                c = o
            if not fused:
                r_pushed = L.A.apply(2, c, L.r, b=rhs, overlap=self.overlap, comm_stream=cs, chan=ch(l, "res"),
                                     after_boundary=early)   # r[:n] = b - A x
            nxt_b = self.levels[l + 1].b if l + 1 < nl else None
            early_b = None
            if nxt_b is None:
                nxt_b = self._tail_local_b()
            elif early_ok and getattr(L, "early_next_res", False):
                f_n, s_n, _ = self._first_op(l + 1, nu1, nu2, chans)
                if f_n:        # the next level's first pass pushes its right-hand side (scaled by dw on the unscaled fused op)
                    Ln = self.levels[l + 1]
                    early_b = (lambda _c=ch(l + 1, "res"), _b=nxt_b, _s=(None if s_n else Ln.dw): _c.push(_b, scale=_s))
            rhs_pushed = bool(L.R.apply(0, L.r, nxt_b, overlap=self.overlap, comm_stream=cs, chan=ch(l, "R"),
                                        prepushed=bool(r_pushed), after_boundary=early_b))   # b_c = R r
            cur.append(c)
            rhs = nxt_b
        # replicated tail: gather the restricted residual, every rank solves, keep my slice
        xc = self._tail_solve(rhs, nu1, nu2, chans["tail"] if chans is not None else None)
        e_pushed = False
        for l in range(nl - 1, -1, -1):
            L = self.levels[l]
            xa, xb = L.x
            n = L.n
            c = cur[l]
            if xc is not None:                       # coarse correction not yet in L.e (tail, or a level without post-smoothing)
                L.e[:L.nc].copy_(xc)
                e_pushed = False
            rhs_l = b if l == 0 else self.levels[l].b
            # the last sweep writes straight into its consumer: the caller's vector (level 0) or the finer level's
            # coarse-correction buffer — no copy on the way up
            target = x_out if l == 0 else self.levels[l - 1].e

            def early_up(last_is_q):
                """push of the finer level's prolongation channel, chained behind this level's last boundary kernel"""
                if not early_ok or l == 0 or nu2 == 0:
                    return None
                Lf = self.levels[l - 1]
                if not getattr(L, "early_up_Q" if last_is_q else "early_up_A", False):
                    return None
                return (lambda _c=ch(l - 1, "P"), _e=Lf.e: _c.push(_e))
            k0 = 0
            this_pushed, e_pushed = e_pushed, False
            if L.Q is not None and nu2 > 0:
                # x + P e followed by one sweep == x + dw.*r + Q e (r is still in L.r): one pass over Q
                o = target if nu2 == 1 else (xb if c is xa else xa)
                ab = early_up(True) if nu2 == 1 else None
                if L.lazy:
                    ran = L.Q.apply(7, L.e, o, b=L.r, dw=L.dw, overlap=self.overlap, comm_stream=cs, chan=ch(l, "P"), aux=rhs_l,
                                    prepushed=this_pushed, after_boundary=ab)
                else:
                    ran = L.Q.apply(5, L.e, o, b=L.r, dw=L.dw, overlap=self.overlap, comm_stream=cs, chan=ch(l, "P"), aux=c,
                                    prepushed=this_pushed, after_boundary=ab)
                if nu2 == 1:
                    e_pushed = bool(ran)
                c = o
                k0 = 1
            else:
                L.P.apply(1, L.e, c, overlap=self.overlap, comm_stream=cs, chan=ch(l, "P"), prepushed=this_pushed)   # x += P e
            for k in range(k0, nu2):
                last = k == nu2 - 1
                o = target if last else (xb if c is xa else xa)
                ran = L.A.apply(3, c, o, b=rhs_l, dw=L.dw, overlap=self.overlap, comm_stream=cs, chan=ch(l, "post", k),
                                after_boundary=early_up(False) if last else None)
                if last:
                    e_pushed = bool(ran)
                c = o
            xc = None if nu2 > 0 else c[:n]
        if xc is not None:
            x_out.copy_(xc)
        return x_out

    def capture(self, b, x_out, nu1=1, nu2=1, checked=False):
        """Capture one cycle (kernels + NCCL exchanges + side-stream fork/joins) into a CUDA graph and return a
        replay callable.  b / x_out are baked in (update them in place).  All ranks must call this together.
        checked=True: the callable synchronises and raises if a halo wait timed out (for callers that read x_out right
        away); the default only enqueues — follow it with check_exchange()."""
        self.vcycle(b, x_out, nu1, nu2)          # warm-up: every lazy buffer exists before capture
        self.vcycle(b, x_out, nu1, nu2)
        torch.cuda.synchronize()
        self.comm.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            self.vcycle(b, x_out, nu1, nu2)
        torch.cuda.synchronize()
        self._graph = g
        if checked:
            def replay_checked():
                g.replay()
                self.check_exchange()
            return replay_checked
        return g.replay

    def _tail_local_b(self):
        lo = int(self.tail_offsets[self.comm.rank])
        hi = int(self.tail_offsets[self.comm.rank + 1])
        return self.tail_b[lo:hi]

    def _tail_solve(self, b_local, nu1, nu2, chan=None):
        comm = self.comm
        lo = int(self.tail_offsets[comm.rank])
        hi = int(self.tail_offsets[comm.rank + 1])
        if chan is not None:
            chan.push(b_local)
            chan.unpack(self.tail_b, self._tail_unpack_idx)
        elif comm.world > 1:
            self._tail_in[:hi - lo].copy_(b_local)
            dist.all_gather_into_tensor(self._tail_out, self._tail_in, group=comm.group)
            check(lib.mlamg_gather(core.dt(self.tail_b), n_tail_all(self), core.ptr(self._tail_map), core.ptr(self._tail_out),
                                   core.ptr(self.tail_b), core.stream()))
        elif b_local.data_ptr() != self.tail_b.data_ptr():
            self.tail_b.copy_(b_local)
        check(lib.mlamg_vcycle(self.tail._h, core.ptr(self.tail_b), core.ptr(self.tail_x), nu1, nu2, 1, core.stream()))
        return self.tail_x[lo:hi]

    # -- solvers --------------------------------------------------------------------------------------
    def matvec(self, x_local, out):
        L = self.levels[0]
        L.x[0][:L.n].copy_(x_local)
        L.A.apply(0, L.x[0], out, overlap=self.overlap, comm_stream=self.comm_stream if self.overlap else None)
        return out

    def gdot(self, x, y):
        return self.comm.allreduce_sum(core.dot(x, y))

    def pcg(self, b, x0=None, tol=1e-8, maxiter=200, nu1=1, nu2=1, lookahead=2):
        """V-cycle preconditioned CG on the local slices; returns (x, residual history, iterations).

        Device resident: alpha, beta, the dot products, the iteration counter, the convergence flag and the residual
        history live in a small state tensor on every rank (csrc/solvers.cu, mlamg_dloop_*).  A local dot product lands
        in a field of the state, `all_reduce` sums that field in place on the device, one-thread kernels advance the
        state identically on all ranks — no `.item()` anywhere in the iteration.  The host only polls a pinned copy of
        the flag `lookahead` iterations behind what it has enqueued (every piece of an iteration that starts after
        convergence returns at once, so x, r and the history are those of the converged iteration)."""
        comm = self.comm
        n = b.numel()
        dtc = core.dt(b)
        s = core.stream
        x = torch.zeros_like(b) if x0 is None else x0.clone()
        ap, p = torch.empty_like(b), torch.zeros_like(b)
        # r and z are kept between calls so that the preconditioner apply z = V(r) is ONE captured CUDA graph that is
        # replayed every iteration (peer transport / single rank; the NCCL transport launches the cycle eagerly)
        key = (n, b.dtype, nu1, nu2, bool(self.overlap), self.halo)
        ws = getattr(self, "_pcg_ws", None)
        if ws is None or ws["key"] != key:
            ws = self._pcg_ws = {"key": key, "r": torch.empty_like(b), "z": torch.empty_like(b), "replay": None}
            if self.halo == "peer" or comm.world == 1:
                ws["replay"] = self.capture(ws["r"], ws["z"], nu1, nu2)
        r, z = ws["r"], ws["z"]
        precond = ws["replay"] if ws["replay"] is not None else (lambda: self.vcycle(r, z, nu1, nu2))
        state = torch.zeros(int(lib.mlamg_dloop_state_bytes()) // 8, dtype=torch.float64, device=b.device)
        flags = state.view(torch.int32)[20:24]                        # it, done, maxiter, first
        res_d = torch.zeros(maxiter + 1, dtype=torch.float64, device=b.device)
        sp, rp = core.ptr(state), core.ptr(res_d)

        def allred(field):
            if comm.world > 1:
                dist.all_reduce(state[field:field + 1], group=comm.group)

        def dot_into(u, v, field):
            check(lib.mlamg_dloop_dot(dtc, n, core.ptr(u), core.ptr(v), sp, field, s()))
            allred(field)

        check(lib.mlamg_dloop_init(sp, float(tol), int(maxiter), s()))
        self.matvec(x, ap)
        r.copy_(b)
        core.axpby(-1.0, ap, 1.0, r)
        dot_into(b, b, 8)
        dot_into(r, r, 2)
        check(lib.mlamg_dloop_scalar(sp, 0, rp, s()))
        M = lookahead + 1
        pins = [torch.zeros(4, dtype=torch.int32).pin_memory() for _ in range(M)]
        evs = [torch.cuda.Event() for _ in range(M)]
        pins[0].copy_(flags, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        stop = bool(pins[0][1] != 0)
        k = 0
        while not stop and k < maxiter:
            precond()
            dot_into(r, z, 8)
            check(lib.mlamg_dloop_scalar(sp, 1, rp, s()))
            check(lib.mlamg_dloop_direction(dtc, n, core.ptr(z), core.ptr(p), sp, s()))
            self.matvec(p, ap)
            dot_into(p, ap, 1)
            check(lib.mlamg_dloop_scalar(sp, 2, rp, s()))
            check(lib.mlamg_dloop_update(dtc, n, core.ptr(p), core.ptr(ap), core.ptr(x), core.ptr(r), sp, s()))
            allred(2)
            check(lib.mlamg_dloop_scalar(sp, 3, rp, s()))
            pins[k % M].copy_(flags, non_blocking=True)
            evs[k % M].record()
            if k >= lookahead:
                j = (k - lookahead) % M
                evs[j].synchronize()
                stop = bool(pins[j][1] != 0)
            k += 1
        torch.cuda.current_stream().synchronize()
        it = int(flags[0].item())
        res = res_d[:it + 1].cpu().numpy()
        self.check_exchange()          # a halo wait that timed out returned stale values: never hand such a result out
        return x, res, it

    def cycle_bytes(self, nu1=1, nu2=1):
        """algorithmic bytes one rank moves per cycle on its distributed levels (+ the replicated tail)"""
        v = 8 if self.dtype == torch.float64 else 4
        tot = 0.0
        for L in self.levels:
            N, nnz, Nc, pn = L.n, L.A.csr.nnz, L.nc, L.P.csr.nnz
            b_jac = nnz * (v + 4) + 4 * (N + 1) + 4 * v * N
            b_res = nnz * (v + 4) + 4 * (N + 1) + 3 * v * N
            pre = (nu1 - 1) * b_jac + 3 * v * N if nu1 > 0 else 0
            if nu1 == 1 and (self.fuse_pre or nnz <= 12 * N) and (self.halo == "peer" or self.comm.world == 1):
                pre, b_res = 0, b_jac          # fused x = dw.*b, r = b - A x: read A, b, dw; write x, r
                if self.fuse_pre and L.Q is not None and nu2 > 0:
                    b_res = nnz * (v + 4) + 4 * (N + 1) + 2 * v * N      # lazy x: read A D_w, b; write r
            post = nu2 * b_jac + (pn * (v + 4) + 4 * (N + 1) + v * Nc + 2 * v * N)
            if L.Q is not None and nu2 > 0:   # fused prolongation + first post sweep: read Q, e, x, r, dw; write x
                post = (nu2 - 1) * b_jac + L.Q.csr.nnz * (v + 4) + 4 * (N + 1) + v * Nc + 4 * v * N
            tot += pre + post + b_res + (pn * (v + 4) + 4 * (Nc + 1) + v * N + v * Nc)
        return tot + self.tail.cycle_bytes(nu1, nu2, True)


def n_tail_all(h):
    return int(h.tail_offsets[-1])


def slab_geometry(n, world, geometry="cube"):
    """(nx, ny, nz_local) of one rank's z-slab, n^3 DOF per rank either way.
    'slab': n x n x n per rank, global n x n x (n*world).
    'cube' (BASELINE.json config 5): (2n) x (2n) x (n/4) per rank for world > 1 — at n = 256 slabs of 512^2 x 64, the
    global grid 512 x 512 x (64*world) is the 512^3 cube at 8 GPUs (134 217 728 DOF)."""
    if geometry == "slab" or world == 1:
        return n, n, n
    if geometry != "cube":
        raise ValueError(f"unknown geometry {geometry!r}")
    if n % 4:
        raise ValueError("cube geometry needs n divisible by 4")
    return 2 * n, 2 * n, n // 4


def poisson_slab(n, world, rank, dtype=torch.float64, geometry="slab"):
    """rows of rank's z-slab of the global Dirichlet 7-point grid (see slab_geometry), GLOBAL column ids"""
    nx, ny, nzl = slab_geometry(n, world, geometry)
    nzg = nzl * world
    z0 = nzl * rank
    N = nx * ny * nzl
    rowptr = torch.empty(N + 1, dtype=torch.int32, device="cuda")
    col = torch.empty(7 * N, dtype=torch.int32, device="cuda")
    val = torch.empty(7 * N, dtype=dtype, device="cuda")
    nnz = ctypes.c_longlong(0)
    check(lib.mlamg_poisson_csr_slab(core.dt(val), nx, ny, nzg, z0, nzl, core.ptr(rowptr), core.ptr(col), core.ptr(val),
                                     ctypes.byref(nnz), core.stream()))
    return rowptr, col[:nnz.value].clone(), val[:nnz.value].clone()


def slab_lambda_max(n, world, geometry="slab"):
    """analytic rho(D^-1 A) of the global Dirichlet 7-point box"""
    nx, ny, nzl = slab_geometry(n, world, geometry)
    return 1.0 + (np.cos(np.pi / (nx + 1)) + np.cos(np.pi / (ny + 1)) + np.cos(np.pi / (nzl * world + 1))) / 3.0
