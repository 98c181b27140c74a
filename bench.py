#!/usr/bin/env python
"""bench.py — headline benchmark of the aggregation-AMG hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 256]

Workload (config.workload): 3D 7-point Dirichlet Poisson, 256^3 DOF per GPU, fp64; full hierarchy
(Lloyd aggregation ratio 0.027 'unit' distances, smoothed-aggregation P, Galerkin RAP), one *step* =
one V(1,1) weighted-Jacobi cycle applied as a preconditioner (zero initial guess) to a resident
right-hand side.  metric = GDOF/s (fine DOFs x cycles / s / 1e9; V-cycles/s reported beside it).

  value     : cycles timed with CUDA events, inputs resident in HBM, CUDA-graph replay of the cycle
  e2e       : the same cycle through the C-ABI host-buffer entry point (mlamg_vcycle_host): pinned host
              b -> H2D -> V-cycle -> D2H -> host x, copies inside the timed region
  roofline  : the dominant kernel (fused Jacobi sweep on the fine level), CUDA-event duration per launch
              from an instrumented repeat of the timed steps, against MEASURED_PEAKS.json hbm_gbs
  setup_parity : the oracle builds the hierarchy from the same A on the host (its own Lloyd, P, P^T A P, ~1 min at 256^3)
              and every level is compared with the GPU's: labels / moved seeds array_equal, P and A_l bit-identical
  cpu_baseline : the oracle's scipy V-cycle on the ORACLE-built hierarchy, 1 host core (scipy sparsetools is
              single-threaded), a bounded number of cycles; its result is also the full-size parity check of the GPU cycle
  --impl reference : the oracle port end to end on the host: its own CPU setup of the SAME workload (256^3), then
              `steps` V(1,1) cycles timed one by one.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "ml-amg_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

RATIO = 0.027
METRIC = "vcycle_gdof_per_s"
UNIT = "GDOF/s"


def workload_name(n, ngpu, geometry="cube"):
    if ngpu == 1:
        return f"poisson3d_7pt_{n}^3_per_gpu_fp64_lloyd{RATIO}_SA_V(1,1)_jacobi"
    nx, ny, nzl = (2 * n, 2 * n, n // 4) if geometry == "cube" else (n, n, n)
    return (f"poisson3d_7pt_{n}^3_per_gpu_fp64_lloyd{RATIO}_SA_V(1,1)_jacobi_x{ngpu}gpu_zslab_global_{nx}x{ny}x{nzl * ngpu}")


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def lam_fn_factory(n, lams_out=None):
    """lambda_max rule of BOTH arms: the fine level uses the analytic rho(D^-1 A) = 1 + cos(pi/(n+1)) of the Dirichlet
    7-point stencil (ARPACK on 16.7 M rows is 160 s+ and is not part of the timed setup); every coarser
    level is solved for — Lanczos to 1e-13 on the device here, ARPACK (`eigsh` on D^-1/2 A D^-1/2) in the reference arm."""
    import mlamg
    exact = 1.0 + np.cos(np.pi / (n + 1))

    def lam(A):
        v = exact if A.shape[0] == n ** 3 else mlamg.lambda_max(A)
        if lams_out is not None:
            lams_out.append(v)
        return v
    return lam


def lam_fn_reference(n):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    exact = 1.0 + np.cos(np.pi / (n + 1))

    def lam(A):
        if A.shape[0] == n ** 3:
            return exact
        d = 1.0 / np.sqrt(A.diagonal())
        B = sp.diags(d) @ A @ sp.diags(d)           # same spectrum as D^-1 A (multigrid.py:105), symmetric -> eigsh
        return float(np.abs(spla.eigsh(B, k=1, which="LA", return_eigenvectors=False)).max())
    return lam


# ------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    import mlamg
    from mlamg import core, distributed as md_
    cpus = md_.bind_cpu_affinity(local)            # host buffers of this rank on the NUMA node of its GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.n
    if world > 1:
        return run_ours_distributed(args, rank, world, local, cpus)
    t_setup0 = time.time()
    A = mlamg.poisson((n, n, n), torch.float64)
    lams = []
    H = mlamg.build_hierarchy(A, aggregates="lloyd", ratio=RATIO, distance="unit", maxiter=10, rand=0,
                              lam_max=lam_fn_factory(n, lams), max_coarse=1000, max_levels=8)
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup0
    N = A.shape[0]
    b = torch.from_numpy(np.random.RandomState(0).randn(N)).cuda()
    x = torch.empty_like(b)
    s = core.stream()

    def cycle():
        core.check(core.lib.mlamg_vcycle(H._h, core.ptr(b), core.ptr(x), 1, 1, 1, s))

    if args.profile:       # ncu --profile-from-start off: only these plain-launch cycles are captured
        for _ in range(max(args.warmup, 1)):
            cycle()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for _ in range(args.steps):
            cycle()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profiled_cycles": args.steps, "workload": workload_name(n, 1)}))
        return

    # kernels per cycle (plain launches are counted by the library)
    c0 = mlamg.launch_count(); cycle(); torch.cuda.synchronize(); kernels_per_cycle = mlamg.launch_count() - c0
    H.use_graph(True)
    clocks = ClockSampler(local); clocks.start()      # samples span warm-up + timed region + instrumented repeat
    t_busy = time.time()
    while time.time() - t_busy < 0.7:                 # let the SM/memory clocks settle under load
        for _ in range(20):
            cycle()
        torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        cycle()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        cycle()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    H.use_graph(False)
    value = N / ms / 1e6

    # --- the fine-level kernels of the cycle, each timed per launch with CUDA events right after a full cycle
    # (instrumented repeat of the timed steps; same arguments and buffers sizes as inside the cycle).  The dominant
    # one (largest share of the step) is the zero-guess sweep fused with the residual: one pass over A.
    A0, P0, R0, dw0 = H._apply[0]
    Nc = P0.shape[1]
    v = 8
    r_ = torch.empty_like(b); t_ = torch.empty_like(b)
    ec = torch.randn(Nc, dtype=torch.float64, device="cuda"); bc = torch.empty_like(ec)
    lazy = bool(H._scaled) and bool(H._Q)          # the cycle's fine level: r = b - (A D_w) b, then x = dw.*(b + r) + Q e
    if lazy:
        As0 = A0.with_values(H._scaled[0])
        first = ("csr_rowop_kernel<double,LANES=1,OP_RESIDUAL> on the column-scaled copy (fine level: r = b - (A D_w) b, "
                 "one pass over A, x = dw.*b never materialised)",
                 lambda: core.residual(As0, b, b, r_), A0.nnz * (v + 4) + 4 * (N + 1) + 2 * v * N)
    elif H._scaled:
        vs0 = H._scaled[0]
        first = ("csr_rowop_kernel<double,LANES=1,OP_RESZERO_S> (fine level: x=dw.*b, r=b-(A D_w)b fused, one pass over A)",
                 lambda: core.jacobi_zero_residual_scaled(A0, vs0, dw0, b, t_, r_), A0.nnz * (v + 4) + 4 * (N + 1) + 4 * v * N)
    else:
        first = ("csr_rowop_kernel<double,LANES=1,OP_RESZERO> (fine level: x=dw.*b, r=b-Ax fused, one pass over A)",
                 lambda: core.jacobi_zero_residual(A0, dw0, b, t_, r_), A0.nnz * (v + 4) + 4 * (N + 1) + 4 * v * N)
    ops = [first,
           ("csr_rowop_kernel<double,LANES=8,OP_SPMV> (fine level: restriction b_c = R r)",
            lambda: core.spmv(R0, r_, bc), P0.nnz * (v + 4) + 4 * (Nc + 1) + v * N + v * Nc)]
    if lazy:
        Q0 = H._Q[0]
        q32 = H._w32.get(0, (None, None))[1] if hasattr(H, "_w32") else None
        if q32 is not None:      # the cycle's kernel: thread-per-row on the W32 (warp-interleaved) copy of Q
            ops.append(("csr_w32_rowop_kernel<double,OP_PSMOOTH0> (fine level: x = dw.*(b + r) + Q e, prolongation + post sweep fused, "
                        "W32 copy of Q)",
                        lambda: core.prolong_smooth_zero_w32(Q0, q32, ec, b, r_, dw0, x), Q0.nnz * (v + 4) + 4 * (N + 1) + v * Nc + 4 * v * N))
        else:
            ops.append(("csr_rowop_kernel<double,LANES=1,OP_PSMOOTH0> (fine level: x = dw.*(b + r) + Q e, prolongation + post sweep fused)",
                        lambda: core.prolong_smooth_zero(Q0, ec, b, r_, dw0, x), Q0.nnz * (v + 4) + 4 * (N + 1) + v * Nc + 4 * v * N))
    elif H._Q:
        Q0 = H._Q[0]
        ops.append(("csr_rowop_kernel<double,LANES=1,OP_PSMOOTH> (fine level: x += dw.*r + Q e, prolongation + post sweep fused)",
                    lambda: core.prolong_smooth(Q0, ec, t_, r_, dw0, x), Q0.nnz * (v + 4) + 4 * (N + 1) + v * Nc + 4 * v * N))
    else:
        ops.append(("csr_rowop_kernel<double,LANES=1,OP_JACOBI> (fine level: Jacobi sweep)",
                    lambda: core.jacobi_sweep(A0, dw0, b, t_, x), A0.nnz * (v + 4) + 4 * (N + 1) + 4 * v * N))
    peak, peak_kind = measured_peak()
    kern = []
    for name, fn, nbytes in ops:
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for k in range(args.steps):
            cycle()
            ev[k][0].record()
            fn()
            ev[k][1].record()
        torch.cuda.synchronize()
        t_ms = float(np.mean([a.elapsed_time(c) for a, c in ev]))
        kern.append({"kernel": name, "ms_per_launch": round(t_ms, 4), "algorithmic_bytes_per_launch": int(nbytes),
                     "achieved": round(nbytes / t_ms / 1e6, 1), "frac": round(nbytes / t_ms / 1e6 / peak, 4),
                     "share_of_step": round(t_ms / ms, 3)})
    # the metric also names plain SpMV and the smoother sweep: timed the same way on the fine operator
    extra = []
    for name, fn, nbytes in (
            ("spmv y = A x (csr_rowop_kernel<double,LANES=1,OP_SPMV>)", lambda: core.spmv(A0, b, t_),
             A0.nnz * (v + 4) + 4 * (N + 1) + 2 * v * N),
            ("weighted-Jacobi sweep x' = x + dw.*(b - A x), fused with its residual (OP_JACOBI; L1-Jacobi is the same kernel "
             "with dw = 1/||row||_1)", lambda: core.jacobi_sweep(A0, dw0, b, t_, x), A0.nnz * (v + 4) + 4 * (N + 1) + 4 * v * N)):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for k in range(args.steps):
            cycle()
            ev[k][0].record()
            fn()
            ev[k][1].record()
        torch.cuda.synchronize()
        t_ms = float(np.mean([a.elapsed_time(c) for a, c in ev]))
        extra.append({"kernel": name, "ms_per_launch": round(t_ms, 4), "algorithmic_bytes_per_launch": int(nbytes),
                      "achieved": round(nbytes / t_ms / 1e6, 1), "frac": round(nbytes / t_ms / 1e6 / peak, 4)})
    clk = clocks.stop()
    dom = max(kern, key=lambda d: d["ms_per_launch"])
    # BASELINE.md's accounting of a V(1,1) cycle: per level 2 B_jacobi + B_residual + B_restrict + B_prolong_add (two passes
    # over A plus P and R) — what the refactored cycle avoids moving; reported beside the bytes it actually needs
    b_std = 0.0
    for (Al, Pl, Rl, dwl) in H._apply[:-1]:
        n_, nz_, nc_, pn_ = Al.shape[0], Al.nnz, Pl.shape[1], Pl.nnz
        b_std += 2 * (nz_ * (v + 4) + 4 * (n_ + 1) + 4 * v * n_) + (nz_ * (v + 4) + 4 * (n_ + 1) + 3 * v * n_) \
            + (pn_ * (v + 4) + 4 * (nc_ + 1) + v * n_ + v * nc_) + (pn_ * (v + 4) + 4 * (n_ + 1) + v * nc_ + 2 * v * n_)
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": peak, "peak_kind": peak_kind,
                "unit": "GB/s", "frac": dom["frac"], "traffic": None, "ms_per_launch": dom["ms_per_launch"],
                "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                "cycle_bytes": H.cycle_bytes(1, 1, True),
                "cycle_frac": round(H.cycle_bytes(1, 1, True) / ms / 1e6 / peak, 4),
                "cycle_bytes_baseline_formula": b_std, "cycle_frac_baseline_formula": round(b_std / ms / 1e6 / peak, 4),
                "fine_level_kernels": kern,
                "spmv_and_smoother": extra}
    roofline["traffic"], roofline["traffic_source"] = ncu_traffic(dom["kernel"])

    # --- e2e: host buffers through the C ABI (H2D + cycle + D2H inside the timed region)
    hb = torch.from_numpy(np.random.RandomState(1).randn(N)).pin_memory()
    hx = torch.empty(N, dtype=torch.float64).pin_memory()
    H.use_graph(True)
    for _ in range(2):
        H.apply_host(hb, hx, 1, 1, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        H.apply_host(hb, hx, 1, 1, 1)
    e2e_s = (time.perf_counter() - t0) / args.steps
    H.use_graph(False)
    e2e = {"value": round(N / e2e_s / 1e9, 4), "unit": UNIT, "h2d_bytes_per_step": N * 8, "d2h_bytes_per_step": N * 8,
           "ms_per_step": round(e2e_s * 1e3, 3)}

    # --- setup parity + CPU baseline: the oracle builds the hierarchy from the same A on the host (its own Lloyd, P, RAP),
    # every level is compared with the GPU's, and the oracle's cycle on ITS hierarchy is timed and compared with the GPU's
    hx_cycle = hx.numpy()                       # result of the last e2e apply: one zero-guess V(1,1) on hb
    parity, cpu = None, None
    if args.cpu_cycles > 0:
        try:
            ref_levels, parity = oracle_setup_parity(H, n, lams)
            cpu = cpu_baseline(ref_levels, hb.numpy(), hx_cycle, args.cpu_cycles)
        except Exception as exc:        # noqa: BLE001  (e.g. the host runs out of memory for the 256^3 oracle): report it
            parity = parity or {"ok": None, "error": f"{type(exc).__name__}: {exc}"[:300]}

    out = {"metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": common_config(n, 1),
           "detail": {"nnz": A0.nnz, "levels": [l.A.shape[0] for l in H.levels],
                      "operator_complexity": round(H.operator_complexity(), 4),
                      "l2_policy": f"inputs larger than L2 (fine operator {A0.nnz * 12 / 1e9:.1f} GB vs 126 MB L2)",
                      "setup_s": round(setup_s, 2), "lambda_max_per_level": [round(v, 12) for v in lams]},
           "setup_parity": parity,
           "vcycles_per_s": round(1e3 / ms, 2), "clocks": clk, "e2e": e2e, "gpu_launches": kernels_per_cycle * args.steps,
           "kernels_per_cycle": kernels_per_cycle, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(out))


def common_config(n, ngpu, geometry="cube"):
    """the keys both arms print identically (the reference arm runs the same workload)"""
    return {"workload": workload_name(n, ngpu, geometry), "dof": n ** 3 * ngpu, "dof_per_gpu": n ** 3, "cycle": "V(1,1) zero-guess",
            "aggregation": f"lloyd ratio {RATIO} unit rand 0 maxiter 10", "max_coarse": 1000}


def ncu_traffic(kernel_label):
    """DRAM bytes per launch of the dominant kernel from the round's `ncu --set full` capture (tools/ncu_traffic.py wrote
    profiles/traffic_r02.json).  The record carries the sha256 of csrc/apply.cu at capture time: if the kernel source has
    changed since, the number is stale and null is reported instead."""
    import hashlib
    path = os.path.join(ROOT, "profiles", "traffic_r02.json")
    try:
        rec = json.load(open(path))
        sha = hashlib.sha256(open(os.path.join(ROOT, "ml-amg_b200", "csrc", "apply.cu"), "rb").read()).hexdigest()
        if rec.get("apply_cu_sha256") != sha:
            return None, "profiles/traffic_r02.json is stale (apply.cu changed since the ncu capture)"
        for op, val in rec["dram_bytes_per_launch"].items():
            if op in kernel_label:
                return val, rec.get("source")
    except Exception as exc:       # noqa: BLE001
        return None, f"no capture ({type(exc).__name__})"
    return None, "kernel not in the capture"


def oracle_setup_parity(H, n, lams):
    """oracle.multilevel.build_hierarchy on the host from the same A (given the omegas the GPU used, so that every level
    can be compared bit for bit) vs the GPU-built hierarchy."""
    from oracle import multilevel as oml
    t0 = time.time()
    ref = oml.build_hierarchy(oml.poisson((n, n, n)), ratio=RATIO, distance="unit", maxiter=10, rand=0, lam_max=list(lams),
                              max_coarse=1000, max_levels=8)
    t_cpu = time.time() - t0

    def same_csr(G, S):
        G = G.to_scipy()
        return bool(G.shape == S.shape and np.array_equal(G.indptr, S.indptr) and np.array_equal(G.indices, S.indices)
                    and np.array_equal(G.data, S.data))
    par = {"levels_equal": len(ref) == len(H.levels), "labels_equal": [], "roots_equal": [], "P_bitwise": [],
           "A_bitwise": [], "oracle_setup_s": round(t_cpu, 1)}
    for Lg, Lr in zip(H.levels, ref):
        par["A_bitwise"].append(same_csr(Lg.A, Lr.A))
        if Lr.P is not None and Lg.P is not None:
            par["labels_equal"].append(bool(np.array_equal(Lg.labels.cpu().numpy(), Lr.labels)))
            par["roots_equal"].append(bool(np.array_equal(Lg.roots.cpu().numpy(), Lr.roots)))
            par["P_bitwise"].append(same_csr(Lg.P, Lr.P))
    par["ok"] = bool(par["levels_equal"] and all(par["labels_equal"]) and all(par["roots_equal"]) and all(par["P_bitwise"])
                     and all(par["A_bitwise"]))
    return ref, par


def dist_parity(args, comm):
    """Before timing: the row-partitioned hierarchy on the SAME N ranks at a reduced per-rank size against the partitioned
    CPU oracle (tests/dist_check.py: Lloyd labels, P and Galerkin operators bit for bit, V-cycle, PCG history and
    iteration count).  Every rank checks its own rows; the verdicts are combined with an allreduce."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import dist_check
    t0 = time.time()
    out = dist_check.run_check(args.parity_n, comm, geometry=args.geometry, verbose=True, extras=False)
    flags = torch.tensor([int(out["ok"]), int(out["labels_equal"]), int(out["P_bitwise"]), int(out["A_bitwise"]),
                          int(out["pcg_iters_equal"])], dtype=torch.int64, device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    errs = torch.tensor([out["vcycle_rel_err"], out["pcg_hist_err"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    f = [bool(v) for v in flags.cpu()]
    return {"ok": f[0], "labels_equal": f[1], "P_bitwise": f[2], "A_bitwise": f[3], "pcg_iters_equal": f[4],
            "vcycle_rel_err": float(errs[0]), "pcg_hist_err": float(errs[1]), "pcg_iterations": out["pcg_iterations"],
            "dof_per_gpu": out["dof_per_gpu"], "ranks": comm.world, "geometry": args.geometry,
            "dist_levels": out["dist_levels"], "tail_levels": out["tail_levels"], "halo_transport": out["halo_transport"],
            "oracle": "oracle.multilevel.build_hierarchy_partitioned (scipy + restated pyamg loops) on every rank",
            "seconds": round(time.time() - t0, 1)}


def run_ours_distributed(args, rank, world, local, cpus=None):
    """Weak scaling: n^3 DOF per GPU in z-slabs (geometry 'cube': (2n) x (2n) x (n/4) per rank, i.e. the 512^3 cube of
    BASELINE.json config 5 at n = 256 on 8 GPUs; 'slab': n x n x n per rank), row-partitioned levels with the peer-memory
    halo exchange overlapped with the interior rows, coarse levels replicated below `--replicate-below` rows.
    value = global DOFs x cycles / max-over-ranks device time."""
    import torch
    import torch.distributed as dist
    import mlamg
    from mlamg import core, distributed as md
    n = args.n
    comm = md.Comm()
    parity = dist_parity(args, comm) if args.parity_n > 0 else None      # an exception here ends the run at once (a rank
    # that swallowed it would leave its peers waiting in the checker's collectives)
    if parity is not None and not parity["ok"]:
        if rank == 0:
            print(json.dumps({"error": "multi-GPU parity check failed; nothing was timed", "parity": parity}), flush=True)
        dist.barrier()
        dist.destroy_process_group()
        sys.exit(3)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.time()
    rowptr, col, val = md.poisson_slab(n, world, rank, geometry=args.geometry)
    lam0 = md.slab_lambda_max(n, world, args.geometry)      # analytic, the global box
    H = md.DistHierarchy(rowptr, col, val, comm, ratio=RATIO, distance="unit", maxiter=10, rand=0,
                         lam_max=[lam0], max_levels=8, max_coarse=1000, replicate_below=args.replicate_below)
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    N_loc = n ** 3
    b = torch.from_numpy(np.random.RandomState(rank).randn(N_loc)).cuda()
    x = torch.empty_like(b)

    def cycle_eager():
        H.vcycle(b, x, 1, 1)

    c0 = mlamg.launch_count(); cycle_eager(); torch.cuda.synchronize(); kernels_per_cycle = mlamg.launch_count() - c0
    x_eager = x.clone()
    cycle, launch_mode = cycle_eager, "eager launches"
    # peer transport: the cycle is kernels only (device-side flags over NVLink), so the whole cycle is captured as one
    # CUDA graph.  NCCL transport: capture is opt-in (a graph holding NCCL work made process-group teardown hang once).
    want_graph = os.environ.get("MLAMG_DIST_GRAPH", "1" if H.halo == "peer" else "0") == "1"
    if want_graph:
        try:
            replay = H.capture(b, x, 1, 1)
            x.zero_(); replay(); torch.cuda.synchronize()
            same = bool(torch.equal(x, x_eager))
            cycle, launch_mode = replay, f"CUDA graph of the whole cycle (bitwise equal to eager: {same})"
        except Exception as exc:               # noqa: BLE001
            launch_mode = f"eager launches (graph capture failed: {type(exc).__name__})"

    def timed(k):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            cycle()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / k

    clocks = ClockSampler(local); clocks.start()
    t_busy = time.time()
    for _ in range(60):                           # fixed count: every rank issues the same collectives
        for _ in range(10):
            cycle()
        torch.cuda.synchronize()
    timed(max(args.warmup, 3))
    ms = timed(args.steps)
    # dominant kernel on this rank: the zero-guess sweep fused with the residual over all local rows of the fine
    # level (one pass over A, no exchange), CUDA events
    L0 = H.levels[0]
    interior = L0.A.interior_range if L0.A.interior_range is not None else (0, 0)
    n_int = interior[1] - interior[0]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        cycle()
        ev[k][0].record()
        if hasattr(L0.A, "csr_scaled"):      # the cycle's pass: r = b - (A D_w) b (interior rows: no halo column)
            core.rowop(L0.A.csr_scaled, 2, b, L0.r, b=b, row_range=interior)
        else:
            L0.A.rowop(4, None, L0.r, b=b, dw=L0.dw, aux=L0.x[1], row_range=interior)
        ev[k][1].record()
    torch.cuda.synchronize()
    jac_ms = float(np.mean([a.elapsed_time(c) for a, c in ev]))
    clk = clocks.stop()
    # e2e: host slices in and out every step
    hb = torch.from_numpy(np.random.RandomState(100 + rank).randn(N_loc)).pin_memory()
    hx = torch.empty(N_loc, dtype=torch.float64).pin_memory()

    def e2e_step():
        b.copy_(hb, non_blocking=True)
        cycle()
        hx.copy_(x, non_blocking=True)
        torch.cuda.synchronize()
    e2e_step()
    dist.barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    dist.barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t1) / args.steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    if rank == 0:
        v = 8
        nnz = int(L0.A.csr.rowptr[interior[1]].item() - L0.A.csr.rowptr[interior[0]].item())
        B_jac = nnz * (v + 4) + 4 * (n_int + 1) + (2 if hasattr(L0.A, "csr_scaled") else 4) * v * n_int
        peak, peak_kind = measured_peak()
        achieved = B_jac / jac_ms / 1e6
        cyc_bytes = H.cycle_bytes(1, 1)
        out = {"metric": METRIC, "value": round(N_loc * world / ms / 1e6, 4), "unit": UNIT, "n_gpus": world,
               "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": common_config(n, world, args.geometry),
               "parity": parity,
               "detail": {"distributed_levels": [int(o[-1]) for o in H.offsets[:-1]],
                          "replicated_levels": [l.A.shape[0] for l in H.tail.levels], "replicate_below": args.replicate_below,
                          "halo_entries_fine": L0.A.plan.n_halo, "geometry": args.geometry,
                          "slab_per_rank": list(md.slab_geometry(n, world, args.geometry)),
                          "halo_transport": ("tagged peer stores into CUDA-IPC windows over NVLink, boundary rows gather the halo in "
                                             "place, interior rows in between (no collective in the cycle)" if H.halo == "peer"
                                             else "NCCL all-to-all on a side stream overlapped with the interior rows"),
                          "launch_mode": launch_mode,
                          "l2_policy": "inputs larger than L2 (fine operator 1.4 GB per GPU vs 126 MB L2)",
                          "setup_s": round(setup_s, 2)},
               "vcycles_per_s": round(1e3 / ms, 2), "clocks": clk,
               "e2e": {"value": round(N_loc * world / e2e_s / 1e9, 4), "unit": UNIT, "h2d_bytes_per_step": N_loc * 8 * world,
                       "d2h_bytes_per_step": N_loc * 8 * world, "ms_per_step": round(e2e_s * 1e3, 3),
                       "cpu_affinity_rank0": (f"{len(cpus)} cores local to the GPU" if cpus else "not set")},
               "gpu_launches": kernels_per_cycle * args.steps * world, "kernels_per_cycle_per_rank": kernels_per_cycle,
               "roofline": {"bound": "hbm", "kernel": "csr_rowop_kernel<double,LANES=1,OP_RESIDUAL> on the column-scaled copy (fine level: r = b - (A D_w) b, one pass over A; interior rows of rank 0)",
                            "achieved": round(achieved, 1), "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                            "frac": round(achieved / peak, 4), "traffic": None, "ms_per_launch": round(jac_ms, 4),
                            "algorithmic_bytes_per_launch": B_jac, "cycle_bytes_per_gpu": cyc_bytes,
                            "cycle_frac": round(cyc_bytes / ms / 1e6 / peak, 4)},
               "cpu_baseline": None}
        print(json.dumps(out), flush=True)
    H.check_exchange()
    dist.barrier()
    if launch_mode.startswith("CUDA graph") and H.halo != "peer":
        os._exit(0)                            # skip NCCL teardown with live captured graphs holding NCCL work
    H._graph = None
    H.close()
    dist.destroy_process_group()


def cpu_baseline(levels, b, x_gpu, cycles):
    """oracle.multilevel.vcycle (scipy, 1 core) on the ORACLE-built hierarchy; also the full-size parity check of the
    GPU cycle (x_gpu = GPU result of one zero-guess V(1,1) on b, on the GPU-built hierarchy)."""
    from oracle import multilevel as oml
    x = oml.vcycle(levels, b.copy(), None, 1, 1)           # warm-up + parity
    rel = float(np.abs(x - x_gpu).max() / np.abs(x).max())
    ts = []
    for _ in range(max(1, cycles)):
        t0 = time.perf_counter()
        oml.vcycle(levels, b.copy(), None, 1, 1)
        ts.append(time.perf_counter() - t0)
    N = levels[0].A.shape[0]
    # the reference's own expressions on the fine level, best of 2: A@x (multigrid.py:181), the MLAMG.jacobi
    # form x + Dinv*(b - A@x) (MLAMG.py:143-146), P.T@r and P@e (multigrid.py:181)
    L0 = levels[0]
    xv = np.random.RandomState(2).randn(N)
    ec = np.random.RandomState(3).randn(L0.P.shape[1])

    def best(fn, k=2):
        out = []
        for _ in range(k):
            t0 = time.perf_counter(); fn(); out.append(time.perf_counter() - t0)
        return min(out)
    v = 8
    nnz, pn, Nc = L0.A.nnz, L0.P.nnz, L0.P.shape[1]
    ops = {"spmv": (lambda: L0.A @ xv, nnz * (v + 4) + 4 * (N + 1) + 2 * v * N),
           "jacobi_sweep": (lambda: xv + L0.dw * (b - L0.A @ xv), nnz * (v + 4) + 4 * (N + 1) + 4 * v * N),
           "restrict": (lambda: L0.R @ xv, pn * (v + 4) + 4 * (Nc + 1) + v * N + v * Nc),
           "prolong_add": (lambda: xv + L0.P @ ec, pn * (v + 4) + 4 * (N + 1) + v * Nc + 2 * v * N)}
    cpu_ops = {}
    for name, (fn, nbytes) in ops.items():
        t = best(fn)
        cpu_ops[name] = {"ms": round(t * 1e3, 1), "GBs": round(nbytes / t / 1e9, 2)}
    return {"value": round(N / min(ts) / 1e9, 5), "unit": UNIT, "cores": 1, "cores_available": os.cpu_count(),
            "fine_level_ops_scipy": cpu_ops,
            "kind": "port", "sample": f"{len(ts)} full-size V(1,1) cycles of the oracle (scipy, single-threaded) on the "
            f"hierarchy the oracle built itself, best of {len(ts)}; {min(ts):.2f} s per cycle", "parity_rel_err_vs_gpu_cycle": rel}


# ------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's CPU implementation of the path (oracle port: scipy + the restated pyamg C loops, single-threaded
    like scipy/pyamg themselves) on the SAME workload: full CPU setup at n^3, then `steps` V(1,1) cycles.  At N > 1 the
    workload is N independent-size slabs of n^3 DOF each (weak scaling): rank 0 runs the per-GPU share (one n^3 problem)
    as the bounded sample — throughput per DOF is what the driver's ratio needs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import multilevel as oml
    n = args.ref_n if args.ref_n > 0 else args.n
    t0 = time.time()
    A = oml.poisson((n, n, n))
    levels = oml.build_hierarchy(A, ratio=RATIO, distance="unit", maxiter=10, rand=0, lam_max=lam_fn_reference(n),
                                 max_coarse=1000, max_levels=8)
    setup_s = time.time() - t0
    N = A.shape[0]
    b = np.random.RandomState(0).randn(N)
    for _ in range(min(max(args.warmup, 1), 2)):       # scipy has no warm-up effects beyond the first touch of the arrays
        oml.vcycle(levels, b.copy(), None, 1, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oml.vcycle(levels, b.copy(), None, 1, 1)
    s = (time.perf_counter() - t0) / args.steps
    val = round(N / s / 1e9, 5)
    sample = (f"each step = one oracle V(1,1) cycle on the full {n}^3 problem" + (f" (= one GPU's share of the {args.gpus}-GPU "
              f"weak-scaling workload)" if args.gpus > 1 else "") + f"; CPU setup {setup_s:.1f} s untimed, like the GPU arm's; "
              "scipy sparsetools and the pyamg loops are single-threaded")
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": round(s * 1e3, 3), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": common_config(args.n, args.gpus, args.geometry),
           "detail": {"sample_dof": N, "levels": [l.A.shape[0] for l in levels], "setup_s": round(setup_s, 2)},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "cores_available": os.cpu_count(), "kind": "port",
                            "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=256, help="grid side per GPU")
    ap.add_argument("--ref-n", type=int, default=0, help="grid side of the reference arm (0 = the same --n as the GPU arm)")
    ap.add_argument("--cpu-cycles", type=int, default=3, help="oracle cycles timed for cpu_baseline (0 = skip)")
    ap.add_argument("--geometry", default="cube", choices=["cube", "slab"],
                    help="N > 1: 'cube' = z-slabs of (2n)x(2n)x(n/4) (512^3 global at n=256 on 8 GPUs, BASELINE config 5), "
                         "'slab' = n x n x n per rank (global n x n x nN)")
    ap.add_argument("--replicate-below", type=int, default=40000,
                    help="N > 1: levels with at most this many GLOBAL rows are replicated on every rank (measured at 256^3 per GPU: a 24 k-row "
                         "level is cheaper replicated, a 98 k-row level cheaper distributed; profiles/r02_dist_scale_tune_n*.jsonl)")
    ap.add_argument("--parity-n", type=int, default=48, help="N > 1: per-GPU grid side of the pre-timing parity check (0 = skip)")
    ap.add_argument("--profile", action="store_true", help="wrap `steps` plain-launch cycles in cudaProfilerStart/Stop (for ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
