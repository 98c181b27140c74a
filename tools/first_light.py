"""Quick device timings of the apply kernels and the setup phases (development aid, not the bench)."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ml-amg_b200")]
import numpy as np
import torch
import mlamg


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ratio = float(sys.argv[2]) if len(sys.argv) > 2 else 0.027
    out = {}
    for dtype in (torch.float64, torch.float32):
        v = 8 if dtype == torch.float64 else 4
        A = mlamg.poisson((n, n, n), dtype)
        N, nnz = A.shape[0], A.nnz
        x = torch.randn(N, dtype=dtype, device="cuda")
        b = torch.randn(N, dtype=dtype, device="cuda")
        y = torch.empty_like(x)
        dw = mlamg.smoother_diag(A, "jacobi", 2 / 3)
        B_spmv = nnz * (v + 4) + 4 * (N + 1) + 2 * v * N
        B_jac = nnz * (v + 4) + 4 * (N + 1) + 4 * v * N
        timeit(lambda: mlamg.spmv(A, x, y), reps=600)    # ~1 s busy: let the clocks settle
        t = timeit(lambda: mlamg.spmv(A, x, y))
        out[f"spmv_{v*8}"] = dict(ms=t, gbs=B_spmv / t / 1e6)
        t = timeit(lambda: mlamg.jacobi_sweep(A, dw, b, x, y))
        out[f"jacobi_{v*8}"] = dict(ms=t, gbs=B_jac / t / 1e6)
        t = timeit(lambda: mlamg.residual(A, x, b, y))
        out[f"residual_{v*8}"] = dict(ms=t, gbs=(B_spmv + v * N) / t / 1e6)
        S = mlamg.DeviceSELL(A)
        t = timeit(lambda: S.spmv(x, y))
        out[f"sell_spmv_{v*8}"] = dict(ms=t, gbs=B_spmv / t / 1e6, padding=S.padding)
        t = timeit(lambda: S.jacobi_sweep(dw, b, x, y))
        out[f"sell_jacobi_{v*8}"] = dict(ms=t, gbs=B_jac / t / 1e6)
        for lanes in (1, 2, 4, 8):
            mlamg.set_csr_lanes(lanes)
            t = timeit(lambda: mlamg.jacobi_sweep(A, dw, b, x, y))
            out[f"csr_jacobi_L{lanes}_{v*8}"] = dict(ms=t, gbs=B_jac / t / 1e6)
        mlamg.set_csr_lanes(-1)
        del S
        t = timeit(lambda: y.copy_(x))
        out[f"copy_{v*8}"] = dict(ms=t, gbs=2 * v * N / t / 1e6)
        print(json.dumps({k: out[k] for k in out if k.endswith(str(v * 8))}), flush=True)
        del A, x, b, y, dw
    # setup phases, fp64
    torch.cuda.synchronize()
    t0 = time.time()
    A = mlamg.poisson((n, n, n), torch.float64)
    torch.cuda.synchronize(); t1 = time.time()
    labels, nc, roots, seeds = mlamg.lloyd_labels(A, ratio=ratio, distance="unit", maxiter=10, rand=0)
    torch.cuda.synchronize(); t2 = time.time()
    Agg = mlamg.agg_from_labels(labels, nc, torch.float64)
    P = mlamg.sa_prolongator(A, Agg, (4 / 3) / 2.0)
    torch.cuda.synchronize(); t3 = time.time()
    R = mlamg.transpose(P)
    torch.cuda.synchronize(); t4 = time.time()
    AH = mlamg.galerkin(A, P, R)
    torch.cuda.synchronize(); t5 = time.time()
    print(json.dumps(dict(gen=t1 - t0, lloyd=t2 - t1, P=t3 - t2, transpose=t4 - t3, rap=t5 - t4, nc=nc, nnzP=P.nnz,
                          nnzAH=AH.nnz, launches=mlamg.launch_count())), flush=True)
    del Agg, P, R, AH
    t0 = time.time()
    H = mlamg.build_hierarchy(A, ratio=ratio, distance="unit", rand=0, lam_max=2.0, max_coarse=1000, max_levels=8)
    torch.cuda.synchronize()
    print("hierarchy setup s", time.time() - t0)
    print(H)
    b = torch.randn(A.shape[0], dtype=torch.float64, device="cuda")
    xo = torch.empty_like(b)
    from mlamg import core
    def cyc():
        core.check(core.lib.mlamg_vcycle(H._h, core.ptr(b), core.ptr(xo), 1, 1, 1, core.stream()))
    t = timeit(cyc, reps=10)
    print(json.dumps(dict(vcycle_ms=t, cycle_bytes=H.cycle_bytes(), gbs=H.cycle_bytes() / t / 1e6, gdofs=A.shape[0] / t / 1e6)))
    H.use_graph(True)
    t = timeit(cyc, reps=10)
    print(json.dumps(dict(vcycle_graph_ms=t, gbs=H.cycle_bytes() / t / 1e6)))
    H.use_graph(False)
    # per-kernel times on the fine level and the transfer operators
    L0 = H.levels[0]
    N = A.shape[0]
    r = torch.empty_like(b); bc = torch.empty(L0.R.shape[0], dtype=b.dtype, device="cuda")
    for name, fn in [("L0 jacobi(sell)" if L0.sell else "L0 jacobi(csr)", (lambda: L0.sell.jacobi_sweep(L0.dw, b, xo, r)) if L0.sell else (lambda: mlamg.jacobi_sweep(L0.A, L0.dw, b, xo, r))),
                     ("L0 restrict R", lambda: mlamg.spmv(L0.R, r, bc)),
                     ("L0 prolong P", lambda: mlamg.spmv_add(L0.P, bc, xo)),
                     ("L0 jacobi_zero", lambda: mlamg.jacobi_zero(L0.dw, b, xo))]:
        print(name, "ms", round(timeit(fn), 4))
    for lanes in (2, 4, 8, 16, 32):
        mlamg.set_csr_lanes(lanes)
        print("R lanes", lanes, "ms", round(timeit(lambda: mlamg.spmv(L0.R, r, bc)), 4))
    for lanes in (1, 2, 4):
        mlamg.set_csr_lanes(lanes)
        print("P lanes", lanes, "ms", round(timeit(lambda: mlamg.spmv_add(L0.P, bc, xo)), 4))
    for lanes in (1, 2, 4):
        mlamg.set_csr_lanes(lanes)
        print("A csr jacobi lanes", lanes, "ms", round(timeit(lambda: mlamg.jacobi_sweep(L0.A, L0.dw, b, xo, r)), 4))
    mlamg.set_csr_lanes(-1)
    L1 = H.levels[1]
    b1 = torch.randn(L1.A.shape[0], dtype=b.dtype, device="cuda"); x1 = torch.randn_like(b1); y1 = torch.empty_like(b1)
    print("L1 jacobi sell" if L1.sell else "L1 (no sell)", "ms", round(timeit(lambda: L1.sell.jacobi_sweep(L1.dw, b1, x1, y1)), 4) if L1.sell else "", "padding", L1.sell.padding if L1.sell else None)
    for lanes in (1, 2, 4, 8):
        mlamg.set_csr_lanes(lanes)
        print("L1 csr jacobi lanes", lanes, "ms", round(timeit(lambda: mlamg.jacobi_sweep(L1.A, L1.dw, b1, x1, y1)), 4))
    mlamg.set_csr_lanes(-1)
    print("P nnz", L0.P.nnz, "mean row", L0.P.nnz / N, "R mean row", L0.R.nnz / L0.R.shape[0])
    x, res = H.solve(b, tol=1e-8, maxiter=100, accel="cg", return_residuals=True)
    print("pcg iters", len(res) - 1, "final rel res", res[-1] / res[0])


if __name__ == "__main__":
    main()
