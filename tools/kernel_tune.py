"""Per-op timing of the V-cycle kernels under every threads-per-row variant (development aid, one GPU).

    python tools/kernel_tune.py [n]

For levels 0 and 1 of the n^3 Poisson hierarchy: fused Jacobi sweep, residual, fused zero-sweep+residual (and its
unfused pair), prolong-add, restriction — each timed as 10 back-to-back launches replayed from a CUDA graph,
with the algorithmic bytes and the resulting GB/s.  One JSON line per (level, op)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ml-amg_b200")]
import numpy as np
import torch


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    import mlamg
    from mlamg import core
    A = mlamg.poisson((n, n, n), torch.float64)
    exact = 1.0 + np.cos(np.pi / (n + 1))
    H = mlamg.build_hierarchy(A, aggregates="lloyd", ratio=0.027, distance="unit", maxiter=10, rand=0,
                              lam_max=lambda M: exact if M.shape[0] == n ** 3 else mlamg.lambda_max(M),
                              max_coarse=1000, max_levels=8)

    def gtime(fn, reps=10):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 5 / reps * 1e3       # us

    v = 8
    for l in ((0, 1, 2) if os.environ.get("TUNE_W32") == "1" else (0, 1)):
        if l >= len(H._apply) - 1:
            continue
        Al, P, R, dw = H._apply[l]
        N, nnz, Nc, pn = Al.shape[0], Al.nnz, P.shape[1], P.nnz
        rs = torch.Generator(device="cuda").manual_seed(l)
        x = torch.randn(N, dtype=torch.float64, device="cuda", generator=rs)
        b = torch.randn(N, dtype=torch.float64, device="cuda", generator=rs)
        y = torch.empty_like(x); r = torch.empty_like(x)
        e = torch.randn(Nc, dtype=torch.float64, device="cuda", generator=rs)
        bc = torch.empty_like(e)
        first = R.col[R.rowptr[:-1].long().clamp(max=max(R.nnz - 1, 0))]
        order = torch.argsort(first, stable=True).to(torch.int32).contiguous()
        vs = core.scaled_values(Al, dw)
        Q = H._Q[l]
        ops = {
            "reszero_scaled": (lambda: core.jacobi_zero_residual_scaled(Al, vs, dw, b, y, r), nnz * (v + 4) + 4 * (N + 1) + 4 * v * N),
            "psmooth": (lambda: core.prolong_smooth(Q, e, x, r, dw, y), Q.nnz * (v + 4) + 4 * (N + 1) + v * Nc + 4 * v * N),
            "psmooth0": (lambda: core.prolong_smooth_zero(Q, e, b, r, dw, y), Q.nnz * (v + 4) + 4 * (N + 1) + v * Nc + 4 * v * N),
            "residual_scaled": (lambda: core.residual(Al.with_values(vs), b, b, r), nnz * (v + 4) + 4 * (N + 1) + 2 * v * N),
            "jacobi": (lambda: core.jacobi_sweep(Al, dw, b, x, y), nnz * (v + 4) + 4 * (N + 1) + 4 * v * N),
            "residual": (lambda: core.residual(Al, x, b, r), nnz * (v + 4) + 4 * (N + 1) + 3 * v * N),
            "reszero_fused": (lambda: core.jacobi_zero_residual(Al, dw, b, y, r), nnz * (v + 4) + 4 * (N + 1) + 4 * v * N),
            "zero_then_residual": (lambda: (core.jacobi_zero(dw, b, y), core.residual(Al, y, b, r)),
                                   nnz * (v + 4) + 4 * (N + 1) + 6 * v * N),
            "prolong_add": (lambda: core.spmv_add(P, e, y), pn * (v + 4) + 4 * (N + 1) + v * Nc + 2 * v * N),
            "restrict": (lambda: core.spmv(R, r, bc), pn * (v + 4) + 4 * (Nc + 1) + v * N + v * Nc),
            "restrict_ordered": (lambda: core.spmv_perm(R, r, order, bc), pn * (v + 4) + 4 * (Nc + 1) + v * N + v * Nc),
        }
        if os.environ.get("TUNE_W32") == "1":
            w32 = core.csr_to_w32(Q)
            As = Al.with_values(vs)
            aw32 = core.csr_to_w32(As)
            ops = {"psmooth0_csr": ops["psmooth0"],
                   "psmooth0_w32": (lambda: core.prolong_smooth_zero_w32(Q, w32, e, b, r, dw, y), ops["psmooth0"][1]),
                   "residual_scaled_csr": ops["residual_scaled"],
                   "residual_scaled_w32": (lambda: core.residual_w32(As, aw32, b, b, r), ops["residual_scaled"][1])}
        for name, (fn, nbytes) in ops.items():
            out = {"level": l, "op": name, "rows": N if "restrict" not in name else Nc,
                   "mean_row": round((pn / Nc) if "restrict" in name else (pn / N if name == "prolong_add" else
                                                                           (Q.nnz / N if name.startswith("psmooth") else nnz / N)), 2),
                   "MB": round(nbytes / 1e6, 1), "us": {}, "GBs": {}}
            mean = out["mean_row"]
            cand = [l for l in (1, 2, 4, 8, 16, 32) if l <= max(1, 2 * mean) and l * 16 >= mean]
            if os.environ.get("TUNE_ALL") == "1":
                cand = [0, 1, 2, 4, 8, 16, 32]
            if os.environ.get("TUNE_W32") == "1":
                cand = []
            if os.environ.get("TUNE_TMA") == "1":          # plain thread-per-row vs the TMA-staged kernel, short-row ops only
                if mean > 12 or name in ("restrict", "restrict_ordered"):
                    continue
                cand = [0, 1]
            for lanes in [-2] + cand:
                for nb in ((0,) if lanes == -2 else (2, 4, 8)):
                    if lanes == 0 and (name == "restrict_ordered" or nb != 4):
                        continue
                    mlamg.set_csr_lanes(lanes)
                    core.set_csr_batch(nb)
                    try:
                        t = gtime(fn)
                    finally:
                        mlamg.set_csr_lanes(-1)
                        core.set_csr_batch(0)
                    key = "heur" if lanes == -2 else ("staged" if lanes == 0 else f"{lanes}x{nb}")
                    out["us"][key] = round(t, 1)
                    out["GBs"][key] = round(nbytes / t / 1e3)
            print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
