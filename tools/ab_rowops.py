"""A/B helper: a few row-ops of the 256^3 hierarchy timed with the library selected by MLAMG_LIB_PATH (development aid)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ml-amg_b200")]
import numpy as np
import torch


def main():
    import mlamg
    from mlamg import core
    n = 256
    A = mlamg.poisson((n, n, n), torch.float64)
    exact = 1.0 + np.cos(np.pi / (n + 1))
    H = mlamg.build_hierarchy(A, aggregates="lloyd", ratio=0.027, distance="unit", maxiter=10, rand=0,
                              lam_max=lambda M: exact if M.shape[0] == n ** 3 else mlamg.lambda_max(M), max_coarse=1000, max_levels=8)

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
        return round(e0.elapsed_time(e1) / 5 / reps * 1e3, 1)
    out = {"lib": os.environ.get("MLAMG_LIB_PATH", "default")}
    for l in (0, 1):
        Al, P, R, dw = H._apply[l]
        N, Nc = Al.shape[0], P.shape[1]
        b = torch.randn(N, dtype=torch.float64, device="cuda"); r = torch.randn(N, dtype=torch.float64, device="cuda")
        y = torch.empty_like(b); e = torch.randn(Nc, dtype=torch.float64, device="cuda"); bc = torch.empty_like(e)
        Q = H._Q[l]
        out[f"L{l}_restrict"] = gtime(lambda: core.spmv(R, r, bc))
        out[f"L{l}_psmooth0_csr"] = gtime(lambda: core.prolong_smooth_zero(Q, e, b, r, dw, y))
        if l in H._w32:
            a32, q32 = H._w32[l]
            out[f"L{l}_psmooth0_w32"] = gtime(lambda: core.prolong_smooth_zero_w32(Q, q32, e, b, r, dw, y))
            if a32 is not None:
                As = Al.with_values(H._scaled[l])
                out[f"L{l}_residual_w32"] = gtime(lambda: core.residual_w32(As, a32, b, b, r))
        out[f"L{l}_residual_csr"] = gtime(lambda: core.residual(Al.with_values(H._scaled[l]), b, b, r))
    x = torch.empty(n ** 3, dtype=torch.float64, device="cuda"); bb = torch.randn(n ** 3, dtype=torch.float64, device="cuda")
    H.use_graph(True)
    cyc = lambda: core.check(core.lib.mlamg_vcycle(H._h, core.ptr(bb), core.ptr(x), 1, 1, 1, core.stream()))
    for _ in range(20):
        cyc()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        cyc()
    e1.record(); torch.cuda.synchronize()
    out["cycle_us"] = round(e0.elapsed_time(e1) / 50 * 1e3, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
