"""GPU: hierarchy SETUP parity at the sizes BASELINE.json names — the GPU builds, the oracle builds from the same A,
and the two are compared entry by entry (round 1 only compared cycles on a hierarchy downloaded from the GPU).

  config 1  2D 5-point Poisson 256^2, Lloyd ratio 0.1 'unit' rand 0, SA P, two-level amg_2_v with the reference's
            protocol (utils/evaluate_dataset.py:91-96: b = 0, x0 = RandomState(0).randn / ||.||, res_tol 1e-10):
            Agg / roots / seeds array_equal, P and P^T A P bit-identical, default-argument P (Lanczos omega) against the
            ARPACK-omega P, Gauss-Seidel and Jacobi histories, iteration counts, convergence factors;
  config 2  3D 7-point 128^3, ratio 0.027: Lloyd labels, moved seeds, P, R, A_H bit-identical on EVERY level of the
            multilevel hierarchy (the oracle is given the omegas the GPU computed; the GPU's lambda_max is separately
            held to scipy's ARPACK value on level 1).
Bars: integer/index work and sparsity patterns bit-exact; P / A_H values bit-identical (stronger than the 1e-12 bar);
histories <= 1e-12 relative to the initial residual, identical iteration counts, convergence factor <= 1e-9.
"""
import time

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

from helpers import assert_csr_bitwise, assert_csr_close, hist_err0, rel_hist_err
from oracle import reference_path as rp, multilevel as oml

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def config1():
    A = oml.poisson((256, 256))
    assert A.shape[0] == 65536 and A.nnz == 326656
    Agg, roots, seeds = rp.lloyd_aggregation(A, ratio=0.1, distance="unit", maxiter=10, rand=0)
    lam = rp.lambda_max_dinv_a(A)                                  # ARPACK, as the reference (multigrid.py:105)
    P = sp.csr_matrix(rp.smoothed_aggregation_jacobi(A, Agg, omega=(4.0 / 3.0) / lam))
    return dict(A=A, Agg=Agg, roots=roots, seeds=seeds, lam=lam, P=P)


def test_config1_lloyd_aggregation_256x256_bit_exact(config1):
    import ns.lib.graph as g
    A = config1["A"]
    Agg, roots, seeds = g.lloyd_aggregation(A, ratio=0.1, distance="unit", maxiter=10, rand=0)
    assert Agg.shape == (65536, 6554) and Agg.dtype == np.int8
    assert np.array_equal(seeds, config1["seeds"])
    assert np.array_equal(roots, config1["roots"]), "moved seeds differ from the oracle at 256^2"
    ref = sp.csr_matrix(config1["Agg"])
    assert np.array_equal(Agg.indptr, ref.indptr) and np.array_equal(Agg.indices, ref.indices)
    assert np.array_equal(Agg.data, ref.data)
    # the 'same' distance on a weighted strength matrix (what evaluate_dataset passes, common.py:56) at this size
    C = sp.csr_matrix((1.0 / np.abs(A.data) + (np.arange(A.nnz) % 7) * 0.125, A.indices, A.indptr), shape=A.shape)
    got = g.lloyd_aggregation(C, ratio=0.1, distance="same", rand=0)
    want = rp.lloyd_aggregation(C, ratio=0.1, distance="same", rand=0)
    assert np.array_equal(got[1], want[1])
    assert np.array_equal(got[0].indices, sp.csr_matrix(want[0]).indices)
    assert np.array_equal(got[0].indptr, sp.csr_matrix(want[0]).indptr)


def test_config1_prolongator_and_galerkin_256x256(config1):
    import mlamg
    import ns.lib.multigrid as mg
    A, Agg, lam, P_ref = config1["A"], config1["Agg"], config1["lam"], config1["P"]
    exact = 1.0 + np.cos(np.pi / 257)
    assert abs(lam - exact) < 1e-10
    P = mg.smoothed_aggregation_jacobi(A, Agg, omega=(4.0 / 3.0) / lam)
    assert_csr_bitwise(P, P_ref)
    # the call the reference's callers make (no omega): Lanczos on the device instead of ARPACK.  lambda agrees to
    # 1e-10, so P agrees to 1e-8 (it is linear in omega); bit-identity needs omega passed in because ARPACK's own
    # last bits vary from run to run (random start vector).
    info = {}
    lam_dev = mlamg.lambda_max(mlamg.DeviceCSR.from_scipy(A), info=info)
    assert abs(lam_dev - lam) <= 1e-10 * lam, (lam_dev, lam, info)
    P_default = mg.smoothed_aggregation_jacobi(A, Agg)
    assert_csr_close(P_default, P_ref, 1e-8)
    AH = mlamg.galerkin(mlamg.DeviceCSR.from_scipy(A), mlamg.DeviceCSR.from_scipy(P_ref)).to_scipy()
    assert_csr_bitwise(AH, rp.canonical_csr(rp.galerkin(A, P_ref)))
    assert AH.shape == (6554, 6554)


@pytest.mark.parametrize("smoother", ["gauss_seidel", "jacobi"])
def test_config1_amg_2_v_protocol_256x256(config1, smoother):
    """utils/evaluate_dataset.py:91-96 / utils/common.py:84-96 through ns.lib.multigrid.amg_2_v."""
    import ns.lib.multigrid as mg
    A, P = config1["A"], config1["P"]
    n = A.shape[1]
    b = np.zeros(n)
    x = np.random.RandomState(0).randn(n)
    x /= np.linalg.norm(x, 2)
    kw = dict(res_tol=1e-10, jacobi_weight=2.0 / 3.0)
    t0 = time.perf_counter()
    ref = rp.amg_2_v(A, P, b, x.copy(), smoother=smoother, **kw)
    t_cpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    got = mg.amg_2_v(A, P, b, x.copy(), smoother=smoother, **kw)
    t_gpu = time.perf_counter() - t0
    assert got[3] == ref[3], (got[3], ref[3])
    assert hist_err0(got[2], ref[2]) < 1e-12
    assert rel_hist_err(got[2][:10], ref[2][:10]) < 1e-11
    assert abs(got[1] - ref[1]) < 1e-9
    assert np.abs(got[0] - ref[0]).max() < 1e-12
    print(f"config 1 amg_2_v[{smoother}]: {got[3]} iterations, conv {got[1]:.4f}; CPU oracle {t_cpu:.2f} s, GPU {t_gpu:.2f} s")


def test_config2_setup_parity_128cubed_every_level():
    import mlamg
    n = 128
    A = oml.poisson((n, n, n))
    exact = 1.0 + np.cos(np.pi / (n + 1))
    infos = []

    def lam_dev(M):
        if M.shape[0] == n ** 3:
            return exact
        info = {}
        lam = mlamg.lambda_max(M, info=info)
        infos.append((M.shape[0], lam, info))
        return lam

    t0 = time.perf_counter()
    H = mlamg.build_hierarchy(A, aggregates="lloyd", ratio=0.027, distance="unit", maxiter=10, rand=0, lam_max=lam_dev,
                              max_coarse=1000, max_levels=8)
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    lams = [(4.0 / 3.0) / L.omega_sa for L in H.levels[:-1]]
    t0 = time.perf_counter()
    ref = oml.build_hierarchy(A, ratio=0.027, distance="unit", maxiter=10, rand=0, lam_max=lams, max_coarse=1000,
                              max_levels=8)
    t_cpu = time.perf_counter() - t0
    assert len(H.levels) == len(ref) >= 3
    for lvl, (Lg, Lr) in enumerate(zip(H.levels, ref)):
        assert_csr_bitwise(Lg.A.to_scipy(), Lr.A)
        if Lr.P is None:
            continue
        assert np.array_equal(Lg.labels.cpu().numpy(), Lr.labels), f"level {lvl}: Lloyd labels differ"
        assert np.array_equal(Lg.roots.cpu().numpy(), Lr.roots), f"level {lvl}: moved seeds differ"
        assert np.array_equal(np.asarray(Lg.seeds), np.asarray(Lr.seeds))
        assert_csr_bitwise(Lg.P.to_scipy(), Lr.P)
        assert_csr_bitwise(Lg.R.to_scipy(), Lr.R)
    # the device eigenvalue of level 1 against ARPACK on the oracle's (identical) level-1 operator
    A1 = ref[1].A
    d = 1.0 / np.sqrt(A1.diagonal())
    B = sp.diags(d) @ A1 @ sp.diags(d)
    lam_arpack = float(np.abs(spla.eigsh(B, k=1, which="LA", return_eigenvectors=False, tol=1e-13)).max())
    assert abs(infos[0][1] - lam_arpack) <= 1e-10 * lam_arpack, (infos[0], lam_arpack)
    print(f"config 2 setup: levels {[l.A.shape[0] for l in ref]}; GPU {t_gpu:.2f} s, CPU oracle {t_cpu:.2f} s; "
          f"lambda_max {[(m, i['steps'], i['residual']) for m, _, i in infos]}")
