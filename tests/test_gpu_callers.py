"""GPU: the reference's evaluation call sequences (utils/common.py:40-111) on the mirrored modules, against the
same sequence on the oracle."""
import importlib.util
import os

import numpy as np
import numpy.linalg as la
import pytest
import scipy.sparse as sp

from helpers import ROOT
from oracle import reference_path as rp, multilevel as oml

pytestmark = pytest.mark.gpu


def _common():
    spec = importlib.util.spec_from_file_location("mlamg_utils_common", os.path.join(ROOT, "ml-amg_b200", "utils", "common.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _G:
    def __init__(self, A):
        self.A = A


def _oracle_conv(A, C, alpha, lam, seeded_rand):
    np.random.seed(0)
    Agg, _, _ = rp.lloyd_aggregation(C, ratio=alpha, distance="same", rand=seeded_rand)
    P = rp.smoothed_aggregation_jacobi(A, Agg, omega=(4.0 / 3.0) / lam)
    x = np.random.RandomState(0).randn(A.shape[1])
    x /= la.norm(x, 2)
    return rp.amg_2_v(A, P, np.zeros(A.shape[1]), x, res_tol=1e-10, jacobi_weight=2. / 3.)[1]


def test_evaluate_ref_conv_and_dataset_match_the_oracle_sequence():
    from mlamg import problems
    common = _common()
    grids = [_G(sp.csr_matrix(oml.poisson((18, 16)))), _G(problems.voronoi_jump_problem(14, seed=5)[0])]
    lam = 2.0
    for name in ("invabs", "unit", "abs"):
        S = common.strength_measure_funcs[name]
        got = common.evaluate_ref_conv(grids, S, alpha=0.2, lam_max=lambda A: lam)
        ref = [_oracle_conv(g.A, sp.csr_matrix(S(g.A)), 0.2, lam, None) for g in grids]
        assert np.allclose(got, ref, rtol=0, atol=1e-9), (name, got, ref)
        assert all(0.0 < c < 1.0 for c in got)
    got = common.evaluate_dataset(None, grids, S=common.strength_measure_funcs["invabs"], alpha=0.2, lam_max=lambda A: lam)
    ref = [_oracle_conv(g.A, common.strength_measure_funcs["invabs"](g.A), 0.2, lam, 0) for g in grids]
    assert np.allclose(got, ref, rtol=0, atol=1e-9)
    C = common.strength_measure_funcs["olson"](grids[0].A)               # parity: tests/test_zz_gpu_strength.py
    assert sp.isspmatrix_csr(C) and C.shape == grids[0].A.shape and C.nnz == grids[0].A.nnz


def test_evaluate_dataset_with_a_model_tail():
    """a stand-in model: random-init network outputs fed through the device-resident tail (agg_interp.py:469-484)"""
    import ns.model.agg_interp as ai
    from mlamg import problems
    common = _common()

    class Model:
        def forward(self, A, alpha):
            top_k, bf, ph = problems.random_gnn_outputs(A, alpha, seed=1)
            agg_T, P_T, labels, P = ai.forward_tail(A, top_k, bf, ph + np.float32(0.1))
            return agg_T, P_T, None, top_k, None

    class Broken:
        def forward(self, A, alpha):
            raise KeyError(-1)             # unreachable node: the reference scores the grid as 1.0

    grids = [_G(sp.csr_matrix(oml.poisson((14, 14))))]
    conv = common.evaluate_dataset(None, grids, model=Model(), alpha=0.15)
    assert 0.0 < conv[0] <= 1.0
    assert common.evaluate_dataset(None, grids, model=Broken(), alpha=0.15)[0] == 1.0


def _train_dataset():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "ml-amg_b200"))
    spec = importlib.util.spec_from_file_location("mlamg_utils_train_dataset", os.path.join(ROOT, "ml-amg_b200", "utils", "train_dataset.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_population_evaluation_reuses_grid_state_and_matches_per_call_results():
    """utils/train_dataset.py:81-138: the GA fitness loop, population x grids.  The grid-major batched evaluation
    (operator uploaded and Gauss-Seidel schedule built once per grid) gives the same convergence factors as
    independent evaluate_dataset calls, and those match the oracle's two-grid solver on the same P."""
    import ns.lib.sparse_tensor as nst
    td = _train_dataset()
    grids = [_G(sp.csr_matrix(oml.poisson((16, 14)))), _G(sp.csr_matrix(oml.poisson((9, 8, 7))))]
    model = td.StandInModel()
    population = [np.full(5, 0.1 * (i + 1)) for i in range(3)]
    conv = td.evaluate_population(population, grids, model, alpha=0.2)
    assert conv.shape == (3, 2) and np.all(conv > 0) and np.all(conv <= 1.0)
    for i, w in enumerate(population):
        single = td.evaluate_dataset(w, grids, model, alpha=0.2)
        assert np.array_equal(single, conv[i])
        # against the oracle: same P (the model is deterministic given the weights), reference two-grid loop
        model.load_flat_weights(w)
        for g, grid in enumerate(grids):
            P = nst.to_scipy(model.forward(grid.A, 0.2)[1]).astype(np.float64)
            x = np.random.RandomState(0).randn(grid.A.shape[1])
            x /= la.norm(x, 2)
            ref = rp.amg_2_v(grid.A, P, np.zeros(grid.A.shape[1]), x, error_tol=1e-6)[1]
            assert abs(conv[i, g] - ref) < 1e-8, (i, g, conv[i, g], ref)
    f = td.fitness(3, population[0], grids, model, alpha=0.2)
    assert abs(f - 1.0 / conv[0].mean()) < 1e-12
    f_rel = td.fitness(3, population[0], grids, model, benchmark=np.array([0.5, 0.25]), alpha=0.2)
    assert abs(f_rel - 1.0 / np.mean(conv[0] / np.array([0.5, 0.25]))) < 1e-12
