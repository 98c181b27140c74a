"""CPU: the call SEQUENCE of the mirrored evaluation drivers (ml-amg_b200/utils/common.py: seeding protocol, which measure,
which aggregation arguments, which start vector, which solver arguments) against the convergence factors the UNMODIFIED
reference drivers returned (tests/golden/make_golden_eval.py).  The device modules the drivers call are replaced — in this
test only — by the oracle's statements of the same functions, so what is exercised is the driver logic itself; the device
modules are held to the same numbers on the GPU (tests/test_zzz_gpu_modes.py)."""
import importlib.util
import os

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import ROOT, load_eval_golden
from oracle import pyamg_restated as pr, reference_path as rp


class _G:
    def __init__(self, A):
        self.A = A


class _Host:
    def __init__(self, M):
        self.M = sp.csr_matrix(M)

    def to_scipy(self):
        return self.M


@pytest.fixture
def common(monkeypatch):
    import ns.lib.graph
    import ns.lib.multigrid
    from mlamg import strength
    monkeypatch.setattr(ns.lib.graph, "lloyd_aggregation", rp.lloyd_aggregation)
    monkeypatch.setattr(ns.lib.multigrid, "smoothed_aggregation_jacobi",
                        lambda A, Agg, omega=None, lam_max=None: rp.smoothed_aggregation_jacobi(A, Agg, omega=(4.0 / 3.0) / lam_max))
    monkeypatch.setattr(ns.lib.multigrid, "amg_2_v", rp.amg_2_v)

    def olson(A, **kw):
        A = sp.csr_matrix(A)
        return _Host(pr.evolution_strength_of_connection(A) + sp.csr_matrix((1.0 / np.abs(A.data), A.indices, A.indptr), A.shape))

    def evolution(A, **kw):
        A = sp.csr_matrix(A)
        return _Host(pr.evolution_strength_of_connection(A) + sp.csr_matrix((np.ones_like(A.data), A.indices, A.indptr), A.shape) * 0.1)
    monkeypatch.setattr(strength, "olson_measure", olson)
    monkeypatch.setattr(strength, "evolution_measure_plus_pattern", evolution)
    spec = importlib.util.spec_from_file_location("mlamg_utils_common_cpu", os.path.join(ROOT, "ml-amg_b200", "utils", "common.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_driver_sequences_reproduce_the_reference_drivers(common):
    z, grids = load_eval_golden()
    data = [_G(A) for A in grids.values()]

    def lam_of(key):
        table = {A.shape[0]: float(v) for A, v in zip(grids.values(), z[key])}
        return lambda A: table[A.shape[0]]

    assert sorted(common.strength_measure_funcs) == ["abs", "evolution", "invabs", "olson", "unit"]
    for measure in ("abs", "evolution", "invabs", "unit", "olson"):
        got = common.evaluate_ref_conv(data, common.strength_measure_funcs[measure], alpha=0.2, lam_max=lam_of(f"ref_conv_{measure}_lam"))
        assert np.allclose(got, z[f"ref_conv_{measure}"], rtol=0, atol=1e-10), (measure, got, z[f"ref_conv_{measure}"])
    got = common.evaluate_dataset(None, data, alpha=0.2, lam_max=lam_of("dataset_conv_default_lam"))
    assert np.allclose(got, z["dataset_conv_default"], rtol=0, atol=1e-10)
    got = common.evaluate_dataset(None, data, S=common.strength_measure_funcs["invabs"], alpha=0.3, omega=0.5,
                                  lam_max=lam_of("dataset_conv_invabs_lam"))
    assert np.allclose(got, z["dataset_conv_invabs"], rtol=0, atol=1e-10)


def test_a_failing_model_scores_one_and_parse_bool(common):
    z, grids = load_eval_golden()

    class Broken:
        def forward(self, A, alpha):
            raise KeyError(-1)
    conv = common.evaluate_dataset(None, [_G(next(iter(grids.values())))], model=Broken(), alpha=0.2)
    assert conv[0] == 1.0                                             # utils/common.py:63-70
    assert common.parse_bool_str("T") and common.parse_bool_str("true") and not common.parse_bool_str("0")
