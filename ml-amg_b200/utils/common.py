"""Mirror of the reference's evaluation helpers (/root/reference/utils/common.py:24-111) — the canonical
call sequences of the hot path — running on the B200 modules under `ml-amg_b200/ns`.

  strength_measure_funcs   :24-30   'abs', 'invabs', 'unit', 'evolution', 'olson' (host CSR in, host CSR out, as in the
                                    reference); the two pyamg-based measures run pyamg's evolution strength of connection
                                    on the device (mlamg/strength.py, csrc/strength.cu), incl. its Arnoldi estimate of
                                    rho(D^-1 A) from numpy's global stream — callers seed it exactly as the reference does
  evaluate_dataset         :40-82   Lloyd aggregates + SA prolongator (or a model's learned P) -> two-grid
                                    convergence factor per grid
  evaluate_ref_conv        :84-111  the same with the baseline strength measure

Differences: an optional trailing `lam_max` (callable A -> |lambda_max(D^-1 A)|) lets a parity run inject the
oracle's spectral radius (the reference calls ARPACK); plotting on failure is dropped.
"""
import traceback

import numpy as np
import numpy.linalg as la
import scipy.sparse as sp

import ns.lib.graph
import ns.lib.multigrid
import ns.lib.sparse_tensor


def _evolution(A):
    """:27  pyamg.strength.evolution_strength_of_connection(A) + 0.1 * pattern(A)"""
    from mlamg import strength
    return strength.evolution_measure_plus_pattern(sp.csr_matrix(A)).to_scipy()


def _olson(A):
    """:30  pyamg.strength.evolution_strength_of_connection(A) + 1/|A|  (the default of evaluate_dataset, :52)"""
    from mlamg import strength
    return strength.olson_measure(sp.csr_matrix(A)).to_scipy()


strength_measure_funcs = {
    'abs': lambda A: abs(A),
    'evolution': _evolution,
    'invabs': lambda A: sp.csr_matrix((1.0 / np.abs(A.data), A.indices, A.indptr), A.shape),
    'unit': lambda A: sp.csr_matrix((np.ones_like(A.data), A.indices, A.indptr), A.shape),
    'olson': _olson,
}


def parse_bool_str(v):
    v = v.lower()
    return v == 't' or v == 'true'


def _conv(A, P, neumann_solve, omega, x):
    res = ns.lib.multigrid.amg_2_v(A, P, np.zeros(A.shape[1]), x, res_tol=1e-10, singular=neumann_solve,
                                   jacobi_weight=omega)[1]
    return 0.0 if np.isnan(res) else res


def evaluate_dataset(weights, dataset, model=None, S=None, neumann_solve=False, alpha=0.3, omega=2. / 3., gen=None,
                     lam_max=None):
    """common.py:40-82.  `model` is any object with `forward(A, alpha) -> (agg_T, P_T, ...)` (weights are loaded by
    the caller's `load_state_dict` when the model offers it)."""
    if model is not None and weights is not None and hasattr(model, 'load_flat_weights'):
        model.load_flat_weights(weights)
    conv = np.zeros(len(dataset))
    for i in range(len(dataset)):
        A = dataset[i].A
        np.random.seed(0)
        if model is not None:
            try:
                agg_T, P_T = model.forward(A, alpha)[:2]
                P = ns.lib.sparse_tensor.to_scipy(P_T)
            except Exception:                                  # noqa: BLE001  (reference: score the grid as 1.0)
                print(f'Could not evaluate grid {i}: {traceback.format_exc()}')
                conv[i] = 1.0
                continue
        else:
            # :51-58 (the reference also builds the Lloyd aggregates when a model is given, and never uses them)
            C = strength_measure_funcs['olson'](A) if S is None else S(A)
            L_Agg, _, _ = ns.lib.graph.lloyd_aggregation(C, ratio=alpha, distance='same', rand=0)
            P = ns.lib.multigrid.smoothed_aggregation_jacobi(A, L_Agg, lam_max=None if lam_max is None else lam_max(A))
        x = np.random.RandomState(0).randn(A.shape[1])
        x /= la.norm(x, 2)
        conv[i] = _conv(A, P, neumann_solve, omega, x)
    return conv


def evaluate_ref_conv(dataset, strength_measure_func, neumann_solve=False, alpha=0.3, omega=2. / 3., lam_max=None):
    """common.py:84-111 (pyamg's lloyd_aggregation there = the repo's own copy with the global numpy RNG)."""
    conv = np.zeros(len(dataset))
    for i in range(len(dataset)):
        A = dataset[i].A
        np.random.seed(0)
        C = strength_measure_func(A)
        Agg, _, _ = ns.lib.graph.lloyd_aggregation(C, ratio=alpha, distance='same')
        P = ns.lib.multigrid.smoothed_aggregation_jacobi(A, Agg, lam_max=None if lam_max is None else lam_max(A))
        np.random.seed(0)
        x = np.random.randn(A.shape[1])
        x /= la.norm(x, 2)
        np.random.seed()
        conv[i] = _conv(A, P, neumann_solve, omega, x)
    return conv
