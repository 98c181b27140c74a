"""Mirror of the GA fitness evaluation of the reference (/root/reference/utils/train_dataset.py:81-138) on the B200
modules: the one caller that runs the hot path population x grids times per generation.

  evaluate_dataset(weights, dataset, model, alpha, omega)   :81-117   one individual over a list of grids
  fitness(generation, weights, ...)                         :120-138  1 / mean(conv) (optionally relative to a benchmark)
  evaluate_population(population, dataset, model, ...)      (added)   the whole population, GRID-major

What is amortised (the reference rebuilds everything for every (individual, grid) pair): per grid, the operator is
uploaded once, its Gauss-Seidel dependency schedule is built once, the start vector `RandomState(0).randn / ||.||`
and the zero right-hand side are kept; per (individual, grid) only the network tail (Bellman-Ford aggregates,
`P = P_hat Agg`), `P^T A P`, its factorisation and the two-grid iteration run.  `model` is any object with
`forward(A, alpha) -> (agg_T, P_T, ...)` and, optionally, `load_flat_weights(weights)` — the GNN layers themselves are
out of scope; `StandInModel` below produces random-init outputs through the device tail of
ns/model/agg_interp.py so that the loop can be exercised and timed.
"""
import traceback

import numpy as np
import numpy.linalg as la
import scipy.sparse as sp
import torch

import ns.lib.multigrid
import ns.lib.sparse_tensor
import ns.model.agg_interp as agg_interp
from ns.lib.profiler import Profiler
from mlamg import core


class GridState:
    """per-grid state shared by every individual of the population"""

    def __init__(self, grid):
        self.A = grid.A if hasattr(grid, 'A') else grid
        self.A = sp.csr_matrix(self.A)
        n = self.A.shape[1]
        x = np.random.RandomState(0).randn(n)                       # :104-105
        self.x0 = x / la.norm(x, 2)
        self.b = np.zeros(n)
        self.solver_state = {}                                      # device copy of A + Gauss-Seidel schedule (amg_2_v fills it)

    @property
    def device_A(self):
        if 'A' not in self.solver_state:
            self.solver_state['A'] = core.DeviceCSR.wrap(self.A)
        return self.solver_state['A']


def _as_states(dataset):
    return [g if isinstance(g, GridState) else GridState(g) for g in dataset]


def _two_grid_conv(state, P, num_pre_relax, num_post_relax):
    res = ns.lib.multigrid.amg_2_v(state.A, P, state.b, state.x0.copy(), error_tol=1e-6, pre_smoothing_steps=num_pre_relax,
                                   post_smoothing_steps=num_post_relax, _state=state.solver_state)[1]        # :114
    return 1.0 if np.isnan(res) else float(res)


def evaluate_dataset(weights, dataset, model=None, alpha=0.3, omega=2. / 3., num_pre_relax=1, num_post_relax=1):
    """train_dataset.py:81-117: convergence factor of the two-grid solver built from the model's P on every grid.
    `dataset` may hold Grid objects / matrices or GridState objects (reused across calls)."""
    if weights is not None and hasattr(model, 'load_flat_weights'):
        model.load_flat_weights(weights)
    states = _as_states(dataset)
    conv = np.zeros(len(states))
    for i, st in enumerate(states):
        try:
            with torch.no_grad():
                P_T = None
                with Profiler('model inferencing'):                 # :97 (wall clock, device time and kernel count of the tail)
                    P_T = model.forward(st.device_A if getattr(model, 'accepts_device_matrix', False) else st.A, alpha)[1]
            P = ns.lib.sparse_tensor.to_scipy(P_T)
        except Exception:                                           # noqa: BLE001  (:99-102: score the grid as 1.0)
            print(f'Could not evaluate grid {i}: {traceback.format_exc()}')
            conv[i] = 1.0
            continue
        conv[i] = _two_grid_conv(st, P, num_pre_relax, num_post_relax)
    return conv


def evaluate_population(population, dataset, model, alpha=0.3, omega=2. / 3., num_pre_relax=1, num_post_relax=1):
    """conv[i, g] for every individual i (flat weight vector) and grid g.  Grid-major: the per-grid state is built once
    and serves the whole population."""
    states = _as_states(dataset)
    conv = np.ones((len(population), len(states)))
    for g, st in enumerate(states):
        for i, w in enumerate(population):
            conv[i, g] = evaluate_dataset(w, [st], model, alpha, omega, num_pre_relax, num_post_relax)[0]
    return conv


def fitness(generation, weights, dataset, model, benchmark=None, batch_size=None, alpha=0.3):
    """train_dataset.py:120-138: 1 / mean convergence factor over the (optionally subsampled) training grids;
    `benchmark` (reference convergence per grid) switches to the relative measure."""
    states = _as_states(dataset)
    idx = np.arange(len(states))
    if batch_size is not None and batch_size < len(states):
        idx = np.random.RandomState(generation).choice(len(states), size=batch_size, replace=False)
    raw = evaluate_dataset(weights, [states[i] for i in idx], model, alpha=alpha)
    if benchmark is not None:
        return 1. / np.average(raw / np.asarray(benchmark)[idx])
    return 1. / np.average(raw)


class StandInModel:
    """Random-init stand-in for FullAggNet (agg_interp.py:432-486): the three network outputs are drawn as
    relu(N(0,1)) (+ a floor on P_hat so that no aggregate is left without weight), seeded by the flat weight vector;
    the tail (Bellman-Ford aggregates, Agg, P = P_hat Agg) is the device tail of ns/model/agg_interp.py."""
    accepts_device_matrix = True

    def __init__(self, seed=0):
        self.seed = int(seed)

    def load_flat_weights(self, weights):
        self.seed = int(abs(float(np.sum(weights))) * 1e6) % (2 ** 31)

    def forward(self, A, alpha):
        Ad = core.DeviceCSR.wrap(A)
        n, nnz = Ad.shape[0], Ad.nnz
        g = torch.Generator().manual_seed(self.seed)
        k = int(np.ceil(alpha * n))
        top_k = torch.sort(torch.topk(torch.rand(n, generator=g), k).indices).values
        bf = torch.relu(torch.randn(nnz, generator=g)).to(torch.float32)
        ph = (torch.relu(torch.randn(nnz, generator=g)) + 0.1).to(torch.float32)
        agg_T, P_T, labels, P = agg_interp.forward_tail(Ad, top_k, bf.cuda(), ph.cuda())
        return agg_T, P_T, bf, top_k, None
