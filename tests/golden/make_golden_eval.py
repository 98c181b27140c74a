"""Generate tests/golden/ref_evaluation_sequences.npz by running the UNMODIFIED reference driver functions
/root/reference/utils/common.py::evaluate_ref_conv / evaluate_dataset (the canonical evaluation of SURVEY.md §3.1:
seed -> strength measure -> Lloyd aggregation -> smoothed aggregation incl. ARPACK -> amg_2_v) on small grids, for every
strength measure incl. the default 'olson'.

Run:  python tests/golden/make_golden_eval.py        (needs /root/reference; CPU only)

What executes unmodified: utils/common.py (loaded by path), ns/lib/graph.py, ns/lib/multigrid.py, ns/lib/sparse*.py.
Shims (third-party packages absent here, PARITY UNPINNED for what they stand for):
  pyamg.graph.lloyd_cluster / bellman_ford, pyamg.relaxation.relaxation.gauss_seidel   -> oracle.pyamg_restated (as make_golden.py)
  pyamg.strength.evolution_strength_of_connection                                       -> oracle.pyamg_restated
  pyamg.aggregation.lloyd_aggregation(C, ratio, distance) -> (AggOp, seeds)             -> the reference's OWN copy of that routine
        (ns/lib/graph.py:156-239 says it is pyamg's code at commit e3fb6fe) with rand=None, i.e. numpy's global stream
  matplotlib, torch_sparse, and the reference modules common.py imports but these functions never touch
  (ns.model.agg_interp, ns.model.data, ns.ga.*): empty modules.
`spla.eigs` is wrapped by a recorder (not altered): ARPACK's lambda_max per call is stored, so that a parity run can inject
the very omega the reference used.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as mg0                               # noqa: E402
from oracle import pyamg_restated as pr                 # noqa: E402
from oracle import multilevel as oml                    # noqa: E402
_spec = importlib.util.spec_from_file_location("mlamg_problems", os.path.join(ROOT, "ml-amg_b200", "mlamg", "problems.py"))
pb = importlib.util.module_from_spec(_spec)             # host-side generators only; loaded by path so that `ns` stays the reference's
_spec.loader.exec_module(pb)


class G:
    def __init__(self, A):
        self.A = A


def grids():
    out = {"poisson2d_16x14": sp.csr_matrix(oml.poisson((16, 14))).astype(np.float64),
           "delaunay_500": sp.csr_matrix(pb.delaunay_laplacian(500, seed=2)[0]),
           "voronoi_jump_delaunay_300": sp.csr_matrix(pb.voronoi_jump_problem(300, seed=4, mesh="delaunay", npts=300)[0])}
    for A in out.values():
        A.sort_indices()
    return out


def main():
    mg0.install_shims()
    pyamg = sys.modules["pyamg"]
    strength = types.ModuleType("pyamg.strength")
    strength.evolution_strength_of_connection = pr.evolution_strength_of_connection
    aggregation = types.ModuleType("pyamg.aggregation")
    pyamg.strength, pyamg.aggregation = strength, aggregation
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.update({"pyamg.strength": strength, "pyamg.aggregation": aggregation, "matplotlib": mpl, "matplotlib.pyplot": plt})
    sys.path.insert(0, REF)
    import ns.lib.graph as rgraph
    import ns.lib.multigrid as rmg
    import scipy.sparse.linalg as spla
    assert rgraph.__file__.startswith(REF) and rmg.__file__.startswith(REF)
    aggregation.lloyd_aggregation = lambda C, ratio=0.03, distance="unit", maxiter=10: rgraph.lloyd_aggregation(C, ratio, distance, maxiter)[::2]
    for name in ("ns.model.agg_interp", "ns.model.data", "ns.ga", "ns.ga.parga", "ns.ga.torch"):
        sys.modules[name] = types.ModuleType(name)
    import ns
    import ns.model
    ns.model.agg_interp, ns.model.data = sys.modules["ns.model.agg_interp"], sys.modules["ns.model.data"]
    ns.ga = sys.modules["ns.ga"]
    ns.ga.parga, ns.ga.torch = sys.modules["ns.ga.parga"], sys.modules["ns.ga.torch"]
    spec = importlib.util.spec_from_file_location("reference_common", os.path.join(REF, "utils", "common.py"))
    common = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(common)

    lam = []
    real_eigs = spla.eigs

    def eigs_recorder(*a, **k):
        out = real_eigs(*a, **k)
        lam.append(float(np.abs(out).item()))
        return out
    rmg.spla.eigs = eigs_recorder

    out = {}
    gs = grids()
    data = [G(A) for A in gs.values()]
    out["grid_names"] = np.array(list(gs))
    for name, A in gs.items():
        out[f"{name}_indptr"], out[f"{name}_indices"], out[f"{name}_data"] = A.indptr, A.indices, A.data
    for measure, S in common.strength_measure_funcs.items():
        lam.clear()
        conv = common.evaluate_ref_conv(data, S, alpha=0.2)
        out[f"ref_conv_{measure}"], out[f"ref_conv_{measure}_lam"] = conv, np.array(lam)
        print("evaluate_ref_conv", measure, conv)
    lam.clear()
    conv = common.evaluate_dataset(None, data, alpha=0.2)                     # S=None -> 'olson'
    out["dataset_conv_default"], out["dataset_conv_default_lam"] = conv, np.array(lam)
    print("evaluate_dataset default", conv)
    lam.clear()
    conv = common.evaluate_dataset(None, data, S=common.strength_measure_funcs["invabs"], alpha=0.3, omega=0.5)
    out["dataset_conv_invabs"], out["dataset_conv_invabs_lam"] = conv, np.array(lam)
    print("evaluate_dataset invabs", conv)
    np.savez_compressed(os.path.join(HERE, "ref_evaluation_sequences.npz"), **out)


if __name__ == "__main__":
    main()
