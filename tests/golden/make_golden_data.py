"""Generate tests/golden/ref_grids_1d.npz by running the UNMODIFIED reference module /root/reference/ns/model/data.py
(`Grid.structured_1d_poisson_dirichlet`, `Grid.structured_1d_poisson_neumann`, `Grid.save` / `Grid.load`).

Run:  python tests/golden/make_golden_data.py        (needs /root/reference; CPU only)

data.py imports plotting / meshing / graph packages at module top that are not installed here (torch_geometric, pyamg,
matplotlib, pygmsh); they are shimmed by EMPTY modules — none of them is touched by the functions exercised.
"""
import os
import sys
import tempfile
import types

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def install_shims():
    for name in ("torch_geometric", "pyamg", "pyamg.gallery", "pyamg.gallery.mesh", "pyamg.gallery.fem", "matplotlib",
                 "matplotlib.pyplot", "pygmsh", "torch_sparse"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    sys.modules["pyamg"].gallery = sys.modules["pyamg.gallery"]
    sys.modules["pyamg.gallery"].mesh = sys.modules["pyamg.gallery.mesh"]
    sys.modules["pyamg.gallery"].fem = sys.modules["pyamg.gallery.fem"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    import ns.model.data as rdata
    assert rdata.__file__.startswith(REF)
    return rdata


def main():
    rdata = install_shims()
    if len(sys.argv) == 3 and sys.argv[1] == "--check-load":      # the reference's own loader on a file written by the mirror
        g = rdata.Grid.load(sys.argv[2])
        A = sp.csr_matrix(g.A)
        print("LOADED", A.shape[0], A.nnz, repr(float(abs(A).sum())), np.asarray(g.x).shape[0], sorted(k for k in g.extra if k != "filename"))
        return
    out = {}
    cases = {"dirichlet_7": (rdata.Grid.structured_1d_poisson_dirichlet, 7, (0, 1)),
             "dirichlet_12_shifted": (rdata.Grid.structured_1d_poisson_dirichlet, 12, (-1.0, 2.5)),
             "neumann_6": (rdata.Grid.structured_1d_poisson_neumann, 6, (0, 1)),
             "neumann_9_shifted": (rdata.Grid.structured_1d_poisson_neumann, 9, (0.5, 3.0))}
    for key, (fn, n, xdim) in cases.items():
        g = fn(n, xdim)
        A = sp.csr_matrix(g.A)
        A.sort_indices()
        out[f"{key}_n"], out[f"{key}_xdim"] = n, np.array(xdim, dtype=float)
        out[f"{key}_indptr"], out[f"{key}_indices"], out[f"{key}_data"] = A.indptr, A.indices, A.data
        out[f"{key}_x"] = np.asarray(g.x)
    # the file format, written by the reference itself
    g = rdata.Grid.structured_1d_poisson_neumann(5)
    g.extra = {"note": "written by the reference"}
    path = os.path.join(HERE, "ref_written_by_reference.grid")
    g.save(path)
    back = rdata.Grid.load(path)
    assert (back.A != g.A).nnz == 0
    np.savez_compressed(os.path.join(HERE, "ref_grids_1d.npz"), **out)
    print("ok", sorted(cases), os.path.getsize(path), "bytes of .grid")


if __name__ == "__main__":
    main()
