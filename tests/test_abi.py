"""CPU: the C-ABI library loads and exports every symbol include/mlamg.h declares; the Python
binding table mirrors the header; without a GPU every compute path fails loudly (no fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from helpers import ROOT

HEADER = os.path.join(ROOT, "include", "mlamg.h")
LIB = os.path.join(ROOT, "ml-amg_b200", "mlamg", "libmlamg_b200.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mlamg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build with `python ml-amg_b200/build.py`"
    lib = ctypes.CDLL(LIB)
    names = declared_symbols()
    assert len(names) >= 40
    for name in names:
        assert hasattr(lib, name), f"{name} declared in mlamg.h but not exported"


def test_binding_table_matches_header():
    import mlamg
    from mlamg import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), f"{name}: header has {len(params)} parameters, binding {len(args)}"
    assert _lib.lib.mlamg_version() >= 100


def test_sass_is_sm100a():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import mlamg
    import ns.lib.graph as g
    import ns.lib.multigrid as mg
    import scipy.sparse as sp
    A = sp.eye(4, format="csr")
    with pytest.raises(RuntimeError):
        mlamg.build_hierarchy(A)
    with pytest.raises(RuntimeError):
        g.lloyd_aggregation(A, ratio=0.5)
    with pytest.raises(RuntimeError):
        mg.jacobi(A, np.ones(4), np.zeros(4))


def test_host_side_argument_errors_match_reference():
    import ns.lib.graph as g
    import ns.lib.multigrid as mg
    import scipy.sparse as sp
    A = sp.eye(4, format="csr")
    with pytest.raises(ValueError):
        g.lloyd_aggregation(A, ratio=0.0)
    with pytest.raises(ValueError):
        g.lloyd_aggregation(A, ratio=1.5)
    with pytest.raises(TypeError):
        g.lloyd_aggregation(A.tocoo(), ratio=0.5)
    with pytest.raises(ValueError):
        g.lloyd_aggregation(A, ratio=0.5, distance="bogus")
    with pytest.raises(TypeError):
        g.lloyd_aggregation(A, ratio=0.5, rand="seed")
    with pytest.raises(RuntimeError, match="One of res_tol or error_tol must be set!"):
        mg.amg_2_v(A, A, np.zeros(4), np.zeros(4))


def test_lloyd_seeding_is_the_reference_rule():
    import mlamg
    s = mlamg.lloyd_seeds(1000, 0.1, 0)
    assert np.array_equal(s, np.random.RandomState(0).permutation(1000)[:100])
    assert len(mlamg.lloyd_seeds(1001, 0.1, 3)) == 101          # ceil
    rs = np.random.RandomState(5)
    ref = np.random.RandomState(5).permutation(50)[:5]
    assert np.array_equal(mlamg.lloyd_seeds(50, 0.1, rs), ref)
