"""CPU: host-side logic of the extension that needs no GPU — the tridiagonal QL step of the Lanczos eigen-solve
(csrc/tridiag.h) compiled with g++ and held to scipy, and a numpy emulation of the device Lanczos loop with the same
stopping rule against the analytic spectrum (the device run is checked in tests/test_gpu_kernels.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg as sl
import scipy.sparse as sp

from helpers import ROOT
from oracle import multilevel as oml

SHIM = r"""
#include "tridiag.h"
extern "C" int ql_last_row(int m, double *d, const double *e, double *z) {
    std::vector<double> dv(d, d + m), ev(e, e + (m > 0 ? m - 1 : 0)), zv;
    const bool ok = mlamg::tridiag_ql_last_row(dv, ev, zv);
    for (int i = 0; i < m; i++) { d[i] = dv[i]; z[i] = zv[i]; }
    return ok ? 0 : 1;
}
"""


@pytest.fixture(scope="module")
def ql(tmp_path_factory):
    d = tmp_path_factory.mktemp("ql")
    src, so = d / "shim.cpp", d / "shim.so"
    src.write_text(SHIM)
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "ml-amg_b200", "csrc"), str(src),
                           "-o", str(so)])
    lib = ctypes.CDLL(str(so))

    def run(diag, off):
        dd = np.array(diag, dtype=np.float64)
        ee = np.array(off, dtype=np.float64)
        z = np.zeros_like(dd)
        vp = ctypes.c_void_p
        rc = lib.ql_last_row(len(dd), dd.ctypes.data_as(vp), ee.ctypes.data_as(vp), z.ctypes.data_as(vp))
        assert rc == 0
        return dd, z
    return run


def test_tridiagonal_ql_last_row_vs_scipy(ql):
    rs = np.random.RandomState(0)
    for m in (1, 2, 3, 17, 200):
        d, e = rs.randn(m), rs.randn(max(m - 1, 0))
        w, z = ql(d, e)
        if m == 1:
            assert w[0] == d[0] and z[0] == 1.0
            continue
        W, V = sl.eigh_tridiagonal(d, e)
        o = np.argsort(w)
        assert np.abs(w[o] - W).max() < 1e-13 * max(1.0, np.abs(W).max())
        assert np.abs(np.abs(z[o]) - np.abs(V[-1])).max() < 1e-13


def test_lanczos_stopping_rule_reaches_machine_precision(ql):
    """same recurrence, start vector, check schedule and stopping rule as csrc/eigen.cu, in numpy"""
    shape = (48, 40)
    A = oml.poisson(shape)
    n = A.shape[0]
    s = 1.0 / np.sqrt(A.diagonal())
    B = (sp.diags(s) @ A @ sp.diags(s)).tocsr()
    i = np.arange(n, dtype=np.uint32)
    h = i * np.uint32(2654435761)
    h ^= h >> np.uint32(15)
    h *= np.uint32(2246822519)
    h ^= h >> np.uint32(13)
    v = (0.5 + (h & 0xffff) / 65536.0) * np.where(h & 0x10000, -1.0, 1.0)
    v /= np.linalg.norm(v)
    vp, al, be, nxt, tol = np.zeros(n), [], [], 20, 1e-13
    theta = None
    for j in range(3000):
        w = B @ v
        al.append(v @ w)
        w = w - al[-1] * v - (be[-1] * vp if be else 0.0)
        be.append(np.linalg.norm(w))
        if j + 1 == nxt:
            d, z = ql(al, be[:-1])
            k = int(np.argmax(np.abs(d)))
            theta = abs(d[k])
            gap = np.min(np.abs(np.delete(np.abs(d), k) - theta))
            res = be[-1] * abs(z[k])
            if min(res, res * res / gap if gap > 0 else res) <= tol * theta:
                break
            nxt = j + 1 + max(10, (j + 1) // 4)
        vp, v = v, w / be[-1]
    exact = 1 + (np.cos(np.pi / 49) + np.cos(np.pi / 41)) / 2
    assert abs(theta - exact) <= 1e-13 * exact and j < 1000


def test_legacy_permutation_head_is_numpys():
    """mlamg_legacy_permutation_head (host routine of the extension) == RandomState(seed).permutation(n)[:k], the Lloyd
    seeding of ns/lib/graph.py:229-231, for many (seed, n) incl. the sizes of configs 1 and 2"""
    from mlamg._lib import lib, check
    for seed, n in [(0, 1), (0, 2), (3, 7), (0, 576), (1, 1000), (7, 65536), (0, 65536), (123456789, 100003),
                    (2 ** 32 - 1, 5000), (0, 2097152)]:
        k = max(1, int(np.ceil(0.1 * n)))
        out = np.empty(k, dtype=np.int32)
        check(lib.mlamg_legacy_permutation_head(seed, n, k, out.ctypes.data_as(ctypes.c_void_p)))
        assert np.array_equal(out, np.random.RandomState(seed).permutation(n)[:k]), (seed, n)


def test_profiler_mirror_nests_and_prints_like_the_reference(capsys):
    """ns/lib/profiler.py:4-52: disabled = silent no-op; enabled = hierarchical print when the root section closes;
    exceptions inside a section are swallowed, as in the reference (__exit__ returns True)"""
    from ns.lib.profiler import Profiler
    with Profiler("off"):
        pass
    assert capsys.readouterr().out == ""
    Profiler.enabled = True
    try:
        with Profiler("root"):
            with Profiler("child a"):
                pass
            with Profiler("child b"):
                with Profiler("grandchild"):
                    raise ValueError("swallowed")
        lines = capsys.readouterr().out.splitlines()
    finally:
        Profiler.enabled = False
        Profiler.current = None
    assert [ln.strip().split("]")[0] + "]" for ln in lines] == ["[root]", "[child a]", "[child b]", "[grandchild]"]
    assert [len(ln) - len(ln.lstrip()) for ln in lines] == [0, 2, 2, 4]
    assert all(ln.split("] ")[1].split("s")[0].replace(".", "").isdigit() for ln in lines)
