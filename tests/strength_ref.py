"""Test helper: numpy statements of the CONTRACT of every kernel in csrc/strength.cu (one function per C entry point,
plain loops, small cases).  Used twice: as stand-ins for the C-ABI wrappers in the CPU wiring test
(tests/test_strength_cpu.py) and as the per-kernel reference of the GPU test (tests/test_zz_gpu_strength.py).
The end-to-end reference is the oracle's restatement of pyamg (oracle.pyamg_restated.evolution_strength_of_connection)."""
import numpy as np

SQRT_EPS = np.sqrt(np.finfo(float).eps)


def rows_of(indptr):
    return np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))


def evolution_step(indptr, indices, data, inv_rho):
    """-> (s_val, dinv_a_val, flag)"""
    n = len(indptr) - 1
    r = rows_of(indptr)
    diag = np.zeros(n)
    present = np.zeros(n, dtype=bool)
    on = indices == r
    diag[r[on]] = data[on]
    present[r[on]] = True
    dinv = np.ones(n)
    nz = diag != 0
    dinv[nz] = 1.0 / diag[nz]
    t = data * dinv[r]
    s = inv_rho * t
    return np.where(on, 1.0 - s, -s), t, int(not present.all())


def incomplete_matmul(Ap, Aj, Ax, Bp, Bj, Bx, Sp, Sj):
    out = np.zeros(len(Sj))
    for row in range(len(Sp) - 1):
        for k in range(Sp[row], Sp[row + 1]):
            col = Sj[k]
            a, b, acc = Ap[row], Bp[col], 0.0
            while a < Ap[row + 1] and b < Bp[col + 1]:
                if Aj[a] == Bj[b]:
                    acc = acc + Ax[a] * Bx[b]
                    a += 1
                    b += 1
                elif Aj[a] < Bj[b]:
                    a += 1
                else:
                    b += 1
            out[k] = acc
    return out


def evolution_measure(indptr, indices, data):
    n = len(indptr) - 1
    r = rows_of(indptr)
    diag = np.zeros(n)
    on = indices == r
    diag[r[on]] = data[on]
    d = diag[r]
    out = np.zeros_like(data)
    nzm = data != 0
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = d[nzm] / data[nzm]
    m = np.abs(1.0 - ratio)
    m[np.abs(ratio) < 1e-4] = 0.0
    m[(d[nzm] * data[nzm]) < 0.0] = 0.0
    small = (m != 0) & (m < SQRT_EPS)
    m[small] = 1e-4
    out[nzm] = m
    return out


def distance_filter(indptr, indices, data, epsilon):
    out = data.copy()
    for i in range(len(indptr) - 1):
        sl = slice(indptr[i], indptr[i + 1])
        off = indices[sl] != i
        mn = out[sl][off].min() if off.any() else np.finfo(float).max
        with np.errstate(over="ignore"):
            thr = epsilon * mn
        kill = off & (out[sl] >= thr)
        out[sl] = np.where(kill, 0.0, out[sl])
    return out


def _lookup(indptr, indices, data, r, c):
    for q in range(indptr[r], indptr[r + 1]):
        if indices[q] == c:
            return True, data[q]
    return False, 0.0


def evolution_symmetrize(Ap, Aj, Mp, Mj, Mx, symmetrize):
    out = np.zeros(len(Aj))
    for i in range(len(Ap) - 1):
        for j in range(Ap[i], Ap[i + 1]):
            c = Aj[j]
            if c == i:
                out[j] = 1.0
                continue
            fa, a = _lookup(Mp, Mj, Mx, i, c)
            if not symmetrize:
                out[j] = a
                continue
            fb, b = _lookup(Mp, Mj, Mx, c, i)
            out[j] = 0.5 * (a + b) if (fa and fb) else 0.5 * (a if fa else b)
    return out


def invert_scale_rows(indptr, data):
    inv = 1.0 / data
    out = np.empty_like(data)
    for i in range(len(indptr) - 1):
        sl = slice(indptr[i], indptr[i + 1])
        mx = max(np.abs(inv[sl]).max(), np.finfo(float).tiny) if indptr[i + 1] > indptr[i] else np.finfo(float).tiny
        out[sl] = inv[sl] * (1.0 / mx)
    return out


def pattern_add(Ap, Aj, w, Ep, Ej, Ex):
    out = np.array(w, dtype=float, copy=True)
    for i in range(len(Ap) - 1):
        for j in range(Ap[i], Ap[i + 1]):
            f, e = _lookup(Ep, Ej, Ex, i, Aj[j])
            if f:
                out[j] = e + w[j]
    return out
