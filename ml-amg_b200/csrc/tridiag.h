// Host-only helper of the Lanczos eigen-solve (eigen.cu); plain C++ so that the CPU test-suite can compile and check it
// without a GPU (tests/test_host_logic.py).
#pragma once
#include <math.h>
#include <vector>

namespace mlamg {

// Eigenvalues of the symmetric tridiagonal (d[0..m), e[0..m-1) off-diagonal) by implicit QL, carrying the LAST row of
// the eigenvector matrix (its entries are the s_m of the Lanczos residual formula).  O(m^2).
inline bool tridiag_ql_last_row(std::vector<double> &d, std::vector<double> &e, std::vector<double> &zlast) {
    const int m = (int)d.size();
    zlast.assign(m, 0.0);
    if (m == 0) return true;
    zlast[m - 1] = 1.0;
    e.resize(m, 0.0);
    for (int l = 0; l < m; l++) {
        int iter = 0, mm;
        do {
            for (mm = l; mm < m - 1; mm++) {
                const double dd = fabs(d[mm]) + fabs(d[mm + 1]);
                if (fabs(e[mm]) <= 2.3e-16 * dd) break;
            }
            if (mm != l) {
                if (iter++ == 200) return false;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = hypot(g, 1.0);
                g = d[mm] - d[l] + e[l] / (g + (g >= 0.0 ? fabs(r) : -fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = mm - 1; i >= l; i--) {
                    double f = s * e[i];
                    const double b = c * e[i];
                    e[i + 1] = (r = hypot(f, g));
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[mm] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    d[i + 1] = g + (p = s * r);
                    g = c * r - b;
                    f = zlast[i + 1];
                    zlast[i + 1] = s * zlast[i] + c * f;
                    zlast[i] = c * zlast[i] - s * f;
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[mm] = 0.0;
            }
        } while (mm != l);
    }
    return true;
}

}  // namespace mlamg
