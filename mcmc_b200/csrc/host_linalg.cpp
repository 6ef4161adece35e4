// One-time host linear algebra on the preconditioner, the counterpart of
//   inv_precond_matrix  = BMO_MATOPS_INV(precond_matrix)         (src/hmc.cpp:58)
//   sqrt_precond_matrix = BMO_MATOPS_CHOL_LOWER(precond_matrix)  (src/hmc.cpp:59)
// O(d^3) once per call; the per-draw work happens on the device.  Column-major d x d.
#include <cmath>
#include <utility>
#include <vector>

#include "engine.h"

namespace mcmcb200
{

// A^-1 by LU with partial pivoting (what Eigen's inverse() uses for dynamic sizes).
bool host_inverse_colmajor(const double* A, int n, double* inv)
{
    std::vector<double> lu(A, A + (size_t)n * n);
    std::vector<int> perm(n);
    for (int i = 0; i < n; ++i) perm[i] = i;
#define LU(i, j) lu[(size_t)(j) * n + (i)]
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double big = std::fabs(LU(k, k));
        for (int i = k + 1; i < n; ++i) {
            const double v = std::fabs(LU(i, k));
            if (v > big) { big = v; piv = i; }
        }
        if (!(big > 0.0) || !std::isfinite(big)) return false;
        if (piv != k) {
            for (int j = 0; j < n; ++j) std::swap(LU(k, j), LU(piv, j));
            std::swap(perm[k], perm[piv]);
        }
        const double pivot = LU(k, k);
        for (int i = k + 1; i < n; ++i) LU(i, k) /= pivot;
        for (int j = k + 1; j < n; ++j) {
            const double f = LU(k, j);
            for (int i = k + 1; i < n; ++i) LU(i, j) -= LU(i, k) * f;
        }
    }
    std::vector<double> col(n);
    for (int c = 0; c < n; ++c) {
        for (int i = 0; i < n; ++i) col[i] = (perm[i] == c) ? 1.0 : 0.0;
        for (int i = 0; i < n; ++i) {
            double s = col[i];
            for (int j = 0; j < i; ++j) s -= LU(i, j) * col[j];
            col[i] = s;
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = col[i];
            for (int j = i + 1; j < n; ++j) s -= LU(i, j) * col[j];
            col[i] = s / LU(i, i);
        }
        for (int i = 0; i < n; ++i) inv[(size_t)c * n + i] = col[i];
    }
#undef LU
    return true;
}

// Lower Cholesky factor.  chol_mode MCMCB200_CHOL_EIGEN_LLT keeps A's entries in the strict upper
// triangle, which is what `(A).llt().matrixLLT()` hands back under the Eigen backend
// (include/BaseMatrixOps/include/core/cholesky.hpp:37); MCMCB200_CHOL_LOWER zeroes it
// (Armadillo backend, cholesky.hpp:31).
bool host_cholesky_colmajor(const double* A, int n, int chol_mode, double* L)
{
    for (size_t k = 0; k < (size_t)n * n; ++k) L[k] = A[k];
#define LL(i, j) L[(size_t)(j) * n + (i)]
    for (int j = 0; j < n; ++j) {
        double diag = LL(j, j);
        for (int k = 0; k < j; ++k) diag -= LL(j, k) * LL(j, k);
        if (!(diag > 0.0)) return false;
        const double r = std::sqrt(diag);
        LL(j, j) = r;
        for (int i = j + 1; i < n; ++i) {
            double v = LL(i, j);
            for (int k = 0; k < j; ++k) v -= LL(i, k) * LL(j, k);
            LL(i, j) = v / r;
        }
    }
    if (chol_mode == MCMCB200_CHOL_LOWER)
        for (int j = 1; j < n; ++j)
            for (int i = 0; i < j; ++i) LL(i, j) = 0.0;
#undef LL
    return true;
}

}  // namespace mcmcb200
