// TEST ONLY (CPU).  The product's one-time host algebra on precond_mat / cov_mat (mcmc_b200/csrc/host_linalg.cpp) against the
// operations the reference performs on the same matrix — BMO_MATOPS_INV = A.inverse(), BMO_MATOPS_CHOL_LOWER = A.llt().matrixLLT()
// (include/BaseMatrixOps/include/core/inv.hpp:34, cholesky.hpp:37; src/hmc.cpp:58-59) — evaluated through the stand-in Eigen the
// reference is compiled against in oracle/_ref.  Bit for bit: with the same inverse and factor on the device, a dense-mass chain in
// STRICT arithmetic starts from the same operands as the reference's.
#include <Eigen/Dense>
#include <cstdio>
#include <random>
#include <vector>

namespace mcmcb200
{
bool host_inverse_colmajor(const double* A, int n, double* inv);
bool host_cholesky_colmajor(const double* A, int n, int chol_mode, double* L);
}

int main()
{
    std::mt19937_64 eng(7);
    std::normal_distribution<double> nd;
    int bad_inv = 0, bad_chol = 0, total = 0;
    const int sizes[] = {1, 2, 3, 7, 16, 33, 64, 128};
    for (int n : sizes)
        for (int rep = 0; rep < 12; ++rep) {
            Eigen::MatrixXd a(n, n), M(n, n);
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) a(i, j) = nd(eng);
            for (int i = 0; i < n; ++i)
                for (int j = 0; j <= i; ++j) {
                    double s = 0;
                    for (int k = 0; k < n; ++k) s += a(i, k) * a(j, k);
                    M(i, j) = M(j, i) = s / n + (i == j ? 0.5 + rep : 0.0);
                }
            const Eigen::MatrixXd Minv = M.inverse();
            const Eigen::MatrixXd L = M.llt().matrixLLT();
            std::vector<double> A(size_t(n) * n), inv(size_t(n) * n), Lc(size_t(n) * n);
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) A[size_t(j) * n + i] = M(i, j);
            if (!mcmcb200::host_inverse_colmajor(A.data(), n, inv.data()) || !mcmcb200::host_cholesky_colmajor(A.data(), n, 1, Lc.data())) return 2;
            bool ei = true, ec = true;
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) {
                    if (inv[size_t(j) * n + i] != Minv(i, j)) ei = false;
                    if (Lc[size_t(j) * n + i] != L(i, j)) ec = false;
                }
            bad_inv += !ei;
            bad_chol += !ec;
            ++total;
        }
    std::printf("cases %d inverse_differs %d cholesky_differs %d\n", total, bad_inv, bad_chol);
    return (bad_inv || bad_chol) ? 1 : 0;
}
