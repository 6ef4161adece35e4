// How far is the CPU baseline's dense algebra from what a real-Eigen build of the reference could do?
// The reference's C2 hot loop is 12 dense 128 x 128 products per draw even with M = I (inv_precond_matrix = eye:
// src/hmc.cpp:57-59,160,171,184).  This probe times exactly those expressions through oracle/standin/Eigen/Dense with the
// baseline's flags (oracle/Makefile FAST) and prints GFLOP/s; tools/standin_gemv_probe.py sets it beside OpenBLAS dgemv
// (numpy) on the same core — an upper bound for a vectorised Eigen GEMV.  Test/measurement tooling, not product code.
#include <Eigen/Dense>
#include <chrono>
#include <cstdio>
#include <cstdlib>

int main(int argc, char** argv)
{
    const int d = argc > 1 ? std::atoi(argv[1]) : 128;
    const int reps = argc > 2 ? std::atoi(argv[2]) : 200000;
    Eigen::MatrixXd M = Eigen::MatrixXd::Identity(d, d);
    Eigen::VectorXd p(d), x(d);
    for (int i = 0; i < d; ++i) { p(i) = 0.001 * (i + 1); x(i) = 0.0; }
    const double eps = 1e-9;
    double K = 0.0;
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) {
        x += eps * M * p;                 // src/hmc.cpp:171
        K += p.dot(M * p) / 2.0;          // src/hmc.cpp:160,184
        p(r % d) += 1e-12;                // keep the loop from being hoisted
    }
    auto t1 = std::chrono::steady_clock::now();
    const double s = std::chrono::duration<double>(t1 - t0).count();
    const double flop = 2.0 * reps * 2.0 * double(d) * double(d);
    std::printf("{\"d\": %d, \"reps\": %d, \"seconds\": %.4f, \"gemv_per_s\": %.1f, \"gflops\": %.2f, \"check\": %.6g}\n", d, reps, s,
                2.0 * reps / s, flop / s * 1e-9, K + x(0));
    return 0;
}
