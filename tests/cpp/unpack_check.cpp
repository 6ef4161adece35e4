// CPU-only check of the draws_out adapter of include/mcmc_b200.hpp: the library returns [C][d][n_keep] (every chain's block
// is the reference's column-major n_keep x d Mat_t, SURVEY Q23, transposed on the device); unpack() copies the blocks into
// the Cube_t's matrices, chains spread over host threads.
#include <cstdio>
#include <random>
#include <vector>

#include "mcmc_b200.hpp"

int main()
{
    std::mt19937_64 gen(7);
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    const size_t shapes[][3] = {{1, 1, 1}, {1, 5, 3}, {3, 20, 128}, {7, 33, 65}, {2, 1, 200}, {5, 64, 1}, {64, 700, 100}, {40, 1000, 128}};
    for (const auto& sh : shapes) {
        const size_t C = sh[0], T = sh[1], d = sh[2];
        std::vector<double> buf(C * T * d);
        for (double& v : buf) v = u(gen);
        mcmc::Cube_t cube;
        mcmc::b200_detail::unpack(buf.data(), C, T, d, cube);
        if (cube.n_mat() != C) { std::printf("FAIL n_mat\n"); return 1; }
        for (size_t c = 0; c < C; ++c) {
            const mcmc::Mat_t& m = cube.mat(c);
            if (size_t(m.rows()) != T || size_t(m.cols()) != d) { std::printf("FAIL shape\n"); return 1; }
            for (size_t t = 0; t < T; ++t)
                for (size_t j = 0; j < d; ++j)
                    if (m(t, j) != buf[(c * d + j) * T + t]) { std::printf("FAIL value C=%zu T=%zu d=%zu\n", C, T, d); return 1; }
        }
    }
    std::printf("unpack ok\n");
    return 0;
}
