// Host side of MCMCB200_RNG_MT19937_TAPE: replay the reference's random stream.
//
// The reference seeds one std::mt19937_64 per call (src/hmc.cpp:68, rand_engine_t =
// std::mt19937_64, include/misc/mcmc_options.hpp:101) and draws
//   * normals with bmo::stats::rnorm: a NEW std::normal_distribution<fp_t> for every variate
//     (include/BaseMatrixOps/include/stats/rnorm.hpp:57-59, vector form :120-128), so the
//     distribution's cached second variate is always thrown away;
//   * uniforms with bmo::stats::runif: a_adj = nextafter(0, 1), then
//     std::uniform_real_distribution<fp_t>(a_adj, 1) (stats/runif.hpp:58-63).
// That stream is serial and standard-library specific, so it cannot be generated on the
// GPU bit-for-bit; for drop-in parity the library generates, per chain, exactly the variates
// the reference would consume and the kernels read them in order:
//   [n_pre_normals normals]  then per draw  [d normals][1 uniform]
// (n_pre_normals = d for NUTS / RM-HMC; 0 for HMC / MALA).
#include <cmath>
#include <random>

#include "engine.h"

namespace mcmcb200
{

void host_mt19937_tape(uint64_t seed, long long n_pre_normals, long long n_draws, int d, double* out)
{
    std::mt19937_64 engine(seed);
    long long pos = 0;
    auto one_normal = [&]() {
        std::normal_distribution<double> dist(0.0, 1.0);
        return 0.0 + 1.0 * dist(engine);
    };
    const double lo = std::nextafter(0.0, 1.0);
    for (long long i = 0; i < n_pre_normals; ++i) out[pos++] = one_normal();
    for (long long t = 0; t < n_draws; ++t) {
        for (int j = 0; j < d; ++j) out[pos++] = one_normal();
        std::uniform_real_distribution<double> ud(lo, 1.0);
        out[pos++] = ud(engine);
    }
}

}  // namespace mcmcb200
