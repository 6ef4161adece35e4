// Drives the C++ drop-in header exactly like reference user code and prints the draws as C99 hex floats so the
// Python test can compare them bit-for-bit with tests/golden/reference_golden.json (cases G2, G3, G4, C2 chain 5).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "mcmc_b200.hpp"

static void dump(const char* name, const mcmc::Mat_t& m, size_t n_accept)
{
    std::printf("%s %zu %zu %zu", name, (size_t)m.rows(), (size_t)m.cols(), n_accept);
    for (size_t t = 0; t < m.rows(); ++t)
        for (size_t j = 0; j < m.cols(); ++j) std::printf(" %a", m(t, j));
    std::printf("\n");
}

int main()
{
    mcmc::Mat_t draws;
    {   // G2: HMC d=3, seed 1
        mcmc::ColVec_t x0(3); x0(0) = 1; x0(1) = -1; x0(2) = 0.5;
        mcmc::algo_settings_t s; s.rng_seed_value = 1; s.b200.arith = MCMCB200_ARITH_STRICT;
        s.hmc_settings.n_burnin_draws = 0; s.hmc_settings.n_keep_draws = 5; s.hmc_settings.n_leap_steps = 10; s.hmc_settings.step_size = 0.1;
        if (!mcmc::hmc(x0, mcmc::device_kernel("iso_gauss"), draws, nullptr, s)) { std::fprintf(stderr, "%s\n", mcmc::last_error()); return 1; }
        dump("G2_hmc_d3", draws, s.hmc_settings.n_accept_draws);
    }
    {   // G3: MALA d=3, seed 1
        mcmc::ColVec_t x0(3); x0(0) = 1; x0(1) = -1; x0(2) = 0.5;
        mcmc::algo_settings_t s; s.rng_seed_value = 1; s.b200.arith = MCMCB200_ARITH_STRICT;
        s.mala_settings.n_burnin_draws = 0; s.mala_settings.n_keep_draws = 5; s.mala_settings.step_size = 0.5;
        if (!mcmc::mala(x0, mcmc::device_kernel("iso_gauss"), draws, nullptr, s)) { std::fprintf(stderr, "%s\n", mcmc::last_error()); return 1; }
        dump("G3_mala_d3", draws, s.mala_settings.n_accept_draws);
    }
    {   // golden rwmh_d3: RWMH d=3, seed 1, par_scale 0.5 (value-only kernel slot, include/mcmc/rwmh.hpp:43-72)
        mcmc::ColVec_t x0(3); x0(0) = 1; x0(1) = -1; x0(2) = 0.5;
        mcmc::algo_settings_t s; s.rng_seed_value = 1; s.b200.arith = MCMCB200_ARITH_STRICT;
        s.rwmh_settings.n_burnin_draws = 0; s.rwmh_settings.n_keep_draws = 8; s.rwmh_settings.par_scale = 0.5;
        if (!mcmc::rwmh(x0, mcmc::device_kernel("iso_gauss"), draws, nullptr, s)) { std::fprintf(stderr, "%s\n", mcmc::last_error()); return 1; }
        dump("rwmh_d3", draws, s.rwmh_settings.n_accept_draws);
    }
    {   // G4: RM-HMC on the Normal model, seed 1
        // data x_k = 2 + 2 sin k, k < 100: {n, mean, sum of squared deviations} exactly as stored in the golden fixture
        const double stats[3] = {100.0, 0x1.00f8824d3cb51p+1, 0x1.901597d089536p+7};
        mcmc::kernel_data dta = {stats, 3};
        mcmc::ColVec_t x0(2); x0(0) = 3; x0(1) = 3;
        mcmc::algo_settings_t s; s.rng_seed_value = 1; s.b200.arith = MCMCB200_ARITH_STRICT;
        s.rmhmc_settings.n_burnin_draws = 0; s.rmhmc_settings.n_keep_draws = 5; s.rmhmc_settings.n_leap_steps = 1; s.rmhmc_settings.step_size = 0.2;
        if (!mcmc::rmhmc(x0, mcmc::device_kernel("normal_model"), mcmc::device_kernel("normal_model"), draws, &dta, &dta, s)) {
            std::fprintf(stderr, "%s\n", mcmc::last_error()); return 1; }
        dump("G4_rmhmc_normal", draws, s.rmhmc_settings.n_accept_draws);
    }
    {   // many chains in one call: 3 chains of the C2 shape, chain 2 must equal a single-chain call with seed+2
        const size_t d = 128, C = 3;
        mcmc::Mat_t x0(d, C);
        for (size_t c = 0; c < C; ++c) for (size_t j = 0; j < d; ++j) x0(j, c) = std::sin(0.37 * c + 0.11 * j);
        mcmc::algo_settings_t s; s.rng_seed_value = 12345; s.b200.arith = MCMCB200_ARITH_STRICT;
        s.hmc_settings.n_burnin_draws = 10; s.hmc_settings.n_keep_draws = 20; s.hmc_settings.n_leap_steps = 10; s.hmc_settings.step_size = 0.1;
        mcmc::Cube_t cube;
        if (!mcmc::hmc(x0, mcmc::device_kernel("iso_gauss"), cube, nullptr, s)) { std::fprintf(stderr, "%s\n", mcmc::last_error()); return 1; }
        mcmc::ColVec_t x2(d); for (size_t j = 0; j < d; ++j) x2(j) = x0(j, 2);
        mcmc::algo_settings_t s2 = s; s2.rng_seed_value = 12347;
        if (!mcmc::hmc(x2, mcmc::device_kernel("iso_gauss"), draws, nullptr, s2)) { std::fprintf(stderr, "%s\n", mcmc::last_error()); return 1; }
        int same = (cube.n_mat() == C) && std::memcmp(cube.mat(2).data(), draws.data(), sizeof(double) * 20 * d) == 0;
        std::printf("multichain_consistent %d %zu\n", same, s.b200.n_accept_per_chain.size());
        // the same call sharded over a device list (one host thread per entry; the list {0, 0, 0} also exercises the
        // re-entrancy of the C ABI: three concurrent calls on one device) must return the same cube
        mcmc::algo_settings_t s3 = s;
        s3.b200.devices = {0, 0, 0};
        mcmc::Cube_t cube3;
        if (!mcmc::hmc(x0, mcmc::device_kernel("iso_gauss"), cube3, nullptr, s3)) { std::fprintf(stderr, "%s\n", mcmc::last_error()); return 1; }
        int same3 = (cube3.n_mat() == C);
        for (size_t c = 0; c < C && same3; ++c) same3 = std::memcmp(cube3.mat(c).data(), cube.mat(c).data(), sizeof(double) * 20 * d) == 0;
        same3 = same3 && s3.b200.n_accept_per_chain == s.b200.n_accept_per_chain;
        std::printf("sharded_consistent %d\n", same3);
    }
    {   // box constraints, written like reference user code (golden case hmc_box_d4: all four bound types)
        const double inf = INFINITY;
        const double w[4] = {1.0, 0.5, 2.0, 1.5};
        mcmc::kernel_data dta = {w, 4};
        mcmc::ColVec_t x0(4); x0(0) = 0.3; x0(1) = 0.7; x0(2) = 0.4; x0(3) = 0.2;
        mcmc::algo_settings_t s; s.rng_seed_value = 31; s.b200.arith = MCMCB200_ARITH_STRICT;
        s.vals_bound = true;
        s.lower_bounds = mcmc::ColVec_t(4); s.upper_bounds = mcmc::ColVec_t(4);
        s.lower_bounds(0) = -inf; s.lower_bounds(1) = 0.0; s.lower_bounds(2) = -inf; s.lower_bounds(3) = -1.0;
        s.upper_bounds(0) = inf; s.upper_bounds(1) = inf; s.upper_bounds(2) = 2.0; s.upper_bounds(3) = 1.5;
        s.hmc_settings.n_burnin_draws = 5; s.hmc_settings.n_keep_draws = 40; s.hmc_settings.n_leap_steps = 6; s.hmc_settings.step_size = 0.2;
        if (!mcmc::hmc(x0, mcmc::device_kernel("diag_gauss"), draws, &dta, s)) { std::fprintf(stderr, "%s\n", mcmc::last_error()); return 1; }
        dump("hmc_box_d4", draws, s.hmc_settings.n_accept_draws);
        // vals_bound without bound vectors of length n_vals is refused, not ignored
        mcmc::algo_settings_t bad; bad.vals_bound = true;
        const bool ok = mcmc::hmc(x0, mcmc::device_kernel("diag_gauss"), draws, &dta, bad);
        std::printf("bounds_refused %d\n", ok ? 0 : 1);
    }
    {   // G5: NUTS, 1-D N(0,1), seed 3 — on the reference's own stream, which is the wrapper's default (one launch per draw)
        mcmc::ColVec_t x0(1); x0(0) = 0.3;
        mcmc::algo_settings_t s; s.rng_seed_value = 3; s.b200.arith = MCMCB200_ARITH_STRICT;
        s.nuts_settings.n_burnin_draws = 0; s.nuts_settings.n_keep_draws = 3; s.nuts_settings.step_size = 0.05; s.nuts_settings.n_adapt_draws = 0;
        if (!mcmc::nuts(x0, mcmc::device_kernel("iso_gauss"), draws, nullptr, s)) { std::fprintf(stderr, "%s\n", mcmc::last_error()); return 1; }
        dump("G5_nuts_1d", draws, s.nuts_settings.n_accept_draws);
    }
    {   // golden de_iso_d3: mcmc::de, one population of 12 members, seed 11 (Cube_t of n_keep matrices n_pop x n_vals)
        mcmc::ColVec_t x0(3); x0(0) = 0.5; x0(1) = -0.5; x0(2) = 1.0;
        mcmc::algo_settings_t s; s.rng_seed_value = 11; s.b200.arith = MCMCB200_ARITH_STRICT;
        s.de_settings.n_pop = 12; s.de_settings.n_burnin_draws = 5; s.de_settings.n_keep_draws = 20;
        mcmc::Cube_t cube;
        if (!mcmc::de(x0, mcmc::device_kernel("iso_gauss"), cube, nullptr, s)) { std::fprintf(stderr, "%s\n", mcmc::last_error()); return 1; }
        std::printf("de_iso_d3 %zu %zu %zu", cube.n_mat() * (size_t)cube.mat(0).rows(), (size_t)cube.mat(0).cols(), s.de_settings.n_accept_draws);
        for (size_t g = 0; g < cube.n_mat(); ++g)
            for (size_t i = 0; i < cube.mat(g).rows(); ++i)
                for (size_t j = 0; j < cube.mat(g).cols(); ++j) std::printf(" %a", cube.mat(g)(i, j));
        std::printf("\n");
    }
    {   // the tensor_fn slot: a registered metric that does not belong to the kernel is refused, not ignored
        mcmc::ColVec_t x0(2); x0(0) = 3; x0(1) = 3;
        const double stats[3] = {100.0, 2.0, 400.0};
        mcmc::kernel_data dta = {stats, 3};
        mcmc::algo_settings_t s; s.rmhmc_settings.n_burnin_draws = 0; s.rmhmc_settings.n_keep_draws = 1;
        const bool ok = mcmc::rmhmc(x0, mcmc::device_kernel("normal_model"), mcmc::device_metric("funnel_softabs"), draws, &dta, nullptr, s);
        std::printf("foreign_metric_refused %d\n", ok ? 0 : 1);
    }
    return 0;
}
