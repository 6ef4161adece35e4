// The reference's fp32 build (MCMC_FPN_TYPE float, include/misc/mcmc_options.hpp:80-99) through the drop-in header.
// Compiled twice by tests/test_cpp_dropin.py — with -DMCMC_FPN_TYPE=float and without — from this one source, written
// against fp_t like reference user code; prints the draws as C99 hex so the test can compare the two builds: the fp32
// build must return the fp64 build's draws narrowed to float (inputs here are exactly representable in fp32).
#include <cmath>
#include <cstdio>
#include <vector>

#include "mcmc_b200.hpp"

static void dump(const char* name, const mcmc::Mat_t& m, size_t n_accept)
{
    std::printf("%s %zu %zu %zu", name, (size_t)m.rows(), (size_t)m.cols(), n_accept);
    for (size_t t = 0; t < m.rows(); ++t)
        for (size_t j = 0; j < m.cols(); ++j) std::printf(" %a", (double)m(t, j));
    std::printf("\n");
}

int main()
{
    using mcmc::fp_t;
    std::printf("sizeof_fp_t %zu\n", sizeof(fp_t));
    {   // one chain, dense mass matrix and a data blob, Philox
        const size_t d = 6;
        std::vector<fp_t> w(d);
        for (size_t j = 0; j < d; ++j) w[j] = fp_t(0.5) + fp_t(0.25) * fp_t(j);
        mcmc::kernel_data dta = {w.data(), d};
        mcmc::ColVec_t x0(d);
        for (size_t j = 0; j < d; ++j) x0(j) = fp_t(0.125) * fp_t((int)j - 2);
        mcmc::algo_settings_t s;
        s.rng_seed_value = 7; s.b200.rng_mode = MCMCB200_RNG_PHILOX;
        s.hmc_settings.n_burnin_draws = 3; s.hmc_settings.n_keep_draws = 12; s.hmc_settings.n_leap_steps = 5; s.hmc_settings.step_size = fp_t(0.25);
        s.hmc_settings.precond_mat = mcmc::Mat_t(d, d);
        for (size_t i = 0; i < d; ++i)
            for (size_t j = 0; j < d; ++j) s.hmc_settings.precond_mat(i, j) = (i == j) ? fp_t(1.5) : fp_t(0.125);
        mcmc::Mat_t draws;
        if (!mcmc::hmc(x0, mcmc::device_kernel("diag_gauss"), draws, &dta, s)) { std::fprintf(stderr, "%s\n", mcmc::last_error()); return 1; }
        dump("hmc_precond_d6", draws, s.hmc_settings.n_accept_draws);
    }
    {   // many chains, box constraints, NUTS
        const size_t d = 4, C = 5;
        mcmc::Mat_t x0(d, C);
        for (size_t c = 0; c < C; ++c)
            for (size_t j = 0; j < d; ++j) x0(j, c) = fp_t(0.25) + fp_t(0.0625) * fp_t(c + j);
        mcmc::algo_settings_t s;
        s.rng_seed_value = 11; s.b200.rng_mode = MCMCB200_RNG_PHILOX;
        s.vals_bound = true;
        s.lower_bounds = mcmc::ColVec_t(d); s.upper_bounds = mcmc::ColVec_t(d);
        for (size_t j = 0; j < d; ++j) { s.lower_bounds(j) = fp_t(-0.5); s.upper_bounds(j) = fp_t(2.0); }
        s.nuts_settings.n_burnin_draws = 5; s.nuts_settings.n_keep_draws = 6; s.nuts_settings.n_adapt_draws = 5; s.nuts_settings.step_size = fp_t(0.5);
        // (the defaults 0.55 / 0.05 are not representable in fp32: an fp32 build adapts with 0.55f / 0.05f, as the reference's does)
        s.nuts_settings.target_accept_rate = fp_t(0.5625); s.nuts_settings.gamma_val = fp_t(0.0625);
        mcmc::Cube_t cube;
        if (!mcmc::nuts(x0, mcmc::device_kernel("iso_gauss"), cube, nullptr, s)) { std::fprintf(stderr, "%s\n", mcmc::last_error()); return 1; }
        for (size_t c = 0; c < C; ++c) {
            char nm[32];
            std::snprintf(nm, sizeof(nm), "nuts_box_chain%zu", c);
            dump(nm, cube.mat(c), s.b200.n_accept_per_chain.size() == C ? (size_t)s.b200.n_accept_per_chain[c] : 0);
        }
    }
    return 0;
}
