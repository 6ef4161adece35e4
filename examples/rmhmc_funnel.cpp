// RM-HMC on Neal's funnel with the SoftAbs metric, many chains in one call — BASELINE config 5 written like reference
// user code (the reference's twin is examples/eigen/rmhmc_normal.cpp: same call shape, a registered kernel + metric in
// place of the two std::function callbacks).
//
//   g++ -std=c++14 -O2 -I include examples/rmhmc_funnel.cpp -o rmhmc_funnel -L mcmc_b200 -lmcmc_b200 -Wl,-rpath,$PWD/mcmc_b200
//
// funnel: x[0] = v ~ N(0, 3^2), x[i] | v ~ N(0, e^v).  b200.rmhmc_metric_id picks the metric registered with the kernel that
// plays the reference's tensor_fn: 1 = Fisher-type diagonal metric, 2 = SoftAbs of the Hessian (alpha = 1e6).
#include <cmath>
#include <cstdio>

#include "mcmc_b200.hpp"

int main(int argc, char** argv)
{
    const size_t d = 64, n_chains = argc > 1 ? size_t(std::atoi(argv[1])) : 256;
    mcmc::Mat_t initial_vals(d, n_chains);   // one COLUMN per chain
    for (size_t c = 0; c < n_chains; ++c)
        for (size_t j = 0; j < d; ++j) initial_vals(j, c) = (j == 0) ? 0.2 * std::sin(0.7 * c) : 0.6 * std::sin(0.37 * c + 0.11 * j);

    mcmc::algo_settings_t settings;
    settings.rng_seed_value = 1;
    settings.rmhmc_settings.step_size = 0.01;
    settings.rmhmc_settings.n_leap_steps = 5;
    settings.rmhmc_settings.n_fp_steps = 5;
    settings.rmhmc_settings.n_burnin_draws = 10;
    settings.rmhmc_settings.n_keep_draws = 20;
    settings.b200.rmhmc_metric_id = 2;                 // SoftAbs
    settings.b200.rng_mode = MCMCB200_RNG_PHILOX;      // in-kernel generator (the default replays the reference's mt19937_64)
    // settings.b200.devices = {0, 1, 2, 3};           // shard the chains over several GPUs from this one call

    mcmc::Cube_t draws_out;   // one n_keep x d matrix per chain
    const mcmc::registered_kernel funnel = mcmc::device_kernel("funnel");
    if (!mcmc::rmhmc(initial_vals, funnel, funnel, draws_out, nullptr, nullptr, settings)) {
        std::fprintf(stderr, "mcmc::rmhmc failed: %s\n", mcmc::last_error());
        return 1;
    }
    double acc = 0.0, v_mean = 0.0;
    for (size_t c = 0; c < n_chains; ++c) {
        acc += double(settings.b200.n_accept_per_chain[c]);
        for (size_t t = 0; t < draws_out.mat(c).rows(); ++t) v_mean += draws_out.mat(c)(t, 0);
    }
    std::printf("rmhmc funnel d=%zu, %zu chains: mean v %g, acceptance rate %g\n", d, n_chains,
                v_mean / (n_chains * settings.rmhmc_settings.n_keep_draws), acc / (n_chains * settings.rmhmc_settings.n_keep_draws));
    return 0;
}
