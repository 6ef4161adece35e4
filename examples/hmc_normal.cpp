// Sampling the posterior of a Normal(mu, sigma) likelihood with HMC — the B200 twin of the reference's
// examples/eigen/hmc_normal.cpp (same model, same settings, same call shape).
//
//   g++ -std=c++14 -O2 -I include examples/hmc_normal.cpp -o hmc_normal -L mcmc_b200 -lmcmc_b200 -Wl,-rpath,$PWD/mcmc_b200
//
// The reference passes a std::function computing log p(x | mu, sigma) from the data vector; here the registered
// __device__ functor "normal_model" evaluates the same density from the sufficient statistics (n, mean, sum of
// squared deviations), which is all the likelihood depends on.
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "mcmc_b200.hpp"

int main()
{
    const int n_data = 1000;
    const double mu = 2.0, sigma = 2.0;
    std::mt19937 gen(12345);
    std::normal_distribution<> dist;
    std::vector<double> x(n_data);
    double xbar = 0.0;
    for (double& v : x) { v = mu + sigma * dist(gen); xbar += v; }
    xbar /= n_data;
    double M2 = 0.0;
    for (double v : x) M2 += (v - xbar) * (v - xbar);
    const double stats[3] = {double(n_data), xbar, M2};
    mcmc::kernel_data dta = {stats, 3};

    mcmc::ColVec_t initial_val(2);
    initial_val(0) = mu + 1;     // mu
    initial_val(1) = sigma + 1;  // sigma

    mcmc::algo_settings_t settings;
    settings.rng_seed_value = 1;
    settings.hmc_settings.step_size = 0.08;
    settings.hmc_settings.n_leap_steps = 5;
    settings.hmc_settings.n_burnin_draws = 2000;
    settings.hmc_settings.n_keep_draws = 2000;

    mcmc::Mat_t draws_out;
    if (!mcmc::hmc(initial_val, mcmc::device_kernel("normal_model"), draws_out, &dta, settings)) {
        std::fprintf(stderr, "mcmc::hmc failed: %s\n", mcmc::last_error());
        return 1;
    }
    double m0 = 0, m1 = 0;
    for (size_t t = 0; t < draws_out.rows(); ++t) { m0 += draws_out(t, 0); m1 += draws_out(t, 1); }
    std::printf("hmc mean: %g %g (data mean %g)\n", m0 / draws_out.rows(), m1 / draws_out.rows(), xbar);
    std::printf("acceptance rate: %g\n", double(settings.hmc_settings.n_accept_draws) / settings.hmc_settings.n_keep_draws);
    return 0;
}
