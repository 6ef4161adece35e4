/* The C ABI from plain C (what any FFI — cgo, JNI, ctypes, N-API — binds): 1024 HMC chains on a 128-dimensional standard
 * Gaussian, draws returned to host memory, then the same run summarised on the device without copying the draws.
 *
 *   gcc -std=c99 -O2 -I include examples/c_abi_hmc.c -o c_abi_hmc -L mcmc_b200 -lmcmc_b200 -Wl,-rpath,$PWD/mcmc_b200 -lm
 *
 * Each entry point replaces one reference interface (see include/mcmc_b200.h): mcmcb200_hmc_run <- bool mcmc::hmc(...)
 * include/mcmc/hmc.hpp:43-72. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mcmc_b200.h"
#include "mcmc_b200_summary.h"

int main(void)
{
    const int64_t n_chains = 1024, n_keep = 200;
    const int32_t d = 128;
    double* x0 = (double*)malloc(sizeof(double) * n_chains * d);
    double* draws = (double*)malloc(sizeof(double) * n_chains * n_keep * d);
    int64_t* n_accept = (int64_t*)malloc(sizeof(int64_t) * n_chains);
    for (int64_t c = 0; c < n_chains; ++c)
        for (int32_t j = 0; j < d; ++j) x0[c * d + j] = sin(0.37 * (double)c + 0.11 * j);

    mcmcb200_problem_t pr;
    memset(&pr, 0, sizeof(pr));
    pr.n_chains = n_chains;
    pr.n_dim = d;
    pr.target_id = mcmcb200_target_lookup("iso_gauss");   /* the registered __device__ functor replacing the callback */
    pr.initial_vals = x0;
    pr.initial_mem = MCMCB200_MEM_HOST;
    pr.device = -1;

    mcmcb200_rng_t rng;
    memset(&rng, 0, sizeof(rng));
    rng.mode = MCMCB200_RNG_PHILOX;   /* MCMCB200_RNG_MT19937_TAPE reproduces the reference's std::mt19937_64 stream */
    rng.seed = 12345;

    mcmcb200_hmc_settings_t st;
    mcmcb200_hmc_settings_default(&st);   /* the reference's defaults (mcmc_structs.hpp:66-78) */
    st.n_burnin_draws = 100;
    st.n_keep_draws = n_keep;
    st.n_leap_steps = 10;
    st.step_size = 0.1;

    mcmcb200_output_t out;
    memset(&out, 0, sizeof(out));
    out.draws_out = draws;
    out.draws_mem = MCMCB200_MEM_HOST;
    out.n_accept_draws = n_accept;

    if (mcmcb200_hmc_run(&pr, &rng, &st, &out) != MCMCB200_OK) {
        fprintf(stderr, "mcmcb200_hmc_run: %s\n", mcmcb200_last_error());   /* e.g. no CUDA device: there is no CPU fallback */
        return 1;
    }
    double acc = 0.0;
    for (int64_t c = 0; c < n_chains; ++c) acc += (double)n_accept[c];
    printf("hmc: %lld chains x %lld draws, kernel %.3f ms, acceptance rate %.3f\n", (long long)n_chains, (long long)n_keep,
           out.kernel_ms, acc / (double)(n_chains * n_keep));

    /* posterior summaries of the host copy (uploaded) — with draws_mem = MCMCB200_MEM_DEVICE nothing would be copied */
    double mean[128], var[128], rhat[128];
    mcmcb200_summary_t sm;
    memset(&sm, 0, sizeof(sm));
    sm.mean = mean; sm.var = var; sm.rhat = rhat;
    if (mcmcb200_summarize_draws(draws, MCMCB200_MEM_HOST, n_chains, n_keep, d, -1, NULL, &sm) != MCMCB200_OK) {
        fprintf(stderr, "mcmcb200_summarize_draws: %s\n", mcmcb200_last_error());
        return 1;
    }
    printf("element 0: mean %.4f var %.4f R-hat %.4f\n", mean[0], var[0], rhat[0]);
    free(x0); free(draws); free(n_accept);
    return 0;
}
