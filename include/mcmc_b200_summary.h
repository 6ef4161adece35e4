/*
 * mcmc_b200 — on-device summaries of draws_out ("summary reductions (means, accept rates, R-hat
 * across chains)" over the chain-major device buffer).
 *
 * The reference hands every kept draw back to the caller (Mat_t& draws_out, src/hmc.cpp:138,196-203) and users
 * compute posterior summaries on the host.  With thousands of chains resident in HBM that transfer is the end-to-end
 * bottleneck (4.19 GB over PCIe for the C2 job against a 2 ms kernel), so this entry point reduces
 * draws_out[n_chains][n_keep][n_dim] where it lies and returns only O(n_chains * n_dim) numbers.
 */
#ifndef MCMC_B200_SUMMARY_H
#define MCMC_B200_SUMMARY_H

#include "mcmc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcmcb200_summary {
    double* mean;       /* HOST [n_dim]: mean over all chains and kept draws                                        */
    double* var;        /* HOST [n_dim]: sample variance (ddof 1) over all n_chains*n_keep draws; may be NULL       */
    double* rhat;       /* HOST [n_dim]: Gelman-Rubin potential scale reduction sqrt(((n-1)/n W + B/n)/W) with
                           W = mean of the within-chain variances, B/n = variance of the chain means (both ddof 1);
                           NaN when n_chains < 2 or n_keep < 2; may be NULL                                           */
    double* chain_mean; /* optional HOST [n_chains][n_dim]: per-chain means                                          */
    double* chain_var;  /* optional HOST [n_chains][n_dim]: per-chain sample variances (ddof 1)                      */
    float kernel_ms;    /* filled by the library: device time of the two reduction kernels                           */
    int32_t reserved0;
} mcmcb200_summary_t;

/* draws: [n_chains][n_keep][n_dim] doubles in `draws_mem` memory (MCMCB200_MEM_DEVICE: reduced in place, nothing is
   copied; MCMCB200_MEM_HOST: uploaded first).  device = CUDA ordinal (-1: current), stream = cudaStream_t or NULL. */
int mcmcb200_summarize_draws(const double* draws, int32_t draws_mem, int64_t n_chains, int64_t n_keep, int32_t n_dim,
                             int32_t device, void* stream, mcmcb200_summary_t* out);

#ifdef __cplusplus
}
#endif

#endif /* MCMC_B200_SUMMARY_H */
