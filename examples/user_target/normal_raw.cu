// A USER-DEFINED target, compiled into the user's own shared library — libmcmc_b200.so is neither edited nor rebuilt.
//
// It is the log-likelihood callback of the reference's HMC example, written by the user for the device:
// /root/reference/examples/eigen/hmc_normal.cpp:44-76 (ll_dens): Normal(mu, sigma) likelihood of n RAW observations,
//   log pi(mu, sigma) = -n (1/2 log 2 pi + log sigma) - sum_i (x_i - mu)^2 / (2 sigma^2)
//   d/dmu = sum_i (x_i - mu) / sigma^2          d/dsigma = sum_i (x_i - mu)^2 / sigma^3 - n / sigma
// target_data = {n, x_0, ..., x_{n-1}} (what the example keeps in norm_data_t).  The 32 lanes of the chain's warp share
// the sum over observations.  Build + registration recipe: include/mcmc_b200_device.cuh; tests/test_user_target.py
// builds this file, loads it and checks the draws against the reference's golden vectors for the same likelihood.
#define MCMCB200_USER_TARGET_TAG normal_raw
#define MCMCB200_USER_MAX_EPL 2   // n_dim = 2: only the smallest tile is instantiated
#define MCMCB200_USER_NO_DE
#include "mcmc_b200_device.cuh"

struct NormalRaw {
    static constexpr bool needs_scratch = false;
    static constexpr bool dense_matrix = false;
    static constexpr bool separable = false;
    static constexpr bool per_element_data = false;
    template <int EPL, bool STRICT, bool WANT_VALUE, bool WANT_GRAD, bool REDUCE = true, class Ctx = mcmcb200::WarpCtx>
    static __device__ __forceinline__ double eval(const double* __restrict__ data, const Ctx& w, const double (&x)[EPL], double (&g)[EPL])
    {
        using namespace mcmcb200;
        typedef Ar<STRICT> A;
        const int n = (int)data[0];
        const double mu = __shfl_sync(FULL, x[0], 0), sigma = __shfl_sync(FULL, x[1], 0);   // elements 0 and 1 live on lane 0
        double s1 = 0.0, s2 = 0.0;
        for (int i = w.lane; i < n; i += 32) {
            const double r = A::sub(data[1 + i], mu);
            s1 = A::add(s1, r);
            s2 = A::mad(r, r, s2);
        }
        warp_sum2<STRICT>(s1, s2);
        const double s2g = A::mul(sigma, sigma);
        if (WANT_GRAD) {
#pragma unroll
            for (int k = 0; k < EPL; ++k) g[k] = 0.0;
            if (w.lane == 0) {
                g[0] = s1 / s2g;
                g[1] = A::sub(s2 / A::mul(s2g, sigma), (double)n / sigma);
            }
        }
        if (WANT_VALUE) {
            const double v = A::sub(A::mul(-(double)n, A::add(0.91893853320467274178, log(sigma))), s2 / A::mul(2.0, s2g));
            return (REDUCE || w.lane == 0) ? v : 0.0;
        }
        return 0.0;
    }
};

// Fisher information of the same model as the RM-HMC metric (examples/eigen/rmhmc_normal.cpp:82-111):
// G = diag(n / sigma^2, 2 n / sigma^2); dG/dmu = 0, dG/dsigma = -2 G / sigma.
struct NormalRawFisher {
    template <bool STRICT> static __device__ __forceinline__ void eval(const double* __restrict__ data, int d, int lane, const double* xs, double* G, double* dG)
    {
        typedef mcmcb200::Ar<STRICT> A;
        if (lane == 0) {
            const double n = data[0], sigma = xs[1];
            const double s2 = A::mul(sigma, sigma);
            const double g0 = n / s2, g3 = A::mul(2.0, n) / s2;
            G[0] = g0; G[1] = 0.0; G[2] = 0.0; G[3] = g3;
            if (dG) { dG[4] = A::mul(-2.0, g0) / sigma; dG[5] = A::mul(-2.0, 0.0) / sigma; dG[6] = A::mul(-2.0, 0.0) / sigma; dG[7] = A::mul(-2.0, g3) / sigma; }
        }
        (void)d;
    }
};

static int64_t normal_raw_data_len(int32_t n_dim) { return n_dim == 2 ? 1 : -1; }   // {n, x_0 ...}: at least the count

#define MCMCB200_USER_FUNCTOR NormalRaw
#define MCMCB200_USER_METRIC_TYPE NormalRawFisher
#define MCMCB200_USER_TARGET_NAME "normal_raw"
#define MCMCB200_USER_DATA_LEN normal_raw_data_len
#include "mcmc_b200_register.cuh"
