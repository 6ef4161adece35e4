// HMC for 512 < n_dim <= 2048: four warps (one CTA) per chain, each warp owning a contiguous 32*EPL-element segment
// of the chain's vectors in registers.  Same algorithm and arithmetic as hmc.cu (src/hmc.cpp:155-205); only the
// reductions change: warp butterfly, then the four warp partials are added in warp order through shared memory.
// Available for separable targets (iso_gauss, diag_gauss) with M = I — the configuration of the dimension sweep in
// BASELINE config 5 (d in {32, 128, 512, 2048} with the HMC iso-Gaussian kernel).
#include "engine.h"
#include "rng.cuh"
#include "targets.cuh"
#include <math_constants.h>

namespace mcmcb200
{

template <class T, int EPL, bool STRICT, int RNGM>
__global__ void __launch_bounds__(128) hmc_wide_kernel(const __grid_constant__ HmcLaunch a)
{
    typedef Ar<STRICT> A;
    __shared__ double2 rng_tab[RNGM == RNG_PHILOX ? RNG_TAB_DOUBLE2 : 1];
    __shared__ double red[2][4];
    __shared__ double u_sh;
    if (RNGM == RNG_PHILOX) {
        build_rng_tables(rng_tab);
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long chain = blockIdx.x;
    const int d = a.d;
    constexpr int SEG = 32 * EPL;
    const int seg_off = warp * SEG;
    const int d_loc = max(0, min(d - seg_off, SEG));
    const double* tdata_seg = a.tdata + (T::per_element_data ? seg_off : 0);
    const WarpCtx w{lane, d_loc, nullptr};
    int parity = 0;
    // sum over the chain: strict-order butterfly per warp, then warps 0..3 in order
    auto chain_sum = [&](double v) -> double {
        v = warp_sum<STRICT>(v);
        if (lane == 0) red[parity][warp] = v;
        __syncthreads();
        const double s = A::add(A::add(A::add(red[parity][0], red[parity][1]), red[parity][2]), red[parity][3]);
        parity ^= 1;
        return s;
    };

    double x[EPL], p[EPL], g[EPL], xs[EPL];
    load_vec<EPL>(a.x0 + chain * d + seg_off, d_loc, lane, x);
    ChainRng<RNGM> rng;
    rng.init(a.rng, chain, a.chain_offset + chain);

    double U = -chain_sum(T::template eval<EPL, STRICT, true, false, false>(tdata_seg, w, x, g));
    int n_acc = 0;
    const int n_total = (int)(a.n_burnin + a.n_keep);
    const int n_burnin = (int)a.n_burnin;
    const double eps = a.eps, heps = 0.5 * eps;
    const int L = a.n_leap;
    double* out_row = a.draws + chain * a.n_keep * d + seg_off;
    double* out_lp = a.logp ? a.logp + chain * a.n_keep : nullptr;

    for (int t = 0; t < n_total; ++t) {
        rng.template normals<EPL, false>(a.rng, t, d_loc, lane, rng_tab, p, seg_off / 2, seg_off, d);
        double K0 = A::mul(0.5, lane_dot<EPL, STRICT>(p, p));
#pragma unroll
        for (int k = 0; k < EPL; ++k) xs[k] = x[k];
        double U1 = 0.0;
        if (L > 0) {
            T::template eval<EPL, STRICT, false, true>(tdata_seg, w, x, g);
#pragma unroll
            for (int k = 0; k < EPL; ++k) p[k] = STRICT ? A::add(p[k], A::mul(A::mul(eps, g[k]), 0.5)) : fma(heps, g[k], p[k]);
            for (int s = 0; s < L; ++s) {
#pragma unroll
                for (int k = 0; k < EPL; ++k) x[k] = A::mad(eps, p[k], x[k]);
                if (s + 1 < L) {
                    T::template eval<EPL, STRICT, false, true>(tdata_seg, w, x, g);
#pragma unroll
                    for (int k = 0; k < EPL; ++k) {
                        if (STRICT) {
                            const double hk = A::mul(A::mul(eps, g[k]), 0.5);
                            p[k] = A::add(A::add(p[k], hk), hk);
                        } else {
                            p[k] = fma(eps, g[k], p[k]);
                        }
                    }
                }
            }
            U1 = -T::template eval<EPL, STRICT, true, true, false>(tdata_seg, w, x, g);
#pragma unroll
            for (int k = 0; k < EPL; ++k) p[k] = STRICT ? A::add(p[k], A::mul(A::mul(eps, g[k]), 0.5)) : fma(heps, g[k], p[k]);
        }
        double K1 = A::mul(0.5, lane_dot<EPL, STRICT>(p, p));

        // uniform #0 of the draw: tape entry after the d normals, or (Philox) the spare bits of blocks 0 and 1, which
        // only warp 0 holds -> broadcast through shared memory, made visible by the barrier inside chain_sum
        double u = 0.0;
        if (RNGM == RNG_TAPE) {
            u = rng.uniform(a.rng, t, 0);
        } else if (warp == 0) {
            u = rng.uniform(a.rng, t, 0);
            if (lane == 0) u_sh = u;
        }
        bool acc;
        if (STRICT) {
            if (L > 0) U1 = chain_sum(U1); else U1 = U;
            K0 = chain_sum(K0);
            K1 = chain_sum(K1);
            if (RNGM == RNG_PHILOX) u = u_sh;
            if (!isfinite(U1)) U1 = CUDART_INF;
            const double comp = fmin(0.01, A::add(-A::add(U1, K1), A::add(U, K0)));
            acc = u < exp(comp);
            if (acc) U = U1;
        } else {
            // FAST: U is the reduced energy of the current state; one chain-wide sum of (K0 - U1 - K1) decides,
            // and the accepted state's energy is reduced only when needed (every warp runs the same two barriers)
            const double part = chain_sum(K0 - (L > 0 ? U1 : 0.0) - K1);
            if (RNGM == RNG_PHILOX) u = u_sh;
            const double dH = (L > 0) ? (U + part) : part;   // with L == 0 the position did not move: U1 = U
            acc = u < 1.0 + dH;   // same rule as hmc.cu (dH = +inf accepts, NaN / -inf reject)
            if (!acc) acc = (fabs(dH) <= 1.7976931348623157e308) && (u < exp(dH));
            const double U1r = chain_sum(L > 0 ? U1 : 0.0);
            if (acc && L > 0) U = U1r;
        }
        if (!acc) {
#pragma unroll
            for (int k = 0; k < EPL; ++k) x[k] = xs[k];
        }
        if (t >= n_burnin) {
            store_vec<EPL>(out_row, d_loc, lane, x);
            out_row += d;
            if (out_lp) {
                if (threadIdx.x == 0) *out_lp = -U;
                ++out_lp;
            }
            n_acc += acc ? 1 : 0;
        }
    }
    if (threadIdx.x == 0 && a.n_accept) a.n_accept[chain] = n_acc;
}

template <class T, int EPL> static int launch_wide_epl(const HmcLaunch& a)
{
    const unsigned blocks = (unsigned)a.n_chains;
    if (a.rng.mode == RNG_PHILOX) {
        if (a.strict) hmc_wide_kernel<T, EPL, true, RNG_PHILOX><<<blocks, 128, 0, a.stream>>>(a);
        else hmc_wide_kernel<T, EPL, false, RNG_PHILOX><<<blocks, 128, 0, a.stream>>>(a);
    } else {
        if (a.strict) hmc_wide_kernel<T, EPL, true, RNG_TAPE><<<blocks, 128, 0, a.stream>>>(a);
        else hmc_wide_kernel<T, EPL, false, RNG_TAPE><<<blocks, 128, 0, a.stream>>>(a);
    }
    MCMCB200_CUDA_TRY(cudaGetLastError());
    return MCMCB200_OK;
}

template <class T> static int launch_wide_target(const HmcLaunch& a)
{
    if (a.d <= 1024) return launch_wide_epl<T, 8>(a);
    return launch_wide_epl<T, 16>(a);
}

bool hmc_wide_supported(int target_id, int d, bool has_precond)
{
    return !has_precond && d > 32 * MAX_EPL && d <= 2048 && (target_id == MCMCB200_TARGET_ISO_GAUSS || target_id == MCMCB200_TARGET_DIAG_GAUSS);
}

int launch_hmc_wide(const HmcLaunch& a)
{
    if (a.broadcast_x0) { set_error("hmc (n_dim > 512): broadcast_initial is not supported"); return MCMCB200_ERR_UNSUPPORTED; }
    switch (a.target_id) {
    case MCMCB200_TARGET_ISO_GAUSS: return launch_wide_target<IsoGauss>(a);
#ifndef MCMCB200_FAST_BUILD
    case MCMCB200_TARGET_DIAG_GAUSS: return launch_wide_target<DiagGauss>(a);
#endif
    default: set_error("hmc (n_dim > 512): target %d is not separable", a.target_id); return MCMCB200_ERR_UNSUPPORTED;
    }
}

}  // namespace mcmcb200
