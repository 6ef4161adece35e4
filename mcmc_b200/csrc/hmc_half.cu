// HMC for small n_dim: TWO chains per warp.
//
// The warp-per-chain layout (warp.cuh: element j on lane (j % 64) / 2) leaves lanes 16..31 idle when n_dim <= 32, and the cost
// of a draw is per WARP (Philox rounds, Box-Muller, the leapfrog's dependent FMAs, the butterfly): BASELINE config 5's sweep
// point n_dim = 32 ran at the same 0.57 ms as n_dim = 128 — 0.17 of the HBM contract roofline.  Here the upper half-warp carries
// a second chain: same per-lane arithmetic as hmc_kernel<..., FAST, Philox> (hmc.cu; /root/reference/src/hmc.cpp:155-205), the
// reductions run over 16 lanes (the xor-16 stage of the warp butterfly only ever added the idle lanes' zeros), Philox
// counters, spare-bit uniform and accept decision are per half.  Draws are bit-identical to hmc_kernel's (tests/test_gpu_hmc.py).
// FAST arithmetic, Philox, identity mass, no box constraints, targets iso_gauss / diag_gauss; MCMCB200_HMC_HALF=0/1 forces it.
#include "engine.h"
#include "rng.cuh"
#include "targets.cuh"
#include "hmc_half.h"
#include <cstdlib>

namespace mcmcb200
{

namespace
{

constexpr int HALF_WARPS = 4;   // warps per CTA = 8 chains

__device__ __forceinline__ double half_sum(double v)   // over the 16 lanes of a half-warp, same order as warp_sum's last four stages
{
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    return v;
}

template <class T> __global__ void __launch_bounds__(HALF_WARPS * 32) hmc_half_kernel(const __grid_constant__ HmcLaunch a)
{
    __shared__ double2 rng_tab[RNG_TAB_DOUBLE2];
    __shared__ double backup[HALF_WARPS * 2 * 32];
    build_rng_tables(rng_tab);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = lane >> 4, hl = lane & 15;
    const long long pair0 = ((long long)blockIdx.x * HALF_WARPS + warp) * 2;
    if (pair0 >= a.n_chains) return;   // whole warp: no block-level barriers below
    const bool active = pair0 + half < a.n_chains;           // an odd chain count leaves the last upper half without a chain:
    const long long chain = active ? pair0 + half : pair0;   // it shadows the lower half's chain and stores nothing
    const int d = a.d;
    double* const bscr = backup + (warp * 2 + half) * 32;
    const WarpCtx w{hl, d, nullptr};

    double x[2], p[2], g[2];
    load_vec<2>(a.x0 + (a.broadcast_x0 ? 0 : chain * d), d, hl, x);
    ChainRng<RNG_PHILOX> rng;
    rng.init(a.rng, chain, a.chain_offset + chain);
    double U = -T::template eval<2, false, true, false, false>(a.tdata, w, x, g);   // this lane's partial sum of -log pi(x)
    int n_acc = 0;
    const int n_total = (int)(a.n_burnin + a.n_keep), n_burnin = (int)a.n_burnin;
    const double eps = a.eps, heps = 0.5 * eps;
    const int L = a.n_leap;
    double* out_row = a.draws + chain * a.n_keep * d;
    double* out_lp = a.logp ? a.logp + chain * a.n_keep : nullptr;

    for (int t = 0; t < n_total; ++t) {
        rng.template normals<2, false>(a.rng, t, d, hl, rng_tab, p);   // pair index = lane within the half, chain id = the half's
        const double K0 = 0.5 * lane_dot<2, false>(p, p);
        double U1;
        if (L > 0) {
            bscr[2 * hl] = x[0];
            bscr[2 * hl + 1] = x[1];
            T::template eval<2, false, false, true, true>(a.tdata, w, x, g);
            p[0] = fma(heps, g[0], p[0]); p[1] = fma(heps, g[1], p[1]);
            for (int s = 0; s < L; ++s) {
                x[0] = fma(eps, p[0], x[0]); x[1] = fma(eps, p[1], x[1]);
                if (s + 1 < L) {
                    T::template eval<2, false, false, true, true>(a.tdata, w, x, g);
                    p[0] = fma(eps, g[0], p[0]); p[1] = fma(eps, g[1], p[1]);
                }
            }
            U1 = -T::template eval<2, false, true, true, false>(a.tdata, w, x, g);
            p[0] = fma(heps, g[0], p[0]); p[1] = fma(heps, g[1], p[1]);
        } else {
            U1 = U;
        }
        const double K1 = 0.5 * lane_dot<2, false>(p, p);
        // uniform #0 of the draw from the spare bits of the half's Philox blocks 0 and 1 (rng.cuh)
        const unsigned s0 = __shfl_sync(FULL, rng.spare, lane & 16), s1 = __shfl_sync(FULL, rng.spare, (lane & 16) + 1);
        const double sd = __hiloint2double(0x43300000 | (s0 >> 8), (s0 << 24) | s1) - 4503599627370496.0;
        const double u = fma(sd, 3.5527136788005009e-15, 1.7763568394002505e-15);
        const double dH = half_sum((U + K0) - (U1 + K1));
        bool acc = u < 1.0 + dH;
        if (!acc) acc = (fabs(dH) <= 1.7976931348623157e308) && (u < exp(dH));
        if (acc) {
            U = U1;
        } else if (L > 0) {
            x[0] = bscr[2 * hl];
            x[1] = bscr[2 * hl + 1];
        }
        if (t >= n_burnin) {
            if (active) store_vec<2>(out_row, d, hl, x);
            out_row += d;
            if (out_lp) {
                const double Ur = half_sum(U);
                if (active && hl == 0) *out_lp = -Ur;
                ++out_lp;
            }
            n_acc += acc ? 1 : 0;
        }
    }
    if (active && hl == 0 && a.n_accept) a.n_accept[chain] = n_acc;
}

}  // namespace

bool hmc_half_supported(const HmcLaunch& a)
{
    const bool target_ok = a.target_id == MCMCB200_TARGET_ISO_GAUSS || a.target_id == MCMCB200_TARGET_DIAG_GAUSS;
    if (!target_ok || a.strict || a.rng.mode != RNG_PHILOX || a.S_cm != nullptr || a.lb != nullptr || a.d > 32 || a.d < 1) return false;
    if (const char* e = std::getenv("MCMCB200_HMC_HALF")) return e[0] != '0';
    return a.n_chains >= 2;
}

int launch_hmc_half(const HmcLaunch& a)
{
    const long long warps = (a.n_chains + 1) / 2;
    const unsigned blocks = (unsigned)((warps + HALF_WARPS - 1) / HALF_WARPS);
    if (a.target_id == MCMCB200_TARGET_ISO_GAUSS) hmc_half_kernel<IsoGauss><<<blocks, HALF_WARPS * 32, 0, a.stream>>>(a);
    else hmc_half_kernel<DiagGauss><<<blocks, HALF_WARPS * 32, 0, a.stream>>>(a);
    MCMCB200_CUDA_TRY(cudaGetLastError());
    return MCMCB200_OK;
}

}  // namespace mcmcb200
