// HMC for FEW chains: two warps per chain.
//
// The production kernel (hmc_pipe_kernel, hmc.cu) gives every chain one warp and fills the fp64 issue slots by keeping ~7
// warps per SM sub-partition resident.  A strong-scaling shard — BASELINE's 4096 chains over 8 GPUs = 512 chains per GPU —
// leaves at most ONE warp per sub-partition, and a lone warp runs at its own dependency latency: ncu / CUDA events show
// 0.48 ms for 1100 draws whether the GPU holds 512 chains or 148 (843 cycles per draw against 524 at full occupancy), which
// caps the kernel-only strong-scaling efficiency at 0.54 for 8 GPUs.  Splitting a chain's ELEMENTS over two warps
// does not help (the per-draw cost is dominated by fixed-length dependency chains: Philox rounds, Box-Muller, 10 dependent
// leapfrog steps, the butterfly).  What does help is splitting the draw's two independent HALVES:
//   producer warp: Philox4x32-10 + Box-Muller for draw t + 1 (depends on nothing but the counter)
//   consumer warp: the trajectory, energy butterfly and accept test of draw t
// coupled through a double-buffered slot in shared memory and ONE 64-thread named barrier per draw.  Same variates, same
// arithmetic in the same order as hmc_pipe_kernel: the draws are bit-identical (tests/test_gpu_hmc.py).
// RESULT (B200, profiles/r2_hmc_duo_probe.txt): at n_leap = 10 it does NOT beat the production kernel — 0.474 vs 0.482 ms at 148
// chains, slower from 512 chains on — because that kernel's software pipelining already hides the variates and what a lone
// warp runs at is the trajectory itself: 10 x 2 dependent DFMA, a 5-stage butterfly and the accept test = ~830 cycles per
// draw, which no second warp shortens.  Strong scaling of C2 at 512 chains per GPU is therefore latency-bound by the
// algorithm's own dependency chain (kernel-only efficiency 0.54 at 8 GPUs).  The kernel is kept for what it does speed up:
// few chains with any other trajectory length (variates generated up front there): 1.2 - 1.5 x.
// Replaces the chain loop over /root/reference/src/hmc.cpp:155-205 for few-chain calls; FAST arithmetic, Philox, identity
// mass, separable targets, full tiles (n_dim = 64, 128, 256).  MCMCB200_HMC_DUO=0/1 forces the choice.
#include "engine.h"
#include "rng.cuh"
#include "targets.cuh"
#include "hmc_duo.h"
#include <cstdlib>

namespace mcmcb200
{

namespace
{

constexpr int DUO_CHAINS = 2;   // chains per CTA: 4 warps = one per SM sub-partition

template <class T, int EPL> __global__ void __launch_bounds__(DUO_CHAINS * 64) hmc_duo_kernel(const __grid_constant__ HmcLaunch a)
{
    constexpr int d = 32 * EPL;
    constexpr int SLOT = d + 34;   // z[d], the 32 lane partials of sum(-2 ln u1) (= sum z^2 up to rounding), the draw's uniform
    extern __shared__ __align__(16) double smem[];
    __shared__ double2 rng_tab[RNG_TAB_DOUBLE2];
    build_rng_tables(rng_tab);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ci = warp >> 1, producer = warp & 1;
    const long long chain = (long long)blockIdx.x * DUO_CHAINS + ci;
    if (chain >= a.n_chains) return;   // both warps of the pair leave together; only pair-wide barriers below
    double* const home = smem + (size_t)ci * (d + 2 * SLOT);
    double* const slot0 = home + d;
    const int n_burnin = (int)a.n_burnin, n_total = (int)(a.n_burnin + a.n_keep);
    auto pair_barrier = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + ci) : "memory"); };

    if (producer) {
        const unsigned gchain = static_cast<unsigned>(a.chain_offset + chain);
        for (int t = 0; t < n_total; ++t) {
            double z[EPL], ksum;
            BmPipe<EPL / 2> bp;
            bp.begin(lane, t, gchain);
            bp.template slice<0, 1>(a.rng, rng_tab, z, ksum);
            const double u = bp.uniform0();
            double* const s = slot0 + (t & 1) * SLOT;   // free: the consumer copied draw t - 2 out of it before barrier t - 1
#pragma unroll
            for (int m = 0; m < EPL / 2; ++m) *reinterpret_cast<double2*>(s + m * 64 + 2 * lane) = make_double2(z[2 * m], z[2 * m + 1]);
            s[d + lane] = ksum;
            if (lane == 0) s[d + 32] = u;
            pair_barrier();   // barrier t: draw t's variates are in place
        }
        return;
    }

    // ---- consumer: hmc_pipe_kernel's transition with the variates read from the slot ----
    const WarpCtx w{lane, d, home};
    double x[EPL], g[EPL], p[EPL];
    load_vec_full<EPL>(a.x0 + (a.broadcast_x0 ? 0 : chain * d), lane, x);
    double U = -T::template eval<EPL, false, true, false, false>(a.tdata, w, x, g);   // this lane's partial sum of -log pi(x)
    int n_acc = 0;
    const double eps = a.eps, heps = 0.5 * eps;
    const int L = a.n_leap;
    double* const out_base = a.draws + chain * a.n_keep * d + 2 * lane;
    double* const out_lp = a.logp ? a.logp + chain * a.n_keep : nullptr;
    for (int t = 0; t < n_total; ++t) {
        pair_barrier();   // barrier t
        const double* const s = slot0 + (t & 1) * SLOT;
#pragma unroll
        for (int m = 0; m < EPL / 2; ++m) {
            const double2 v = *reinterpret_cast<const double2*>(s + m * 64 + 2 * lane);
            p[2 * m] = v.x;
            p[2 * m + 1] = v.y;
        }
        const double ksum = s[d + lane], u = s[d + 32];
        double dH = fma(0.5, ksum, U);   // U0 + K0 (lane partial)
        bool acc = true;
        if (L > 0) {
#pragma unroll
            for (int m = 0; m < EPL / 2; ++m) *reinterpret_cast<double2*>(home + m * 64 + 2 * lane) = make_double2(x[2 * m], x[2 * m + 1]);
            T::template eval<EPL, false, false, true, false>(a.tdata, w, x, g);
#pragma unroll
            for (int k = 0; k < EPL; ++k) p[k] = fma(heps, g[k], p[k]);
            for (int st = 0; st < L; ++st) {
#pragma unroll
                for (int k = 0; k < EPL; ++k) x[k] = fma(eps, p[k], x[k]);
                if (st + 1 < L) {
                    T::template eval<EPL, false, false, true, false>(a.tdata, w, x, g);
#pragma unroll
                    for (int k = 0; k < EPL; ++k) p[k] = fma(eps, g[k], p[k]);
                }
            }
            const double U1 = -T::template eval<EPL, false, true, true, false>(a.tdata, w, x, g);
#pragma unroll
            for (int k = 0; k < EPL; ++k) p[k] = fma(heps, g[k], p[k]);
            // dH = (U0 + K0) - (U1 + K1), one butterfly; u < exp(min(0.01, dH)) holds whenever u < 1 + dH (src/hmc.cpp:187)
            dH = fma(-0.5, lane_dot<EPL, false>(p, p), dH - U1);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) dH += __shfl_xor_sync(FULL, dH, off);
            acc = u < 1.0 + dH;
            if (!acc) acc = (fabs(dH) <= 1.7976931348623157e308) && (u < exp(dH));
            if (!acc) {
#pragma unroll
                for (int m = 0; m < EPL / 2; ++m) {
                    const double2 v = *reinterpret_cast<const double2*>(home + m * 64 + 2 * lane);
                    x[2 * m] = v.x;
                    x[2 * m + 1] = v.y;
                }
            }
            U = acc ? U1 : U;
        }
        if (t >= n_burnin) {
            double* row = out_base + (size_t)(t - n_burnin) * d;
#pragma unroll
            for (int m = 0; m < EPL / 2; ++m) *reinterpret_cast<double2*>(row + m * 64) = make_double2(x[2 * m], x[2 * m + 1]);
            if (out_lp) {
                const double Ur = warp_sum<false>(U);
                if (lane == 0) out_lp[t - n_burnin] = -Ur;
            }
            n_acc += acc ? 1 : 0;
        }
    }
    if (lane == 0 && a.n_accept) a.n_accept[chain] = n_acc;
}

template <class T, int EPL> int launch_duo(const HmcLaunch& a)
{
    const long long blocks = (a.n_chains + DUO_CHAINS - 1) / DUO_CHAINS;
    const size_t smem = (size_t)DUO_CHAINS * (a.d + 2 * (a.d + 34)) * sizeof(double);
    hmc_duo_kernel<T, EPL><<<(unsigned)blocks, DUO_CHAINS * 64, smem, a.stream>>>(a);
    MCMCB200_CUDA_TRY(cudaGetLastError());
    return MCMCB200_OK;
}

template <class T> int launch_duo_target(const HmcLaunch& a)
{
    switch (a.d) {
    case 64: return launch_duo<T, 2>(a);
    case 128: return launch_duo<T, 4>(a);
    default: return launch_duo<T, 8>(a);
    }
}

}  // namespace

bool hmc_duo_supported(const HmcLaunch& a)
{
    const bool target_ok = a.target_id == MCMCB200_TARGET_ISO_GAUSS || a.target_id == MCMCB200_TARGET_DIAG_GAUSS;
    if (!target_ok || a.strict || a.rng.mode != RNG_PHILOX || a.S_cm != nullptr || a.lb != nullptr) return false;
    if (!(a.d == 64 || a.d == 128 || a.d == 256)) return false;
    if ((reinterpret_cast<uintptr_t>(a.x0) | reinterpret_cast<uintptr_t>(a.draws)) & 15) return false;
    if (const char* e = std::getenv("MCMCB200_HMC_DUO")) return e[0] != '0';
    // Measured on B200 (profiles/r2_hmc_duo_probe.txt): with n_leap = 10 the production kernel already hides the variates under the
    // trajectory (software pipelining) and a lone warp's time IS the trajectory's dependency chain — the two-warp kernel ties at
    // <= 296 chains and loses above; for every other trajectory length (variates generated up front) it is 1.2 - 1.5 x faster up
    // to ~2000 chains.
    return a.n_leap != 10 && a.n_chains <= 2048;
}

int launch_hmc_duo(const HmcLaunch& a)
{
    if (a.target_id == MCMCB200_TARGET_ISO_GAUSS) return launch_duo_target<IsoGauss>(a);
    return launch_duo_target<DiagGauss>(a);
}

}  // namespace mcmcb200
