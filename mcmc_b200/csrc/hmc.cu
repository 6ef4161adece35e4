// Many-chain HMC: one persistent kernel, one warp per chain, all draws of a chain in one launch.
//
// Replaces internal::hmc_impl (/root/reference/src/hmc.cpp:30-227) run once per chain.  Per draw
// (SURVEY Appendix E):
//   z ~ N(0,I)                       src/hmc.cpp:156   (Philox in-kernel, or the reference's tape)
//   p = sqrtM z ; K0 = p.(M^-1 p)/2  :158-160
//   L leapfrog steps                 :164-176  p += (eps grad)/2 ; x += (eps M^-1) p ; p += (eps grad)/2
//   U1 = -log pi(x) (non-finite -> +inf), K1                         :178-184
//   accept iff u < exp(min(0.01, -(U1+K1) + (U0+K0)))                :188-191
//   kept draws written to draws_out, post-burn-in accepts counted    :196-203
// State (x, p, grad) never leaves registers during a trajectory; HBM traffic is the initial x
// (d*8 B per chain, once) and the draws_out row (d*8 B per kept draw): <= 2*d*8 B per transition,
// the contract figure of SURVEY §8(d).  The gradient at the end of leapfrog step k is the gradient
// at the start of step k+1, so it is evaluated L+1 times per draw instead of the reference's 2L
// calls + 1 value call; in STRICT mode the two half-kicks are still applied as two
// separately rounded updates, so element-wise results are bit-identical to the reference order.
#include "engine.h"
#include "rng.cuh"
#include "targets.cuh"
#include "box.cuh"
#include <math_constants.h>

namespace mcmcb200
{

// Half / full momentum kicks.  STRICT keeps the reference's rounding sequence p + (eps*g)/2 (a full kick is two
// separately rounded half kicks, src/hmc.cpp:167,175); FAST fuses them.
// With box constraints the force is J(v) (diagonal) times the raw gradient: p + ((eps*J)*grad)/2 (src/hmc.cpp:122).
template <int EPL, bool STRICT, bool BOX>
__device__ __forceinline__ void kick_half(double (&p)[EPL], const double (&g)[EPL], const double (&J)[EPL], double eps, double heps)
{
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
        if (BOX) p[k] = Ar<STRICT>::add(p[k], Ar<STRICT>::mul(Ar<STRICT>::mul(J[k], Ar<STRICT>::mul(eps, g[k])), 0.5));
        else p[k] = STRICT ? Ar<STRICT>::add(p[k], Ar<STRICT>::mul(Ar<STRICT>::mul(eps, g[k]), 0.5)) : fma(heps, g[k], p[k]);
    }
}
template <int EPL, bool STRICT, bool BOX>
__device__ __forceinline__ void kick_full(double (&p)[EPL], const double (&g)[EPL], const double (&J)[EPL], double eps)
{
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
        if (BOX) {
            const double hk = Ar<STRICT>::mul(Ar<STRICT>::mul(J[k], Ar<STRICT>::mul(eps, g[k])), 0.5);
            p[k] = Ar<STRICT>::add(Ar<STRICT>::add(p[k], hk), hk);
        } else if (STRICT) {
            const double hk = Ar<STRICT>::mul(Ar<STRICT>::mul(eps, g[k]), 0.5);
            p[k] = Ar<STRICT>::add(Ar<STRICT>::add(p[k], hk), hk);
        } else {
            p[k] = fma(eps, g[k], p[k]);
        }
    }
}

// resident CTAs per SM the register allocator should aim for (4096 chains = 1024 CTAs = 6.9 per SM at EPL <= 4)
constexpr int hmc_min_blocks(int epl) { return epl <= 4 ? 7 : (epl == 8 ? 4 : 2); }

// FT ("full tile"): n_dim == 32*EPL and 16-byte aligned rows, so no padding predicates anywhere.
// LS: number of leapfrog steps fixed at compile time (0 = runtime a.n_leap).  With LS > 0 the whole draw is straight-line
// code, so ptxas interleaves the next Box-Muller polynomial chains, the Philox rounds and the leapfrog DFMAs freely.
// BOX: box constraints (vals_bound), see box.cuh; available with M = I on the generic (runtime-L) kernels.
template <class T, int EPL, bool DENSE_M, bool STRICT, int RNGM, bool FT, int LS = 0, bool BOX = false>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, hmc_min_blocks(EPL)) hmc_kernel(const __grid_constant__ HmcLaunch a)
{
    extern __shared__ double smem[];
    __shared__ double2 log_tab[RNGM == RNG_PHILOX ? RNG_TAB_DOUBLE2 : 1];
    typedef Ar<STRICT> A;
    if (RNGM == RNG_PHILOX) {
        build_rng_tables(log_tab);
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long chain = (long long)blockIdx.x * WARPS_PER_BLOCK + warp;
    if (chain >= a.n_chains) return;  // whole warp exits together; no block-level barriers below
    const int d = a.d;
    const int dpad = (d + 1) & ~1;
    constexpr bool NEED_SCR = T::needs_scratch || DENSE_M;  // must match the launcher's shared-memory size
    double* bscr = smem + (size_t)warp * (NEED_SCR ? 3 : 1) * dpad;  // backup of the current state while a trajectory runs in place
    double* tscr = bscr + dpad;                      // target functor scratch
    double* mscr = tscr + dpad;                      // mass-matrix scratch
    const WarpCtx w{lane, d, tscr};

    double x[EPL], p[EPL], g[EPL], Jr[EPL];  // Jr is dead (no registers) unless BOX
    if (FT) load_vec_full<EPL>(a.x0 + (a.broadcast_x0 ? 0 : chain * d), lane, x);
    else load_vec<EPL>(a.x0 + (a.broadcast_x0 ? 0 : chain * d), d, lane, x);
    BoxLane<BOX ? EPL : 1> bx;
    if (BOX) {
        bx.load(a.lb, a.ub, d, lane);
#pragma unroll
        for (int k = 0; k < EPL; ++k) x[k] = bx.transform(BOX ? k : 0, x[k]);   // first_draw = transform(initial_vals), src/hmc.cpp:132-136
    }

    ChainRng<RNGM> rng;
    rng.init(a.rng, chain, a.chain_offset + chain);

    // U = -log pi(x): STRICT carries the reduced scalar (src/hmc.cpp:140); FAST carries this lane's partial sum,
    // so that one butterfly per draw reduces (U0 + K0) - (U1 + K1) directly.
    double U = -box_eval<T, EPL, STRICT, BOX, true, false, STRICT>(a.tdata, w, bx, x, g, Jr);
    int n_acc = 0;
    const int n_total = (int)(a.n_burnin + a.n_keep);
    const int n_burnin = (int)a.n_burnin;
    const double eps = a.eps;
    const double heps = 0.5 * eps;
    const int L = LS ? LS : a.n_leap;
    double* out_row = a.draws + chain * a.n_keep * d;
    double* out_lp = a.logp ? a.logp + chain * a.n_keep : nullptr;

    // Software pipelining (compile-time-L Philox kernels only): the variates of draw t+1 are generated in the same
    // straight-line block as the trajectory of draw t — they do not depend on the chain state, so the integer Philox
    // rounds and the Box-Muller chains fill the issue slots left by the dependent leapfrog DFMAs.
    constexpr bool PIPE = (LS > 0) && (RNGM == RNG_PHILOX) && !DENSE_M;
    double zn[PIPE ? EPL : 1];
    double un = 0.0;
    if (PIPE) {
        double ztmp[EPL];
        rng.template normals<EPL, FT>(a.rng, 0, d, lane, log_tab, ztmp);
        un = rng.uniform(a.rng, 0, 0);
#pragma unroll
        for (int k = 0; k < (PIPE ? EPL : 1); ++k) zn[k] = ztmp[k];
    }

    for (int t = 0; t < n_total; ++t) {
        // ---- momentum refresh: p = sqrtM z, K0 = p.(M^-1 p)/2 (lane partial in FAST) ----
        double u_pipe = un;
        if (PIPE) {
#pragma unroll
            for (int k = 0; k < EPL; ++k) p[k] = zn[k < (PIPE ? EPL : 1) ? k : 0];
            double ztmp[EPL];
            rng.template normals<EPL, FT>(a.rng, t + 1, d, lane, log_tab, ztmp);   // one spare draw past the end: harmless
            un = rng.uniform(a.rng, t + 1, 0);
#pragma unroll
            for (int k = 0; k < (PIPE ? EPL : 1); ++k) zn[k] = ztmp[k];
        } else {
            rng.template normals<EPL, FT>(a.rng, t, d, lane, log_tab, p);
        }
        double K0;
        if (DENSE_M) {
            double tmp[EPL];
            stage_vec<EPL>(mscr, d, lane, p);
            gemv_cm<EPL, STRICT>(a.S_cm, d, lane, mscr, 1.0, tmp);
#pragma unroll
            for (int k = 0; k < EPL; ++k) p[k] = tmp[k];
            stage_vec<EPL>(mscr, d, lane, p);
            gemv_cm<EPL, STRICT>(a.Minv_cm, d, lane, mscr, 1.0, tmp);
            K0 = A::mul(0.5, STRICT ? warp_dot<EPL, STRICT>(p, tmp) : lane_dot<EPL, STRICT>(p, tmp));
        } else {
            K0 = A::mul(0.5, STRICT ? warp_dot<EPL, STRICT>(p, p) : lane_dot<EPL, STRICT>(p, p));
        }

        // ---- trajectory in place on x; the current state is parked in shared memory (2 x STS.128 per lane at d=128)
        //      and only read back on a rejection ----
        double U1;
        if (L > 0) {
            if (FT) store_vec_full<EPL>(bscr, lane, x);
            else store_vec<EPL>(bscr, d, lane, x);
            box_eval<T, EPL, STRICT, BOX, false, true, true>(a.tdata, w, bx, x, g, Jr);
            kick_half<EPL, STRICT, BOX>(p, g, Jr, eps, heps);
            auto step = [&](int s) {
                if (DENSE_M) {
                    double tmp[EPL];
                    stage_vec<EPL>(mscr, d, lane, p);
                    gemv_cm<EPL, STRICT>(a.Minv_cm, d, lane, mscr, eps, tmp);  // (eps M^-1) p
#pragma unroll
                    for (int k = 0; k < EPL; ++k) x[k] = A::add(x[k], tmp[k]);
                } else {
#pragma unroll
                    for (int k = 0; k < EPL; ++k) x[k] = A::mad(eps, p[k], x[k]);
                }
                if (s + 1 < L) {
                    box_eval<T, EPL, STRICT, BOX, false, true, true>(a.tdata, w, bx, x, g, Jr);
                    kick_full<EPL, STRICT, BOX>(p, g, Jr, eps);  // end of step s and start of step s+1 share this gradient
                }
            };
            if (LS > 0) {
#pragma unroll
                for (int s = 0; s < LS; ++s) step(s);
            } else {
                for (int s = 0; s < L; ++s) step(s);
            }
            U1 = -box_eval<T, EPL, STRICT, BOX, true, true, STRICT>(a.tdata, w, bx, x, g, Jr);  // value-only call of :178 fused in
            kick_half<EPL, STRICT, BOX>(p, g, Jr, eps, heps);
        } else {
            U1 = U;
        }

        double K1;
        if (DENSE_M) {
            double tmp[EPL];
            stage_vec<EPL>(mscr, d, lane, p);
            gemv_cm<EPL, STRICT>(a.Minv_cm, d, lane, mscr, 1.0, tmp);
            K1 = A::mul(0.5, STRICT ? warp_dot<EPL, STRICT>(p, tmp) : lane_dot<EPL, STRICT>(p, tmp));
        } else {
            K1 = A::mul(0.5, STRICT ? warp_dot<EPL, STRICT>(p, p) : lane_dot<EPL, STRICT>(p, p));
        }

        // ---- Metropolis test ----
        const double u = PIPE ? u_pipe : rng.uniform(a.rng, t, 0);
        bool acc;
        if (STRICT) {
            // the reference's expression, evaluated literally (src/hmc.cpp:180-191)
            if (!isfinite(U1)) U1 = CUDART_INF;
            const double comp = fmin(0.01, A::add(-A::add(U1, K1), A::add(U, K0)));
            acc = u < exp(comp);
        } else {
            // dH = (U0 + K0) - (U1 + K1) in one butterfly.  u < exp(min(0.01, dH)) holds trivially for dH >= 0 (u < 1)
            // and whenever u < 1 + dH (<= exp(dH)), so exp() is evaluated only in the thin band 1 + dH <= u: the
            // decision is always that of the exact test.  A non-finite energy rejects (src/hmc.cpp:180-182).
            const double dH = warp_sum<false>((U + K0) - (U1 + K1));
            acc = false;
            if (fabs(dH) <= 1.7976931348623157e308) acc = (u < 1.0 + dH) ? true : (u < exp(dH));
        }
        if (acc) {
            U = U1;
        } else if (L > 0) {
            if (FT) load_vec_full<EPL>(bscr, lane, x);
            else load_vec<EPL>(bscr, d, lane, x);
        }
        if (t >= n_burnin) {
            if (BOX) {   // draws_out rows are mapped back with inv_transform (src/hmc.cpp:211-218)
                double xo[EPL];
#pragma unroll
                for (int k = 0; k < EPL; ++k) xo[k] = bx.inv(BOX ? k : 0, x[k]);
                store_vec<EPL>(out_row, d, lane, xo);
            } else if (FT) store_vec_full<EPL>(out_row, lane, x);
            else store_vec<EPL>(out_row, d, lane, x);
            out_row += d;
            if (out_lp) {
                const double Ur = STRICT ? U : warp_sum<false>(U);
                if (lane == 0) *out_lp = -Ur;
                ++out_lp;
            }
            n_acc += acc ? 1 : 0;
        }
    }
    if (lane == 0 && a.n_accept) a.n_accept[chain] = n_acc;
}

template <class T, int EPL, bool DENSE_M, bool STRICT, int RNGM, bool FT, int LS = 0, bool BOX = false> static int launch_one(const HmcLaunch& a)
{
    const long long blocks = (a.n_chains + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    const int dpad = (a.d + 1) & ~1;
    const size_t smem = (size_t)WARPS_PER_BLOCK * ((T::needs_scratch || DENSE_M) ? 3 : 1) * dpad * sizeof(double);
    auto kern = hmc_kernel<T, EPL, DENSE_M, STRICT, RNGM, FT, LS, BOX>;
    if (smem > 40 * 1024) MCMCB200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, WARPS_PER_BLOCK * 32, smem, a.stream>>>(a);
    MCMCB200_CUDA_TRY(cudaGetLastError());
    return MCMCB200_OK;
}

template <class T, int EPL, bool DENSE_M> static int launch_mass(const HmcLaunch& a)
{
    if (a.lb != nullptr) {   // box constraints: generic kernels, M = I
        if (DENSE_M) {
            set_error("hmc: vals_bound together with precond_mat is not supported on the device path");
            return MCMCB200_ERR_UNSUPPORTED;
        }
        if (a.rng.mode == RNG_PHILOX)
            return a.strict ? launch_one<T, EPL, false, true, RNG_PHILOX, false, 0, true>(a) : launch_one<T, EPL, false, false, RNG_PHILOX, false, 0, true>(a);
        return a.strict ? launch_one<T, EPL, false, true, RNG_TAPE, false, 0, true>(a) : launch_one<T, EPL, false, false, RNG_TAPE, false, 0, true>(a);
    }
    if (a.rng.mode == RNG_PHILOX) {
        // the unpredicated full-tile kernels exist for the production configuration: Philox, identity mass
        const bool ft = !DENSE_M && a.d == 32 * EPL && ((reinterpret_cast<uintptr_t>(a.x0) | reinterpret_cast<uintptr_t>(a.draws)) & 15) == 0;
        if (!DENSE_M && ft) {
            if (a.strict) return launch_one<T, EPL, false, true, RNG_PHILOX, true>(a);
            // the most common trajectory lengths get fully unrolled, software-pipelined kernels
            if (a.n_leap == 10) return launch_one<T, EPL, false, false, RNG_PHILOX, true, 10>(a);
            return launch_one<T, EPL, false, false, RNG_PHILOX, true>(a);
        }
        return a.strict ? launch_one<T, EPL, DENSE_M, true, RNG_PHILOX, false>(a) : launch_one<T, EPL, DENSE_M, false, RNG_PHILOX, false>(a);
    }
    return a.strict ? launch_one<T, EPL, DENSE_M, true, RNG_TAPE, false>(a) : launch_one<T, EPL, DENSE_M, false, RNG_TAPE, false>(a);
}

template <class T, int EPL> static int launch_epl(const HmcLaunch& a)
{
    return (a.S_cm != nullptr) ? launch_mass<T, EPL, true>(a) : launch_mass<T, EPL, false>(a);
}

template <class T> static int launch_target(const HmcLaunch& a)
{
    switch (epl_for_dim(a.d)) {
    case 2: return launch_epl<T, 2>(a);
    case 4: return launch_epl<T, 4>(a);
    case 8: return launch_epl<T, 8>(a);
    case 16: return launch_epl<T, 16>(a);
    default:
        set_error("hmc: n_dim=%d exceeds the register-resident kernels (max %d)", a.d, 32 * MAX_EPL);
        return MCMCB200_ERR_UNSUPPORTED;
    }
}

int launch_hmc(const HmcLaunch& a)
{
    switch (a.target_id) {
#define X(ID, TYPE) \
    case ID: return launch_target<TYPE>(a);
        MCMCB200_FOREACH_TARGET(X)
#undef X
    default:
        set_error("hmc: unknown target id %d", a.target_id);
        return MCMCB200_ERR_UNKNOWN_TARGET;
    }
}

}  // namespace mcmcb200
