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
#include <type_traits>

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
// BOX: box constraints (vals_bound), see box.cuh; on the generic (runtime-L) kernels, with or without a dense mass matrix.
template <class T, int EPL, bool DENSE_M, bool STRICT, int RNGM, bool FT, bool BOX = false>
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
    const int L = a.n_leap;
    double* out_row = a.draws + chain * a.n_keep * d;
    double* out_lp = a.logp ? a.logp + chain * a.n_keep : nullptr;

    for (int t = 0; t < n_total; ++t) {
        // ---- momentum refresh: p = sqrtM z, K0 = p.(M^-1 p)/2 (lane partial in FAST) ----
        rng.template normals<EPL, FT>(a.rng, t, d, lane, log_tab, p);
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
            for (int s = 0; s < L; ++s) step(s);
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
        const double u = rng.uniform(a.rng, t, 0);
        bool acc;
        if (STRICT) {
            // the reference's expression, evaluated literally (src/hmc.cpp:180-191)
            if (!isfinite(U1)) U1 = CUDART_INF;
            const double comp = fmin(0.01, A::add(-A::add(U1, K1), A::add(U, K0)));
            acc = u < exp(comp);
        } else {
            // dH = (U0 + K0) - (U1 + K1) in one butterfly.  u < exp(min(0.01, dH)) holds trivially for dH >= 0 (u < 1)
            // and whenever u < 1 + dH (<= exp(dH)), so exp() is evaluated only in the thin band 1 + dH <= u: the
            // decision is always that of the exact test.  A non-finite proposal energy rejects (dH = -inf or NaN,
            // src/hmc.cpp:180-182); a chain that starts where log pi = -inf (U0 = +inf, dH = +inf) accepts its first
            // finite proposal like the reference (min(0.01, +inf) = 0.01).  Same rule in every FAST kernel.
            const double dH = warp_sum<false>((U + K0) - (U1 + K1));
            acc = u < 1.0 + dH;
            if (!acc) acc = (fabs(dH) <= 1.7976931348623157e308) && (u < exp(dH));
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

// ---------------------------------------------------------------------------------------------------------------
// Production kernel: FAST arithmetic, in-kernel Philox, identity mass, full tiles (n_dim == 32*EPL) — the
// configuration of the headline benchmark.  Same per-draw algorithm as hmc_kernel; what differs is how the draw
// loop is laid out for the issue-bound fp64 pipe:
//   * software pipelining: the variates of draw t+1 (independent of the chain state) are generated in the same
//     straight-line block as the trajectory of draw t; the draw loop is unrolled by two so the "current" and "next"
//     variate registers swap roles without copies;
//   * K0 = sum z^2 / 2 comes from the Box-Muller radii (z0^2 + z1^2 = -2 ln u1), not from a second dot product;
//   * no control flow on the accept decision: the trajectory runs in place on x, the pre-trajectory state is parked
//     in shared memory and reloaded by a PREDICATED ld.shared into the same registers on a rejection, so neither
//     path needs register moves; exp() is evaluated only in the thin band 1 + dH <= u;
//   * burn-in and kept draws are separate loops (no per-draw store predicate, no carried row pointer).
// LS = compile-time number of leapfrog steps that carry the slices of the next draw's variates (0 = none: runtime a.n_leap,
// variates generated up front), UNR = draw-loop unroll (1 or 2), MORE = the trajectory continues after the LS unrolled steps
// with a.n_leap - LS plain steps (so every trajectory length >= LS keeps the interleaved variate generation, not only LS).
template <int EPL> __device__ __forceinline__ void restore_if(const double* home, int lane, double (&x)[EPL], bool pred)
{
    const unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(home + 2 * lane));
#pragma unroll
    for (int m = 0; m < EPL / 2; ++m)
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t@q ld.volatile.shared.v2.f64 {%0, %1}, [%2];\n\t}"
                     : "+d"(x[2 * m]), "+d"(x[2 * m + 1])
                     : "r"(addr + m * 512), "r"(static_cast<int>(pred))
                     : "memory");
}

// the matching store is opaque too, so ptxas cannot forward the stored registers into the predicated reload (it
// would keep a second copy of the state live in registers and pay a move per word)
template <int EPL> __device__ __forceinline__ void park(double* home, int lane, const double (&x)[EPL])
{
    const unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(home + 2 * lane));
#pragma unroll
    for (int m = 0; m < EPL / 2; ++m)
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" : : "r"(addr + m * 512), "d"(x[2 * m]), "d"(x[2 * m + 1]) : "memory");
}

template <class T, int EPL, int LS, int UNR, bool MORE = false>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, hmc_min_blocks(EPL)) hmc_pipe_kernel(const __grid_constant__ HmcLaunch a)
{
    extern __shared__ double smem[];
    __shared__ double2 rng_tab[RNG_TAB_DOUBLE2];
    build_rng_tables(rng_tab);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long chain = (long long)blockIdx.x * WARPS_PER_BLOCK + warp;
    if (chain >= a.n_chains) return;  // whole warp exits together; no block-level barriers below
    constexpr int d = 32 * EPL;
    double* home = smem + (size_t)warp * (T::needs_scratch ? 2 : 1) * d;  // the chain's current state while a trajectory runs
    const WarpCtx w{lane, d, home + d};

    double x[EPL], g[EPL];
    load_vec_full<EPL>(a.x0 + (a.broadcast_x0 ? 0 : chain * d), lane, x);
    ChainRng<RNG_PHILOX> rng;
    rng.init(a.rng, chain, a.chain_offset + chain);
    double U = -T::template eval<EPL, false, true, false, false>(a.tdata, w, x, g);  // this lane's partial sum of -log pi(x)
    int n_acc = 0;
    const int n_burnin = (int)a.n_burnin, n_total = (int)(a.n_burnin + a.n_keep);
    const double eps = a.eps, heps = 0.5 * eps;
    const int L = a.n_leap;
    double* const out_base = a.draws + chain * a.n_keep * d + 2 * lane;
    double* const out_lp = a.logp ? a.logp + chain * a.n_keep : nullptr;

    double zA[EPL], zB[EPL], uA, uB, kA, kB;  // variates of the current / next draw (roles alternate)
    {
        BmPipe<EPL / 2> bp;
        bp.begin(lane, 0, rng.chain);
        bp.template slice<0, 1>(a.rng, rng_tab, zA, kA);
        uA = bp.uniform0();
    }

    // one transition; p holds z_t on entry (consumed), (zn, un, kn) receive the variates of draw t + 1
#ifndef MCMCB200_PIPE_TAIL
#define MCMCB200_PIPE_TAIL 5   // measured on B200 (tools/c2_variants.cu): 0 -> 2.23 ms, 2 -> 2.18, 3/4 -> 2.10, 5 -> 2.08 ms
#endif
    constexpr int TAIL = LS > 0 ? MCMCB200_PIPE_TAIL : 0;   // slices of RNG work deferred to the butterfly
    constexpr int NS = (LS > 0 ? LS : 1) + TAIL;
    auto draw = [&](int t, double (&p)[EPL], double u, double ksum, double (&zn)[EPL], double& un, double& kn, auto keep) {
        // the variates of draw t + 1 (one spare draw past the end: harmless), generated slice by slice between the
        // leapfrog steps below when the trajectory length is a compile-time constant
        BmPipe<EPL / 2> bp;
        bp.begin(lane, t + 1, rng.chain);
        if (LS == 0) bp.template slice<0, 1>(a.rng, rng_tab, zn, kn);
        double dH = fma(0.5, ksum, U);   // U0 + K0 (lane partial)
        bool acc = true;
        if (LS > 0 || L > 0) {
            park<EPL>(home, lane, x);
            T::template eval<EPL, false, false, true, false>(a.tdata, w, x, g);
            kick_half<EPL, false, false>(p, g, g, eps, heps);
            auto step = [&](int s) {
#pragma unroll
                for (int k = 0; k < EPL; ++k) x[k] = fma(eps, p[k], x[k]);
                if ((LS > 0 && !MORE) ? (s + 1 < LS) : (s + 1 < L)) {
                    T::template eval<EPL, false, false, true, false>(a.tdata, w, x, g);
                    kick_full<EPL, false, false>(p, g, g, eps);
                }
            };
            if (LS > 0) {
                static_for<0, LS>([&](auto sc) {
                    constexpr int s = decltype(sc)::value;
                    bp.template slice<s, NS>(a.rng, rng_tab, zn, kn);
                    step(s);
                });
                if (MORE)
                    for (int s = LS; s < L; ++s) step(s);
            } else {
                for (int s = 0; s < L; ++s) step(s);
            }
            un = bp.uniform0();
            const double U1 = -T::template eval<EPL, false, true, true, false>(a.tdata, w, x, g);
            kick_half<EPL, false, false>(p, g, g, eps, heps);
            // dH = (U0 + K0) - (U1 + K1), one butterfly.  u < exp(min(0.01, dH)) holds whenever u < 1 + dH (<= exp(dH);
            // also dH = +inf, src/hmc.cpp:187), so exp() is evaluated only in the thin band 1 + dH <= u; a NaN rejects.
            dH = fma(-0.5, lane_dot<EPL, false>(p, p), dH - U1);
            static_for<0, 5>([&](auto st) {   // the butterfly, with the last slices of the next draw's variates in its shadow
                constexpr int k = decltype(st)::value;
                dH += __shfl_xor_sync(FULL, dH, 16 >> k);
                if constexpr (k < TAIL) bp.template slice<(LS > 0 ? LS : 1) + k, NS>(a.rng, rng_tab, zn, kn);
            });
            acc = u < 1.0 + dH;
            if (!acc) acc = (fabs(dH) <= 1.7976931348623157e308) && (u < exp(dH));
            restore_if<EPL>(home, lane, x, !acc);
            U = acc ? U1 : U;
        }
        if (LS == 0 && L == 0) un = bp.uniform0();
        if (decltype(keep)::value) {
            double* row = out_base + (size_t)(t - n_burnin) * d;
#pragma unroll
            for (int m = 0; m < EPL / 2; ++m) *reinterpret_cast<double2*>(row + m * 64) = make_double2(x[2 * m], x[2 * m + 1]);
            if (out_lp) {
                const double Ur = warp_sum<false>(U);
                if (lane == 0) out_lp[t - n_burnin] = -Ur;
            }
            n_acc += acc ? 1 : 0;
        }
    };
    auto run = [&](int t0, int t1, auto keep) {
        for (int t = t0; t < t1; t += UNR) {
            draw(t, zA, uA, kA, zB, uB, kB, keep);
            if (UNR == 2 && t + 1 < t1) {
                draw(t + 1, zB, uB, kB, zA, uA, kA, keep);
            } else {
#pragma unroll
                for (int k = 0; k < EPL; ++k) zA[k] = zB[k];
                uA = uB;
                kA = kB;
            }
        }
    };
    run(0, n_burnin, std::false_type());
    run(n_burnin, n_total, std::true_type());
    if (lane == 0 && a.n_accept) a.n_accept[chain] = n_acc;
}

#ifndef MCMCB200_KERNEL_ONLY   // tools/c2_variants.cu instantiates the headline kernel alone
template <class T, int EPL, bool DENSE_M, bool STRICT, int RNGM, bool FT, bool BOX = false> static int launch_one(const HmcLaunch& a)
{
    const long long blocks = (a.n_chains + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    const int dpad = (a.d + 1) & ~1;
    const size_t smem = (size_t)WARPS_PER_BLOCK * ((T::needs_scratch || DENSE_M) ? 3 : 1) * dpad * sizeof(double);
    auto kern = hmc_kernel<T, EPL, DENSE_M, STRICT, RNGM, FT, BOX>;
    if (smem > 16 * 1024) MCMCB200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, WARPS_PER_BLOCK * 32, smem, a.stream>>>(a);
    MCMCB200_CUDA_TRY(cudaGetLastError());
    return MCMCB200_OK;
}

template <class T, int EPL, int LS, bool MORE = false> static int launch_pipe(const HmcLaunch& a)
{
    const long long blocks = (a.n_chains + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    const size_t smem = (size_t)WARPS_PER_BLOCK * (T::needs_scratch ? 2 : 1) * a.d * sizeof(double);
    auto kern = hmc_pipe_kernel<T, EPL, LS, 1, MORE>;   // UNR = 2 is slower on B200 (instruction cache): 2.33 vs 2.23 ms
    if (smem > 16 * 1024) MCMCB200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, WARPS_PER_BLOCK * 32, smem, a.stream>>>(a);
    MCMCB200_CUDA_TRY(cudaGetLastError());
    return MCMCB200_OK;
}

template <class T, int EPL, bool DENSE_M> static int launch_mass(const HmcLaunch& a)
{
    if (a.lb != nullptr) {   // box constraints: generic kernels; the kick uses J o grad, the drift (eps M^-1) p (src/hmc.cpp:107-122,171)
        if (a.rng.mode == RNG_PHILOX)
            return a.strict ? launch_one<T, EPL, DENSE_M, true, RNG_PHILOX, false, true>(a) : launch_one<T, EPL, DENSE_M, false, RNG_PHILOX, false, true>(a);
        return a.strict ? launch_one<T, EPL, DENSE_M, true, RNG_TAPE, false, true>(a) : launch_one<T, EPL, DENSE_M, false, RNG_TAPE, false, true>(a);
    }
    if (a.rng.mode == RNG_PHILOX) {
        // the unpredicated full-tile kernels exist for the production configuration: Philox, identity mass
        const bool ft = !DENSE_M && a.d == 32 * EPL && ((reinterpret_cast<uintptr_t>(a.x0) | reinterpret_cast<uintptr_t>(a.draws)) & 15) == 0;
        if (!DENSE_M && ft) {
            if (a.strict) return launch_one<T, EPL, false, true, RNG_PHILOX, true>(a);
            // production configuration: the software-pipelined kernel; the most common trajectory length is unrolled
            if constexpr (T::separable && EPL <= 8) {   // EPL = 16: the unrolled loop outgrows the instruction cache (2.67 vs 2.55 ms)
                if (a.n_leap == 10) return launch_pipe<T, EPL, 10>(a);
                // (a variant that interleaves the variates with the first five steps of ANY trajectory length >= 5 — LS = 5,
                //  MORE = true — measured no faster than the up-front generation on B200: L = 5 2.197 vs 2.204 ms; not instantiated)
            }
            return launch_pipe<T, EPL, 0>(a);
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
    MCMCB200_EPL_CASE(2, (launch_epl<T, 2>(a)))
    MCMCB200_EPL_CASE(4, (launch_epl<T, 4>(a)))
    MCMCB200_EPL_CASE(8, (launch_epl<T, 8>(a)))
    MCMCB200_EPL_CASE(16, (launch_epl<T, 16>(a)))
    default:
        set_error("hmc: n_dim=%d exceeds the register-resident kernels (max %d)", a.d, 32 * MAX_EPL);
        return MCMCB200_ERR_UNSUPPORTED;
    }
}

int MCMCB200_SLICED(launch_hmc)(const HmcLaunch& a)
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
#endif  // MCMCB200_KERNEL_ONLY

}  // namespace mcmcb200
