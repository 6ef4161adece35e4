// Many-chain random-walk Metropolis-Hastings: one persistent kernel, one warp per chain.
//
// Replaces internal::rwmh_impl (/root/reference/src/rwmh.cpp:30-172) run once per chain — the gradient-free
// sibling of MALA, same boundary and output format.  Per draw:
//   z ~ N(0,I);  y = x + (par_scale chol(cov)) z                              src/rwmh.cpp:133-135
//   LP1 = log pi(y)  (+ log-Jacobian with box constraints; non-finite -> -inf) :137-141
//   accept iff u < exp(min(0, LP1 - LP))                                       :145-155
//   kept draws written to draws_out (mapped back with inv_transform when bounded, :166-173)
// The chain state never leaves registers; HBM traffic is the initial x once and one d*8-byte row per kept draw.
// With cov_mat empty the reference multiplies by par_scale * chol(I), a dense matrix whose only non-zero entries
// are par_scale on the diagonal; adding the exact zeros changes nothing, so DENSE_C = false computes
// x + par_scale * z (two roundings in STRICT mode, as the reference's product-then-sum; one FMA in FAST mode).
#include "engine.h"
#include "rng.cuh"
#include "targets.cuh"
#include "box.cuh"
#include <math_constants.h>

namespace mcmcb200
{

constexpr int rwmh_min_blocks(int epl) { return epl <= 4 ? 7 : (epl == 8 ? 4 : 2); }

template <class T, int EPL, bool DENSE_C, bool STRICT, int RNGM, bool BOX = false>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, rwmh_min_blocks(EPL)) rwmh_kernel(const __grid_constant__ RwmhLaunch a)
{
    extern __shared__ double smem[];
    __shared__ double2 rng_tab[RNGM == RNG_PHILOX ? RNG_TAB_DOUBLE2 : 1];
    typedef Ar<STRICT> A;
    if (RNGM == RNG_PHILOX) {
        build_rng_tables(rng_tab);
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long chain = (long long)blockIdx.x * WARPS_PER_BLOCK + warp;
    if (chain >= a.n_chains) return;  // whole warp exits together; no block-level barriers below
    const int d = a.d;
    const int dpad = (d + 1) & ~1;
    double* tscr = smem + (size_t)warp * 2 * dpad;  // target functor scratch
    double* mscr = tscr + dpad;                      // staged z for the dense product
    const WarpCtx w{lane, d, tscr};
    const double scale = a.par_scale;

    BoxLane<BOX ? EPL : 1> bx;
    if (BOX) bx.load(a.lb, a.ub, d, lane);
    double x[EPL], y[EPL], z[EPL], gdummy[EPL], Jdummy[EPL];   // the dummies are never written (WANT_GRAD = false)
    load_vec<EPL>(a.x0 + (a.broadcast_x0 ? 0 : chain * d), d, lane, x);
    if (BOX) {
#pragma unroll
        for (int k = 0; k < EPL; ++k) x[k] = bx.transform(BOX ? k : 0, x[k]);   // src/rwmh.cpp:105-107
    }
    ChainRng<RNGM> rng;
    rng.init(a.rng, chain, a.chain_offset + chain);

    // LP: STRICT carries the reduced scalar (src/rwmh.cpp:111), FAST this lane's partial sum (one butterfly per draw)
    double LP = box_eval<T, EPL, STRICT, BOX, true, false, STRICT>(a.tdata, w, bx, x, gdummy, Jdummy);
    int n_acc = 0;
    const int n_total = (int)(a.n_burnin + a.n_keep);
    const int n_burnin = (int)a.n_burnin;
    double* out_row = a.draws + chain * a.n_keep * d;
    double* out_lp = a.logp ? a.logp + chain * a.n_keep : nullptr;

    for (int t = 0; t < n_total; ++t) {
        rng.template normals<EPL, false>(a.rng, t, d, lane, rng_tab, z);
        if (DENSE_C) {
            double tz[EPL];
            stage_vec<EPL>(mscr, d, lane, z);
            gemv_cm<EPL, STRICT>(a.S_cm, d, lane, mscr, 1.0, tz);   // S = par_scale * chol(cov), scaled on the host
#pragma unroll
            for (int k = 0; k < EPL; ++k) y[k] = A::add(x[k], tz[k]);
        } else {
#pragma unroll
            for (int k = 0; k < EPL; ++k) y[k] = STRICT ? A::add(x[k], A::mul(scale, z[k])) : fma(scale, z[k], x[k]);
        }
        double LP1 = box_eval<T, EPL, STRICT, BOX, true, false, STRICT>(a.tdata, w, bx, y, gdummy, Jdummy);
        const double u = rng.uniform(a.rng, t, 0);
        bool acc;
        if (STRICT) {
            if (!isfinite(LP1)) LP1 = -CUDART_INF;                    // src/rwmh.cpp:139-141
            const double comp = fmin(0.0, A::sub(LP1, LP));           // :145 (min(0, NaN) = 0, like std::min)
            acc = u < exp(comp);
        } else {
            // u < exp(min(0, dl)) holds whenever u < 1 + dl (<= exp(dl); also dl = +inf), so exp() is evaluated only
            // in the thin band 1 + dl <= u; a NaN difference rejects.
            const double dl = warp_sum<false>(LP1 - LP);
            acc = u < 1.0 + dl;
            if (!acc) acc = (fabs(dl) <= 1.7976931348623157e308) && (u < exp(dl));
        }
        if (acc) {
            LP = LP1;
#pragma unroll
            for (int k = 0; k < EPL; ++k) x[k] = y[k];
        }
        if (t >= n_burnin) {
            if (BOX) {
                double xo[EPL];
#pragma unroll
                for (int k = 0; k < EPL; ++k) xo[k] = bx.inv(BOX ? k : 0, x[k]);
                store_vec<EPL>(out_row, d, lane, xo);
            } else {
                store_vec<EPL>(out_row, d, lane, x);
            }
            out_row += d;
            if (out_lp) {
                const double lr = STRICT ? LP : warp_sum<false>(LP);
                if (lane == 0) *out_lp = lr;
                ++out_lp;
            }
            n_acc += acc ? 1 : 0;
        }
    }
    if (lane == 0 && a.n_accept) a.n_accept[chain] = n_acc;
}

template <class T, int EPL, bool DENSE_C, bool STRICT, int RNGM, bool BOX> static int launch_one(const RwmhLaunch& a)
{
    const long long blocks = (a.n_chains + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    const int dpad = (a.d + 1) & ~1;
    const size_t smem = (T::needs_scratch || DENSE_C) ? (size_t)WARPS_PER_BLOCK * 2 * dpad * sizeof(double) : 0;
    auto kern = rwmh_kernel<T, EPL, DENSE_C, STRICT, RNGM, BOX>;
    if (smem > 16 * 1024) MCMCB200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, WARPS_PER_BLOCK * 32, smem, a.stream>>>(a);
    MCMCB200_CUDA_TRY(cudaGetLastError());
    return MCMCB200_OK;
}

template <class T, int EPL, bool DENSE_C, bool BOX> static int launch_mode(const RwmhLaunch& a)
{
    if (a.rng.mode == RNG_PHILOX)
        return a.strict ? launch_one<T, EPL, DENSE_C, true, RNG_PHILOX, BOX>(a) : launch_one<T, EPL, DENSE_C, false, RNG_PHILOX, BOX>(a);
    return a.strict ? launch_one<T, EPL, DENSE_C, true, RNG_TAPE, BOX>(a) : launch_one<T, EPL, DENSE_C, false, RNG_TAPE, BOX>(a);
}

template <class T, int EPL> static int launch_epl(const RwmhLaunch& a)
{
    const bool dense = a.S_cm != nullptr, box = a.lb != nullptr;
    if (dense) return box ? launch_mode<T, EPL, true, true>(a) : launch_mode<T, EPL, true, false>(a);
    return box ? launch_mode<T, EPL, false, true>(a) : launch_mode<T, EPL, false, false>(a);
}

template <class T> static int launch_target(const RwmhLaunch& a)
{
    switch (epl_for_dim(a.d)) {
    MCMCB200_EPL_CASE(2, (launch_epl<T, 2>(a)))
    MCMCB200_EPL_CASE(4, (launch_epl<T, 4>(a)))
    MCMCB200_EPL_CASE(8, (launch_epl<T, 8>(a)))
    MCMCB200_EPL_CASE(16, (launch_epl<T, 16>(a)))
    default:
        set_error("rwmh: n_dim=%d exceeds the register-resident kernels (max %d)", a.d, 32 * MAX_EPL);
        return MCMCB200_ERR_UNSUPPORTED;
    }
}

int MCMCB200_SLICED(launch_rwmh)(const RwmhLaunch& a)
{
    switch (a.target_id) {
#define X(ID, TYPE) \
    case ID: return launch_target<TYPE>(a);
        MCMCB200_FOREACH_TARGET(X)
#undef X
    default:
        set_error("rwmh: unknown target id %d", a.target_id);
        return MCMCB200_ERR_UNKNOWN_TARGET;
    }
}

}  // namespace mcmcb200
