// Many-chain MALA: one persistent kernel, one warp per chain.
//
// Replaces internal::mala_impl (/root/reference/src/mala.cpp:30-208) + mala_prop_adjustment
// (include/mcmc/mala.ipp:30-70) + stats_mcmc::dmvnorm (include/stats/dmvnorm.hpp:28-54), run once per chain.
// Per draw, with mu(v) = v + ((eps^2 M) grad log pi(v))/2  (src/mala.cpp:123):
//   z ~ N(0,I);  y = mu(x) + (eps sqrtM) z                                  :150-159
//   LP1 = log pi(y)  (non-finite -> -inf)                                    :162-166
//   adj = log N(x; mu(y), eps^2 M) - log N(y; mu(x), eps^2 M)                mala.ipp:63-64
//   accept iff u < exp(min(0.01, LP1 - LP + adj))                            src/mala.cpp:170-173
// What the reference spends per draw — 3 gradient calls + 1 value call and two dmvnorm() evaluations, each a
// Cholesky log-det plus a pivoted-QR solve of the d x d matrix eps^2 M (O(d^3)) — collapses to ONE
// fused value+gradient evaluation at the proposal: the gradient / mean at x are carried from the draw that
// accepted x, the two log-dets and the -d/2 log(2 pi) constants cancel exactly, and Sigma^-1 = (eps^2 M)^-1 is
// constant (inverted once on the host), leaving adj = -1/2 (q(x - mu(y)) - q(y - mu(x))), q(r) = r' Sigma^-1 r.
// This is the "cancelled form" the oracle restates (oracle.cpp run_mala, mala_exact_dmvnorm = 0) and checks
// against the literal reference.
#include "engine.h"
#include "rng.cuh"
#include "targets.cuh"
#include "box.cuh"
#include <math_constants.h>

namespace mcmcb200
{

constexpr int mala_min_blocks(int epl) { return epl <= 4 ? 7 : (epl == 8 ? 3 : 1); }

// BOX: box constraints (src/mala.cpp:104-118,152-157, mala.ipp:50-56): drift ((eps^2 J) M) grad / 2, noise
// ((eps sqrt(J)) sqrtM) z, and BOTH proposal densities use the covariance eps^2 J(proposal) M (not symmetric
// for a dense M — the reference's dmvnorm takes it as it is, and so does the cancelled form: Sigma^-1 r = M^-1 (r / (J eps^2)),
// the two log-dets are the same number).  With DENSE_M the launch's SigInv_cm holds M^-1 (not (eps^2 M)^-1).
template <class T, int EPL, bool DENSE_M, bool STRICT, int RNGM, bool BOX = false>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, mala_min_blocks(EPL)) mala_kernel(const __grid_constant__ MalaLaunch a)
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
    if (chain >= a.n_chains) return;
    const int d = a.d;
    const int dpad = (d + 1) & ~1;
    double* tscr = smem + (size_t)warp * 2 * dpad;
    double* mscr = tscr + dpad;
    const WarpCtx w{lane, d, tscr};
    const double eps = a.eps;
    const double e2 = A::mul(eps, eps);  // step_size * step_size

    BoxLane<BOX ? EPL : 1> bx;
    if (BOX) bx.load(a.lb, a.ub, d, lane);
    // mean = v + ((eps^2 M) g)/2 ; bounded (M = I): v + ((J eps^2) g)/2
    auto mala_mean = [&](const double (&v)[EPL], const double (&g)[EPL], const double (&Jv)[EPL], double (&out)[EPL]) {
        if (BOX && DENSE_M) {
            double t[EPL];
            stage_vec<EPL>(mscr, d, lane, g);
            gemv_cm_rowscaled<EPL, STRICT>(a.M_cm, d, lane, mscr, Jv, e2, t);
#pragma unroll
            for (int k = 0; k < EPL; ++k) out[k] = A::add(v[k], A::mul(t[k], 0.5));
        } else if (BOX) {
#pragma unroll
            for (int k = 0; k < EPL; ++k) out[k] = A::add(v[k], A::mul(A::mul(A::mul(Jv[k], e2), g[k]), 0.5));
        } else if (DENSE_M) {
            double t[EPL];
            stage_vec<EPL>(mscr, d, lane, g);
            gemv_cm<EPL, STRICT>(a.M_cm, d, lane, mscr, e2, t);
#pragma unroll
            for (int k = 0; k < EPL; ++k) out[k] = A::add(v[k], A::mul(t[k], 0.5));
        } else {
#pragma unroll
            for (int k = 0; k < EPL; ++k) out[k] = STRICT ? A::add(v[k], A::mul(A::mul(e2, g[k]), 0.5)) : fma(0.5 * e2, g[k], v[k]);
        }
    };
    // lane partial of q(r) = r' Sigma^-1 r ; bounded: Sigma = diag(J eps^2) with J at the proposal
    auto quad_lane = [&](const double (&r)[EPL], const double (&Jp)[EPL]) -> double {
        if (BOX) {
            double t[EPL];
#pragma unroll
            for (int k = 0; k < EPL; ++k) t[k] = r[k] / A::mul(Jp[k], e2);
            if (DENSE_M) {
                double s2[EPL];
                stage_vec<EPL>(mscr, d, lane, t);
                gemv_cm<EPL, STRICT>(a.SigInv_cm, d, lane, mscr, 1.0, s2);   // M^-1 (r / (J eps^2))
                return lane_dot<EPL, STRICT>(r, s2);
            }
            return lane_dot<EPL, STRICT>(r, t);
        }
        if (DENSE_M) {
            double t[EPL];
            stage_vec<EPL>(mscr, d, lane, r);
            gemv_cm<EPL, STRICT>(a.SigInv_cm, d, lane, mscr, 1.0, t);
            return lane_dot<EPL, STRICT>(r, t);
        }
        return lane_dot<EPL, STRICT>(r, r);
    };

    double x[EPL], mx[EPL], y[EPL], my[EPL], g[EPL], r[EPL], Jx[EPL], Jy[EPL];   // Jx, Jy dead unless BOX
    load_vec<EPL>(a.x0 + (a.broadcast_x0 ? 0 : chain * d), d, lane, x);
    if (BOX) {
#pragma unroll
        for (int k = 0; k < EPL; ++k) x[k] = bx.transform(BOX ? k : 0, x[k]);
    }
    ChainRng<RNGM> rng;
    rng.init(a.rng, chain, a.chain_offset + chain);

    // LP: STRICT carries the reduced scalar, FAST the lane partial (see hmc.cu)
    double LP = box_eval<T, EPL, STRICT, BOX, true, true, STRICT>(a.tdata, w, bx, x, g, Jx);  // src/mala.cpp:138
    mala_mean(x, g, Jx, mx);
    int n_acc = 0;
    const int n_total = (int)(a.n_burnin + a.n_keep);
    const int n_burnin = (int)a.n_burnin;
    double* out_row = a.draws + chain * a.n_keep * d;
    double* out_lp = a.logp ? a.logp + chain * a.n_keep : nullptr;

    for (int t = 0; t < n_total; ++t) {
        rng.template normals<EPL, false>(a.rng, t, d, lane, rng_tab, r);  // z
        if (BOX && DENSE_M) {   // mean + ((eps chol(J)) sqrtM) z: rows of sqrtM scaled by sqrt(J_ii), then by eps
            double tz[EPL], sj[EPL];
#pragma unroll
            for (int k = 0; k < EPL; ++k) sj[k] = sqrt(Jx[k]);
            stage_vec<EPL>(mscr, d, lane, r);
            gemv_cm_rowscaled<EPL, STRICT>(a.S_cm, d, lane, mscr, sj, eps, tz);
#pragma unroll
            for (int k = 0; k < EPL; ++k) y[k] = A::add(mx[k], tz[k]);
        } else if (BOX) {   // mean + ((eps chol(J)) sqrtM) z with chol of the diagonal J = sqrt(J_ii), sqrtM = I
#pragma unroll
            for (int k = 0; k < EPL; ++k) y[k] = A::mad(A::mul(sqrt(Jx[k]), eps), r[k], mx[k]);
        } else if (DENSE_M) {
            double tz[EPL];
            stage_vec<EPL>(mscr, d, lane, r);
            gemv_cm<EPL, STRICT>(a.S_cm, d, lane, mscr, eps, tz);  // (eps sqrtM) z
#pragma unroll
            for (int k = 0; k < EPL; ++k) y[k] = A::add(mx[k], tz[k]);
        } else {
#pragma unroll
            for (int k = 0; k < EPL; ++k) y[k] = A::mad(eps, r[k], mx[k]);
        }
        double LP1 = box_eval<T, EPL, STRICT, BOX, true, true, STRICT>(a.tdata, w, bx, y, g, Jy);
        mala_mean(y, g, Jy, my);

        // q1 = q(x - mu(y)), q2 = q(y - mu(x))
#pragma unroll
        for (int k = 0; k < EPL; ++k) r[k] = A::sub(x[k], my[k]);
        double q1 = quad_lane(r, Jy);
#pragma unroll
        for (int k = 0; k < EPL; ++k) r[k] = A::sub(y[k], mx[k]);
        double q2 = quad_lane(r, Jy);

        const double u = rng.uniform(a.rng, t, 0);
        bool acc;
        if (STRICT) {
            q1 = warp_sum<true>(q1);
            q2 = warp_sum<true>(q2);
            if (!DENSE_M && !BOX) {  // Sigma^-1 = I / eps^2
                q1 = q1 / e2;
                q2 = q2 / e2;
            }
            if (!isfinite(LP1)) LP1 = -CUDART_INF;                            // src/mala.cpp:164-166
            const double adj = A::mul(-0.5, A::sub(q1, q2));
            const double comp = fmin(0.01, A::add(A::sub(LP1, LP), adj));      // :170
            acc = u < exp(comp);
        } else {
            const double qs = (DENSE_M || BOX) ? 1.0 : 1.0 / e2;
            const double dl = warp_sum<false>((LP1 - LP) - 0.5 * qs * (q1 - q2));
            acc = u < 1.0 + dl;   // dl = +inf (chain started where log pi = -inf) accepts like the reference; NaN / -inf reject
            if (!acc) acc = (fabs(dl) <= 1.7976931348623157e308) && (u < exp(dl));
        }
        if (acc) {
            LP = LP1;
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                x[k] = y[k];
                mx[k] = my[k];
                if (BOX) Jx[k] = Jy[k];
            }
        }
        if (t >= n_burnin) {
            if (BOX) {   // src/mala.cpp:192-199
                double xo[EPL];
#pragma unroll
                for (int k = 0; k < EPL; ++k) xo[k] = bx.inv(BOX ? k : 0, x[k]);
                store_vec<EPL>(out_row, d, lane, xo);
            } else
            store_vec<EPL>(out_row, d, lane, x);
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

template <class T, int EPL, bool DENSE_M, bool STRICT, int RNGM, bool BOX = false> static int launch_one(const MalaLaunch& a)
{
    const long long blocks = (a.n_chains + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    const int dpad = (a.d + 1) & ~1;
    const size_t smem = (T::needs_scratch || DENSE_M) ? (size_t)WARPS_PER_BLOCK * 2 * dpad * sizeof(double) : 0;
    auto kern = mala_kernel<T, EPL, DENSE_M, STRICT, RNGM, BOX>;
    if (smem > 16 * 1024) MCMCB200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, WARPS_PER_BLOCK * 32, smem, a.stream>>>(a);
    MCMCB200_CUDA_TRY(cudaGetLastError());
    return MCMCB200_OK;
}

template <class T, int EPL, bool DENSE_M> static int launch_mass(const MalaLaunch& a)
{
    if (a.lb != nullptr) {   // box constraints, with or without a dense precond_mat
        if (a.rng.mode == RNG_PHILOX)
            return a.strict ? launch_one<T, EPL, DENSE_M, true, RNG_PHILOX, true>(a) : launch_one<T, EPL, DENSE_M, false, RNG_PHILOX, true>(a);
        return a.strict ? launch_one<T, EPL, DENSE_M, true, RNG_TAPE, true>(a) : launch_one<T, EPL, DENSE_M, false, RNG_TAPE, true>(a);
    }
    if (a.rng.mode == RNG_PHILOX)
        return a.strict ? launch_one<T, EPL, DENSE_M, true, RNG_PHILOX>(a) : launch_one<T, EPL, DENSE_M, false, RNG_PHILOX>(a);
    return a.strict ? launch_one<T, EPL, DENSE_M, true, RNG_TAPE>(a) : launch_one<T, EPL, DENSE_M, false, RNG_TAPE>(a);
}

template <class T> static int launch_target(const MalaLaunch& a)
{
    const bool dense = a.S_cm != nullptr;
    switch (epl_for_dim(a.d)) {
    MCMCB200_EPL_CASE(2, (dense ? launch_mass<T, 2, true>(a) : launch_mass<T, 2, false>(a)))
    MCMCB200_EPL_CASE(4, (dense ? launch_mass<T, 4, true>(a) : launch_mass<T, 4, false>(a)))
    MCMCB200_EPL_CASE(8, (dense ? launch_mass<T, 8, true>(a) : launch_mass<T, 8, false>(a)))
    MCMCB200_EPL_CASE(16, (dense ? launch_mass<T, 16, true>(a) : launch_mass<T, 16, false>(a)))
    default:
        set_error("mala: n_dim=%d exceeds the register-resident kernels (max %d)", a.d, 32 * MAX_EPL);
        return MCMCB200_ERR_UNSUPPORTED;
    }
}

int MCMCB200_SLICED(launch_mala)(const MalaLaunch& a)
{
    switch (a.target_id) {
#define X(ID, TYPE) \
    case ID: return launch_target<TYPE>(a);
        MCMCB200_FOREACH_TARGET(X)
#undef X
    default:
        set_error("mala: unknown target id %d", a.target_id);
        return MCMCB200_ERR_UNKNOWN_TARGET;
    }
}

}  // namespace mcmcb200
