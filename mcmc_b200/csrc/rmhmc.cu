// Many-chain RM-HMC (Riemannian-manifold HMC) for small, dense metrics: one THREAD per chain.
//
// Replaces internal::rmhmc_impl (/root/reference/src/rmhmc.cpp:30-294) run once per chain, bug-compatibly
// (Appendix E).  The reference's user callbacks
//     target_log_kernel(vals, grad_out, data)  and  tensor_fn(vals, Cube_t* tensor_deriv_out, tensor_data)
// (include/mcmc/rmhmc.hpp:47-53) become a registered __device__ metric functor that returns G(x) (D x D) and, on
// request, the D derivative matrices dG/dx_i.  Per leapfrog step the reference performs n_fp fixed-point iterations
// of the implicit momentum half-step (each with D products of D x D matrices, src/rmhmc.cpp:132-140), n_fp
// iterations of the implicit position step (each a metric evaluation + D x D inverse, :224-228), a metric+derivative
// evaluation, an inverse and a final half-step: at the D = 2..4 of the registered metrics everything fits in one
// thread's registers, so chains map to threads (32 chains per warp) instead of warps.
//
// Registered metric: the Fisher information of the 2-parameter Normal(mu, sigma) model of
// examples/eigen/rmhmc_normal.cpp:82-111 (target MCMCB200_TARGET_NORMAL_MODEL).  Larger metrics — Neal's funnel with
// its Fisher-type or SoftAbs metric up to n_dim = 64 (BASELINE config 5) — run warp-per-chain in rmhmc_general.cu.
#include "engine.h"
#include "rng.cuh"
#include "box.cuh"
#include <math_constants.h>

namespace mcmcb200
{

// ---- D x D column-major linear algebra in registers, in the oracle's operation order -------------------------
template <int D, bool STRICT> struct SmallLA {
    typedef Ar<STRICT> A;
    // y = M v (column-major axpy order, j increasing)
    static __device__ __forceinline__ void gemv(const double (&M)[D * D], const double (&v)[D], double (&y)[D], double alpha = 1.0)
    {
#pragma unroll
        for (int i = 0; i < D; ++i) y[i] = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const double t = A::mul(alpha, v[j]);
#pragma unroll
            for (int i = 0; i < D; ++i) y[i] = A::mad(M[j * D + i], t, y[i]);
        }
    }
    static __device__ __forceinline__ double dot(const double (&a)[D], const double (&b)[D])
    {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) s = A::mad(a[i], b[i], s);
        return s;
    }
    // C = A B
    static __device__ __forceinline__ void matmul(const double (&Am)[D * D], const double (&B)[D * D], double (&C)[D * D])
    {
#pragma unroll
        for (int k = 0; k < D * D; ++k) C[k] = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j)
#pragma unroll
            for (int l = 0; l < D; ++l) {
                const double t = B[j * D + l];
#pragma unroll
                for (int i = 0; i < D; ++i) C[j * D + i] = A::mad(Am[l * D + i], t, C[j * D + i]);
            }
    }
    // inverse by LU with partial pivoting (BMO_MATOPS_INV -> Eigen inverse(), src/rmhmc.cpp:181,226,233)
    static __device__ __forceinline__ void inverse(const double (&Am)[D * D], double (&inv)[D * D])
    {
        double lu[D * D];
        int piv[D];
#pragma unroll
        for (int k = 0; k < D * D; ++k) lu[k] = Am[k];
#pragma unroll
        for (int i = 0; i < D; ++i) piv[i] = i;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            int p = k;
            double best = fabs(lu[k * D + k]);
#pragma unroll
            for (int i = k + 1; i < D; ++i)
                if (fabs(lu[k * D + i]) > best) { best = fabs(lu[k * D + i]); p = i; }
            if (p != k) {
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const double tmp = lu[j * D + k];
#pragma unroll
                    for (int pp = 0; pp < D; ++pp)
                        if (pp == p) { lu[j * D + k] = lu[j * D + pp]; lu[j * D + pp] = tmp; }
                }
#pragma unroll
                for (int pp = 0; pp < D; ++pp)
                    if (pp == p) { const int tp = piv[k]; piv[k] = piv[pp]; piv[pp] = tp; }
            }
            const double dd = lu[k * D + k];
#pragma unroll
            for (int i = k + 1; i < D; ++i) lu[k * D + i] = lu[k * D + i] / dd;
#pragma unroll
            for (int j = k + 1; j < D; ++j) {
                const double t = lu[j * D + k];
#pragma unroll
                for (int i = k + 1; i < D; ++i) lu[j * D + i] = A::sub(lu[j * D + i], A::mul(lu[k * D + i], t));
            }
        }
#pragma unroll
        for (int c = 0; c < D; ++c) {
            double y[D];
#pragma unroll
            for (int i = 0; i < D; ++i) y[i] = (piv[i] == c) ? 1.0 : 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                double s = y[i];
#pragma unroll
                for (int j = 0; j < i; ++j) s = A::sub(s, A::mul(lu[j * D + i], y[j]));
                y[i] = s;
            }
#pragma unroll
            for (int i = D - 1; i >= 0; --i) {
                double s = y[i];
#pragma unroll
                for (int j = i + 1; j < D; ++j) s = A::sub(s, A::mul(lu[j * D + i], y[j]));
                y[i] = s / lu[i * D + i];
            }
#pragma unroll
            for (int i = 0; i < D; ++i) inv[c * D + i] = y[i];
        }
    }
    // lower Cholesky in place; chol_mode MCMCB200_CHOL_EIGEN_LLT keeps A's strict upper triangle
    static __device__ __forceinline__ void chol(const double (&Am)[D * D], int chol_mode, double (&L)[D * D])
    {
#pragma unroll
        for (int k = 0; k < D * D; ++k) L[k] = Am[k];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            double s = L[j * D + j];
#pragma unroll
            for (int k = 0; k < j; ++k) s = A::sub(s, A::mul(L[k * D + j], L[k * D + j]));
            const double dd = sqrt(s);
            L[j * D + j] = dd;
#pragma unroll
            for (int i = j + 1; i < D; ++i) {
                double t = L[j * D + i];
#pragma unroll
                for (int k = 0; k < j; ++k) t = A::sub(t, A::mul(L[k * D + i], L[k * D + j]));
                L[j * D + i] = t / dd;
            }
        }
        if (chol_mode == MCMCB200_CHOL_LOWER) {
#pragma unroll
            for (int j = 1; j < D; ++j)
#pragma unroll
                for (int i = 0; i < j; ++i) L[j * D + i] = 0.0;
        }
    }
    // (diag(llt).log() * 2).sum()  (core/log_det.hpp:35)
    static __device__ __forceinline__ double logdet(const double (&Am)[D * D])
    {
        double L[D * D];
        chol(Am, MCMCB200_CHOL_EIGEN_LLT, L);
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) s = A::add(s, A::mul(log(L[i * D + i]), 2.0));
        return s;
    }
};

// ---- registered target + metric for thread-per-chain kernels --------------------------------------------------
// Normal(mu, sigma) likelihood on sufficient statistics; metric G = diag(n/sigma^2, 2n/sigma^2), dG/dmu = 0,
// dG/dsigma = -2 G / sigma  (examples/eigen/rmhmc_normal.cpp:82-111; host twin: oracle/host_targets.hpp)
template <bool STRICT> struct NormalModelRM {
    static constexpr int D = 2;
    typedef Ar<STRICT> A;
    static __device__ __forceinline__ double logp(const double* __restrict__ data, const double (&x)[2], double* g)
    {
        const double n = __ldg(data), xbar = __ldg(data + 1), M2 = __ldg(data + 2);
        const double mu = x[0], sigma = x[1];
        const double dm = A::sub(xbar, mu);
        const double ss = A::mad(n, A::mul(dm, dm), M2);
        const double s2 = A::mul(sigma, sigma);
        if (g) {
            g[0] = A::mul(n, dm) / s2;
            g[1] = A::sub(ss / A::mul(s2, sigma), n / sigma);
        }
        const double a = A::mul(-n, A::add(0.91893853320467274178, log(sigma)));
        return A::sub(a, ss / A::mul(2.0, s2));
    }
    static __device__ __forceinline__ void metric(const double* __restrict__ data, const double (&x)[2], double (&G)[4], double (*dG)[4])
    {
        const double n = __ldg(data);
        const double sigma = x[1];
        const double s2 = A::mul(sigma, sigma);
        G[0] = n / s2; G[1] = 0.0; G[2] = 0.0; G[3] = A::mul(2.0, n) / s2;
        if (dG) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                dG[0][k] = 0.0;
                dG[1][k] = A::mul(-2.0, G[k]) / sigma;
            }
        }
    }
};

// BOX: box constraints (src/rmhmc.cpp:99-168,278-285): the chain runs in the transformed space, log pi gets the
// log-Jacobian, the metric is evaluated at inv_transform(v) and the momentum update is scaled by the diagonal J (Q9).
template <class TM, bool STRICT, int RNGM, bool BOX = false>
__global__ void __launch_bounds__(128) rmhmc_kernel(const __grid_constant__ RmhmcLaunch a)
{
    constexpr int D = TM::D;
    typedef Ar<STRICT> A;
    typedef SmallLA<D, STRICT> LA;
    __shared__ double2 rng_tab[RNGM == RNG_PHILOX ? RNG_TAB_DOUBLE2 : 1];
    if (RNGM == RNG_PHILOX) {
        build_rng_tables(rng_tab);
        __syncthreads();
    }
    const long long chain = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= a.n_chains) return;
    const unsigned gchain = (unsigned)(a.chain_offset + chain);
    const double* tape = (RNGM == RNG_TAPE) ? a.rng.tape + chain * a.rng.tape_stride : nullptr;
    long long cursor = 0;

    // D normals of draw t (D <= 4: at most two Philox blocks), then uniform #0 — same stream definition as rng.cuh
    auto draw_normals = [&](long long t, double (&z)[D], unsigned& spare48_hi, unsigned& spare48_lo) {
        if (RNGM == RNG_PHILOX) {
            unsigned r0[4], r1[4];
            philox4x32_10(0u, (unsigned)(t + 1), gchain, 0u, a.rng, r0);
            philox4x32_10(1u, (unsigned)(t + 1), gchain, 0u, a.rng, r1);
            spare48_hi = ((r0[1] & 0xfffu) << 12) | (r0[3] & 0xfffu);
            spare48_lo = ((r1[1] & 0xfffu) << 12) | (r1[3] & 0xfffu);
            BmPair b[2];
            double z0[2], z1[2];
            b[0].setup(r0, rng_tab);
            b[1].setup(r1, rng_tab);
            bm_eval<2>(b, z0, z1);
            z[0] = z0[0];
            if (D > 1) z[1] = z1[0];
            if (D > 2) z[2] = z0[1];
            if (D > 3) z[3] = z1[1];
        } else {
#pragma unroll
            for (int i = 0; i < D; ++i) z[i] = tape[cursor + i];
            cursor += D;
        }
    };
    auto uniform0 = [&](unsigned s0, unsigned s1) -> double {
        if (RNGM == RNG_PHILOX) {
            const double sd = __hiloint2double(0x43300000 | (s0 >> 8), (s0 << 24) | s1) - 4503599627370496.0;
            return fma(sd, 3.5527136788005009e-15, 1.7763568394002505e-15);
        }
        return tape[cursor++];
    };

    BoxLane<BOX ? D : 1> bx;
    if (BOX) bx.load_seq(a.lb, a.ub, D);
    auto to_x = [&](const double (&v)[D], double (&x)[D]) {
#pragma unroll
        for (int i = 0; i < D; ++i) x[i] = BOX ? bx.inv(BOX ? i : 0, v[i]) : v[i];
    };
    // log pi in the sampler's space: log pi(inv(v)) + log_jacobian(v)
    auto logp_v = [&](const double (&v)[D]) -> double {
        if (!BOX) return TM::logp(a.tdata, v, nullptr);
        double x[D];
        to_x(v, x);
        double lj = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) lj = A::add(lj, bx.logjac(BOX ? i : 0, v[i]));
        return A::add(TM::logp(a.tdata, x, nullptr), lj);
    };
    auto metric_v = [&](const double (&v)[D], double (&G)[D * D], double (*dG)[D * D]) {
        if (!BOX) { TM::metric(a.tdata, v, G, dG); return; }
        double x[D];
        to_x(v, x);
        TM::metric(a.tdata, x, G, dG);
    };

    // returns (eps * F)/2, F_i = -grad_i + 1/2 (tr(Ainv dG_i) - ((Ainv dG_i)' q).(Ainv q))   (src/rmhmc.cpp:132-146, Q16)
    auto mntm_update = [&](const double (&y)[D], const double (&q)[D], const double (&Ainv)[D * D], const double (&dG)[D][D * D],
                           double (&out)[D]) {
        double g[D], Aq[D], yx[D];
        to_x(y, yx);
        TM::logp(a.tdata, yx, g);
        LA::gemv(Ainv, q, Aq);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double Tm[D * D], Tt[D * D], tq[D];
            LA::matmul(Ainv, dG[i], Tm);
            double tr = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) tr = A::add(tr, Tm[k * D + k]);
#pragma unroll
            for (int r = 0; r < D; ++r)
#pragma unroll
                for (int c = 0; c < D; ++c) Tt[c * D + r] = Tm[r * D + c];
            LA::gemv(Tt, q, tq);
            const double dp = LA::dot(tq, Aq);
            g[i] = A::add(-g[i], A::mul(0.5, A::sub(tr, dp)));
        }
#pragma unroll
        for (int i = 0; i < D; ++i)
            out[i] = BOX ? A::mul(A::mul(bx.invjac(BOX ? i : 0, y[i]), A::mul(a.eps, g[i])), 0.5) : A::mul(A::mul(a.eps, g[i]), 0.5);
    };

    double prev[D], cur[D], p[D], z[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        prev[i] = a.x0[(a.broadcast_x0 ? 0 : chain * D) + i];
        if (BOX) prev[i] = bx.transform(BOX ? i : 0, prev[i]);   // :166-168
    }
    unsigned sp0 = 0, sp1 = 0;
    draw_normals(-1, z, sp0, sp1);   // src/rmhmc.cpp:176 (value unused, advances the stream: Q3)

    double newG[D * D], prevG[D * D], invNew[D * D], invPrev[D * D], newdG[D][D * D], prevdG[D][D * D];
    metric_v(prev, newG, newdG);              // :179
    LA::inverse(newG, invNew);                // :181
#pragma unroll
    for (int k = 0; k < D * D; ++k) { prevG[k] = newG[k]; invPrev[k] = invNew[k]; }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int k = 0; k < D * D; ++k) prevdG[i][k] = newdG[i][k];
    const double cons_term = a.cons_term;   // 0.5 * n_vals * MCMC_LOG_2PI evaluated in long double on the host (:188, Q19)
    double prev_U = A::add(A::sub(cons_term, logp_v(prev)), A::mul(0.5, LA::logdet(newG)));   // :190

    int n_acc = 0;
    const int n_total = (int)(a.n_burnin + a.n_keep);
    const int n_burnin = (int)a.n_burnin;
    double* out_row = a.draws + chain * a.n_keep * D;
    double* out_lp = a.logp ? a.logp + chain * a.n_keep : nullptr;

    for (int t = 0; t < n_total; ++t) {
        draw_normals(t, z, sp0, sp1);                 // :200
        double prev_K;
        {
            double L[D * D], tmp[D];
            LA::chol(prevG, a.chol_mode, L);          // :202 (Q8)
            LA::gemv(L, z, p);
            LA::gemv(invPrev, p, tmp);
            prev_K = A::mul(LA::dot(p, tmp), 0.5);    // :204
        }
#pragma unroll
        for (int i = 0; i < D; ++i) cur[i] = prev[i];
        for (int k = 0; k < a.n_leap; ++k) {
            double q[D], upd[D], wv[D];
#pragma unroll
            for (int i = 0; i < D; ++i) q[i] = p[i];
            for (int kk = 0; kk < a.n_fp; ++kk) {     // :213-215 (Q17: start-of-trajectory metric at every step)
                mntm_update(cur, q, invPrev, prevdG, upd);
#pragma unroll
                for (int i = 0; i < D; ++i) q[i] = A::add(p[i], upd[i]);
            }
#pragma unroll
            for (int i = 0; i < D; ++i) { p[i] = q[i]; wv[i] = cur[i]; }
            for (int kk = 0; kk < a.n_fp; ++kk) {     // :224-228
                double Gw[D * D], sumM[D * D], tv[D];
                metric_v(wv, Gw, nullptr);
                LA::inverse(Gw, invNew);
#pragma unroll
                for (int m = 0; m < D * D; ++m) sumM[m] = A::add(invPrev[m], invNew[m]);
                LA::gemv(sumM, p, tv, A::mul(0.5, a.eps));
#pragma unroll
                for (int i = 0; i < D; ++i) wv[i] = A::add(cur[i], tv[i]);
            }
#pragma unroll
            for (int i = 0; i < D; ++i) cur[i] = wv[i];
            metric_v(cur, newG, newdG);               // :232
            LA::inverse(newG, invNew);                // :233
            mntm_update(cur, p, invNew, newdG, upd);  // :237
#pragma unroll
            for (int i = 0; i < D; ++i) p[i] = A::add(p[i], upd[i]);
        }
        double prop_U = A::add(A::sub(cons_term, logp_v(cur)), A::mul(0.5, LA::logdet(newG)));   // :240
        if (!isfinite(prop_U)) prop_U = CUDART_INF;
        double tmp[D];
        LA::gemv(invNew, p, tmp);
        const double prop_K = A::mul(LA::dot(p, tmp), 0.5);   // :246
        const double comp = fmin(0.01, A::add(-A::add(prop_U, prop_K), A::add(prev_U, prev_K)));   // :250
        const double u = uniform0(sp0, sp1);
        const bool acc = u < exp(comp);
        if (acc) {   // :254-261
            prev_U = prop_U;
#pragma unroll
            for (int i = 0; i < D; ++i) prev[i] = cur[i];
#pragma unroll
            for (int k = 0; k < D * D; ++k) { prevG[k] = newG[k]; invPrev[k] = invNew[k]; }
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int k = 0; k < D * D; ++k) prevdG[i][k] = newdG[i][k];
        }
        if (t >= n_burnin) {
#pragma unroll
            for (int i = 0; i < D; ++i) out_row[i] = BOX ? bx.inv(BOX ? i : 0, prev[i]) : prev[i];   // :278-285
            out_row += D;
            if (out_lp) *out_lp++ = -(prev_U - cons_term);
            n_acc += acc ? 1 : 0;
        }
    }
    if (a.n_accept) a.n_accept[chain] = n_acc;
}

template <bool STRICT, int RNGM> static int launch_nm(const RmhmcLaunch& a)
{
    const int threads = 128;
    const long long blocks = (a.n_chains + threads - 1) / threads;
    if (a.lb != nullptr)
        rmhmc_kernel<NormalModelRM<STRICT>, STRICT, RNGM, true><<<(unsigned)blocks, threads, 0, a.stream>>>(a);
    else
        rmhmc_kernel<NormalModelRM<STRICT>, STRICT, RNGM><<<(unsigned)blocks, threads, 0, a.stream>>>(a);
    MCMCB200_CUDA_TRY(cudaGetLastError());
    return MCMCB200_OK;
}

int launch_rmhmc(const RmhmcLaunch& a)
{
    if (a.target_id != MCMCB200_TARGET_NORMAL_MODEL) {
        set_error("rmhmc: target %d has no registered metric functor (only normal_model carries one)", a.target_id);
        return MCMCB200_ERR_UNSUPPORTED;
    }
    if (a.rng.mode == RNG_PHILOX) return a.strict ? launch_nm<true, RNG_PHILOX>(a) : launch_nm<false, RNG_PHILOX>(a);
    return a.strict ? launch_nm<true, RNG_TAPE>(a) : launch_nm<false, RNG_TAPE>(a);
}

}  // namespace mcmcb200
