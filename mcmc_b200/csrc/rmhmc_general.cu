// Many-chain RM-HMC for general n_dim <= 64: one WARP per chain, dense d x d metric algebra in a per-chain scratch
// area in global memory (L1/L2-resident), vectors lane-striped like everywhere else.
//
// Replaces internal::rmhmc_impl (/root/reference/src/rmhmc.cpp:30-294) for metrics too large for rmhmc.cu's
// thread-per-chain registers (BASELINE config 5: Neal's funnel, d = 64).  It follows the oracle's restatement
// (oracle/oracle.cpp run_rmhmc / rm_mntm_update, bit-identical to the unmodified reference) operation by operation:
//   p = chol(G_prev) z ; K0 = p.(G_prev^-1 p)/2                                            src/rmhmc.cpp:200-204
//   per leapfrog step: n_fp x  q = p + (eps/2) F(x, q; G_prev^-1, dG_prev)   (Q16, Q17)    :211-217
//                      n_fp x  w = x + ((eps/2)(G_prev^-1 + G(w)^-1)) p                    :221-230
//                      G, dG at the new x ; G^-1 ; p += (eps/2) F(x, p; G^-1, dG)          :232-237
//   F_i = -grad_i + 1/2 (tr(A D_i) - ((A D_i)' q).(A q)),  A = G^-1, D_i = dG/dx_i         :132-146
//   U = d/2 log 2 pi - log pi + 1/2 logdet G ; accept iff u < exp(min(0.01, dH))           :240-253
// Linear algebra in the oracle's order (so STRICT arithmetic reproduces it bit for bit up to libm's log/exp/sqrt):
// LU inverse with partial pivoting (first maximum wins), column Cholesky with the Eigen matrixLLT storage quirk
// selectable (Q8), log-det from the Cholesky diagonal, products accumulated over the inner index in increasing order.
// STRICT arithmetic performs the reference's O(d^4) momentum update literally (d products of d x d matrices); FAST
// arithmetic contracts it to O(d^3): one d x d matrix per update, then one Frobenius product with each dG/dx_i.
//
// Scratch per chain (doubles): 9 d^2 (G_new, G_prev, G_w, inv_new, inv_prev, sumM, LU, Lchol, T) + 2 d^3 (dG_new,
// dG_prev; the metric functors only write their fixed sparsity pattern into zero-initialised cubes).
#include "engine.h"
#include "rng.cuh"
#include "targets.cuh"
#include "rmhmc_metrics.cuh"
#include <math_constants.h>

namespace mcmcb200
{

constexpr int RG_EPL = 2;       // n_dim <= 64
constexpr int RG_WARPS = 4;

// ---- registered metrics: G (d x d, column-major) and, on request, the d derivative matrices dG/dx_i ----------
// x is the lane-striped position; every lane may read any element through xs (the staged copy in shared memory).
// dG buffers are zero-initialised once by the kernel; eval() must only write entries of its own fixed pattern.
struct NormalModelMetric {   // Fisher information of Normal(mu, sigma), examples/eigen/rmhmc_normal.cpp:82-111 (d = 2)
    template <bool STRICT> static __device__ __forceinline__ void eval(const double* __restrict__ data, int d, int lane, const double* xs,
                                                                        double* G, double* dG)
    {
        typedef Ar<STRICT> A;
        if (lane == 0) {
            const double n = __ldg(data), sigma = xs[1];
            const double s2 = A::mul(sigma, sigma);
            const double g0 = n / s2, g3 = A::mul(2.0, n) / s2;
            G[0] = g0; G[1] = 0.0; G[2] = 0.0; G[3] = g3;
            if (dG) {
                dG[4] = A::mul(-2.0, g0) / sigma; dG[5] = A::mul(-2.0, 0.0) / sigma; dG[6] = A::mul(-2.0, 0.0) / sigma; dG[7] = A::mul(-2.0, g3) / sigma;
            }
        }
        (void)d;
    }
};
struct FunnelFisherMetric {   // minus the expected Hessian of Neal's funnel over x | v (oracle/host_targets.hpp metric_funnel_fisher)
    template <bool STRICT> static __device__ __forceinline__ void eval(const double* __restrict__, int d, int lane, const double* xs, double* G,
                                                                        double* dG)
    {
        const double ev = exp(-xs[0]);
        for (int k = lane; k < d * d; k += 32) G[k] = 0.0;
        __syncwarp();
        for (int i = lane; i < d; i += 32) G[(size_t)i * d + i] = (i == 0) ? 1.0 / 9.0 + (double)(d - 1) / 2.0 : ev;
        if (dG)
            for (int i = lane; i < d; i += 32)
                if (i > 0) dG[(size_t)i * d + i] = -ev;   // block 0 = dG/dv; everything else stays zero
    }
};

struct FunnelSoftabsMetric {
    template <bool STRICT> static __device__ __forceinline__ void eval(const double* __restrict__, int d, int lane, const double* xs, double* G,
                                                                        double* dG)
    {
        typedef Ar<STRICT> A;
        FunnelSoftabsScalars<STRICT> sc;
        sc.compute(xs, d);
        const auto g11 = sc.g11, w = sc.w, P = sc.P, fa = sc.fa;
        for (int j = 0; j < d; ++j)
            for (int i = lane; i < d; i += 32) {
                double g;
                if (i == 0 && j == 0) g = g11.v;
                else if (i == 0 || j == 0) g = A::mul(w.v, xs[i + j]);
                else g = A::add((i == j) ? fa.v : 0.0, A::mul(A::mul(P.v, xs[i]), xs[j]));
                G[(size_t)j * d + i] = g;
            }
        if (!dG) return;
        const size_t dd = (size_t)d * d;
        for (int k = 0; k < d; ++k) {
            const double s2 = A::mul(2.0, xs[k]);
            for (int j = 0; j < d; ++j)
                for (int i = lane; i < d; i += 32) {
                    double g;
                    if (k == 0) {
                        if (i == 0 && j == 0) g = g11.dv;
                        else if (i == 0 || j == 0) g = A::mul(w.dv, xs[i + j]);
                        else g = A::add((i == j) ? fa.dv : 0.0, A::mul(A::mul(P.dv, xs[i]), xs[j]));
                    } else {
                        if (i == 0 && j == 0) g = A::mul(g11.ds, s2);
                        else if (i == 0 || j == 0) g = A::add(A::mul(A::mul(w.ds, s2), xs[i + j]), (i + j == k) ? w.v : 0.0);
                        else g = A::add(A::mul(A::mul(A::mul(P.ds, s2), xs[i]), xs[j]),
                                        A::add((i == k) ? A::mul(P.v, xs[j]) : 0.0, (j == k) ? A::mul(P.v, xs[i]) : 0.0));
                    }
                    dG[(size_t)k * dd + (size_t)j * d + i] = g;
                }
        }
    }
};

// ---- warp-collective dense algebra on column-major d x d matrices in global scratch ---------------------------
template <bool STRICT> struct RG {
    typedef Ar<STRICT> A;
    // y = M (alpha v), v read from `vs` (shared or global, d entries), j increasing
    static __device__ __forceinline__ void gemv(const double* M, int d, int lane, const double* vs, double alpha, double (&y)[RG_EPL])
    {
        y[0] = 0.0; y[1] = 0.0;
        const int i = 2 * lane;
        for (int j = 0; j < d; ++j) {
            const double t = A::mul(alpha, vs[j]);
            const double* col = M + (size_t)j * d;
            if (i < d) y[0] = A::mad(col[i], t, y[0]);
            if (i + 1 < d) y[1] = A::mad(col[i + 1], t, y[1]);
        }
    }
    static __device__ __forceinline__ void stage(double* vs, int d, int lane, const double (&v)[RG_EPL])
    {
        __syncwarp();
        if (2 * lane < d) vs[2 * lane] = v[0];
        if (2 * lane + 1 < d) vs[2 * lane + 1] = v[1];
        __syncwarp();
    }
    // inv = A^-1 by LU with partial pivoting (oracle mat_inverse); lu: d*d scratch, piv: d ints in shared memory
    static __device__ void inverse(const double* Am, int d, int lane, double* lu, int* piv, double* inv)
    {
        for (int k = lane; k < d * d; k += 32) lu[k] = Am[k];
        for (int i = lane; i < d; i += 32) piv[i] = i;
        __syncwarp();
        for (int k = 0; k < d; ++k) {
            // pivot (oracle mat_inverse): p = k, best = |lu(k,k)|; a later row replaces it only if STRICTLY larger, so
            // p is the smallest index attaining the maximum; a NaN never wins, a NaN at (k,k) keeps p = k
            int p = k;
            const double vk = fabs(lu[(size_t)k * d + k]);
            if (vk == vk) {
                double bv = -2.0;
                int bi = 0x7fffffff;
                for (int i = 2 * lane; i < 2 * lane + 2; ++i)
                    if (i >= k && i < d) {
                        double v = fabs(lu[(size_t)k * d + i]);
                        if (!(v == v)) v = -1.0;
                        if (v > bv) { bv = v; bi = i; }
                    }
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const double ov = __shfl_xor_sync(FULL, bv, off);
                    const int oi = __shfl_xor_sync(FULL, bi, off);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                p = __shfl_sync(FULL, bi, 0);
            }
            if (p != k) {
                for (int j = lane; j < d; j += 32) {
                    const double a = lu[(size_t)j * d + k], b = lu[(size_t)j * d + p];
                    lu[(size_t)j * d + k] = b;
                    lu[(size_t)j * d + p] = a;
                }
                if (lane == 0) { const int t = piv[k]; piv[k] = piv[p]; piv[p] = t; }
            }
            __syncwarp();
            const double dd = lu[(size_t)k * d + k];
            for (int i = 2 * lane; i < 2 * lane + 2; ++i)
                if (i > k && i < d) lu[(size_t)k * d + i] = lu[(size_t)k * d + i] / dd;
            __syncwarp();
            // trailing update, four columns in flight per lane (row k of those columns is not modified in this step, and
            // every element receives exactly one update, so the grouping does not change any result)
            for (int j = k + 1; j < d; j += 4) {
                double t[4], v[4][2];
                const double lk0 = (2 * lane > k && 2 * lane < d) ? lu[(size_t)k * d + 2 * lane] : 0.0;
                const double lk1 = (2 * lane + 1 > k && 2 * lane + 1 < d) ? lu[(size_t)k * d + 2 * lane + 1] : 0.0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int jj = (j + u < d) ? j + u : d - 1;
                    t[u] = lu[(size_t)jj * d + k];
                    v[u][0] = (2 * lane < d) ? lu[(size_t)jj * d + 2 * lane] : 0.0;
                    v[u][1] = (2 * lane + 1 < d) ? lu[(size_t)jj * d + 2 * lane + 1] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (j + u < d) {
                        if (2 * lane > k && 2 * lane < d) lu[(size_t)(j + u) * d + 2 * lane] = A::sub(v[u][0], A::mul(lk0, t[u]));
                        if (2 * lane + 1 > k && 2 * lane + 1 < d) lu[(size_t)(j + u) * d + 2 * lane + 1] = A::sub(v[u][1], A::mul(lk1, t[u]));
                    }
            }
            __syncwarp();
        }
        if constexpr (!STRICT) {
            // FAST: lanes over ROWS, four right-hand sides at a time in registers, column-oriented (axpy) substitution.  The
            // forward sweep applies the same operations in the same order as the oracle; the backward sweep applies the terms
            // of a row in decreasing instead of increasing column order (rounding-level difference, FAST mode only).
            const int i0 = 2 * lane, i1 = 2 * lane + 1;
            for (int c0 = 0; c0 < d; c0 += 4) {
                double y[4][2];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    y[cc][0] = (i0 < d && piv[i0] == c0 + cc) ? 1.0 : 0.0;
                    y[cc][1] = (i1 < d && piv[i1] == c0 + cc) ? 1.0 : 0.0;
                }
                for (int j = 0; j < d; ++j) {   // L y = P e_c, unit lower triangle
                    const double l0 = (i0 > j && i0 < d) ? lu[(size_t)j * d + i0] : 0.0;
                    const double l1 = (i1 > j && i1 < d) ? lu[(size_t)j * d + i1] : 0.0;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const double yj = __shfl_sync(FULL, (j & 1) ? y[cc][1] : y[cc][0], j >> 1);
                        y[cc][0] = fma(-l0, yj, y[cc][0]);
                        y[cc][1] = fma(-l1, yj, y[cc][1]);
                    }
                }
                for (int j = d - 1; j >= 0; --j) {   // U x = y
                    const double ujj = lu[(size_t)j * d + j];
                    const double u0 = (i0 < j) ? lu[(size_t)j * d + i0] : 0.0;
                    const double u1 = (i1 < j) ? lu[(size_t)j * d + i1] : 0.0;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const double xj = __shfl_sync(FULL, (j & 1) ? y[cc][1] : y[cc][0], j >> 1) / ujj;
                        if (i0 == j) y[cc][0] = xj;
                        if (i1 == j) y[cc][1] = xj;
                        y[cc][0] = fma(-u0, xj, y[cc][0]);
                        y[cc][1] = fma(-u1, xj, y[cc][1]);
                    }
                }
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    if (c0 + cc < d) {
                        if (i0 < d) inv[(size_t)(c0 + cc) * d + i0] = y[cc][0];
                        if (i1 < d) inv[(size_t)(c0 + cc) * d + i1] = y[cc][1];
                    }
            }
            __syncwarp();
            return;
        }
        // one right-hand side (column of the identity) per lane slot, substitution in the oracle's order
        for (int cc = 0; cc < 2; ++cc) {
            const int c = 2 * lane + cc;
            if (c < d) {
                double* y = inv + (size_t)c * d;
                for (int i = 0; i < d; ++i) y[i] = (piv[i] == c) ? 1.0 : 0.0;
                for (int i = 0; i < d; ++i) {
                    double s = y[i];
                    for (int j = 0; j < i; ++j) s = A::sub(s, A::mul(lu[(size_t)j * d + i], y[j]));
                    y[i] = s;
                }
                for (int i = d - 1; i >= 0; --i) {
                    double s = y[i];
                    for (int j = i + 1; j < d; ++j) s = A::sub(s, A::mul(lu[(size_t)j * d + i], y[j]));
                    y[i] = s / lu[(size_t)i * d + i];
                }
            }
        }
        __syncwarp();
    }
    // lower Cholesky in place (oracle mat_chol); chol_mode MCMCB200_CHOL_EIGEN_LLT keeps A's strict upper triangle (Q8)
    static __device__ void chol(const double* Am, int d, int lane, int chol_mode, double* L)
    {
        for (int k = lane; k < d * d; k += 32) L[k] = Am[k];
        __syncwarp();
        for (int j = 0; j < d; ++j) {
            double s = L[(size_t)j * d + j];
            for (int k = 0; k < j; ++k) s = A::sub(s, A::mul(L[(size_t)k * d + j], L[(size_t)k * d + j]));
            const double dd = sqrt(s);
            for (int i = 2 * lane; i < 2 * lane + 2; ++i)
                if (i > j && i < d) {
                    double t = L[(size_t)j * d + i];
                    for (int k = 0; k < j; ++k) t = A::sub(t, A::mul(L[(size_t)k * d + i], L[(size_t)k * d + j]));
                    L[(size_t)j * d + i] = t / dd;
                }
            __syncwarp();
            if (lane == 0) L[(size_t)j * d + j] = dd;
            __syncwarp();
        }
        if (chol_mode == MCMCB200_CHOL_LOWER) {
            for (int k = lane; k < d * d; k += 32)
                if (k % d < k / d) L[k] = 0.0;   // row < column
            __syncwarp();
        }
    }
    // (diag(llt).log() * 2).sum()  (core/log_det.hpp:35)
    static __device__ double logdet(const double* Am, int d, int lane, double* L)
    {
        chol(Am, d, lane, MCMCB200_CHOL_EIGEN_LLT, L);
        double s = 0.0;
        for (int i = 0; i < d; ++i) s = A::add(s, A::mul(log(L[(size_t)i * d + i]), 2.0));
        return s;
    }
};

template <class T, class M, bool STRICT, int RNGM>
__global__ void __launch_bounds__(RG_WARPS * 32) rmhmc_general_kernel(const __grid_constant__ RmhmcLaunch a)
{
    extern __shared__ double smem[];
    __shared__ double2 rng_tab[RNGM == RNG_PHILOX ? RNG_TAB_DOUBLE2 : 1];
    __shared__ int pivs[RG_WARPS][64];
    typedef Ar<STRICT> A;
    typedef RG<STRICT> LA;
    if (RNGM == RNG_PHILOX) {
        build_rng_tables(rng_tab);
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long chain = (long long)blockIdx.x * RG_WARPS + warp;
    if (chain >= a.n_chains) return;
    const int d = a.d;
    const int dp = (d + 1) & ~1;
    double* vs = smem + (size_t)warp * 5 * dp;   // staged vector for products
    double* xs = vs + dp;                         // staged position for the functors
    double* tscr = xs + dp;                       // target functor scratch
    double* us = tscr + dp;                       // FAST momentum update: u = A q
    double* ups = us + dp;                        //                       u' = A' q
    const WarpCtx w{lane, d, tscr};
    int* piv = pivs[warp];

    const size_t dd2 = (size_t)d * d, dd3 = dd2 * d;
    double* W = a.work + (size_t)chain * (size_t)a.work_stride;
    double* newG = W; double* prevG = W + dd2; double* Gw = W + 2 * dd2;
    double* invNew = W + 3 * dd2; double* invPrev = W + 4 * dd2; double* sumM = W + 5 * dd2;
    double* lu = W + 6 * dd2; double* Lc = W + 7 * dd2; double* Tm = W + 8 * dd2;
    double* newdG = W + 9 * dd2; double* prevdG = newdG + dd3;
    for (size_t k = lane; k < 2 * dd3; k += 32) newdG[k] = 0.0;   // the metric functors keep the zero pattern
    __syncwarp();

    auto metric_at = [&](const double (&x)[RG_EPL], double* G, double* dG) {
        LA::stage(xs, d, lane, x);
        M::template eval<STRICT>(a.tdata, d, lane, xs, G, dG);
        __syncwarp();
    };
    // (eps * F)/2 with F_i = -grad_i + 1/2 (tr(Ainv D_i) - ((Ainv D_i)' q).(Ainv q))
    auto mntm_update = [&](const double (&y)[RG_EPL], const double (&q)[RG_EPL], const double* Ainv, const double* dG, double (&out)[RG_EPL]) {
        double g[RG_EPL], Aq[RG_EPL], tq[RG_EPL];
        T::template eval<RG_EPL, STRICT, false, true, true>(a.tdata, w, y, g);
        LA::stage(vs, d, lane, q);
        LA::gemv(Ainv, d, lane, vs, 1.0, Aq);
        if constexpr (!STRICT) {
            // FAST: tr(A D_i) - ((A D_i)' q).(A q) = sum_{a,b} D_i[b,a] (A[a,b] - u'_b u_a) with u = A q, u' = A' q: one d x d
            // matrix Mt built once per update, then one Frobenius product per i — O(d^3) instead of the reference's O(d^4),
            // and the cube is streamed exactly once, coalesced
            double up[RG_EPL];
            up[0] = 0.0; up[1] = 0.0;
            for (int b = 0; b < d; ++b) {   // u'_a = sum_b A[b,a] q_b: column a of A is contiguous
                const double qb = vs[b];
                if (2 * lane < d) up[0] = fma(Ainv[(size_t)(2 * lane) * d + b], qb, up[0]);
                if (2 * lane + 1 < d) up[1] = fma(Ainv[(size_t)(2 * lane + 1) * d + b], qb, up[1]);
            }
            LA::stage(us, d, lane, Aq);
            LA::stage(ups, d, lane, up);
            for (int aa = 0; aa < d; ++aa) {
                const double ua = us[aa];
                for (int b = lane; b < d; b += 32) Tm[(size_t)aa * d + b] = fma(-ups[b], ua, Ainv[(size_t)b * d + aa]);
            }
            __syncwarp();
            for (int i = 0; i < d; ++i) {
                const double* Di = dG + (size_t)i * dd2;
                // eight independent loads in flight per lane and array: the cube is streamed from L2/HBM and a dependent
                // two-term loop left the warp waiting one memory latency per 64 elements
                double acc[4] = {0.0, 0.0, 0.0, 0.0};
                int k = lane;
                for (; k + 7 * 32 < (int)dd2; k += 8 * 32) {
                    double dv[8], tv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { dv[u] = Di[k + 32 * u]; tv[u] = Tm[k + 32 * u]; }
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc[u & 3] = fma(dv[u], tv[u], acc[u & 3]);
                }
                for (; k < (int)dd2; k += 32) acc[0] = fma(Di[k], Tm[k], acc[0]);
                const double s0 = acc[0] + acc[1], s1 = acc[2] + acc[3];
                const double sft = warp_sum<false>(s0 + s1);
                if (lane == i / 2) g[i & 1] = fma(0.5, sft, -g[i & 1]);
            }
            __syncwarp();
        } else {
        for (int i = 0; i < d; ++i) {
            const double* Di = dG + (size_t)i * dd2;
            // Tm = Ainv * D_i, one column at a time (inner index increasing)
            for (int j = 0; j < d; ++j) {
                double col[RG_EPL];
                LA::gemv(Ainv, d, lane, Di + (size_t)j * d, 1.0, col);
                if (2 * lane < d) Tm[(size_t)j * d + 2 * lane] = col[0];
                if (2 * lane + 1 < d) Tm[(size_t)j * d + 2 * lane + 1] = col[1];
            }
            __syncwarp();
            double tr = 0.0;
            for (int k = 0; k < d; ++k) tr = A::add(tr, Tm[(size_t)k * d + k]);
            // tq_a = sum_b Tm(b, a) q_b  (the materialised transpose times q)
            tq[0] = 0.0; tq[1] = 0.0;
            for (int b = 0; b < d; ++b) {
                const double qb = vs[b];
                if (2 * lane < d) tq[0] = A::mad(Tm[(size_t)(2 * lane) * d + b], qb, tq[0]);
                if (2 * lane + 1 < d) tq[1] = A::mad(Tm[(size_t)(2 * lane + 1) * d + b], qb, tq[1]);
            }
            const double dpv = warp_dot<RG_EPL, STRICT>(tq, Aq);
            const double gi = A::mul(0.5, A::sub(tr, dpv));
            if (lane == i / 2) g[i & 1] = A::add(-g[i & 1], gi);
            __syncwarp();
        }
        }
#pragma unroll
        for (int k = 0; k < RG_EPL; ++k) out[k] = A::mul(A::mul(a.eps, g[k]), 0.5);
    };

    double prev[RG_EPL], cur[RG_EPL], p[RG_EPL], q[RG_EPL], z[RG_EPL], upd[RG_EPL], wv[RG_EPL], t[RG_EPL], gd[RG_EPL];
    load_vec<RG_EPL>(a.x0 + (a.broadcast_x0 ? 0 : chain * d), d, lane, prev);
    ChainRng<RNGM> rng;
    rng.init(a.rng, chain, a.chain_offset + chain);
    rng.template normals<RG_EPL, false>(a.rng, -1, d, lane, rng_tab, z);   // src/rmhmc.cpp:176 (value unused: Q3)

    metric_at(prev, newG, newdG);                                  // :179
    LA::inverse(newG, d, lane, lu, piv, invNew);                   // :181
    for (size_t k = lane; k < dd2; k += 32) { prevG[k] = newG[k]; invPrev[k] = invNew[k]; }
    for (size_t k = lane; k < dd3; k += 32) prevdG[k] = newdG[k];
    __syncwarp();
    const double cons_term = a.cons_term;
    double prev_U = A::add(A::sub(cons_term, T::template eval<RG_EPL, STRICT, true, false, true>(a.tdata, w, prev, gd)),
                           A::mul(0.5, LA::logdet(newG, d, lane, Lc)));   // :190

    int n_acc = 0;
    const int n_total = (int)(a.n_burnin + a.n_keep);
    const int n_burnin = (int)a.n_burnin;
    double* out_row = a.draws + chain * a.n_keep * d;
    double* out_lp = a.logp ? a.logp + chain * a.n_keep : nullptr;
    const double heps = A::mul(0.5, a.eps);

    for (int it = 0; it < n_total; ++it) {
        rng.template normals<RG_EPL, false>(a.rng, it, d, lane, rng_tab, z);   // :200
        LA::chol(prevG, d, lane, a.chol_mode, Lc);                               // :202 (Q8)
        LA::stage(vs, d, lane, z);
        LA::gemv(Lc, d, lane, vs, 1.0, p);
        LA::stage(vs, d, lane, p);
        LA::gemv(invPrev, d, lane, vs, 1.0, t);
        const double prev_K = warp_dot<RG_EPL, STRICT>(p, t) / 2.0;             // :204
#pragma unroll
        for (int k = 0; k < RG_EPL; ++k) cur[k] = prev[k];
        for (int s = 0; s < a.n_leap; ++s) {
#pragma unroll
            for (int k = 0; k < RG_EPL; ++k) q[k] = p[k];
            for (int kk = 0; kk < a.n_fp; ++kk) {                                // :213-215 (Q17)
                mntm_update(cur, q, invPrev, prevdG, upd);
#pragma unroll
                for (int k = 0; k < RG_EPL; ++k) q[k] = A::add(p[k], upd[k]);
            }
#pragma unroll
            for (int k = 0; k < RG_EPL; ++k) { p[k] = q[k]; wv[k] = cur[k]; }
            for (int kk = 0; kk < a.n_fp; ++kk) {                                // :224-228
                metric_at(wv, Gw, nullptr);
                LA::inverse(Gw, d, lane, lu, piv, invNew);
                for (size_t k = lane; k < dd2; k += 32) sumM[k] = A::add(invPrev[k], invNew[k]);
                __syncwarp();
                LA::stage(vs, d, lane, p);
                LA::gemv(sumM, d, lane, vs, heps, t);
#pragma unroll
                for (int k = 0; k < RG_EPL; ++k) wv[k] = A::add(cur[k], t[k]);
            }
#pragma unroll
            for (int k = 0; k < RG_EPL; ++k) cur[k] = wv[k];
            metric_at(cur, newG, newdG);                                         // :232
            LA::inverse(newG, d, lane, lu, piv, invNew);                         // :233
            mntm_update(cur, p, invNew, newdG, upd);                             // :237
#pragma unroll
            for (int k = 0; k < RG_EPL; ++k) p[k] = A::add(p[k], upd[k]);
        }
        double prop_U = A::add(A::sub(cons_term, T::template eval<RG_EPL, STRICT, true, false, true>(a.tdata, w, cur, gd)),
                               A::mul(0.5, LA::logdet(newG, d, lane, Lc)));     // :240
        if (!isfinite(prop_U)) prop_U = CUDART_INF;
        LA::stage(vs, d, lane, p);
        LA::gemv(invNew, d, lane, vs, 1.0, t);
        const double prop_K = warp_dot<RG_EPL, STRICT>(p, t) / 2.0;             // :246
        const double comp = fmin(0.01, A::add(-A::add(prop_U, prop_K), A::add(prev_U, prev_K)));   // :250
        const double u = rng.uniform(a.rng, it, 0);
        const bool acc = u < exp(comp);
        if (acc) {                                                               // :254-261 (buffers swap roles instead of being copied)
#pragma unroll
            for (int k = 0; k < RG_EPL; ++k) prev[k] = cur[k];
            prev_U = prop_U;
            double* tp;
            tp = prevG; prevG = newG; newG = tp;
            tp = invPrev; invPrev = invNew; invNew = tp;
            tp = prevdG; prevdG = newdG; newdG = tp;
        }
        if (it >= n_burnin) {
            store_vec<RG_EPL>(out_row, d, lane, prev);
            out_row += d;
            if (out_lp) {
                if (lane == 0) *out_lp = -A::sub(prev_U, cons_term);
                ++out_lp;
            }
            n_acc += acc ? 1 : 0;
        }
    }
    if (lane == 0 && a.n_accept) a.n_accept[chain] = n_acc;
}

#ifndef MCMCB200_USER_TARGET_TYPE
long long rmhmc_general_work_doubles(int d) { return 9ll * d * d + 2ll * d * d * d; }
#endif

template <class T, class M> static int launch_tm(const RmhmcLaunch& a)
{
    const long long blocks = (a.n_chains + RG_WARPS - 1) / RG_WARPS;
    const int dp = (a.d + 1) & ~1;
    const size_t smem = (size_t)RG_WARPS * 5 * dp * sizeof(double);
#define RG_LAUNCH(S, R)                                                                        \
    do {                                                                                       \
        rmhmc_general_kernel<T, M, S, R><<<(unsigned)blocks, RG_WARPS * 32, smem, a.stream>>>(a); \
        MCMCB200_CUDA_TRY(cudaGetLastError());                                                 \
        return MCMCB200_OK;                                                                    \
    } while (0)
    if (a.rng.mode == RNG_PHILOX) { if (a.strict) RG_LAUNCH(true, RNG_PHILOX); else RG_LAUNCH(false, RNG_PHILOX); }
    if (a.strict) RG_LAUNCH(true, RNG_TAPE); else RG_LAUNCH(false, RNG_TAPE);
#undef RG_LAUNCH
}

#ifndef MCMCB200_USER_TARGET_TYPE   // the library's own registry of metrics; a user's translation unit only needs launch_tm<T, M>
bool rmhmc_general_supported(int target_id, int metric_id, int d)
{
    if (d < 1 || d > 32 * RG_EPL) return false;
    if (target_id == MCMCB200_TARGET_NORMAL_MODEL) return d == 2 && metric_id <= 0;
    if (target_id == MCMCB200_TARGET_FUNNEL) return d >= 2 && metric_id >= 0 && metric_id <= 2;
    return false;
}

int launch_rmhmc_general(const RmhmcLaunch& a)
{
    if (a.lb != nullptr) {
        set_error("rmhmc: vals_bound is only available with the 2-parameter thread-per-chain kernel");
        return MCMCB200_ERR_UNSUPPORTED;
    }
    if (a.target_id == MCMCB200_TARGET_NORMAL_MODEL) return launch_tm<NormalModel, NormalModelMetric>(a);
    if (a.target_id == MCMCB200_TARGET_FUNNEL)
        return a.metric_id == 2 ? launch_tm<Funnel, FunnelSoftabsMetric>(a) : launch_tm<Funnel, FunnelFisherMetric>(a);
    set_error("rmhmc: target %d has no registered metric", a.target_id);
    return MCMCB200_ERR_UNSUPPORTED;
}
#endif

}  // namespace mcmcb200
