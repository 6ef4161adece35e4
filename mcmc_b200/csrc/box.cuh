// Box constraints on the device (algo_settings_t::vals_bound / lower_bounds / upper_bounds).
//
// Restates, element-wise, include/misc/determine_bounds_type.hpp:27-57 (types 1 none, 2 lower, 3 upper, 4 both),
// transform_vals.hpp:25-119 (transform / inv_transform), log_jacobian.hpp:25-58 and the diagonal of
// inv_jacobian_adjust.hpp:25-56, with eps_dbl = DBL_EPSILON (mcmc_options.hpp:103).  The samplers run in the
// transformed space v; the target is evaluated at x = inv_transform(v), the potential gets + log_jacobian(v) and the
// force is the raw gradient times the diagonal "inverse Jacobian" J(v) (SURVEY Q9 — bug-compatible, not the chain rule).
// exp/log come from the CUDA math library, so bounded runs track the CPU reference to the contract tolerance, not
// bit for bit.
#pragma once

#include "warp.cuh"

namespace mcmcb200
{

constexpr double BOX_EPS = 2.220446049250313e-16;

template <int EPL> struct BoxLane {
    double lb[EPL], ub[EPL];
    int type[EPL];

    __device__ __forceinline__ void load(const double* __restrict__ lower, const double* __restrict__ upper, int d, int lane, int elem_off = 0)
    {
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
            const int j = elem_index(lane, k);
            if (j < d) {
                lb[k] = __ldg(lower + elem_off + j);
                ub[k] = __ldg(upper + elem_off + j);
                const bool fl = isfinite(lb[k]), fu = isfinite(ub[k]);
                type[k] = (fl && fu) ? 4 : (fl ? 2 : (fu ? 3 : 1));
            } else {
                lb[k] = 0.0; ub[k] = 0.0; type[k] = 1;
            }
        }
    }
    // thread-per-chain kernels: slot k is element k
    __device__ __forceinline__ void load_seq(const double* __restrict__ lower, const double* __restrict__ upper, int n)
    {
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
            if (k < n) {
                lb[k] = __ldg(lower + k);
                ub[k] = __ldg(upper + k);
                const bool fl = isfinite(lb[k]), fu = isfinite(ub[k]);
                type[k] = (fl && fu) ? 4 : (fl ? 2 : (fu ? 3 : 1));
            } else {
                lb[k] = 0.0; ub[k] = 0.0; type[k] = 1;
            }
        }
    }
    __device__ __forceinline__ double transform(int k, double x) const
    {
        switch (type[k]) {
        case 2: return log(x - lb[k] + BOX_EPS);
        case 3: return -log(ub[k] - x + BOX_EPS);
        case 4: return log(x - lb[k] + BOX_EPS) - log(ub[k] - x + BOX_EPS);
        default: return x;
        }
    }
    __device__ __forceinline__ double inv(int k, double v) const
    {
        switch (type[k]) {
        case 2: return !isfinite(v) ? lb[k] + BOX_EPS : lb[k] + BOX_EPS + exp(v);
        case 3: return !isfinite(v) ? ub[k] - BOX_EPS : ub[k] - BOX_EPS - exp(-v);
        case 4: {
            if (!isfinite(v)) return isnan(v) ? (ub[k] - lb[k]) / 2 : (v < 0.0 ? lb[k] + BOX_EPS : ub[k] - BOX_EPS);
            const double e = exp(v);
            const double r = (lb[k] - BOX_EPS + (ub[k] + BOX_EPS) * e) / (1.0 + e);
            return isfinite(r) ? r : ub[k] - BOX_EPS;
        }
        default: return v;
        }
    }
    // term of log_jacobian (0 for unbounded elements)
    __device__ __forceinline__ double logjac(int k, double v) const
    {
        switch (type[k]) {
        case 2: return v;
        case 3: return -v;
        case 4: {
            const double e = exp(v);
            return isfinite(e) ? log(ub[k] - lb[k]) + v - 2 * log(1 + e) : log(ub[k] - lb[k]) - v;
        }
        default: return 0.0;
        }
    }
    // diagonal entry of inv_jacobian_adjust
    __device__ __forceinline__ double invjac(int k, double v) const
    {
        switch (type[k]) {
        case 2: return exp(-v);
        case 3: return exp(v);
        case 4: { const double e = exp(v); return ((e + 1) * (e + 1)) / (e * (ub[k] - lb[k])); }
        default: return 1.0;
        }
    }
};

// Target evaluation in the transformed space.  Returns log pi(inv(v)) + log_jacobian(v) (REDUCE: warp-uniform total,
// else this lane's partial) when WANT_VALUE; g = raw gradient at inv(v) and J = diagonal of inv_jacobian_adjust(v)
// when WANT_GRAD.  With BOX = false it is exactly T::eval.
template <class T, int EPL, bool STRICT, bool BOX, bool WANT_VALUE, bool WANT_GRAD, bool REDUCE, class Ctx>
__device__ __forceinline__ double box_eval(const double* __restrict__ tdata, const Ctx& w, const BoxLane<BOX ? EPL : 1>& bx,
                                           const double (&v)[EPL], double (&g)[EPL], double (&J)[EPL])
{
    if (!BOX) return T::template eval<EPL, STRICT, WANT_VALUE, WANT_GRAD, REDUCE>(tdata, w, v, g);
    double x[EPL];
#pragma unroll
    for (int k = 0; k < EPL; ++k) x[k] = bx.inv(BOX ? k : 0, v[k]);
    double val = T::template eval<EPL, STRICT, WANT_VALUE, WANT_GRAD, REDUCE>(tdata, w, x, g);
    if (WANT_GRAD) {
#pragma unroll
        for (int k = 0; k < EPL; ++k) J[k] = bx.invjac(BOX ? k : 0, v[k]);
    }
    if (WANT_VALUE) {
        double lj = 0.0;
#pragma unroll
        for (int k = 0; k < EPL; ++k) lj = Ar<STRICT>::add(lj, bx.logjac(BOX ? k : 0, v[k]));
        if (REDUCE) lj = warp_sum<STRICT>(lj);
        val = Ar<STRICT>::add(val, lj);
    }
    return val;
}

}  // namespace mcmcb200
