// Registered __device__ log-density functors.
//
// These replace the reference's host callback
//   std::function<fp_t(const ColVec_t& vals_inp, ColVec_t* grad_out, void* target_data)>
// (include/mcmc/hmc.hpp:46, contract shown by examples/eigen/hmc_normal.cpp:44-76: grad_out
// may be null = "value only"; the return value is log pi).  A functor is warp-cooperative:
// x and grad are lane-striped (warp.cuh), `data` is the target's blob in global memory, and
//   eval<EPL, STRICT, WANT_VALUE, WANT_GRAD, REDUCE>(data, ctx, x, g)
// returns log pi(x) (warp-uniform) when WANT_VALUE, writes d log pi/dx into g when
// WANT_GRAD.  With REDUCE = false the return value is this lane's partial sum (the warp
// total is log pi), so a caller can fold it into one butterfly with its other terms.
// One fused evaluation may serve several reference calls.
// To add a target WITHOUT rebuilding the library: include/mcmc_b200_device.cuh (a user's .cu defines the struct and
// registers it at load time; examples/user_target/).  To add a built-in: write a struct with the same eval<> signature,
// give it an id in include/mcmc_b200.h and add it to MCMCB200_FOREACH_TARGET below.
#pragma once

#include "warp.cuh"

#include <type_traits>

namespace mcmcb200
{

// A functor whose n_dim is bounded may say so (static constexpr int max_epl = 2, 4, 8 or 16: n_dim <= 32 * max_epl); larger
// tiles are then not instantiated for it (the 2-parameter Normal model would otherwise compile every tile width for nothing).
template <class T, class = void> struct target_max_epl { static constexpr int value = 16; };
template <class T> struct target_max_epl<T, std::void_t<decltype(T::max_epl)>> { static constexpr int value = T::max_epl; };

// log pi = -1/2 |x|^2                       (SURVEY §8d C1/C2 target)
struct IsoGauss {
    static constexpr bool needs_scratch = false;
    static constexpr bool dense_matrix = false;   // data starts with a d x d matrix applied through dense_matvec
    static constexpr bool separable = true;         // log pi is a sum over elements: a chain may be split across warps
    static constexpr bool per_element_data = false;
    template <int EPL, bool STRICT, bool WANT_VALUE, bool WANT_GRAD, bool REDUCE = true, class Ctx = WarpCtx>
    static __device__ __forceinline__ double eval(const double*, const Ctx&, const double (&x)[EPL], double (&g)[EPL])
    {
        if (WANT_GRAD) {
#pragma unroll
            for (int k = 0; k < EPL; ++k) g[k] = -x[k];
        }
        if (WANT_VALUE) {
            const double s = REDUCE ? warp_dot<EPL, STRICT>(x, x) : lane_dot<EPL, STRICT>(x, x);
            return -Ar<STRICT>::mul(0.5, s);
        }
        return 0.0;
    }
};

// log pi = -1/2 sum_i w_i x_i^2, data = w[d]
struct DiagGauss {
    static constexpr bool needs_scratch = false;
    static constexpr bool dense_matrix = false;   // data starts with a d x d matrix applied through dense_matvec
    static constexpr bool separable = true;
    static constexpr bool per_element_data = true;  // data[j] belongs to element j
    template <int EPL, bool STRICT, bool WANT_VALUE, bool WANT_GRAD, bool REDUCE = true, class Ctx = WarpCtx>
    static __device__ __forceinline__ double eval(const double* __restrict__ data, const Ctx& w, const double (&x)[EPL],
                                                  double (&g)[EPL])
    {
        double t[EPL];
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
            const int j = elem_index(w.lane, k);
            const double wj = (j < w.d) ? __ldg(data + j) : 0.0;
            t[k] = Ar<STRICT>::mul(wj, x[k]);
        }
        if (WANT_GRAD) {
#pragma unroll
            for (int k = 0; k < EPL; ++k) g[k] = -t[k];
        }
        if (WANT_VALUE) {
            const double s = REDUCE ? warp_dot<EPL, STRICT>(t, x) : lane_dot<EPL, STRICT>(t, x);
            return -Ar<STRICT>::mul(0.5, s);
        }
        return 0.0;
    }
};

// log pi = -1/2 x' P x, data = P[d*d] symmetric          (SURVEY §8d C4 target; functor data = Sigma^-1)
struct DenseGauss {
    static constexpr bool needs_scratch = true;
    static constexpr bool dense_matrix = true;   // data starts with a d x d matrix applied through dense_matvec
    static constexpr bool separable = false;
    static constexpr bool per_element_data = false;
    template <int EPL, bool STRICT, bool WANT_VALUE, bool WANT_GRAD, bool REDUCE = true, class Ctx = WarpCtx>
    static __device__ __forceinline__ double eval(const double* __restrict__ data, const Ctx& w, const double (&x)[EPL],
                                                  double (&g)[EPL])
    {
        double y[EPL];
        dense_matvec<EPL, STRICT>(data, w, x, y);   // per warp, or CTA-cooperative when w.coop_nw > 0 (NUTS, many chains)
        if (WANT_GRAD) {
#pragma unroll
            for (int k = 0; k < EPL; ++k) g[k] = -y[k];
        }
        if (WANT_VALUE) {
            const double s = REDUCE ? warp_dot<EPL, STRICT>(x, y) : lane_dot<EPL, STRICT>(x, y);
            return -Ar<STRICT>::mul(0.5, s);
        }
        return 0.0;
    }
};

// Bayesian linear regression on sufficient statistics: log pi = -1/2 t'At + b't, data = A[d*d] (sym), b[d]
// (SURVEY §8d C3 target; grad = b - A t)
struct LinReg {
    static constexpr bool needs_scratch = true;
    static constexpr bool dense_matrix = true;   // data starts with a d x d matrix applied through dense_matvec
    static constexpr bool separable = false;
    static constexpr bool per_element_data = false;
    template <int EPL, bool STRICT, bool WANT_VALUE, bool WANT_GRAD, bool REDUCE = true, class Ctx = WarpCtx>
    static __device__ __forceinline__ double eval(const double* __restrict__ data, const Ctx& w, const double (&x)[EPL],
                                                  double (&g)[EPL])
    {
        double y[EPL], b[EPL];
        dense_matvec<EPL, STRICT>(data, w, x, y);   // per warp, or CTA-cooperative when w.coop_nw > 0 (NUTS, many chains)
        const double* __restrict__ bp = data + (size_t)w.d * (size_t)w.d;
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
            const int j = elem_index(w.lane, k);
            b[k] = (j < w.d) ? __ldg(bp + j) : 0.0;
        }
        if (WANT_GRAD) {
#pragma unroll
            for (int k = 0; k < EPL; ++k) g[k] = Ar<STRICT>::sub(b[k], y[k]);
        }
        if (WANT_VALUE) {
            double t[EPL];
#pragma unroll
            for (int k = 0; k < EPL; ++k) t[k] = Ar<STRICT>::sub(b[k], Ar<STRICT>::mul(0.5, y[k]));
            return REDUCE ? warp_dot<EPL, STRICT>(x, t) : lane_dot<EPL, STRICT>(x, t);
        }
        return 0.0;
    }
};

// 2-parameter Normal(mu, sigma) likelihood (examples/eigen/hmc_normal.cpp:44-76) on the sufficient
// statistics data = {n, xbar, M2 = sum (x_k - xbar)^2}:  sum (x_k - mu)^2 = M2 + n (xbar - mu)^2.
struct NormalModel {
    static constexpr int max_epl = 2;   // n_dim = 2
    static constexpr bool needs_scratch = false;
    static constexpr bool dense_matrix = false;   // data starts with a d x d matrix applied through dense_matvec
    static constexpr bool separable = false;
    static constexpr bool per_element_data = false;
    template <int EPL, bool STRICT, bool WANT_VALUE, bool WANT_GRAD, bool REDUCE = true, class Ctx = WarpCtx>
    static __device__ __forceinline__ double eval(const double* __restrict__ data, const Ctx& w, const double (&x)[EPL],
                                                  double (&g)[EPL])
    {
        const double n = __ldg(data), xbar = __ldg(data + 1), M2 = __ldg(data + 2);
        const double mu = __shfl_sync(FULL, x[0], 0), sigma = __shfl_sync(FULL, x[1], 0);
        const double dm = Ar<STRICT>::sub(xbar, mu);
        const double ss = Ar<STRICT>::mad(n, Ar<STRICT>::mul(dm, dm), M2);
        const double s2 = Ar<STRICT>::mul(sigma, sigma);
        if (WANT_GRAD) {
#pragma unroll
            for (int k = 0; k < EPL; ++k) g[k] = 0.0;
            if (w.lane == 0) {
                g[0] = Ar<STRICT>::mul(n, dm) / s2;
                g[1] = Ar<STRICT>::sub(ss / Ar<STRICT>::mul(s2, sigma), n / sigma);
            }
        }
        if (WANT_VALUE) {
            const double a = Ar<STRICT>::mul(-n, Ar<STRICT>::add(0.91893853320467274178, log(sigma)));
            const double v = Ar<STRICT>::sub(a, ss / Ar<STRICT>::mul(2.0, s2));
            return (REDUCE || w.lane == 0) ? v : 0.0;
        }
        return 0.0;
    }
};

// Neal's funnel (BASELINE config 5): x[0] = v ~ N(0, 3^2), x[i] | v ~ N(0, e^v), i >= 1; constants dropped:
//   log pi = -v^2/18 - (d-1) v/2 - 1/2 e^-v S,  S = sum_{i>=1} x_i^2     (host twin: oracle/host_targets.hpp TGT_FUNNEL)
struct Funnel {
    static constexpr bool needs_scratch = false;
    static constexpr bool dense_matrix = false;
    static constexpr bool separable = false;
    static constexpr bool per_element_data = false;
    template <int EPL, bool STRICT, bool WANT_VALUE, bool WANT_GRAD, bool REDUCE = true, class Ctx = WarpCtx>
    static __device__ __forceinline__ double eval(const double*, const Ctx& w, const double (&x)[EPL], double (&g)[EPL])
    {
        typedef Ar<STRICT> A;
        const double v = __shfl_sync(FULL, x[0], 0);   // element 0 lives in slot 0 of lane 0
        const double ev = exp(-v);
        double xx[EPL];
#pragma unroll
        for (int k = 0; k < EPL; ++k) xx[k] = (elem_index(w.lane, k) == 0) ? 0.0 : x[k];
        const double S = warp_dot<EPL, STRICT>(xx, xx);   // the gradient of v needs the total, so it is always reduced
        const double dm1 = (double)(w.d - 1);
        const double hes = A::mul(A::mul(0.5, ev), S);
        if (WANT_GRAD) {
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                const int j = elem_index(w.lane, k);
                g[k] = (j == 0) ? A::add(A::sub((-v) / 9.0, dm1 / 2.0), hes) : ((j < w.d) ? -A::mul(ev, x[k]) : 0.0);
            }
        }
        if (WANT_VALUE) {
            const double val = A::sub(A::sub((-A::mul(v, v)) / 18.0, A::mul(dm1, v) / 2.0), hes);
            return (REDUCE || w.lane == 0) ? val : 0.0;
        }
        return 0.0;
    }
};

}  // namespace mcmcb200

// X-macro: (enum id, functor type).  MCMCB200_FAST_BUILD (developer builds for kernel tuning, see
// mcmc_b200/build.py) keeps only the headline target so the library compiles in seconds.
// MCMCB200_TARGET_SLICE = k (build.py compiles hmc.cu / mala.cu / nuts.cu once per target, in parallel): the
// translation unit instantiates the kernels of target k only and exports launch_<sampler>_slice<k>; the
// by-target dispatch lives in dispatch.cu.
// MCMCB200_USER_TARGET_TYPE (include/mcmc_b200_device.cuh): the translation unit belongs to a USER's shared library; the
// kernels are instantiated for the user's functor only, under the id MCMCB200_TARGET_USER, and exported as
// launch_<sampler>_user_<tag> for the registration call.
#define MCMCB200_N_TARGETS 6
#define MCMCB200_CAT2(a, b) a##b
#define MCMCB200_CAT(a, b) MCMCB200_CAT2(a, b)
#if defined(MCMCB200_USER_TARGET_TYPE)
#define MCMCB200_FOREACH_TARGET(X) X(MCMCB200_TARGET_USER, MCMCB200_USER_TARGET_TYPE)
#define MCMCB200_SLICED(name) MCMCB200_CAT(MCMCB200_CAT(name, _user_), MCMCB200_USER_TARGET_TAG)
#elif defined(MCMCB200_TARGET_SLICE)
#if MCMCB200_TARGET_SLICE == 0
#define MCMCB200_FOREACH_TARGET(X) X(MCMCB200_TARGET_ISO_GAUSS, IsoGauss)
#elif MCMCB200_TARGET_SLICE == 1
#define MCMCB200_FOREACH_TARGET(X) X(MCMCB200_TARGET_DIAG_GAUSS, DiagGauss)
#elif MCMCB200_TARGET_SLICE == 2
#define MCMCB200_FOREACH_TARGET(X) X(MCMCB200_TARGET_DENSE_GAUSS, DenseGauss)
#elif MCMCB200_TARGET_SLICE == 3
#define MCMCB200_FOREACH_TARGET(X) X(MCMCB200_TARGET_LINREG, LinReg)
#elif MCMCB200_TARGET_SLICE == 4
#define MCMCB200_FOREACH_TARGET(X) X(MCMCB200_TARGET_NORMAL_MODEL, NormalModel)
#elif MCMCB200_TARGET_SLICE == 5
#define MCMCB200_FOREACH_TARGET(X) X(MCMCB200_TARGET_FUNNEL, Funnel)
#else
#error "MCMCB200_TARGET_SLICE out of range"
#endif
#define MCMCB200_SLICED(name) MCMCB200_CAT(MCMCB200_CAT(name, _slice), MCMCB200_TARGET_SLICE)
#elif defined(MCMCB200_FAST_BUILD)
#define MCMCB200_FOREACH_TARGET(X) X(MCMCB200_TARGET_ISO_GAUSS, IsoGauss)
#define MCMCB200_SLICED(name) name
#else
#define MCMCB200_FOREACH_TARGET(X)          \
    X(MCMCB200_TARGET_ISO_GAUSS, IsoGauss)   \
    X(MCMCB200_TARGET_DIAG_GAUSS, DiagGauss) \
    X(MCMCB200_TARGET_DENSE_GAUSS, DenseGauss) \
    X(MCMCB200_TARGET_LINREG, LinReg)        \
    X(MCMCB200_TARGET_NORMAL_MODEL, NormalModel) \
    X(MCMCB200_TARGET_FUNNEL, Funnel)
#define MCMCB200_SLICED(name) name
#endif
