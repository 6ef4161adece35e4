// Scalar part of the registered RM-HMC metrics, shared by the warp-per-chain kernel (rmhmc_general.cu) and the CTA-per-chain
// kernel (rmhmc_cta.cu).
#pragma once

#include "warp.cuh"

namespace mcmcb200
{

// SoftAbs metric of Neal's funnel, alpha = 1e6 (BASELINE config 5): G = Q f(Lambda) Q' for the Hessian of log pi,
// f(l) = l coth(alpha l).  The Hessian is an arrow matrix, so the spectrum is closed-form (no eigensolver): see
// oracle/host_targets.hpp metric_funnel_softabs for the derivation; this functor repeats the same operations in the
// same order with forward-mode dual numbers (value, d/dv, d/dS), S = sum_{i>=1} x_i^2.
template <bool STRICT> struct Dual2Ops {
    typedef Ar<STRICT> A;
    struct D2 { double v, dv, ds; };
    static __device__ __forceinline__ D2 mk(double v, double dv = 0.0, double ds = 0.0) { D2 r = {v, dv, ds}; return r; }
    static __device__ __forceinline__ D2 add(D2 a, D2 b) { return mk(A::add(a.v, b.v), A::add(a.dv, b.dv), A::add(a.ds, b.ds)); }
    static __device__ __forceinline__ D2 sub(D2 a, D2 b) { return mk(A::sub(a.v, b.v), A::sub(a.dv, b.dv), A::sub(a.ds, b.ds)); }
    static __device__ __forceinline__ D2 mul(D2 a, D2 b)
    {
        return mk(A::mul(a.v, b.v), A::add(A::mul(a.dv, b.v), A::mul(a.v, b.dv)), A::add(A::mul(a.ds, b.v), A::mul(a.v, b.ds)));
    }
    static __device__ __forceinline__ D2 scale(D2 a, double c) { return mk(A::mul(a.v, c), A::mul(a.dv, c), A::mul(a.ds, c)); }
    static __device__ __forceinline__ D2 div(D2 a, D2 b)
    {
        const double q = a.v / b.v;
        return mk(q, A::sub(a.dv, A::mul(q, b.dv)) / b.v, A::sub(a.ds, A::mul(q, b.ds)) / b.v);
    }
    static __device__ __forceinline__ D2 sqrt_(D2 a)
    {
        const double r = sqrt(a.v);
        return mk(r, a.dv / A::mul(2.0, r), a.ds / A::mul(2.0, r));
    }
    static __device__ __forceinline__ D2 softabs(D2 l, double alpha)
    {
        const double z = A::mul(alpha, l.v);
        double f, fp;
        if (fabs(z) < 1e-4) {
            f = A::add(1.0 / alpha, A::mul(z, l.v) / 3.0);
            fp = A::mul(2.0, z) / 3.0;
        } else if (fabs(z) > 300.0) {
            f = fabs(l.v);
            fp = (l.v > 0.0) ? 1.0 : -1.0;
        } else {
            const double ct = 1.0 / tanh(z), sh = sinh(z);
            f = A::mul(l.v, ct);
            fp = A::sub(ct, z / A::mul(sh, sh));
        }
        return mk(f, A::mul(fp, l.dv), A::mul(fp, l.ds));
    }
};
// the four dual numbers (value, d/dv, d/dS) every entry of G and of its derivative cube is built from
template <bool STRICT> struct FunnelSoftabsScalars {
    typename Dual2Ops<STRICT>::D2 g11, w, P, fa;
    // xs: the position (every thread may read any element); S = sum_{i>=1} x_i^2 accumulated in index order
    __device__ __forceinline__ void compute(const double* xs, int d)
    {
        typedef Ar<STRICT> A;
        typedef Dual2Ops<STRICT> O;
        typedef typename O::D2 D2;
        const double alpha = 1e6;
        double S = 0.0;
        for (int i = 1; i < d; ++i) S = A::add(S, A::mul(xs[i], xs[i]));
        const double e = exp(-xs[0]);
        const D2 ev = O::mk(e, -e, 0.0), Sd = O::mk(S, 0.0, 1.0);
        const D2 h = O::sub(O::mk(-1.0 / 9.0), O::scale(O::mul(ev, Sd), 0.5));
        const D2 a = O::mk(-ev.v, -ev.dv, 0.0);
        const D2 beta2 = O::mul(O::mul(ev, ev), Sd);
        const D2 delta = O::scale(O::sub(h, a), 0.5), mu = O::scale(O::add(h, a), 0.5);
        const D2 r = O::sqrt_(O::add(O::mul(delta, delta), beta2));
        const D2 f1 = O::softabs(O::add(mu, r), alpha), f2 = O::softabs(O::sub(mu, r), alpha);
        fa = O::softabs(a, alpha);
        const D2 Sig = O::scale(O::add(f1, f2), 0.5);
        const D2 Del = O::div(O::sub(f1, f2), O::scale(r, 2.0));
        const D2 Dd = O::mul(Del, delta);
        g11 = O::add(Sig, Dd);
        w = O::mul(Del, ev);
        P = O::div(O::sub(O::sub(Sig, Dd), fa), Sd);
    }
};

}  // namespace mcmcb200
