// TEST INFRASTRUCTURE ONLY (oracle/).  Host (CPU) definitions of the synthetic
// log-density targets named in SURVEY.md §8(d).  They play the role of the
// user's `target_log_kernel(vals_inp, grad_out, target_data)` callback
// (contract: /root/reference/examples/eigen/hmc_normal.cpp:44-76 — grad_out may
// be null, the return value is log pi(x)).  Both the unmodified-reference
// driver (ref_driver.cpp) and the restated oracle (oracle.cpp) call these, so
// the two CPU paths see bit-identical callbacks.
//
// `sum_mode` selects the order of the d-term reductions:
//   SUM_SEQ  : plain index order (what the stand-in Eigen does),
//   SUM_WARP : the order the CUDA kernels use — element j lives on lane
//              (j % 64) / 2 of a 32-lane warp (128-bit lane-pair striping), each
//              lane adds its elements in increasing j, then a 5-stage xor
//              butterfly (offsets 16,8,4,2,1) combines the lanes.
// Element-wise results never depend on sum_mode.
#ifndef MCMC_B200_ORACLE_HOST_TARGETS_HPP
#define MCMC_B200_ORACLE_HOST_TARGETS_HPP

#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>

namespace otgt
{

enum { SUM_SEQ = 0, SUM_WARP = 1 };

enum {
    TGT_ISO_GAUSS    = 0,  // log pi = -1/2 |x|^2                              (C1, C2)
    TGT_DIAG_GAUSS   = 1,  // log pi = -1/2 sum_i w_i x_i^2, data = w[d]
    TGT_DENSE_GAUSS  = 2,  // log pi = -1/2 x' P x, data = P[d*d] row-major     (C4)
    TGT_LINREG       = 3,  // log pi = -1/2 t' A t + b' t, data = A[d*d], b[d]  (C3)
    TGT_NORMAL_MODEL = 4,  // 2-parameter Normal(mu, sigma) likelihood on sufficient statistics
                           // data = {n, xbar, M2 = sum (x - xbar)^2}  (examples/eigen/*_normal.cpp)
    TGT_FUNNEL       = 5   // Neal's funnel (C5): x[0] = v ~ N(0, 3^2), x[i] | v ~ N(0, e^v), i >= 1; no data
                           // log pi = -v^2/18 - (d-1) v/2 - 1/2 e^-v sum_{i>=1} x_i^2   (constants dropped)
};

inline int lane_of(int j) { return (j % 64) / 2; }

// sum of t[0..n) in the requested order
inline double reduce_sum(const double* t, int n, int sum_mode)
{
    if (sum_mode == SUM_SEQ) {
        double s = 0.0;
        for (int j = 0; j < n; ++j) s = s + t[j];
        return s;
    }
    double lane[32];
    bool has[32];
    for (int l = 0; l < 32; ++l) { lane[l] = 0.0; has[l] = false; }
    for (int j = 0; j < n; ++j) {
        const int l = lane_of(j);
        // first term initialises the lane partial (the kernel starts from 0.0 + t,
        // which is exact, so both conventions agree bit-for-bit)
        lane[l] = has[l] ? lane[l] + t[j] : 0.0 + t[j];
        has[l] = true;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        double nxt[32];
        for (int l = 0; l < 32; ++l) nxt[l] = lane[l] + lane[l ^ off];
        for (int l = 0; l < 32; ++l) lane[l] = nxt[l];
    }
    return lane[0];
}

inline double dot(const double* a, const double* b, int n, int sum_mode)
{
    std::vector<double> t(static_cast<size_t>(n));
    for (int j = 0; j < n; ++j) t[size_t(j)] = a[j] * b[j];
    return reduce_sum(t.data(), n, sum_mode);
}

// y = A x for row-major A (n x n); each y_i accumulates in increasing j
inline void gemv_rowmajor(const double* A, const double* x, double* y, int n)
{
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        const double* row = A + size_t(i) * size_t(n);
        for (int j = 0; j < n; ++j) s = s + row[j] * x[j];
        y[i] = s;
    }
}

// log pi(x); if grad != nullptr also d log pi / dx.
inline double value_and_grad(int target_id, const double* data, const double* x, double* grad, int d, int sum_mode)
{
    std::vector<double> t(static_cast<size_t>(d));
    switch (target_id) {
    case TGT_ISO_GAUSS: {
        for (int j = 0; j < d; ++j) t[size_t(j)] = x[j] * x[j];
        const double s = reduce_sum(t.data(), d, sum_mode);
        if (grad)
            for (int j = 0; j < d; ++j) grad[j] = -x[j];
        return -(0.5 * s);
    }
    case TGT_DIAG_GAUSS: {
        for (int j = 0; j < d; ++j) t[size_t(j)] = (data[j] * x[j]) * x[j];
        const double s = reduce_sum(t.data(), d, sum_mode);
        if (grad)
            for (int j = 0; j < d; ++j) grad[j] = -(data[j] * x[j]);
        return -(0.5 * s);
    }
    case TGT_DENSE_GAUSS: {
        std::vector<double> y(static_cast<size_t>(d));
        gemv_rowmajor(data, x, y.data(), d);
        for (int j = 0; j < d; ++j) t[size_t(j)] = x[j] * y[size_t(j)];
        const double s = reduce_sum(t.data(), d, sum_mode);
        if (grad)
            for (int j = 0; j < d; ++j) grad[j] = -y[size_t(j)];
        return -(0.5 * s);
    }
    case TGT_LINREG: {
        const double* A = data;
        const double* b = data + size_t(d) * size_t(d);
        std::vector<double> y(static_cast<size_t>(d));
        gemv_rowmajor(A, x, y.data(), d);
        // log pi = sum_j x_j (b_j - y_j / 2)
        for (int j = 0; j < d; ++j) t[size_t(j)] = x[j] * (b[j] - 0.5 * y[size_t(j)]);
        const double s = reduce_sum(t.data(), d, sum_mode);
        if (grad)
            for (int j = 0; j < d; ++j) grad[j] = b[j] - y[size_t(j)];
        return s;
    }
    case TGT_NORMAL_MODEL: {
        const double n = data[0], xbar = data[1], M2 = data[2];
        const double mu = x[0], sigma = x[1];
        const double dm = xbar - mu;
        const double ss = M2 + n * (dm * dm);  // sum (x_k - mu)^2
        const double s2 = sigma * sigma;
        const double ret = -n * (0.91893853320467274178 + std::log(sigma)) - ss / (2.0 * s2);
        if (grad) {
            grad[0] = (n * dm) / s2;
            grad[1] = ss / (s2 * sigma) - n / sigma;
        }
        return ret;
    }
    case TGT_FUNNEL: {
        const double v = x[0];
        const double ev = std::exp(-v);
        t[0] = 0.0;
        for (int j = 1; j < d; ++j) t[size_t(j)] = x[j] * x[j];
        const double S = reduce_sum(t.data(), d, sum_mode);
        const double dm1 = double(d - 1);
        if (grad) {
            grad[0] = ((-v) / 9.0 - dm1 / 2.0) + (0.5 * ev) * S;
            for (int j = 1; j < d; ++j) grad[j] = -(ev * x[j]);
        }
        return ((-(v * v)) / 18.0 - (dm1 * v) / 2.0) - (0.5 * ev) * S;
    }
    default:
        return std::nan("");
    }
}

// Fisher-information metric for TGT_NORMAL_MODEL (examples/eigen/rmhmc_normal.cpp:82-111):
// G = diag(n/sigma^2, 2n/sigma^2); dG/dmu = 0; dG/dsigma = -2 G / sigma.
// G: d*d column-major; dG: d blocks of d*d column-major (may be null).
inline void metric_normal_model(const double* data, const double* x, double* G, double* dG)
{
    const double n = data[0];
    const double sigma = x[1];
    const double s2 = sigma * sigma;
    G[0] = n / s2; G[1] = 0.0; G[2] = 0.0; G[3] = 2.0 * n / s2;
    if (dG) {
        for (int k = 0; k < 8; ++k) dG[k] = 0.0;
        for (int k = 0; k < 4; ++k) dG[4 + k] = (-2.0 * G[k]) / sigma;
    }
}

// Metric for TGT_FUNNEL, id 1 ("funnel_fisher"): minus the expected Hessian over x | v, a diagonal position-dependent
// metric:  G = diag(1/9 + (d-1)/2, e^-v, ..., e^-v);  dG/dv = diag(0, -e^-v, ..., -e^-v);  dG/dx_i = 0.
inline void metric_funnel_fisher(const double* x, int d, double* G, double* dG)
{
    const double ev = std::exp(-x[0]);
    for (size_t k = 0; k < size_t(d) * d; ++k) G[k] = 0.0;
    G[0] = 1.0 / 9.0 + double(d - 1) / 2.0;
    for (int i = 1; i < d; ++i) G[size_t(i) * d + i] = ev;
    if (dG) {
        for (size_t k = 0; k < size_t(d) * d * d; ++k) dG[k] = 0.0;
        for (int i = 1; i < d; ++i) dG[size_t(i) * d + i] = -ev;   // block 0 = dG/dv
    }
}

// ---- SoftAbs metric for TGT_FUNNEL, id 2 ("funnel_softabs", BASELINE config 5) -------------------------------------
// G = Q f(Lambda) Q' for the Hessian H = Q Lambda Q' of log pi, f(l) = l coth(alpha l), alpha = 1e6 (Betancourt 2013).
// The funnel's Hessian is an arrow matrix — H_vv = h = -1/9 - e^-v S/2, H_vi = e^-v x_i, H_ij = a delta_ij, a = -e^-v,
// S = sum x_i^2 — so its spectrum is closed-form: eigenvalue a on the complement of span{e_v, x/|x|} and the two
// eigenvalues mu +- r of [[h, beta], [beta, a]] (beta = e^-v sqrt S, mu = (h+a)/2, delta = (h-a)/2, r = sqrt(delta^2 +
// beta^2)).  With Sig = (f1+f2)/2 and Del = (f1-f2)/(2r):
//   G_vv = Sig + Del delta,   G_vi = (Del e^-v) x_i,   G_ij = f(a) delta_ij + P x_i x_j,   P = (Sig - Del delta - f(a))/S.
// The scalars are functions of (v, S); their partial derivatives come from forward-mode dual numbers (value, d/dv, d/dS),
// and dG/dx_k follows from dS/dx_k = 2 x_k.  The device functor (mcmc_b200/csrc/rmhmc_general.cu) repeats the same
// operations in the same order.
struct Dual2 {
    double v, dv, ds;
};
inline Dual2 d2(double v, double dv = 0.0, double ds = 0.0) { Dual2 r = {v, dv, ds}; return r; }
inline Dual2 d2_add(Dual2 a, Dual2 b) { return d2(a.v + b.v, a.dv + b.dv, a.ds + b.ds); }
inline Dual2 d2_sub(Dual2 a, Dual2 b) { return d2(a.v - b.v, a.dv - b.dv, a.ds - b.ds); }
inline Dual2 d2_mul(Dual2 a, Dual2 b) { return d2(a.v * b.v, a.dv * b.v + a.v * b.dv, a.ds * b.v + a.v * b.ds); }
inline Dual2 d2_scale(Dual2 a, double c) { return d2(a.v * c, a.dv * c, a.ds * c); }
inline Dual2 d2_div(Dual2 a, Dual2 b)
{
    const double q = a.v / b.v;
    return d2(q, (a.dv - q * b.dv) / b.v, (a.ds - q * b.ds) / b.v);
}
inline Dual2 d2_sqrt(Dual2 a)
{
    const double r = std::sqrt(a.v);
    return d2(r, a.dv / (2.0 * r), a.ds / (2.0 * r));
}
// f(l) = l coth(alpha l) and f'(l) = coth(alpha l) - alpha l / sinh^2(alpha l); series near 0, saturated for |alpha l| > 300
inline void softabs_f(double l, double alpha, double* f, double* fp)
{
    const double z = alpha * l;
    if (std::fabs(z) < 1e-4) {
        *f = 1.0 / alpha + (z * l) / 3.0;
        *fp = (2.0 * z) / 3.0;
    } else if (std::fabs(z) > 300.0) {
        *f = std::fabs(l);
        *fp = (l > 0.0) ? 1.0 : -1.0;
    } else {
        const double ct = 1.0 / std::tanh(z), sh = std::sinh(z);
        *f = l * ct;
        *fp = ct - z / (sh * sh);
    }
}
inline Dual2 d2_softabs(Dual2 l, double alpha)
{
    double f, fp;
    softabs_f(l.v, alpha, &f, &fp);
    return d2(f, fp * l.dv, fp * l.ds);
}
struct FunnelSoftabsScalars {
    Dual2 g11, w, fa, P;
};
inline FunnelSoftabsScalars funnel_softabs_scalars(double v, double S, double alpha)
{
    const double e = std::exp(-v);
    const Dual2 ev = d2(e, -e, 0.0), Sd = d2(S, 0.0, 1.0);
    const Dual2 h = d2_sub(d2(-1.0 / 9.0), d2_scale(d2_mul(ev, Sd), 0.5));
    const Dual2 a = d2(-ev.v, -ev.dv, 0.0);
    const Dual2 beta2 = d2_mul(d2_mul(ev, ev), Sd);
    const Dual2 delta = d2_scale(d2_sub(h, a), 0.5), mu = d2_scale(d2_add(h, a), 0.5);
    const Dual2 r = d2_sqrt(d2_add(d2_mul(delta, delta), beta2));
    const Dual2 f1 = d2_softabs(d2_add(mu, r), alpha), f2 = d2_softabs(d2_sub(mu, r), alpha), fa = d2_softabs(a, alpha);
    const Dual2 Sig = d2_scale(d2_add(f1, f2), 0.5);
    const Dual2 Del = d2_div(d2_sub(f1, f2), d2_scale(r, 2.0));
    const Dual2 Dd = d2_mul(Del, delta);
    FunnelSoftabsScalars o;
    o.g11 = d2_add(Sig, Dd);
    o.w = d2_mul(Del, ev);
    o.fa = fa;
    o.P = d2_div(d2_sub(d2_sub(Sig, Dd), fa), Sd);
    return o;
}
inline void metric_funnel_softabs(const double* x, int d, double* G, double* dG)
{
    const double alpha = 1e6;
    double S = 0.0;
    for (int i = 1; i < d; ++i) S = S + x[i] * x[i];
    const FunnelSoftabsScalars c = funnel_softabs_scalars(x[0], S, alpha);
    const size_t dd = size_t(d) * d;
    for (int j = 0; j < d; ++j)
        for (int i = 0; i < d; ++i) {
            double g;
            if (i == 0 && j == 0) g = c.g11.v;
            else if (i == 0 || j == 0) g = c.w.v * x[i + j];
            else g = ((i == j) ? c.fa.v : 0.0) + (c.P.v * x[i]) * x[j];
            G[size_t(j) * d + i] = g;
        }
    if (!dG) return;
    for (int k = 0; k < d; ++k)
        for (int j = 0; j < d; ++j)
            for (int i = 0; i < d; ++i) {
                double g;
                if (k == 0) {   // d/dv
                    if (i == 0 && j == 0) g = c.g11.dv;
                    else if (i == 0 || j == 0) g = c.w.dv * x[i + j];
                    else g = ((i == j) ? c.fa.dv : 0.0) + (c.P.dv * x[i]) * x[j];
                } else {        // d/dx_k = (d/dS) 2 x_k + explicit dependence on x_k
                    const double s2 = 2.0 * x[k];
                    if (i == 0 && j == 0) g = c.g11.ds * s2;
                    else if (i == 0 || j == 0) g = (c.w.ds * s2) * x[i + j] + ((i + j == k) ? c.w.v : 0.0);
                    else g = ((c.P.ds * s2) * x[i]) * x[j] + (((i == k) ? c.P.v * x[j] : 0.0) + ((j == k) ? c.P.v * x[i] : 0.0));
                }
                dG[size_t(k) * dd + size_t(j) * d + i] = g;
            }
}

// metric registered with a target (metric_id 0 = the target's default)
inline bool metric(int target_id, int metric_id, const double* data, const double* x, int d, double* G, double* dG)
{
    if (target_id == TGT_NORMAL_MODEL && d == 2) { metric_normal_model(data, x, G, dG); return true; }
    if (target_id == TGT_FUNNEL && (metric_id == 0 || metric_id == 1)) { metric_funnel_fisher(x, d, G, dG); return true; }
    if (target_id == TGT_FUNNEL && metric_id == 2) { metric_funnel_softabs(x, d, G, dG); return true; }
    return false;
}

}  // namespace otgt

#endif
