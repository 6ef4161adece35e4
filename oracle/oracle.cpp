// TEST INFRASTRUCTURE ONLY (oracle/) — never linked, imported or executed by the
// product path (mcmc_b200/, include/).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load liboracle.so.
//
// A plain-C++ CPU restatement of the reference's four gradient samplers (+ rwmh) with a pluggable
// RNG source and a pluggable reduction order, written from SURVEY.md Appendices
// C/D/E and checked bit-for-bit against the UNMODIFIED reference sources
// (oracle/_ref, built by oracle/Makefile) in tests/test_oracle_vs_reference.py.
// Each function cites the reference lines it follows.
//
// PARITY PINNING: the reference ships no tests, golden vectors or KATs for this
// path, so the oracle is pinned against outputs of the
// reference itself run here (oracle/_ref) and against tests/golden/*.json, which
// tests/golden/make_golden.py generated from oracle/_ref.
//
// RNG sources
//   RNG_MT     std::mt19937_64 consumed exactly like BaseMatrixOps does
//              (include/BaseMatrixOps/include/stats/rnorm.hpp:46-60,120-128: a FRESH
//               std::normal_distribution per variate; runif.hpp:46-64: nextafter(0,1)
//               then uniform_real_distribution) — SURVEY Q1, Q2.
//   RNG_TAPE   a flat per-chain stream of doubles consumed in order (what the CUDA
//              kernels read in tape mode).
//   RNG_PHILOX Philox4x32-10, key = (seed_lo, seed_hi), counter =
//              (index, draw+1, chain, stream); stream 0 = normals by Box-Muller
//              pair q -> elements (2q, 2q+1) (uniform #0 of a draw comes from the
//              unused bits of blocks 0 and 1), stream 1 = uniforms #1.. .  This is
//              the engine's own production RNG (not in the reference); the oracle
//              restates its definition (mcmc_b200/csrc/rng.cuh header)
//              independently with libm log / sin / cos.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <random>
#include <vector>

#include "host_targets.hpp"

namespace
{

typedef std::vector<double> vec;

enum { RNG_MT = 0, RNG_TAPE = 1, RNG_PHILOX = 2 };
enum { S_HMC = 0, S_MALA = 1, S_NUTS = 2, S_RMHMC = 3, S_RWMH = 4 };

// ---------------------------------------------------------------- Philox4x32-10

inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = uint64_t(M0) * c0;
        const uint64_t p1 = uint64_t(M1) * c2;
        const uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = uint32_t(p1);
        const uint32_t n2 = uint32_t(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = uint32_t(p0);
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 52 random bits k -> (k + 1/2) 2^-52 in the open interval (0,1); exact in double
inline uint64_t top52(uint32_t hi, uint32_t lo) { return ((uint64_t(hi) << 32) | lo) >> 12; }
inline double u52_open(uint64_t k) { return (double(k) + 0.5) * 2.220446049250313080847e-16; }

// sin(pi t), cos(pi t) for t in [0,2) with exact octant reduction
inline void sincospi_host(double t, double* s, double* c)
{
    // t = q/2 + r, q in {0,1,2,3}, r in [-1/4, 1/4]
    const double q = std::floor(t * 2.0 + 0.5);
    const double r = t - q * 0.5;
    const double a = 3.14159265358979323846 * r;
    const double sr = std::sin(a), cr = std::cos(a);
    switch (int(q) & 3) {
    case 0: *s = sr; *c = cr; break;
    case 1: *s = cr; *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
    }
}

// ---------------------------------------------------------------- RNG sources

struct Rng {
    int mode;
    std::mt19937_64 eng;
    const double* tape; long tape_len; long cursor;
    uint32_t k0, k1; uint32_t chain;
    double* rec; long rec_cap; long rec_n;  // optional recording of every variate returned

    void record(double v) { if (rec && rec_n < rec_cap) rec[rec_n] = v; ++rec_n; }

    // d normals for draw `draw` (draw = -1: the pre-loop draw of NUTS / RM-HMC)
    void normals(long draw, int d, double* z)
    {
        if (mode == RNG_MT) {
            for (int i = 0; i < d; ++i) {
                std::normal_distribution<double> nd(0.0, 1.0);  // fresh per variate (rnorm.hpp:57)
                z[i] = 0.0 + 1.0 * nd(eng);                     // mu_par + sigma_par*norm_dist(engine) (:59)
            }
        } else if (mode == RNG_TAPE) {
            for (int i = 0; i < d; ++i) z[i] = (cursor < tape_len) ? tape[cursor++] : std::nan("");
        } else {
            // Box-Muller pair q -> elements (2q, 2q+1); definition in mcmc_b200/csrc/rng.cuh / DESIGN.md
            for (int q = 0; 2 * q < d; ++q) {
                uint32_t r[4];
                philox4x32_10(uint32_t(q), uint32_t(draw + 1), chain, 0u, k0, k1, r);
                const double u1 = u52_open(top52(r[0], r[1]));
                const uint64_t k2 = top52(r[2], r[3]);
                const double rad = std::sqrt(-2.0 * std::log(u1));
                double s, c;
                sincospi_host((double(k2) + 0.5) * 4.4408920985006261617e-16, &s, &c);  // phi/pi = (k2+1/2) 2^-51
                z[2 * q] = rad * c;
                if (2 * q + 1 < d) z[2 * q + 1] = rad * s;
            }
        }
        for (int i = 0; i < d; ++i) record(z[i]);
    }

    // k-th uniform of draw `draw`
    double uniform(long draw, int k)
    {
        double u;
        if (mode == RNG_MT) {
            const double a_adj = std::nextafter(0.0, 1.0);  // runif.hpp:60
            std::uniform_real_distribution<double> ud(a_adj, 1.0);
            u = ud(eng);
        } else if (mode == RNG_TAPE) {
            u = (cursor < tape_len) ? tape[cursor++] : std::nan("");
        } else {
            if (k == 0) {
                // uniform #0 of a draw: the 24 unused bits of normal blocks q = 0 and q = 1
                uint32_t a[4], b[4];
                philox4x32_10(0u, uint32_t(draw + 1), chain, 0u, k0, k1, a);
                philox4x32_10(1u, uint32_t(draw + 1), chain, 0u, k0, k1, b);
                const uint64_t sbits = (uint64_t(a[1] & 0xfffu) << 36) | (uint64_t(a[3] & 0xfffu) << 24) |
                                       (uint64_t(b[1] & 0xfffu) << 12) | uint64_t(b[3] & 0xfffu);
                u = (double(sbits) + 0.5) * 3.5527136788005009294e-15;  // 2^-48
            } else {
                uint32_t r[4];
                philox4x32_10(uint32_t(k), uint32_t(draw + 1), chain, 1u, k0, k1, r);
                u = u52_open(top52(r[0], r[1]));
            }
        }
        record(u);
        return u;
    }
};

// ---------------------------------------------------------------- small dense linear algebra
// column-major d x d, mirroring the stand-in Eigen's operation order (oracle/standin/Eigen/Dense)

struct Ctx {
    int target_id; const double* tdata; int d; int sum_mode;
    bool identity;            // precond empty -> M = I (src/hmc.cpp:57)
    vec M, Minv, S;           // precond, inverse, "sqrt" factor (CHOL_LOWER semantics per chol_mode)
    int metric_id = 0;        // RM-HMC: which of the target's registered metrics (0 = default)
    bool bounded;             // algo_settings_t::vals_bound
    std::vector<int> btype;   // determine_bounds_type: 1 none, 2 lower, 3 upper, 4 both
    vec lb, ub;
    // The reference forms its diagonal operators as FULL d x d matrices and multiplies by them: inv_jacobian_adjust
    // (src/hmc.cpp:114-122, src/nuts.cpp:121-129, src/rmhmc.cpp:122-130, src/mala.cpp:111-119) and, with no precond_mat,
    // the identity mass matrix (inv_precond_matrix = eye: src/hmc.cpp:57-59,171,160,184; src/nuts.cpp:64-66,148,204;
    // nuts.ipp:51,66,84,140; bounded MALA: J * M, chol(J) * sqrtM and the proposal covariance, src/mala.cpp:111-119,155-157).  The products equal the element-wise ones — until an element of the multiplied vector is
    // non-finite, when the exact zeros of every OTHER row turn it into NaN there (0 * inf).  That is observable only for a
    // bounded chain whose trajectory diverged (inv_transform maps non-finite coordinates back to finite ones, so such a
    // proposal can be accepted).  Literal-reference mode only (oracle_cfg_t::dense_jacobian): the comparator modes and the
    // device kernels keep the element-wise products.
    bool dense_jac = false;
};

// rows of (J_dense * w) that the off-diagonal zeros of J turn into NaN: row i if any OTHER element of w is non-finite
static void dense_jacobian_poison(const Ctx& c, const double* w, double* out)
{
    if (!c.dense_jac || !c.bounded) return;
    int n_bad = 0;
    for (int j = 0; j < c.d; ++j) n_bad += std::isfinite(w[j]) ? 0 : 1;
    if (n_bad == 0) return;
    for (int i = 0; i < c.d; ++i)
        if (n_bad - (std::isfinite(w[i]) ? 0 : 1) > 0) out[i] = std::numeric_limits<double>::quiet_NaN();
}

const double EPS_DBL = std::numeric_limits<double>::epsilon();   // mcmc::eps_dbl (mcmc_options.hpp:103)

// include/misc/determine_bounds_type.hpp:27-57
void setup_bounds(Ctx& c, int vals_bound, const double* lower, const double* upper, int dense_jacobian = 0)
{
    c.bounded = vals_bound != 0;
    c.dense_jac = c.bounded && dense_jacobian != 0;
    c.btype.assign(c.d, 1);
    if (!c.bounded) return;
    c.lb.assign(lower, lower + c.d);
    c.ub.assign(upper, upper + c.d);
    for (int i = 0; i < c.d; ++i) {
        const bool fl = std::isfinite(lower[i]), fu = std::isfinite(upper[i]);
        c.btype[i] = (fl && fu) ? 4 : (fl ? 2 : (fu ? 3 : 1));
    }
}
// include/misc/transform_vals.hpp:25-59
void box_transform(const Ctx& c, const double* x, double* v)
{
    for (int i = 0; i < c.d; ++i) switch (c.btype[i]) {
        case 1: v[i] = x[i]; break;
        case 2: v[i] = std::log(x[i] - c.lb[i] + EPS_DBL); break;
        case 3: v[i] = -std::log(c.ub[i] - x[i] + EPS_DBL); break;
        default: v[i] = std::log(x[i] - c.lb[i] + EPS_DBL) - std::log(c.ub[i] - x[i] + EPS_DBL); break;
    }
}
// include/misc/transform_vals.hpp:61-119
void box_inv_transform(const Ctx& c, const double* v, double* x)
{
    for (int i = 0; i < c.d; ++i) switch (c.btype[i]) {
        case 1: x[i] = v[i]; break;
        case 2: x[i] = !std::isfinite(v[i]) ? c.lb[i] + EPS_DBL : c.lb[i] + EPS_DBL + std::exp(v[i]); break;
        case 3: x[i] = !std::isfinite(v[i]) ? c.ub[i] - EPS_DBL : c.ub[i] - EPS_DBL - std::exp(-v[i]); break;
        default:
            if (!std::isfinite(v[i])) {
                if (std::isnan(v[i])) x[i] = (c.ub[i] - c.lb[i]) / 2;
                else if (v[i] < 0.0) x[i] = c.lb[i] + EPS_DBL;
                else x[i] = c.ub[i] - EPS_DBL;
            } else {
                x[i] = (c.lb[i] - EPS_DBL + (c.ub[i] + EPS_DBL) * std::exp(v[i])) / (1.0 + std::exp(v[i]));
                if (!std::isfinite(x[i])) x[i] = c.ub[i] - EPS_DBL;
            }
            break;
    }
}
// include/misc/log_jacobian.hpp:25-58 (terms added in index order; SUM_WARP reduces them like every other dot product)
double box_log_jacobian(const Ctx& c, const double* v)
{
    vec t(c.d, 0.0);
    for (int i = 0; i < c.d; ++i) switch (c.btype[i]) {
        case 2: t[i] = v[i]; break;
        case 3: t[i] = -v[i]; break;
        case 4: {
            const double e = std::exp(v[i]);
            t[i] = std::isfinite(e) ? std::log(c.ub[i] - c.lb[i]) + v[i] - 2 * std::log(1 + e) : std::log(c.ub[i] - c.lb[i]) - v[i];
            break;
        }
        default: break;
    }
    if (c.sum_mode == otgt::SUM_SEQ) {   // ret_val starts at 0.0 and only bounded entries are added
        double s = 0.0;
        for (int i = 0; i < c.d; ++i) if (c.btype[i] != 1) s += t[i];
        return s;
    }
    return otgt::reduce_sum(t.data(), c.d, c.sum_mode);
}
// diagonal of include/misc/inv_jacobian_adjust.hpp:25-56
void box_inv_jac_diag(const Ctx& c, const double* v, double* J)
{
    for (int i = 0; i < c.d; ++i) switch (c.btype[i]) {
        case 2: J[i] = std::exp(-v[i]); break;
        case 3: J[i] = std::exp(v[i]); break;
        case 4: { const double e = std::exp(v[i]); J[i] = ((e + 1) * (e + 1)) / (e * (c.ub[i] - c.lb[i])); break; }
        default: J[i] = 1.0; break;
    }
}

void mat_inverse(const vec& A, int n, vec& inv)
{
    vec lu(A);
    std::vector<int> piv(n);
    for (int i = 0; i < n; ++i) piv[i] = i;
    auto at = [&](vec& m, int i, int j) -> double& { return m[size_t(j) * n + i]; };
    for (int k = 0; k < n; ++k) {
        int p = k; double best = std::abs(at(lu, k, k));
        for (int i = k + 1; i < n; ++i)
            if (std::abs(at(lu, i, k)) > best) { best = std::abs(at(lu, i, k)); p = i; }
        if (p != k) {
            for (int j = 0; j < n; ++j) std::swap(at(lu, k, j), at(lu, p, j));
            std::swap(piv[k], piv[p]);
        }
        const double dd = at(lu, k, k);
        for (int i = k + 1; i < n; ++i) at(lu, i, k) /= dd;
        for (int j = k + 1; j < n; ++j) {
            const double t = at(lu, k, j);
            for (int i = k + 1; i < n; ++i) at(lu, i, j) -= at(lu, i, k) * t;
        }
    }
    inv.assign(size_t(n) * n, 0.0);
    vec y(n);
    for (int c = 0; c < n; ++c) {
        for (int i = 0; i < n; ++i) y[i] = (piv[i] == c) ? 1.0 : 0.0;
        for (int i = 0; i < n; ++i) { double s = y[i]; for (int j = 0; j < i; ++j) s -= at(lu, i, j) * y[j]; y[i] = s; }
        for (int i = n - 1; i >= 0; --i) { double s = y[i]; for (int j = i + 1; j < n; ++j) s -= at(lu, i, j) * y[j]; y[i] = s / at(lu, i, i); }
        for (int i = 0; i < n; ++i) at(inv, i, c) = y[i];
    }
}

// in-place lower Cholesky; chol_mode 1 keeps A's strict upper triangle (Eigen matrixLLT storage),
// chol_mode 0 zeroes it (Armadillo chol(A,"lower"), core/cholesky.hpp:31)
void mat_chol(const vec& A, int n, int chol_mode, vec& L)
{
    L = A;
    auto at = [&](int i, int j) -> double& { return L[size_t(j) * n + i]; };
    for (int j = 0; j < n; ++j) {
        double s = at(j, j);
        for (int k = 0; k < j; ++k) s -= at(j, k) * at(j, k);
        const double dd = std::sqrt(s);
        at(j, j) = dd;
        for (int i = j + 1; i < n; ++i) {
            double t = at(i, j);
            for (int k = 0; k < j; ++k) t -= at(i, k) * at(j, k);
            at(i, j) = t / dd;
        }
    }
    if (chol_mode == 0)
        for (int j = 1; j < n; ++j)
            for (int i = 0; i < j; ++i) at(i, j) = 0.0;
}

// y_i = sum_j A_ij * (alpha * v_j), j increasing (ScaledMat * vector in the stand-in)
void gemv_scaled(const vec& A, int n, double alpha, const double* v, double* y)
{
    for (int i = 0; i < n; ++i) y[i] = 0.0;
    for (int j = 0; j < n; ++j) {
        const double t = alpha * v[j];
        const double* col = &A[size_t(j) * n];
        for (int i = 0; i < n; ++i) y[i] += col[i] * t;
    }
}
void gemv_plain(const vec& A, int n, const double* v, double* y)
{
    for (int i = 0; i < n; ++i) y[i] = 0.0;
    for (int j = 0; j < n; ++j) {
        const double t = v[j];
        const double* col = &A[size_t(j) * n];
        for (int i = 0; i < n; ++i) y[i] += col[i] * t;
    }
}
void matmul(const vec& A, const vec& B, int n, vec& C)
{
    C.assign(size_t(n) * n, 0.0);
    for (int j = 0; j < n; ++j)
        for (int l = 0; l < n; ++l) {
            const double t = B[size_t(j) * n + l];
            for (int i = 0; i < n; ++i) C[size_t(j) * n + i] += A[size_t(l) * n + i] * t;
        }
}

double logdet_llt(const vec& A, int n)
{
    vec L; mat_chol(A, n, 1, L);
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += std::log(L[size_t(i) * n + i]) * 2;  // (diag.log()*2).sum(), core/log_det.hpp:35
    return s;
}

// x = A^-1 b by Householder QR with column pivoting (stand-in's colPivHouseholderQr().solve)
void qr_solve(const vec& A, int n, const double* b, double* x)
{
    vec q(A), rhs(b, b + n);
    std::vector<int> perm(n);
    for (int j = 0; j < n; ++j) perm[j] = j;
    auto at = [&](int i, int j) -> double& { return q[size_t(j) * n + i]; };
    for (int k = 0; k < n; ++k) {
        int p = k; double best = -1.0;
        for (int j = k; j < n; ++j) {
            double s = 0.0;
            for (int i = k; i < n; ++i) s += at(i, j) * at(i, j);
            if (s > best) { best = s; p = j; }
        }
        if (p != k) { for (int i = 0; i < n; ++i) std::swap(at(i, k), at(i, p)); std::swap(perm[k], perm[p]); }
        double nrm = 0.0;
        for (int i = k; i < n; ++i) nrm += at(i, k) * at(i, k);
        nrm = std::sqrt(nrm);
        if (nrm == 0.0) continue;
        const double alpha = (at(k, k) > 0.0) ? -nrm : nrm;
        vec v(n - k);
        for (int i = k; i < n; ++i) v[i - k] = at(i, k);
        v[0] -= alpha;
        double vn2 = 0.0;
        for (size_t i = 0; i < v.size(); ++i) vn2 += v[i] * v[i];
        if (vn2 == 0.0) continue;
        for (int j = k; j < n; ++j) {
            double s = 0.0;
            for (int i = k; i < n; ++i) s += v[i - k] * at(i, j);
            const double f = 2.0 * s / vn2;
            for (int i = k; i < n; ++i) at(i, j) -= f * v[i - k];
        }
        double s = 0.0;
        for (int i = k; i < n; ++i) s += v[i - k] * rhs[i];
        const double f = 2.0 * s / vn2;
        for (int i = k; i < n; ++i) rhs[i] -= f * v[i - k];
    }
    vec z(n);
    for (int i = n - 1; i >= 0; --i) {
        double s = rhs[i];
        for (int j = i + 1; j < n; ++j) s -= at(i, j) * z[j];
        z[i] = s / at(i, i);
    }
    for (int j = 0; j < n; ++j) x[perm[j]] = z[j];
}

// ---------------------------------------------------------------- shared pieces

double logp(const Ctx& c, const double* x, double* grad)
{
    return otgt::value_and_grad(c.target_id, c.tdata, x, grad, c.d, c.sum_mode);
}

// box_log_kernel (src/hmc.cpp:84-95): value of the transformed density at v
double box_logp(const Ctx& c, const double* v)
{
    if (!c.bounded) return logp(c, v, nullptr);
    vec x(c.d);
    box_inv_transform(c, v, x.data());
    return logp(c, x.data(), nullptr) + box_log_jacobian(c, v);
}
// gradient callback of mntm_update_fn (src/hmc.cpp:99-128): raw gradient at inv_transform(v) and the diagonal J (Q9)
void box_grad(const Ctx& c, const double* v, double* g, double* J)
{
    if (!c.bounded) { logp(c, v, g); return; }
    vec x(c.d);
    box_inv_transform(c, v, x.data());
    logp(c, x.data(), g);
    box_inv_jac_diag(c, v, J);
}

// p = sqrt_precond * z   (src/hmc.cpp:158)
void momentum_from_normals(const Ctx& c, const double* z, double* p)
{
    if (c.identity) { for (int i = 0; i < c.d; ++i) p[i] = z[i]; return; }
    gemv_plain(c.S, c.d, z, p);
}

// K = p.(M^-1 p)/2   (src/hmc.cpp:160,184)
double kinetic(const Ctx& c, const double* p)
{
    if (c.identity && !c.dense_jac) return otgt::dot(p, p, c.d, c.sum_mode) / 2.0;
    vec t(c.d);
    if (c.identity) {                             // eye * p as a dense product (src/hmc.cpp:160,184)
        for (int i = 0; i < c.d; ++i) t[i] = p[i];
        dense_jacobian_poison(c, p, t.data());
        return otgt::dot(p, t.data(), c.d, c.sum_mode) / 2.0;
    }
    gemv_plain(c.Minv, c.d, p, t.data());
    return otgt::dot(p, t.data(), c.d, c.sum_mode) / 2.0;
}

// one leapfrog step of size eps, the reference's operation order (src/hmc.cpp:164-176, src/nuts.cpp:139-154):
//   p <- p + (eps*grad(x))/2 ; x <- x + (eps*M^-1) p ; p <- p + (eps*grad(x))/2       
void leapfrog(const Ctx& c, double eps, double* x, double* p)
{
    const int d = c.d;
    vec g(d), t(d), J(d);
    box_grad(c, x, g.data(), J.data());
    if (c.bounded) for (int i = 0; i < d; ++i) p[i] = p[i] + (J[i] * (eps * g[i])) / 2.0;   // (step*J)*grad/2: J diagonal
    else for (int i = 0; i < d; ++i) p[i] = p[i] + (eps * g[i]) / 2.0;
    dense_jacobian_poison(c, g.data(), p);
    if (c.identity) {
        for (int i = 0; i < d; ++i) x[i] = x[i] + eps * p[i];
        dense_jacobian_poison(c, p, x);           // (eps * eye) * p as a dense product (src/hmc.cpp:171)
    } else {
        gemv_scaled(c.Minv, d, eps, p, t.data());
        for (int i = 0; i < d; ++i) x[i] = x[i] + t[i];
    }
    box_grad(c, x, g.data(), J.data());
    if (c.bounded) for (int i = 0; i < d; ++i) p[i] = p[i] + (J[i] * (eps * g[i])) / 2.0;
    else for (int i = 0; i < d; ++i) p[i] = p[i] + (eps * g[i]) / 2.0;
    dense_jacobian_poison(c, g.data(), p);
}

void setup_precond(Ctx& c, const double* precond, int chol_mode)
{
    const int d = c.d;
    c.identity = (precond == nullptr);
    if (!c.identity) {
        c.M.assign(precond, precond + size_t(d) * d);
        mat_inverse(c.M, d, c.Minv);
        mat_chol(c.M, d, chol_mode, c.S);
    }
}

}  // namespace

// Optional diagnostic for the parity tests: the margin u - exp(comp) of every accept test of the chain being run (negative
// = accepted).  A GPU chain that leaves the oracle's path must do so at a draw whose margin is at rounding level — anything
// else is a defect, not a flipped coin (tests/parity_util.py).
static thread_local double* g_margin_buf = nullptr;
static thread_local long g_margin_cap = 0;
static inline void note_margin(long it, double u, double comp)
{
    if (g_margin_buf && it >= 0 && it < g_margin_cap) g_margin_buf[it] = u - std::exp(comp);
}

extern "C" {

void oracle_set_margin_buffer(double* buf, long cap) { g_margin_buf = buf; g_margin_cap = cap; }


struct oracle_cfg_t {
    int sampler, target_id;
    const double* tdata;
    int d;
    long n_burnin, n_keep, n_leap_steps;
    double step_size;
    const double* precond;  // d*d column-major or null
    int chol_mode;          // 1 = Eigen matrixLLT storage (Q8), 0 = true lower factor
    long n_fp_steps;
    long n_adapt_draws;
    double target_accept_rate, gamma_val, t0_val, kappa_val;
    long max_tree_depth;
    int rng_mode;
    unsigned long seed;
    const double* tape; long tape_len;
    long chain_id;
    int sum_mode;
    int mala_exact_dmvnorm;  // 1: two dmvnorm() with LLT log-det + QR solve (mala.ipp:63-64); 0: cancelled form
    double* tape_out; long tape_out_cap;  // optional: every variate consumed, in order
    int vals_bound; const double* lower; const double* upper;   // algo_settings_t::vals_bound / lower_bounds / upper_bounds
    int metric_id;           // rmhmc: metric registered with the target (0 = default)
    int dense_jacobian;      // 1: literal src/*.cpp semantics of the dense inv_jacobian_adjust product for non-finite gradients (Ctx::dense_jac)
};

struct oracle_res_t {
    long n_accept;
    long tape_used;      // variates consumed
    double final_step;   // nuts: step size after the last draw
    long n_leapfrog;     // reference-order leapfrog count (nuts: with multiplicity)
};

static void init_rng(Rng& r, const oracle_cfg_t* cfg)
{
    r.mode = cfg->rng_mode;
    r.eng.seed(cfg->seed);
    r.tape = cfg->tape; r.tape_len = cfg->tape_len; r.cursor = 0;
    r.k0 = uint32_t(cfg->seed); r.k1 = uint32_t(uint64_t(cfg->seed) >> 32);
    r.chain = uint32_t(cfg->chain_id);
    r.rec = cfg->tape_out; r.rec_cap = cfg->tape_out_cap; r.rec_n = 0;
}

// ------------------------------------------------------------------ HMC (src/hmc.cpp:30-227, Appendix E)
static int run_hmc(const oracle_cfg_t* cfg, const double* x0, double* draws, double* logp_out, oracle_res_t* res)
{
    Ctx c; c.target_id = cfg->target_id; c.tdata = cfg->tdata; c.d = cfg->d; c.sum_mode = cfg->sum_mode;
    setup_precond(c, cfg->precond, cfg->chol_mode);
    setup_bounds(c, cfg->vals_bound, cfg->lower, cfg->upper, cfg->dense_jacobian);
    Rng rng; init_rng(rng, cfg);
    const int d = c.d;
    const long n_total = cfg->n_burnin + cfg->n_keep;
    const double eps = cfg->step_size;
    const unsigned n_leap = unsigned(cfg->n_leap_steps);  // Q22

    vec prev(x0, x0 + d), cur(d), p(d), z(d);
    if (c.bounded) box_transform(c, x0, prev.data());   // :132-136
    double prev_U = -box_logp(c, prev.data());          // :140
    long n_accept = 0, n_lf = 0;

    for (long t = 0; t < n_total; ++t) {
        rng.normals(t, d, z.data());                 // :156
        momentum_from_normals(c, z.data(), p.data()); // :158
        const double prev_K = kinetic(c, p.data());   // :160
        cur = prev;                                   // :162
        for (unsigned k = 0; k < n_leap; ++k) { leapfrog(c, eps, cur.data(), p.data()); ++n_lf; }  // :164-176
        double prop_U = -box_logp(c, cur.data());       // :178
        if (!std::isfinite(prop_U)) prop_U = std::numeric_limits<double>::infinity();  // :180-182
        const double prop_K = kinetic(c, p.data());     // :184
        const double comp = std::min(0.01, -(prop_U + prop_K) + (prev_U + prev_K));  // :188 (Q6)
        const double u = rng.uniform(t, 0);             // :189
        note_margin(t, u, comp);
        const bool acc = u < std::exp(comp);            // :191
        if (acc) { prev = cur; prev_U = prop_U; }
        if (t >= cfg->n_burnin) {
            const long row = t - cfg->n_burnin;
            for (int j = 0; j < d; ++j) draws[row * d + j] = prev[j];
            if (logp_out) logp_out[row] = -prev_U;
            if (acc) ++n_accept;                        // :198 (Q7)
        }
    }
    if (c.bounded)   // :211-218
        for (long r = 0; r < cfg->n_keep; ++r) { vec tmp(draws + r * d, draws + (r + 1) * d); box_inv_transform(c, tmp.data(), draws + r * d); }
    res->n_accept = n_accept; res->tape_used = rng.rec_n; res->final_step = eps; res->n_leapfrog = n_lf;
    return 0;
}

// diag(J) as a full d x d matrix times B, the way the reference forms it (inv_jacobian_adjust returns a Mat_t): with a
// non-finite J_ii the zeros of B (or of J's own off-diagonal) produce NaN entries — literal mode only (Ctx::dense_jac)
static void dense_diag_times(const double* J, const vec& B, int d, vec& C)
{
    vec Jd(size_t(d) * d, 0.0);
    for (int i = 0; i < d; ++i) Jd[size_t(i) * d + i] = J[i];
    matmul(Jd, B, d, C);
}

// ------------------------------------------------------------------ MALA (src/mala.cpp:30-208, mala.ipp:30-70, dmvnorm.hpp:28-54)
static void mala_mean(const Ctx& c, double eps, const double* v, double* out, double* J_out = nullptr)
{
    // v + (((eps*eps)*M)*grad)/2   (src/mala.cpp:123); bounded: v + ((((eps*eps)*J)*M)*grad)/2   (:118, M = I only here)
    const int d = c.d;
    vec g(d), t(d), J(d);
    box_grad(c, v, g.data(), J.data());
    const double e2 = eps * eps;
    if (c.bounded && c.dense_jac) {
        // literal: J as a full matrix, (J * M) * e2 as a matrix, times grad (run_mala materialises M = eye for this mode)
        vec JM;
        dense_diag_times(J.data(), c.M, d, JM);
        for (double& a : JM) a *= e2;
        gemv_plain(JM, d, g.data(), t.data());
        for (int i = 0; i < d; ++i) out[i] = v[i] + t[i] / 2.0;
        if (J_out) for (int i = 0; i < d; ++i) J_out[i] = J[i];
        return;
    }
    if (c.bounded) {
        // ((e2*J) * M) -> matrix product then "*= e2" (ScaledMat * Matrix in the stand-in): entries (J_ii * M_ij) * e2; times
        // grad (j increasing), /2, added to v.  M = I: (J_ii * 1) * e2 on the diagonal, exact zeros elsewhere.
        if (c.identity) {
            for (int i = 0; i < d; ++i) out[i] = v[i] + ((J[i] * e2) * g[i]) / 2.0;
            dense_jacobian_poison(c, g.data(), out);
        } else {
            for (int i = 0; i < d; ++i) t[i] = 0.0;
            for (int j = 0; j < d; ++j)
                for (int i = 0; i < d; ++i) t[i] += ((J[i] * c.M[size_t(j) * d + i]) * e2) * g[j];
            for (int i = 0; i < d; ++i) out[i] = v[i] + t[i] / 2.0;
        }
        if (J_out) for (int i = 0; i < d; ++i) J_out[i] = J[i];
        return;
    }
    if (c.identity) {
        for (int i = 0; i < d; ++i) out[i] = v[i] + (e2 * g[i]) / 2.0;
    } else {
        gemv_scaled(c.M, d, e2, g.data(), t.data());
        for (int i = 0; i < d; ++i) out[i] = v[i] + t[i] / 2.0;
    }
}

static double dmvnorm_log(const double* X, const double* mu, const vec& Sigma, int d, int sum_mode)
{
    const double cons_term = -0.5 * double(size_t(d)) * double(1.83787706640934548356L);  // dmvnorm.hpp:36
    vec xc(d), sol(d);
    for (int i = 0; i < d; ++i) xc[i] = X[i] - mu[i];
    qr_solve(Sigma, d, xc.data(), sol.data());
    const double quad = otgt::dot(xc.data(), sol.data(), d, sum_mode);
    return cons_term - 0.5 * (logdet_llt(Sigma, d) + quad);  // :41
}

static int run_mala(const oracle_cfg_t* cfg, const double* x0, double* draws, double* logp_out, oracle_res_t* res)
{
    Ctx c; c.target_id = cfg->target_id; c.tdata = cfg->tdata; c.d = cfg->d; c.sum_mode = cfg->sum_mode;
    setup_precond(c, cfg->precond, cfg->chol_mode);
    setup_bounds(c, cfg->vals_bound, cfg->lower, cfg->upper, cfg->dense_jacobian);
    Rng rng; init_rng(rng, cfg);
    const int d = c.d;
    const long n_total = cfg->n_burnin + cfg->n_keep;
    const double eps = cfg->step_size;
    const double e2 = eps * eps;
    if (c.dense_jac && c.identity) {   // literal mode: precond_matrix = eye as a full matrix (src/mala.cpp:57-58), chol(eye) = eye
        c.M.assign(size_t(d) * d, 0.0);
        for (int i = 0; i < d; ++i) c.M[size_t(i) * d + i] = 1.0;
        c.S = c.M; c.Minv = c.M;
        c.identity = false;
    }

    // Sigma = eps^2 M (materialised like ScaledMat -> Mat_t) and, for the cancelled form, its inverse
    vec Sigma(size_t(d) * d, 0.0), SigInv;
    if (c.identity) for (int i = 0; i < d; ++i) Sigma[size_t(i) * d + i] = 1.0 * e2;
    else for (size_t k = 0; k < Sigma.size(); ++k) Sigma[k] = c.M[k] * e2;
    if (!cfg->mala_exact_dmvnorm && !c.identity) mat_inverse(Sigma, d, SigInv);

    vec prev(x0, x0 + d), cur(d), z(d), mean_prev(d), mean_prop(d), t(d), r(d), Jprev(d), Jprop(d);
    if (c.bounded) box_transform(c, x0, prev.data());
    double prev_LP = box_logp(c, prev.data());   // src/mala.cpp:138
    long n_accept = 0;

    for (long it = 0; it < n_total; ++it) {
        rng.normals(it, d, z.data());                 // :150
        mala_mean(c, eps, prev.data(), mean_prev.data(), Jprev.data());
        if (c.bounded && c.dense_jac) {   // literal :155-157: chol of the FULL matrix J, ((eps * L) * sqrtM) as a matrix, times z
            vec Jd(size_t(d) * d, 0.0), L, P;
            for (int i = 0; i < d; ++i) Jd[size_t(i) * d + i] = Jprev[i];
            mat_chol(Jd, d, cfg->chol_mode, L);
            matmul(L, c.S, d, P);
            for (double& a : P) a *= eps;
            gemv_plain(P, d, z.data(), t.data());
            for (int i = 0; i < d; ++i) cur[i] = mean_prev[i] + t[i];
        }
        else if (c.bounded && c.identity)   // :155-157: mean + ((eps*chol(J)) * sqrtM) * z, chol of the diagonal J = sqrt(J_ii), sqrtM = I
            for (int i = 0; i < d; ++i) cur[i] = mean_prev[i] + (std::sqrt(Jprev[i]) * eps) * z[i];
        else if (c.bounded) {          // dense sqrtM: entries (sqrt(J_ii) * S_ij) * eps, times z (j increasing)
            for (int i = 0; i < d; ++i) t[i] = 0.0;
            for (int j = 0; j < d; ++j)
                for (int i = 0; i < d; ++i) t[i] += ((std::sqrt(Jprev[i]) * c.S[size_t(j) * d + i]) * eps) * z[j];
            for (int i = 0; i < d; ++i) cur[i] = mean_prev[i] + t[i];
        }
        else if (c.identity) for (int i = 0; i < d; ++i) cur[i] = mean_prev[i] + eps * z[i];   // :159
        else { gemv_scaled(c.S, d, eps, z.data(), t.data()); for (int i = 0; i < d; ++i) cur[i] = mean_prev[i] + t[i]; }
        double prop_LP = box_logp(c, cur.data());  // :162
        if (!std::isfinite(prop_LP)) prop_LP = -std::numeric_limits<double>::infinity();  // :164-166
        mala_mean(c, eps, cur.data(), mean_prop.data(), Jprop.data());   // mala.ipp:60 (mean at prev is recomputed identically, :61)
        double adj;
        if (c.bounded) {
            // both densities use Sigma = eps^2 * J(prop) * M (mala.ipp:55-56); M = I -> diagonal entries J_ii * e2
            if (cfg->mala_exact_dmvnorm || !c.identity) {
                // (with a dense M there is no cancelled form that keeps the reference's behaviour: Sigma is not symmetric, the
                //  LLT log-det of its lower triangle can be NaN — in BOTH densities — and then min(0.01, NaN) accepts; only the
                //  literal evaluation reproduces that.  The device path refuses this combination.)
                vec Sg(size_t(d) * d, 0.0);
                if (c.dense_jac) { dense_diag_times(Jprop.data(), c.M, d, Sg); for (double& a : Sg) a *= e2; }
                else if (c.identity) for (int i = 0; i < d; ++i) Sg[size_t(i) * d + i] = Jprop[i] * e2;
                else   // (J(prop) * M) * e2: NOT symmetric; LLT reads its lower triangle, the QR solve the whole matrix — as the reference does
                    for (int j = 0; j < d; ++j)
                        for (int i = 0; i < d; ++i) Sg[size_t(j) * d + i] = (Jprop[i] * c.M[size_t(j) * d + i]) * e2;
                adj = dmvnorm_log(prev.data(), mean_prop.data(), Sg, d, c.sum_mode) - dmvnorm_log(cur.data(), mean_prev.data(), Sg, d, c.sum_mode);
            } else {
                for (int i = 0; i < d; ++i) { r[i] = prev[i] - mean_prop[i]; t[i] = r[i] / (Jprop[i] * e2); }
                const double q1 = otgt::dot(r.data(), t.data(), d, c.sum_mode);
                for (int i = 0; i < d; ++i) { r[i] = cur[i] - mean_prev[i]; t[i] = r[i] / (Jprop[i] * e2); }
                const double q2 = otgt::dot(r.data(), t.data(), d, c.sum_mode);
                adj = -0.5 * (q1 - q2);
            }
        } else if (cfg->mala_exact_dmvnorm) {
            adj = dmvnorm_log(prev.data(), mean_prop.data(), Sigma, d, c.sum_mode)
                - dmvnorm_log(cur.data(), mean_prev.data(), Sigma, d, c.sum_mode);   // mala.ipp:63-64
        } else {
            // constants and log-dets cancel: adj = -1/2 (q(prev - mean_prop) - q(cur - mean_prev)), q(r) = r' Sigma^-1 r
            double q1, q2;
            for (int i = 0; i < d; ++i) r[i] = prev[i] - mean_prop[i];
            if (c.identity) q1 = otgt::dot(r.data(), r.data(), d, c.sum_mode) / e2;
            else { gemv_plain(SigInv, d, r.data(), t.data()); q1 = otgt::dot(r.data(), t.data(), d, c.sum_mode); }
            for (int i = 0; i < d; ++i) r[i] = cur[i] - mean_prev[i];
            if (c.identity) q2 = otgt::dot(r.data(), r.data(), d, c.sum_mode) / e2;
            else { gemv_plain(SigInv, d, r.data(), t.data()); q2 = otgt::dot(r.data(), t.data(), d, c.sum_mode); }
            adj = -0.5 * (q1 - q2);
        }
        const double comp = std::min(0.01, prop_LP - prev_LP + adj);   // src/mala.cpp:170
        const double u = rng.uniform(it, 0);                            // :171
        note_margin(it, u, comp);
        const bool acc = u < std::exp(comp);                            // :173
        if (acc) { prev = cur; prev_LP = prop_LP; }
        if (it >= cfg->n_burnin) {
            const long row = it - cfg->n_burnin;
            for (int j = 0; j < d; ++j) draws[row * d + j] = prev[j];
            if (logp_out) logp_out[row] = prev_LP;
            if (acc) ++n_accept;
        }
    }
    if (c.bounded)   // src/mala.cpp:192-199
        for (long rr = 0; rr < cfg->n_keep; ++rr) { vec tmp(draws + rr * d, draws + (rr + 1) * d); box_inv_transform(c, tmp.data(), draws + rr * d); }
    res->n_accept = n_accept; res->tape_used = rng.rec_n; res->final_step = eps; res->n_leapfrog = 0;
    return 0;
}

// ------------------------------------------------------------------ RWMH (src/rwmh.cpp:30-172)
// cfg->step_size carries rwmh_settings_t::par_scale, cfg->precond carries rwmh_settings_t::cov_mat.
static int run_rwmh(const oracle_cfg_t* cfg, const double* x0, double* draws, double* logp_out, oracle_res_t* res)
{
    Ctx c; c.target_id = cfg->target_id; c.tdata = cfg->tdata; c.d = cfg->d; c.sum_mode = cfg->sum_mode;
    c.identity = (cfg->precond == nullptr);   // cov_mat empty -> EYE (:57)
    setup_bounds(c, cfg->vals_bound, cfg->lower, cfg->upper);
    Rng rng; init_rng(rng, cfg);
    const int d = c.d;
    const long n_total = cfg->n_burnin + cfg->n_keep;
    const double par_scale = cfg->step_size;

    // cov_mcmc_chol = par_scale * CHOL_LOWER(cov_mcmc), materialised as a Mat_t (:116)
    vec S;
    if (!c.identity) {
        vec cov(cfg->precond, cfg->precond + size_t(d) * d);
        mat_chol(cov, d, cfg->chol_mode, S);
        for (size_t k = 0; k < S.size(); ++k) S[k] = par_scale * S[k];
    }
    vec prev(x0, x0 + d), cur(d), z(d), t(d);
    if (c.bounded) box_transform(c, x0, prev.data());    // :105-107
    double prev_LP = box_logp(c, prev.data());           // :111
    long n_accept = 0;

    for (long it = 0; it < n_total; ++it) {
        rng.normals(it, d, z.data());                     // :124
        if (c.identity) {
            // dense product with par_scale * I: every off-diagonal term is an exact zero
            for (int i = 0; i < d; ++i) cur[i] = prev[i] + par_scale * z[i];
        } else {
            gemv_plain(S, d, z.data(), t.data());
            for (int i = 0; i < d; ++i) cur[i] = prev[i] + t[i];   // :125
        }
        double prop_LP = box_logp(c, cur.data());         // :127
        if (!std::isfinite(prop_LP)) prop_LP = -std::numeric_limits<double>::infinity();   // :129-131
        const double comp = std::min(0.0, prop_LP - prev_LP);   // :135
        const double u = rng.uniform(it, 0);              // :136
        note_margin(it, u, comp);
        const bool acc = u < std::exp(comp);              // :138
        if (acc) { prev = cur; prev_LP = prop_LP; }
        if (it >= cfg->n_burnin) {
            const long row = it - cfg->n_burnin;
            if (acc) ++n_accept;                          // :142-144
            for (int j = 0; j < d; ++j) draws[row * d + j] = prev[j];   // :149-151
            if (logp_out) logp_out[row] = prev_LP;
        }
    }
    if (c.bounded)   // :158-165
        for (long r = 0; r < cfg->n_keep; ++r) { vec tmp(draws + r * d, draws + (r + 1) * d); box_inv_transform(c, tmp.data(), draws + r * d); }
    res->n_accept = n_accept; res->tape_used = rng.rec_n; res->final_step = par_scale; res->n_leapfrog = 0;
    return 0;
}

// ------------------------------------------------------------------ NUTS (src/nuts.cpp:30-332, nuts.ipp:30-241, Appendix C)
struct NutsEnv {
    const Ctx* c; Rng* rng; long draw; int* ucount; long* n_lf;
    double log_u, prev_U, prev_K;
};

// literal restatement of nuts_build_tree (nuts.ipp:97-241): same slot aliasing (Q13)
static void build_tree(NutsEnv& e, int dir, double eps, const vec& draw_vec, const vec& mntm_vec, long depth,
                       vec& new_draw, vec& pos, vec& neg, vec& mpos, vec& mneg,
                       long& n_val, long& s_val, double& alpha, long& n_alpha)
{
    const Ctx& c = *e.c;
    const int d = c.d;
    if (depth == 0) {
        new_draw = draw_vec;                       // :127
        vec new_mntm(mntm_vec);                    // :128
        leapfrog(c, dir * eps, new_draw.data(), new_mntm.data()); ++*e.n_lf;   // :132
        double prop_U = -box_logp(c, new_draw.data());                         // :134
        if (!std::isfinite(prop_U)) prop_U = std::numeric_limits<double>::infinity();
        const double prop_K = kinetic(c, new_mntm.data());                     // :140
        n_val = (e.log_u <= -prop_U - prop_K);                                 // :146
        s_val = (e.log_u < 1000.0 - prop_U - prop_K);                          // :147
        pos = new_draw; neg = new_draw; mpos = new_mntm; mneg = new_mntm;      // :151-155
        alpha = std::exp(std::min(0.0, -(prop_U + prop_K) + (e.prev_U + e.prev_K)));  // :157
        n_alpha = 1;
        return;
    }
    long n_p, s_p, n_alpha_p; double alpha_p; vec new_draw_p;
    build_tree(e, dir, eps, draw_vec, mntm_vec, depth - 1, new_draw_p, pos, neg, mpos, mneg, n_p, s_p, alpha_p, n_alpha_p);  // :166-171
    if (s_p == 1) {
        long n_pp, s_pp, n_alpha_pp; double alpha_pp; vec new_draw_pp;
        if (dir == -1) {
            vec dummy_draw(pos), dummy_mntm(mpos), start_x(neg), start_p(mneg);   // :186-189
            // callee's pos slots alias OUR neg slots (:195)
            build_tree(e, dir, eps, start_x, start_p, depth - 1, new_draw_pp, neg, dummy_draw, mneg, dummy_mntm,
                       n_pp, s_pp, alpha_pp, n_alpha_pp);
        } else {
            vec dummy_draw(neg), dummy_mntm(mneg), start_x(pos), start_p(mpos);   // :198-201
            build_tree(e, dir, eps, start_x, start_p, depth - 1, new_draw_pp, dummy_draw, pos, dummy_mntm, mpos,
                       n_pp, s_pp, alpha_pp, n_alpha_pp);                          // :203-208
        }
        const double prob = double(n_pp) / double(n_p + n_pp);   // :213
        const double zz = e.rng->uniform(e.draw, (*e.ucount)++);  // :214
        if (zz < prob) new_draw_p = new_draw_pp;                  // :216-218
        n_p += n_pp; alpha_p += alpha_pp; n_alpha_p += n_alpha_pp;  // :220-222
        vec diff(d);
        for (int i = 0; i < d; ++i) diff[i] = pos[i] - neg[i];
        const int chk1 = otgt::dot(diff.data(), mneg.data(), d, c.sum_mode) >= 0.0;   // :226
        const int chk2 = otgt::dot(diff.data(), mpos.data(), d, c.sum_mode) >= 0.0;   // :227
        s_p = s_pp * chk1 * chk2;                                                     // :229
    }
    n_val = n_p; s_val = s_p; alpha = alpha_p; n_alpha = n_alpha_p; new_draw = new_draw_p;   // :234-239
}

static int run_nuts(const oracle_cfg_t* cfg, const double* x0, double* draws, double* logp_out, oracle_res_t* res)
{
    Ctx c; c.target_id = cfg->target_id; c.tdata = cfg->tdata; c.d = cfg->d; c.sum_mode = cfg->sum_mode;
    setup_precond(c, cfg->precond, cfg->chol_mode);
    setup_bounds(c, cfg->vals_bound, cfg->lower, cfg->upper, cfg->dense_jacobian);
    Rng rng; init_rng(rng, cfg);
    const int d = c.d;
    const long n_total = cfg->n_burnin + cfg->n_keep;
    const long n_adapt = (cfg->n_adapt_draws <= n_total) ? cfg->n_adapt_draws : n_total;   // src/nuts.cpp:54
    const double delta = cfg->target_accept_rate;
    const long max_depth = cfg->max_tree_depth;
    double eps_bar = cfg->step_size;                                                        // :59
    const double gamma = cfg->gamma_val, t0 = cfg->t0_val, kappa = cfg->kappa_val;
    long n_lf = 0;
    const double inf = std::numeric_limits<double>::infinity();

    vec first(x0, x0 + d), z(d), mntm(d);
    if (c.bounded) box_transform(c, x0, first.data());  // :158-162
    rng.normals(-1, d, z.data());                       // :166 (Q3)
    momentum_from_normals(c, z.data(), mntm.data());    // :168

    // nuts_find_initial_step_size (nuts.ipp:30-93, Q14)
    double eps = 1.0;
    {
        double pU = -box_logp(c, first.data());
        if (!std::isfinite(pU)) pU = inf;
        const double pK = kinetic(c, mntm.data());
        vec nx(first), np(mntm);
        leapfrog(c, eps, nx.data(), np.data()); ++n_lf;
        double qU = -box_logp(c, nx.data());
        if (!std::isfinite(qU)) qU = inf;
        double qK = kinetic(c, np.data());
        int a_val = 2 * (-(qU + qK) + (pU + pK) > std::log(0.5)) - 1;
        bool cond = (-(qU + qK) + (pU + pK)) > -std::log(2);
        while (cond) {
            eps *= std::pow(2, a_val);
            leapfrog(c, eps, nx.data(), np.data()); ++n_lf;
            qU = -box_logp(c, nx.data());
            if (!std::isfinite(qU)) qU = inf;
            qK = kinetic(c, np.data());
            a_val = 2 * ((-(qU + qK) + (pU + pK)) > std::log(0.5)) - 1;
            cond = (-(qU + qK) + (pU + pK)) > -std::log(2);
        }
    }
    const double mu = std::log(10 * eps);   // src/nuts.cpp:174
    double h = 0.0;

    double prev_U = -box_logp(c, first.data());   // :181
    vec prev(first), new_draw(first), dpos(first), dneg(first), mpos(mntm), mneg(mntm);
    long n_accept = 0;

    for (long t = 0; t < n_total; ++t) {
        int ucount = 0;
        rng.normals(t, d, z.data());                         // :200
        momentum_from_normals(c, z.data(), mntm.data());     // :202
        const double prev_K = kinetic(c, mntm.data());       // :204
        const double log_u = std::log(rng.uniform(t, ucount++)) - prev_U - prev_K;   // :206
        new_draw = prev; dpos = prev; dneg = prev; mpos = mntm; mneg = mntm;          // :210-215
        long depth = 0, n_val = 1, s_val = 1;
        double alpha = 0.0; long n_alpha = 0; int good_round = 0;

        while (s_val == 1 && depth < max_depth) {            // :227
            long n_p = 0, s_p = 0;
            const double zz = rng.uniform(t, ucount++);      // :233
            const int dir = (zz <= 0.5) ? -1 : 1;            // :235
            NutsEnv e; e.c = &c; e.rng = &rng; e.draw = t; e.ucount = &ucount; e.n_lf = &n_lf;
            e.log_u = log_u; e.prev_U = prev_U; e.prev_K = prev_K;
            if (dir == -1) {
                vec dummy_draw(dpos), dummy_mntm(mpos);
                build_tree(e, dir, eps, prev, mntm, depth, new_draw, dummy_draw, dneg, dummy_mntm, mneg, n_p, s_p, alpha, n_alpha);  // :241-246 (Q12)
            } else {
                vec dummy_draw(dneg), dummy_mntm(mneg);
                build_tree(e, dir, eps, prev, mntm, depth, new_draw, dpos, dummy_draw, mpos, dummy_mntm, n_p, s_p, alpha, n_alpha);  // :251-255
            }
            if (s_p == 1) {
                const double z2 = rng.uniform(t, ucount++);   // :261
                if (z2 < double(n_p) / double(n_val)) {       // :263
                    double prop_U = -box_logp(c, new_draw.data());   // :264
                    if (!std::isfinite(prop_U)) prop_U = inf;
                    prev = new_draw; prev_U = prop_U; good_round = 1;     // :272-277
                }
            }
            n_val += n_p; depth += 1;                                      // :283-284
            vec diff(d);
            for (int i = 0; i < d; ++i) diff[i] = dpos[i] - dneg[i];
            const int chk1 = otgt::dot(diff.data(), mneg.data(), d, c.sum_mode) >= 0.0;   // :286
            const int chk2 = otgt::dot(diff.data(), mpos.data(), d, c.sum_mode) >= 0.0;   // :287
            s_val = s_p * chk1 * chk2;                                                    // :289
        }

        if (t < n_adapt) {   // :294-299 (Q15)
            h += (1 / double(double(t + 1) + t0)) * (delta - (double(alpha) / double(n_alpha)) - h);
            eps = std::exp(mu - h * std::sqrt(double(t + 1)) / gamma);
            eps_bar *= std::exp(std::pow(double(t + 1), -kappa) * (std::log(eps) - std::log(eps_bar)));
        } else {
            eps = eps_bar;   // :301
        }
        if (t >= cfg->n_burnin) {
            const long row = t - cfg->n_burnin;
            for (int j = 0; j < d; ++j) draws[row * d + j] = prev[j];
            if (logp_out) logp_out[row] = -prev_U;
            n_accept += good_round;   // :308
        }
    }
    if (c.bounded)   // :316-323
        for (long r = 0; r < cfg->n_keep; ++r) { vec tmp(draws + r * d, draws + (r + 1) * d); box_inv_transform(c, tmp.data(), draws + r * d); }
    res->n_accept = n_accept; res->tape_used = rng.rec_n; res->final_step = eps; res->n_leapfrog = n_lf;
    return 0;
}

// ------------------------------------------------------------------ RM-HMC (src/rmhmc.cpp:30-294, Appendix E)
// metrics registered with a target: otgt::metric (TGT_NORMAL_MODEL: examples/eigen/rmhmc_normal.cpp; TGT_FUNNEL: C5)
static void metric(const Ctx& c, const double* v, vec& G, vec* dG)
{
    // box_tensor_fn (src/rmhmc.cpp:150-161): the metric is evaluated at inv_transform(v) when bounded
    const int d = c.d;
    G.assign(size_t(d) * d, 0.0);
    if (dG) dG->assign(size_t(d) * d * d, 0.0);
    vec x(v, v + d);
    if (c.bounded) box_inv_transform(c, v, x.data());
    otgt::metric(c.target_id, c.metric_id, c.tdata, x.data(), d, G.data(), dG ? dG->data() : nullptr);
}

// returns (eps * F)/2 with F_i = -grad_i + 1/2 (tr(A D_i) - ((A D_i)' q).(A q))   (src/rmhmc.cpp:132-146; Q16 sign)
static void rm_mntm_update(const Ctx& c, double eps, const double* y, const double* q, const vec& A, const vec& dG, double* out)
{
    const int d = c.d;
    vec g(d), Aq(d), tq(d), T, Tt(size_t(d) * d), J(d);
    box_grad(c, y, g.data(), J.data());
    for (int i = 0; i < d; ++i) {
        vec Di(dG.begin() + size_t(i) * d * d, dG.begin() + size_t(i + 1) * d * d);
        matmul(A, Di, d, T);                      // tmp_mat = inv_tensor * deriv.mat(i)
        double tr = 0.0;
        for (int k = 0; k < d; ++k) tr += T[size_t(k) * d + k];
        for (int a = 0; a < d; ++a)
            for (int b = 0; b < d; ++b) Tt[size_t(b) * d + a] = T[size_t(a) * d + b];   // transpose (materialised)
        gemv_plain(Tt, d, q, tq.data());
        gemv_plain(A, d, q, Aq.data());
        const double dp = otgt::dot(tq.data(), Aq.data(), d, c.sum_mode);
        g[i] = -g[i] + 0.5 * (tr - dp);
    }
    if (c.bounded) for (int i = 0; i < d; ++i) out[i] = (J[i] * (eps * g[i])) / 2.0;   // step*J*grad/2 (src/rmhmc.cpp:130)
    else for (int i = 0; i < d; ++i) out[i] = (eps * g[i]) / 2.0;
    dense_jacobian_poison(c, g.data(), out);
}

static int run_rmhmc(const oracle_cfg_t* cfg, const double* x0, double* draws, double* logp_out, oracle_res_t* res)
{
    Ctx c; c.target_id = cfg->target_id; c.tdata = cfg->tdata; c.d = cfg->d; c.sum_mode = cfg->sum_mode;
    c.identity = true;   // precond_mat is never read (Q18)
    c.metric_id = cfg->metric_id;
    setup_bounds(c, cfg->vals_bound, cfg->lower, cfg->upper, cfg->dense_jacobian);
    Rng rng; init_rng(rng, cfg);
    const int d = c.d;
    const long n_total = cfg->n_burnin + cfg->n_keep;
    const double eps = cfg->step_size;
    const unsigned n_leap = unsigned(cfg->n_leap_steps), n_fp = unsigned(cfg->n_fp_steps);
    const double inf = std::numeric_limits<double>::infinity();

    vec prev(x0, x0 + d), cur(d), z(d), p(d), q(d), upd(d), w(d), t(d);
    if (c.bounded) box_transform(c, x0, prev.data());   // :166-168
    rng.normals(-1, d, z.data());   // :176 (Q3: value unused)

    vec newG, newdG, prevG, invNew, invPrev, prevdG, L, Gw, invW, sumM(size_t(d) * d);
    cur = prev;
    metric(c, cur.data(), newG, &newdG);   // :179
    prevG = newG; mat_inverse(newG, d, invNew); invPrev = invNew; prevdG = newdG;   // :181-186
    const double cons_term = double(0.5 * double(size_t(d)) * 1.83787706640934548356L);   // :188 (Q19)
    double prev_U = cons_term - box_logp(c, prev.data()) + 0.5 * logdet_llt(newG, d);  // :190
    long n_accept = 0, n_lf = 0;

    for (long it = 0; it < n_total; ++it) {
        rng.normals(it, d, z.data());               // :200
        mat_chol(prevG, d, cfg->chol_mode, L);      // :202 (Q8)
        gemv_plain(L, d, z.data(), p.data());
        gemv_plain(invPrev, d, p.data(), t.data());
        const double prev_K = otgt::dot(p.data(), t.data(), d, c.sum_mode) / 2.0;   // :204
        cur = prev;                                 // :206
        for (unsigned k = 0; k < n_leap; ++k) {
            q = p;                                  // :211
            for (unsigned kk = 0; kk < n_fp; ++kk) {   // :213-215 (Q17: start-of-trajectory metric)
                rm_mntm_update(c, eps, cur.data(), q.data(), invPrev, prevdG, upd.data());
                for (int i = 0; i < d; ++i) q[i] = p[i] + upd[i];
            }
            p = q;                                  // :217
            w = cur;                                // :221
            for (unsigned kk = 0; kk < n_fp; ++kk) {   // :224-228
                metric(c, w.data(), Gw, nullptr);
                mat_inverse(Gw, d, invW);
                invNew = invW;
                for (size_t m = 0; m < sumM.size(); ++m) sumM[m] = invPrev[m] + invW[m];
                gemv_scaled(sumM, d, 0.5 * eps, p.data(), t.data());
                for (int i = 0; i < d; ++i) w[i] = cur[i] + t[i];
            }
            cur = w;                                // :230
            metric(c, cur.data(), newG, &newdG);    // :232
            mat_inverse(newG, d, invNew);           // :233
            rm_mntm_update(c, eps, cur.data(), p.data(), invNew, newdG, upd.data());   // :237
            for (int i = 0; i < d; ++i) p[i] += upd[i];
            ++n_lf;
        }
        double prop_U = cons_term - box_logp(c, cur.data()) + 0.5 * logdet_llt(newG, d);   // :240
        if (!std::isfinite(prop_U)) prop_U = inf;
        gemv_plain(invNew, d, p.data(), t.data());
        const double prop_K = otgt::dot(p.data(), t.data(), d, c.sum_mode) / 2.0;               // :246
        const double comp = std::min(0.01, -(prop_U + prop_K) + (prev_U + prev_K));             // :250
        const double u = rng.uniform(it, 0);
        note_margin(it, u, comp);
        const bool acc = u < std::exp(comp);
        if (acc) { prev = cur; prev_U = prop_U; prevG = newG; invPrev = invNew; prevdG = newdG; }   // :254-261
        if (it >= cfg->n_burnin) {
            const long row = it - cfg->n_burnin;
            for (int j = 0; j < d; ++j) draws[row * d + j] = prev[j];
            if (logp_out) logp_out[row] = -(prev_U - cons_term);   // includes the +1/2 logdet G term
            if (acc) ++n_accept;
        }
    }
    if (c.bounded)   // :278-285
        for (long r = 0; r < cfg->n_keep; ++r) { vec tmp(draws + r * d, draws + (r + 1) * d); box_inv_transform(c, tmp.data(), draws + r * d); }
    res->n_accept = n_accept; res->tape_used = rng.rec_n; res->final_step = eps; res->n_leapfrog = n_lf;
    return 0;
}

int oracle_run_chain(const oracle_cfg_t* cfg, const double* x0, double* draws_out, double* logp_out, oracle_res_t* res)
{
    switch (cfg->sampler) {
    case S_HMC: return run_hmc(cfg, x0, draws_out, logp_out, res);
    case S_MALA: return run_mala(cfg, x0, draws_out, logp_out, res);
    case S_NUTS: return run_nuts(cfg, x0, draws_out, logp_out, res);
    case S_RMHMC: return run_rmhmc(cfg, x0, draws_out, logp_out, res);
    case S_RWMH: return run_rwmh(cfg, x0, draws_out, logp_out, res);
    default: return -1;
    }
}

// target value / gradient as the samplers see it (used by tests to check the CUDA functors)
double oracle_target(int target_id, const double* tdata, int d, const double* x, double* grad, int sum_mode)
{
    return otgt::value_and_grad(target_id, tdata, x, grad, d, sum_mode);
}

// ------------------------------------------------------------------ DE (src/de.cpp:30-271), SURVEY §8f item 4
// Differential evolution MCMC (ter Braak): ONE population of n_pop members per call; in every generation member i proposes
// X_i + gamma (X_c1 - X_c2) + U(-b, b)^d with two other members chosen at random.  The member loop updates X in place, so
// member i sees the NEW rows of members < i (the reference's OpenMP loop races on exactly that; its single-threaded order is
// the deterministic semantics restated here).  Random stream (single thread): engine(seed) -> one runif -> thread engine
// seed = size_t((u + 0 + 1) * 1000) (stats/seed_values.hpp:27); per member and generation: c1 (redrawn while == i), c2
// (redrawn while == i or == c1), d uniforms in (-b, b), one uniform z; rind(0, n-1) = size_t(U(nextafter(0, n), n))
// (stats/rind.hpp:36).  tape_out (optional) records the variates a device kernel needs, in order: the n_pop*d initial
// uniforms, then per generation and member {c1, c2, d proposal uniforms, z} (indices as doubles, rejected index draws are
// not recorded); rng_mode RNG_TAPE replays such a tape.
struct de_cfg_t {
    int target_id;
    const double* tdata;
    int d;
    long n_pop, n_burnin, n_keep;
    int jumps;
    double par_b, par_gamma_jump;
    const double* init_lb;
    const double* init_ub;
    int vals_bound;
    const double* lower;
    const double* upper;
    int rng_mode;             // RNG_MT or RNG_TAPE
    unsigned long seed;
    const double* tape; long tape_len;
    int sum_mode;
    double* tape_out; long tape_out_cap;
};

static double de_runif(std::mt19937_64& eng, double a, double b)
{
    const double a_adj = std::nextafter(a, b);   // runif.hpp:60
    std::uniform_real_distribution<double> ud(a_adj, b);
    return ud(eng);
}

int oracle_run_de(const de_cfg_t* cfg, const double* x0, double* draws, oracle_res_t* res)
{
    Ctx c; c.target_id = cfg->target_id; c.tdata = cfg->tdata; c.d = cfg->d; c.sum_mode = cfg->sum_mode; c.identity = true;
    setup_bounds(c, cfg->vals_bound, cfg->lower, cfg->upper);
    const int d = c.d;
    const long n_pop = cfg->n_pop, n_total = cfg->n_burnin + cfg->n_keep;
    const double par_b = cfg->par_b;
    const double par_gamma = 2.38 / std::sqrt(2.0 * double(size_t(d)));   // :59 (settings.par_gamma is never read)
    vec lb(d), ub(d);
    for (int j = 0; j < d; ++j) {   // :70-71
        lb[j] = cfg->init_lb ? cfg->init_lb[j] : x0[j] + (-0.5);
        ub[j] = cfg->init_ub ? cfg->init_ub[j] : x0[j] + 0.5;
    }
    if (c.bounded)   // sampling_bounds_check, misc/bounds_check.hpp:27-57
        for (int j = 0; j < d; ++j) {
            if (c.btype[j] == 4 || c.btype[j] == 2) lb[j] = std::max(c.lb[j], lb[j]);
            if (c.btype[j] == 4 || c.btype[j] == 3) ub[j] = std::min(c.ub[j], ub[j]);
        }
    std::mt19937_64 master(cfg->seed), eng;
    long cursor = 0, rec_n = 0;
    auto record = [&](double v) { if (cfg->tape_out && rec_n < cfg->tape_out_cap) cfg->tape_out[rec_n] = v; ++rec_n; };
    auto from_tape = [&]() { return (cursor < cfg->tape_len) ? cfg->tape[cursor++] : std::nan(""); };
    if (cfg->rng_mode == RNG_MT) {
        const double u0 = de_runif(master, 0.0, 1.0);
        eng.seed(static_cast<size_t>((u0 + 0 + 1) * 1000));   // generate_seed_value(0, 1, rand_engine)
    }
    auto unif = [&](double a, double b) {
        const double v = (cfg->rng_mode == RNG_MT) ? de_runif(eng, a, b) : from_tape();
        record(v);
        return v;
    };
    auto index_other = [&](long i, long c1) -> long {   // do { c = rind(0, n_pop-1) } while (c == i [|| c == c1])
        long cidx;
        if (cfg->rng_mode == RNG_MT) {
            do { cidx = long(static_cast<size_t>(de_runif(eng, 0.0, double(n_pop - 1) + 1.0))); } while (cidx == i || cidx == c1);
        } else {
            cidx = long(from_tape());
        }
        record(double(cidx));
        return cidx;
    };

    vec X(size_t(n_pop) * d), tv(n_pop), prop(d), rv(d);
    for (long i = 0; i < n_pop; ++i) {   // :118-137
        for (int j = 0; j < d; ++j) rv[j] = unif(0.0, 1.0);
        for (int j = 0; j < d; ++j) X[size_t(i) * d + j] = lb[j] + (ub[j] - lb[j]) * rv[j];
        double v = box_logp(c, &X[size_t(i) * d]);
        if (!std::isfinite(v)) v = -std::numeric_limits<double>::infinity();
        tv[i] = v;
    }
    long n_accept = 0;
    double gamma_run = par_gamma;
    for (long g = 0; g < n_total; ++g) {
        const double temperature = 1.0;   // de_cooling_schedule, de.hpp:86-89
        if (cfg->jumps && ((g + 1) % 10 == 0)) gamma_run = cfg->par_gamma_jump;   // :147-149
        for (long i = 0; i < n_pop; ++i) {
            const long c1 = index_other(i, -1);          // :166-168
            const long c2 = index_other(i, c1);          // :170-172
            for (int j = 0; j < d; ++j) rv[j] = unif(-par_b, par_b);   // :177
            for (int j = 0; j < d; ++j)                  // :179: (X_i + (X_c1 - X_c2) * gamma) + rand
                prop[j] = (X[size_t(i) * d + j] + (X[size_t(c1) * d + j] - X[size_t(c2) * d + j]) * gamma_run) + rv[j];
            double pv = box_logp(c, prop.data());        // :181
            if (!std::isfinite(pv)) pv = -std::numeric_limits<double>::infinity();
            const double comp = pv - tv[i];              // :189
            const double z = unif(0.0, 1.0);             // :190
            if (comp > temperature * std::log(z)) {      // :192
                for (int j = 0; j < d; ++j) X[size_t(i) * d + j] = prop[j];
                tv[i] = pv;
                if (g >= cfg->n_burnin) ++n_accept;
            }
        }
        if (g >= cfg->n_burnin) {                        // :210-212
            double* out = draws + size_t(g - cfg->n_burnin) * size_t(n_pop) * d;
            for (size_t k = 0; k < size_t(n_pop) * d; ++k) out[k] = X[k];
        }
        if (cfg->jumps && ((g + 1) % 10 == 0)) gamma_run = par_gamma;   // :214-216
    }
    if (c.bounded)   // :223-232
        for (long r = 0; r < cfg->n_keep * n_pop; ++r) { vec tmp(draws + r * d, draws + (r + 1) * d); box_inv_transform(c, tmp.data(), draws + r * d); }
    res->n_accept = n_accept; res->tape_used = rec_n; res->final_step = par_gamma; res->n_leapfrog = 0;
    return 0;
}

// a target's registered metric and its derivative cube (tests check the closed forms against finite differences)
int oracle_metric(int target_id, int metric_id, const double* tdata, int d, const double* x, double* G, double* dG)
{
    return otgt::metric(target_id, metric_id, tdata, x, d, G, dG) ? 0 : -1;
}

// raw RNG streams, for checking the engine's host-side tape generator and device Philox
void oracle_rng_stream(int rng_mode, unsigned long seed, long chain_id, long draw, int d, int n_unif, double* out)
{
    oracle_cfg_t cfg; std::memset(&cfg, 0, sizeof(cfg));
    cfg.rng_mode = rng_mode; cfg.seed = seed; cfg.chain_id = chain_id;
    Rng r; init_rng(r, &cfg);
    r.normals(draw, d, out);
    for (int k = 0; k < n_unif; ++k) out[d + k] = r.uniform(draw, k);
}

}  // extern "C"
