// RM-HMC, FAST arithmetic, one CTA (128 threads) per chain, all metric algebra in SHARED memory, no derivative cube.
//
// Same algorithm as rmhmc_general.cu (/root/reference/src/rmhmc.cpp:30-294; Q16/Q17 semantics: the fixed-point momentum
// iterations use the START-of-trajectory metric, the position iterations G_prev^-1 + G(w)^-1) — what changes is where the
// data lives and what is never materialised:
//   * The reference's metric callback returns the d x d x d derivative cube and the momentum update multiplies d matrices
//     (src/rmhmc.cpp:136-140, O(d^4)).  Only its contractions enter the force:
//         F_i = -grad_i + 1/2 (tr(A D_i) - ((A D_i)' q).(A q)) = -grad_i + 1/2 (<D_i, A>_F - u'^T D_i u),  u = A q, u' = A^T q,
//     so a registered metric provides `contract()`: all d values <D_i, A>_F - u'^T D_i u from its own structure.  For the
//     funnel's SoftAbs metric (arrow Hessian => G = [[g11, w x~'], [w x~, fa I + P x~ x~']]) every D_i is a rank-structured
//     matrix and the d contractions cost two GEMVs with A — O(d^2) instead of streaming a 2 MB cube per update (the cube
//     kernel moved 114 MB of DRAM traffic per chain-draw at d = 64; this one keeps the chain's whole state on chip).
//   * Matrices: ONE d-column scratch matrix `W` in shared memory (even leading dimension d + 2: row pairs move as 128-bit
//     words) in which a metric is built and inverted IN PLACE (Gauss-Jordan with partial pivoting, 2 d^3 flop) and in which the
//     draw's Cholesky factor / log det are taken.  `Ainv0` = G_prev^-1, constant during a draw and only ever read as the
//     operand of matrix-vector products, lives in the chain's global area (L2-resident, streamed with coalesced loads) next
//     to G at the accepted / proposed point (kept for the next draw's Cholesky): one shared matrix instead of two lets three
//     CTAs share an SM, and the eliminations are latency chains that only more resident warps can hide.
//   * 128 threads share each matrix operation (thread = row x column-parity), warp 0 evaluates the warp-cooperative target
//     functor and the random variates.
// FAST arithmetic only (operation orders differ from the reference's LU / Cholesky at rounding level, within the 1e-10
// contract); STRICT arithmetic and metrics without contract() run on rmhmc_general.cu.
#include "engine.h"
#include "rng.cuh"
#include "targets.cuh"
#include "rmhmc_metrics.cuh"
#include "rmhmc_cta.h"
#include <math_constants.h>
#include <cstdlib>
#include <type_traits>

namespace mcmcb200
{

constexpr int RC_THREADS = 128;
constexpr int RC_MAXD = 64;      // the tuned thread mappings (two threads per row / column half, register-tile elimination)
constexpr int RC_MAXD_WIDE = 128;   // generic mappings (one thread per row): 64 < n_dim <= 128, one CTA per SM (the matrix is 133 KB)
// WIDE is a template parameter of the kernel and its helpers; MCMCB200_RMHMC_WIDE=1 selects it at n_dim <= 64 too (tests: identical bits)

__device__ __forceinline__ void rc_sync() { __syncthreads(); }

// sum over the CTA of one value per thread (result in every thread); red: RC_THREADS/32 doubles of shared memory
__device__ __forceinline__ double rc_block_sum(double v, double* red)
{
    v = warp_sum<false>(v);
    rc_sync();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    rc_sync();
    return (red[0] + red[1]) + (red[2] + red[3]);
}

// ---- metrics: G into a shared matrix (leading dimension ld), and the contractions of their derivative ------------
// build(): every thread calls it; thread t writes rows i = t % 64.. of columns j = t / 64 (mod 2).
// contract(): c[k] = <D_k, A>_F - u'^T D_k u for k = 0..d-1 into c (shared); vs: scratch of >= 4 d doubles; red: 4 doubles.
struct FunnelSoftabsCta {
    FunnelSoftabsScalars<false> sc;
    __device__ __forceinline__ void prepare(const double* xs, int d) { sc.compute(xs, d); }
    __device__ __forceinline__ double entry(const double* xs, int i, int j) const
    {
        if (i == 0 && j == 0) return sc.g11.v;
        if (i == 0 || j == 0) return sc.w.v * xs[i + j];
        return ((i == j) ? sc.fa.v : 0.0) + (sc.P.v * xs[i]) * xs[j];
    }
    template <bool WIDE> __device__ __forceinline__ void build(const double* xs, int d, double* G, int ld) const
    {
        if (WIDE) {   // one thread per row
            const int i = threadIdx.x;
            if (i < d)
                for (int j = 0; j < d; ++j) G[(size_t)j * ld + i] = entry(xs, i, j);
            return;
        }
        const int i = threadIdx.x & 63;
        if (i < d)
            for (int j = threadIdx.x >> 6; j < d; j += 2) G[(size_t)j * ld + i] = entry(xs, i, j);
    }
    template <bool WIDE>
    __device__ __forceinline__ void contract(const double* xs, int d, const double* Am, int ld, const double* u, const double* up, double* vs, double* red,
                                             double* c) const
    {
        // yr = A x~, yc = A^T x~ (x~ = x with element 0 zeroed): threads 0..63 rows of A, threads 64..127 rows of A^T
        double* yr = vs;
        double* yc = vs + d;
        const int t = threadIdx.x, i = t & 63;
        if (WIDE) {   // one thread per row: both products
            if (t < d) {
                double acc = 0.0, acc2 = 0.0;
                for (int j = 1; j < d; ++j) acc = fma(Am[(size_t)j * ld + t], xs[j], acc);
                for (int j = 1; j < d; ++j) acc2 = fma(Am[(size_t)t * ld + j], xs[j], acc2);
                yr[t] = acc;
                yc[t] = acc2;
            }
        } else if (i < d) {
            double acc = 0.0;
            if (t < 64) { for (int j = 1; j < d; ++j) acc = fma(Am[(size_t)j * ld + i], xs[j], acc); yr[i] = acc; }
            else { for (int j = 1; j < d; ++j) acc = fma(Am[(size_t)i * ld + j], xs[j], acc); yc[i] = acc; }
        }
        rc_sync();
        // scalars: a_rc = sum_{i>=1} x_i (A(0,i) + A(i,0)), trA1 = sum_{i>=1} A(i,i), xAx = x~' A x~, ux = u.x~, upx = u'.x~, uu1 = sum_{i>=1} u'_i u_i
        double p_arc = 0.0, p_tr = 0.0, p_xax = 0.0, p_ux = 0.0, p_upx = 0.0, p_uu = 0.0;
        if (t >= 1 && t < d) {
            p_arc = xs[t] * (Am[(size_t)t * ld] + Am[t]);
            p_tr = Am[(size_t)t * ld + t];
            p_xax = xs[t] * yr[t];
            p_ux = u[t] * xs[t];
            p_upx = up[t] * xs[t];
            p_uu = up[t] * u[t];
        }
        const double a_rc = rc_block_sum(p_arc, red), trA1 = rc_block_sum(p_tr, red), xAx = rc_block_sum(p_xax, red);
        const double ux = rc_block_sum(p_ux, red), upx = rc_block_sum(p_upx, red), uu1 = rc_block_sum(p_uu, red);
        const double A00 = Am[0], u0 = u[0], up0 = up[0];
        const double cross = up0 * ux + upx * u0;
        if (t == 0) {
            const double fa_ = sc.g11.dv * A00 + sc.w.dv * a_rc + sc.fa.dv * trA1 + sc.P.dv * xAx;
            const double fu_ = sc.g11.dv * (up0 * u0) + sc.w.dv * cross + sc.fa.dv * uu1 + sc.P.dv * (upx * ux);
            c[0] = fa_ - fu_;
        } else if (t < d) {
            const double s2 = 2.0 * xs[t];
            const double fa_ = s2 * (sc.g11.ds * A00 + sc.w.ds * a_rc + sc.P.ds * xAx) + sc.w.v * (Am[(size_t)t * ld] + Am[t]) + sc.P.v * (yc[t] + yr[t]);
            const double fu_ = s2 * (sc.g11.ds * (up0 * u0) + sc.w.ds * cross + sc.P.ds * (upx * ux)) + sc.w.v * (up0 * u[t] + up[t] * u0) +
                               sc.P.v * (up[t] * ux + upx * u[t]);
            c[t] = fa_ - fu_;
        }
        rc_sync();
    }
};

struct FunnelFisherCta {   // G = diag(1/9 + (d-1)/2, e^-v, ..., e^-v); dG/dv = diag(0, -e^-v, ...), everything else zero
    double ev;
    __device__ __forceinline__ void prepare(const double* xs, int) { ev = exp(-xs[0]); }
    template <bool WIDE> __device__ __forceinline__ void build(const double*, int d, double* G, int ld) const
    {
        if (WIDE) {
            const int i = threadIdx.x;
            if (i < d)
                for (int j = 0; j < d; ++j) G[(size_t)j * ld + i] = (i != j) ? 0.0 : ((i == 0) ? 1.0 / 9.0 + (double)(d - 1) / 2.0 : ev);
            return;
        }
        const int i = threadIdx.x & 63;
        if (i < d)
            for (int j = threadIdx.x >> 6; j < d; j += 2) G[(size_t)j * ld + i] = (i != j) ? 0.0 : ((i == 0) ? 1.0 / 9.0 + (double)(d - 1) / 2.0 : ev);
    }
    template <bool WIDE>
    __device__ __forceinline__ void contract(const double*, int d, const double* Am, int ld, const double* u, const double* up, double*, double* red,
                                             double* c) const
    {
        const int t = threadIdx.x;
        double p = 0.0;
        if (t >= 1 && t < d) p = Am[(size_t)t * ld + t] - up[t] * u[t];
        const double s = rc_block_sum(p, red);
        if (t < d) c[t] = (t == 0) ? -ev * s : 0.0;
        rc_sync();
    }
};

// ---- CTA-wide dense algebra on a shared d x d matrix with leading dimension ld (even, >= d + 1) ---------------------
// In-place inverse by Gauss-Jordan elimination with partial pivoting (2 d^3 flop).  Thread (rp, cg) = (t % 32, t / 32) owns
// the row PAIR (2 rp, 2 rp + 1) in a contiguous block of columns: one 128-bit shared load brings both rows of a column, one
// brings two entries of the pivot row, so the rank-1 update costs ~2.3 instructions per multiply-add (the first version,
// one row per thread with 64-bit accesses and per-element special cases, needed 11 and was issue-bound).  There are no
// special rows inside the update: the pivot row is exchanged physically first, and its own scaling is folded into the
// update by giving it the multiplier pivot - 1 (W(k,j) - (pivot - 1) W(k,j) / pivot = W(k,j) / pivot).
// piv: d ints; rc: 2 (RC_MAXD + 2) doubles (the scaled pivot row and the multiplier column), 16-byte aligned.
// A zero / NaN pivot propagates NaN like the reference's LU does.
// WIDE (64 < n_dim <= 128): 64 row pairs x 2 column groups instead of 32 x 4; same operations per element.
template <bool WIDE> __device__ void rc_inverse_inplace(double* W, int d, int ld, int* piv, double* rc)
{
    constexpr int NRP = WIDE ? 64 : 32, NCG = RC_THREADS / NRP, MAXD = WIDE ? RC_MAXD_WIDE : RC_MAXD;
    double* rowbuf = rc;
    double* colbuf = rc + MAXD + 2;
    const int t = threadIdx.x, rp = t % NRP, cg = t / NRP;
    const int i0 = 2 * rp;
    const bool rows_ok = i0 < d;
    const int cb = ((d + 2 * NCG - 1) / (2 * NCG)) * 2;   // columns per thread group (even)
    const int jb = cg * cb, je = (jb + cb < d) ? jb + cb : d;
    for (int k = 0; k < d; ++k) {
        // pivot: first maximum of |W(r, k)|, r >= k (warp 0; a NaN never wins, a NaN at (k,k) keeps r = k)
        if (t < 32) {
            double bv = -2.0;
            int bi = 0x7fffffff;
            const double vk = fabs(W[(size_t)k * ld + k]);
            if (vk == vk) {
                for (int r = t; r < d; r += 32)
                    if (r >= k) {
                        double v = fabs(W[(size_t)k * ld + r]);
                        if (!(v == v)) v = -1.0;
                        if (v > bv) { bv = v; bi = r; }
                    }
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const double ov = __shfl_xor_sync(FULL, bv, off);
                    const int oi = __shfl_xor_sync(FULL, bi, off);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
            } else {
                bi = k;
            }
            if (t == 0) piv[k] = bi;
        }
        rc_sync();
        const int p = piv[k];
        if (p != k) {   // exchange rows k and p (threads over columns)
            if (t < d) {
                const double a = W[(size_t)t * ld + k], b = W[(size_t)t * ld + p];
                W[(size_t)t * ld + k] = b;
                W[(size_t)t * ld + p] = a;
            }
            rc_sync();
        }
        const double pv = W[(size_t)k * ld + k];
        const double rinv = 1.0 / pv;
        if (WIDE) {
            if (t < d) rowbuf[t] = W[(size_t)t * ld + k] * rinv;                              // scaled pivot row
            for (int r = t; r < d + 1; r += RC_THREADS)
                colbuf[r] = (r == k) ? pv - 1.0 : ((r < d) ? W[(size_t)k * ld + r] : 0.0);    // multipliers (see above); padding row: 0
        } else {
            if (t < d) rowbuf[t] = W[(size_t)t * ld + k] * rinv;                              // scaled pivot row
            else if (t >= 64 && t - 64 < d + 1) {
                const int r = t - 64;
                colbuf[r] = (r == k) ? pv - 1.0 : ((r < d) ? W[(size_t)k * ld + r] : 0.0);    // multipliers (see above); padding row: 0
            }
        }
        rc_sync();
        if (rows_ok) {
            const double f0 = colbuf[i0], f1 = colbuf[i0 + 1];
            int j = jb;
            for (; j + 1 < je; j += 2) {
                const double2 r2 = *reinterpret_cast<const double2*>(rowbuf + j);
                double2 wa = *reinterpret_cast<double2*>(W + (size_t)j * ld + i0);
                double2 wb = *reinterpret_cast<double2*>(W + (size_t)(j + 1) * ld + i0);
                wa.x = fma(-f0, r2.x, wa.x); wa.y = fma(-f1, r2.x, wa.y);
                wb.x = fma(-f0, r2.y, wb.x); wb.y = fma(-f1, r2.y, wb.y);
                *reinterpret_cast<double2*>(W + (size_t)j * ld + i0) = wa;
                *reinterpret_cast<double2*>(W + (size_t)(j + 1) * ld + i0) = wb;
            }
            if (j < je) {
                const double r1 = rowbuf[j];
                double2 wa = *reinterpret_cast<double2*>(W + (size_t)j * ld + i0);
                wa.x = fma(-f0, r1, wa.x); wa.y = fma(-f1, r1, wa.y);
                *reinterpret_cast<double2*>(W + (size_t)j * ld + i0) = wa;
            }
            if (k >= jb && k < je) {   // column k: 1 / pivot in the pivot row, -multiplier / pivot elsewhere
                double2 wk;
                wk.x = (i0 == k) ? rinv : -f0 * rinv;
                wk.y = (i0 + 1 == k) ? rinv : -f1 * rinv;
                *reinterpret_cast<double2*>(W + (size_t)k * ld + i0) = wk;
            }
        }
        rc_sync();
    }
    // row swaps of A are column swaps of A^-1, undone in reverse order
    for (int k = d - 1; k >= 0; --k) {
        const int p = piv[k];
        if (p != k) {
            if (t < d) {
                const double a = W[(size_t)k * ld + t], b = W[(size_t)p * ld + t];
                W[(size_t)k * ld + t] = b;
                W[(size_t)p * ld + t] = a;
            }
            rc_sync();
        }
    }
}

// The same elimination with the matrix in REGISTERS (n_dim > 32).  Thread t owns the 8 x 4 tile of rows 8 (t % 8) .. + 7 and
// columns 4 (t / 8) .. + 3 of the (padded) 64 x 64 matrix for all d pivot steps; a step exchanges only what the rank-1 update
// needs through shared memory — column k (the multipliers and the pivot candidates) and the scaled pivot row — so it costs
// 32 multiply-adds against ~12 shared accesses per thread instead of 40 for 32 (rc_inverse_inplace) and two CTA barriers
// instead of four.  Rows are never exchanged: the permutation of partial pivoting is tracked (`lof` = logical position of a
// physical row, `pof` = its inverse) and applied once when the tile is stored — A^-1(i, r) = W(pof[i], lof[r]) read the other
// way round.  Pivot choice (first maximum by LOGICAL position, NaN rules), multipliers and update formulas are those of
// rc_inverse_inplace: the two produce the same bits.
// buf: 3 * RC_MAXD doubles (column k double-buffered, scaled pivot row), perm: 2 * RC_MAXD ints; both 16-byte aligned.
// (The tile is 32 named scalars, not an array: with an array the compiler merges the per-column / per-row cases of a step into
//  one dynamically addressed access and the whole tile moves to local memory — measured 30 ms per C5 draw instead of 16.)
#define RC_R8(OP, c) OP(c, 0) OP(c, 1) OP(c, 2) OP(c, 3) OP(c, 4) OP(c, 5) OP(c, 6) OP(c, 7)
#define RC_C4R8(OP) RC_R8(OP, 0) RC_R8(OP, 1) RC_R8(OP, 2) RC_R8(OP, 3)
#define RC_A(c, r) a##c##r
#define RC_F(r) f##r
__device__ __noinline__ void rc_inverse_regtile(double* W, int d, int ld, double* buf, int* perm)
{
    constexpr int N = RC_MAXD;
    double* const rowbuf = buf + 2 * N;
    int* const lof = perm;       // logical position of physical row r
    int* const pof = perm + N;   // physical row at logical position l
    const int t = threadIdx.x, lane = t & 31;
    const int rb = t & 7, cbk = t >> 3;
    const int i0 = 8 * rb, j0 = 4 * cbk;
    // RC_A(c, r) = element (i0 + r, j0 + c); padding rows / columns: identity
#define RC_DECL(c, r) double RC_A(c, r);
    RC_C4R8(RC_DECL)
#undef RC_DECL
#define RC_LOAD2(c, q)                                                                                   \
    {                                                                                                    \
        const int j = j0 + c, i = i0 + 2 * q;                                                            \
        double2 v = make_double2((i == j) ? 1.0 : 0.0, (i + 1 == j) ? 1.0 : 0.0);                        \
        if (j < d && i < d) {                                                                            \
            v = *reinterpret_cast<const double2*>(W + (size_t)j * ld + i);                               \
            if (i + 1 >= d) v.y = 0.0;                                                                   \
        }                                                                                                \
        RC_A(c, 2 * q) = v.x;                                                                            \
        RC_A(c, 2 * q + 1) = v.y;                                                                        \
    }
    // (token pasting needs literal row numbers)
#define RC_LOADC(c)                                                                                      \
    {                                                                                                    \
        const int j = j0 + c;                                                                            \
        double2 v0 = make_double2((i0 == j) ? 1.0 : 0.0, (i0 + 1 == j) ? 1.0 : 0.0), v1 = make_double2((i0 + 2 == j) ? 1.0 : 0.0, (i0 + 3 == j) ? 1.0 : 0.0); \
        double2 v2 = make_double2((i0 + 4 == j) ? 1.0 : 0.0, (i0 + 5 == j) ? 1.0 : 0.0), v3 = make_double2((i0 + 6 == j) ? 1.0 : 0.0, (i0 + 7 == j) ? 1.0 : 0.0); \
        if (j < d) {                                                                                     \
            const double* col = W + (size_t)j * ld + i0;                                                 \
            if (i0 < d) { v0 = *reinterpret_cast<const double2*>(col); if (i0 + 1 >= d) v0.y = 0.0; }              \
            if (i0 + 2 < d) { v1 = *reinterpret_cast<const double2*>(col + 2); if (i0 + 3 >= d) v1.y = 0.0; }      \
            if (i0 + 4 < d) { v2 = *reinterpret_cast<const double2*>(col + 4); if (i0 + 5 >= d) v2.y = 0.0; }      \
            if (i0 + 6 < d) { v3 = *reinterpret_cast<const double2*>(col + 6); if (i0 + 7 >= d) v3.y = 0.0; }      \
        }                                                                                                \
        RC_A(c, 0) = v0.x; RC_A(c, 1) = v0.y; RC_A(c, 2) = v1.x; RC_A(c, 3) = v1.y;                      \
        RC_A(c, 4) = v2.x; RC_A(c, 5) = v2.y; RC_A(c, 6) = v3.x; RC_A(c, 7) = v3.y;                      \
    }
    RC_LOADC(0) RC_LOADC(1) RC_LOADC(2) RC_LOADC(3)
#undef RC_LOADC
#undef RC_LOAD2
    if (t < N) { lof[t] = t; pof[t] = t; }
    // one column of the tile -> 8 consecutive doubles of shared memory
#define RC_STORE_COL(c, dst)                                                                             \
    {                                                                                                    \
        *reinterpret_cast<double2*>((dst)) = make_double2(RC_A(c, 0), RC_A(c, 1));                       \
        *reinterpret_cast<double2*>((dst) + 2) = make_double2(RC_A(c, 2), RC_A(c, 3));                   \
        *reinterpret_cast<double2*>((dst) + 4) = make_double2(RC_A(c, 4), RC_A(c, 5));                   \
        *reinterpret_cast<double2*>((dst) + 6) = make_double2(RC_A(c, 6), RC_A(c, 7));                   \
    }
    if (cbk == 0) RC_STORE_COL(0, buf + i0)   // column 0 for the first step
    rc_sync();
    for (int k = 0; k < d; ++k) {
        const double* colbuf = buf + (k & 1) * N;
        // pivot: first maximum of |W(l, k)| over logical rows l >= k (a NaN never wins, a NaN at logical (k, k) keeps l = k);
        // every warp finds it redundantly
        // (|v| of a non-NaN double orders like its bit pattern: key = bits + 2, NaN -> 1, not a candidate -> 0; the maximum of
        //  the 64-bit keys and then the smallest logical position among its holders take three warp reductions (redux.sync))
        int p;
        {
            const int pk = pof[k];
            const double vk = colbuf[pk];
            p = pk;
            if (vk == vk) {
                unsigned long long key = 0ull;
                int bl = 0x7fffffff;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = lane + 32 * h;
                    const int l = lof[r];
                    if (r < d && l >= k) {
                        const double v = colbuf[r];
                        const unsigned long long kv = (v == v) ? (unsigned long long)(__double_as_longlong(v) & 0x7fffffffffffffffll) + 2ull : 1ull;
                        if (kv > key || (kv == key && l < bl)) { key = kv; bl = l; }
                    }
                }
                const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
                const unsigned mhi = __reduce_max_sync(FULL, hi);
                const unsigned mlo = __reduce_max_sync(FULL, (hi == mhi) ? lo : 0u);
                const unsigned ml = __reduce_min_sync(FULL, (hi == mhi && lo == mlo) ? (unsigned)bl : 0x7fffffffu);
                p = pof[ml];
            }
        }
        const double pv = colbuf[p];
        const double rinv = 1.0 / pv;
        // the scaled pivot row, from the threads that hold physical row p (p is uniform: no divergence inside the switch)
        if (rb == (p >> 3)) {
            double r0, r1, r2, r3;
            switch (p & 7) {
#define RC_ROW(R) case R: r0 = RC_A(0, R); r1 = RC_A(1, R); r2 = RC_A(2, R); r3 = RC_A(3, R); break;
                RC_ROW(0) RC_ROW(1) RC_ROW(2) RC_ROW(3) RC_ROW(4) RC_ROW(5) RC_ROW(6)
                default: r0 = RC_A(0, 7); r1 = RC_A(1, 7); r2 = RC_A(2, 7); r3 = RC_A(3, 7); break;
#undef RC_ROW
            }
            *reinterpret_cast<double2*>(rowbuf + j0) = make_double2(r0 * rinv, r1 * rinv);
            *reinterpret_cast<double2*>(rowbuf + j0 + 2) = make_double2(r2 * rinv, r3 * rinv);
        }
        rc_sync();
        if (t == 0) {   // logical rows k and lof[p] change places (read again only after the next barrier)
            const int lp = lof[p], pk = pof[k];
            pof[k] = p; pof[lp] = pk;
            lof[p] = k; lof[pk] = lp;
        }
        // multipliers: column k; the pivot row gets pivot - 1 (its scaling folded into the update, see rc_inverse_inplace)
        const double2 c01 = *reinterpret_cast<const double2*>(colbuf + i0), c23 = *reinterpret_cast<const double2*>(colbuf + i0 + 2);
        const double2 c45 = *reinterpret_cast<const double2*>(colbuf + i0 + 4), c67 = *reinterpret_cast<const double2*>(colbuf + i0 + 6);
        const int pr = p - i0;   // the pivot row's place in this tile (0..7) or outside
        const double f0 = (pr == 0) ? pv - 1.0 : c01.x, f1 = (pr == 1) ? pv - 1.0 : c01.y, f2 = (pr == 2) ? pv - 1.0 : c23.x, f3 = (pr == 3) ? pv - 1.0 : c23.y;
        const double f4 = (pr == 4) ? pv - 1.0 : c45.x, f5 = (pr == 5) ? pv - 1.0 : c45.y, f6 = (pr == 6) ? pv - 1.0 : c67.x, f7 = (pr == 7) ? pv - 1.0 : c67.y;
        const double2 ra = *reinterpret_cast<const double2*>(rowbuf + j0), rb2 = *reinterpret_cast<const double2*>(rowbuf + j0 + 2);
        const double rr0 = ra.x, rr1 = ra.y, rr2 = rb2.x, rr3 = rb2.y;
#define RC_UPD(c, r) RC_A(c, r) = fma(-RC_F(r), rr##c, RC_A(c, r));
        RC_C4R8(RC_UPD)
#undef RC_UPD
        if (cbk == (k >> 2)) {   // column k: 1 / pivot in the pivot row, -multiplier / pivot elsewhere (k is uniform)
#define RC_FIX(c, r) RC_A(c, r) = (pr == r) ? rinv : -RC_F(r) * rinv;
            switch (k & 3) {
            case 0: RC_R8(RC_FIX, 0) break;
            case 1: RC_R8(RC_FIX, 1) break;
            case 2: RC_R8(RC_FIX, 2) break;
            default: RC_R8(RC_FIX, 3) break;
            }
#undef RC_FIX
        }
        // column k + 1 for the next step (the other half of the double buffer: this step's is still being read)
        if (k + 1 < d && cbk == ((k + 1) >> 2)) {
            double* nb = buf + ((k + 1) & 1) * N + i0;
            switch ((k + 1) & 3) {
            case 0: RC_STORE_COL(0, nb) break;
            case 1: RC_STORE_COL(1, nb) break;
            case 2: RC_STORE_COL(2, nb) break;
            default: RC_STORE_COL(3, nb) break;
            }
        }
        rc_sync();
    }
    // A^-1(i, r) = tile value at (physical row pof[i], column lof[r]): physical row pr, column c goes to (lof[pr], pof[c])
    {
        int li[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) li[r] = (i0 + r < d) ? lof[i0 + r] : -1;
#define RC_OUT(c, r) if (li[r] >= 0) col[li[r]] = RC_A(c, r);
#define RC_OUTC(c)                                          \
    if (j0 + c < d) {                                       \
        double* col = W + (size_t)pof[j0 + c] * ld;         \
        RC_R8(RC_OUT, c)                                    \
    }
        RC_OUTC(0) RC_OUTC(1) RC_OUTC(2) RC_OUTC(3)
#undef RC_OUTC
#undef RC_OUT
    }
    rc_sync();
}
#undef RC_STORE_COL
#undef RC_F
#undef RC_A
#undef RC_C4R8
#undef RC_R8

// in-place lower Cholesky of the shared matrix (right-looking); the strict upper triangle keeps the input's entries — the
// Eigen matrixLLT storage the reference multiplies with in full; returns nothing, L in the lower triangle
template <bool WIDE> __device__ void rc_cholesky_inplace(double* W, int d, int ld)
{
    constexpr bool wide = WIDE;   // one thread per row instead of two
    const int t = threadIdx.x, i = wide ? t : (t & 63), jpar = t >> 6;
    for (int j = 0; j < d; ++j) {
        const double dj = sqrt(W[(size_t)j * ld + j]);
        rc_sync();
        if (t < d && t >= j) W[(size_t)j * ld + t] = (t == j) ? dj : W[(size_t)j * ld + t] / dj;
        rc_sync();
        // trailing update of the lower triangle: W(i, k) -= L(i, j) L(k, j), k > j, i >= k
        if (i < d && i > j) {
            const double lij = W[(size_t)j * ld + i];
            if (wide) for (int k = j + 1; k <= i; ++k) W[(size_t)k * ld + i] = fma(-lij, W[(size_t)j * ld + k], W[(size_t)k * ld + i]);
            else for (int k = j + 1 + ((jpar + j + 1) & 1); k <= i; k += 2) W[(size_t)k * ld + i] = fma(-lij, W[(size_t)j * ld + k], W[(size_t)k * ld + i]);
        }
        rc_sync();
    }
}

// n_dim > 32: the register-tile elimination; smaller matrices would mostly multiply padding there
__device__ int rc_force_shared_gj = 0;   // MCMCB200_RMHMC_REGTILE=0 (tests: the two eliminations must agree bit for bit)
template <bool WIDE> __device__ __forceinline__ void rc_invert(double* W, int d, int ld, int* piv, double* buf, int* perm)
{
    if (WIDE) rc_inverse_inplace<true>(W, d, ld, piv, buf);
    else if (d > 32 && !rc_force_shared_gj) rc_inverse_regtile(W, d, ld, buf, perm);
    else rc_inverse_inplace<false>(W, d, ld, piv, buf);
}

__device__ double2 rc_rng_tab_g[RNG_TAB_DOUBLE2];
__global__ void rc_build_rng_tab() { build_rng_tables(rc_rng_tab_g); }   // same values from every launch: concurrent calls do not conflict

// 128 registers: four CTAs (16 warps) per SM — the eliminations are latency chains that only more resident warps hide
template <class T, class MC, int RNGM, bool WIDE>
__global__ void __launch_bounds__(RC_THREADS, WIDE ? 1 : 4) rmhmc_cta_kernel(const __grid_constant__ RmhmcLaunch a)
{
    extern __shared__ __align__(16) double smem[];
    // Box-Muller tables: read from global memory (built once per launch by rc_build_rng_tab; d normals per draw use them) —
    // 16 KB of shared memory per CTA would cap the SM at three CTAs
    const double2* const rng_tab = rc_rng_tab_g;
    __shared__ int piv[RC_MAXD_WIDE];
    __shared__ double red[4];
    __shared__ __align__(16) double rcbuf[2 * (RC_MAXD_WIDE + 2)];   // >= 3 RC_MAXD (register-tile elimination)
    __shared__ __align__(16) int rtperm[2 * RC_MAXD];
    __shared__ double sc_u, sc_lp;   // broadcast scalars (uniform, log-density)
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const long long chain = blockIdx.x;
    const int d = a.d;
    const int ld = (d + 2) & ~1;   // even (128-bit row-pair accesses), > d (one padding row), transposed reads at most 4-way conflicted
    const int dp = (d + 1) & ~1;
    double* W = smem;                        // scratch matrix
    double* vec = W + (size_t)d * ld;        // vectors, dp doubles each
    double* xprev = vec;            double* xcur = vec + dp;        double* pv = vec + 2 * dp;   double* qv = vec + 3 * dp;
    double* wv = vec + 4 * dp;      double* gv = vec + 5 * dp;      double* uv = vec + 6 * dp;   double* upv = vec + 7 * dp;
    double* cv = vec + 8 * dp;      double* zv = vec + 9 * dp;      double* tv = vec + 10 * dp;  double* scr = vec + 11 * dp;   // scr: 4 dp
    double* tscr = scr + 4 * dp;    // target functor scratch (dp)
    const WarpCtx wctx{lane, d, tscr};
    // the chain's global area: G at the accepted point and at the proposal (column-major, leading dimension d)
    double* Gacc = a.work + (size_t)chain * (size_t)a.work_stride;
    double* Gnew = Gacc + (size_t)d * d;
    double* Ainv0 = Gnew + (size_t)d * d;    // G_prev^-1 (leading dimension ld), fixed during a draw; written on an accept only
    const double eps = a.eps, heps = 0.5 * eps;
    rc_sync();

    ChainRng<RNGM> rng;
    rng.init(a.rng, chain, a.chain_offset + chain);
    // warp 0: d normals of `draw` into dst (shared)
    auto normals_to = [&](long long draw, double* dst) {
        if (warp == 0) {
            if (!WIDE || d <= 64) {
                double z[2];
                rng.template normals<2, false>(a.rng, draw, d, lane, rng_tab, z);
                if (2 * lane < d) dst[2 * lane] = z[0];
                if (2 * lane + 1 < d) dst[2 * lane + 1] = z[1];
            } else {
                double z[4];
                rng.template normals<4, false>(a.rng, draw, d, lane, rng_tab, z);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (elem_index(lane, k) < d) dst[elem_index(lane, k)] = z[k];
            }
        }
    };
    // warp 0: log pi and/or gradient at the shared vector xin -> sc_lp / gv
    auto target_at = [&](const double* xin, bool want_value, bool want_grad) {
        if (warp == 0) {
            auto run = [&](auto epl_tag) {
                constexpr int E = decltype(epl_tag)::value;
                double x[E], g[E];
#pragma unroll
                for (int k = 0; k < E; ++k) x[k] = (elem_index(lane, k) < d) ? xin[elem_index(lane, k)] : 0.0;
                double v = 0.0;
                if (want_value && want_grad) v = T::template eval<E, false, true, true, true>(a.tdata, wctx, x, g);
                else if (want_value) v = T::template eval<E, false, true, false, true>(a.tdata, wctx, x, g);
                else T::template eval<E, false, false, true, true>(a.tdata, wctx, x, g);
                if (want_grad) {
#pragma unroll
                    for (int k = 0; k < E; ++k)
                        if (elem_index(lane, k) < d) gv[elem_index(lane, k)] = g[k];
                }
                if (want_value && lane == 0) sc_lp = v;
            };
            if constexpr (!WIDE) run(std::integral_constant<int, 2>());
            else if (d <= 64) run(std::integral_constant<int, 2>());
            else run(std::integral_constant<int, 4>());
        }
        rc_sync();
    };
    // y = Am v (threads 0..63) and optionally y2 = Am^T v (threads 64..127)
    constexpr bool wide = WIDE;   // generic thread mappings: one thread per row
    auto gemv2 = [&](const double* Am, const double* v, double* y, double* y2) {
        const int i = t & 63;
        if (wide) {
            if (t < d) {
                double acc = 0.0, acc2 = 0.0;
                int j = 0;
                for (; j + 1 < d; j += 2) { acc = fma(Am[(size_t)j * ld + t], v[j], acc); acc2 = fma(Am[(size_t)(j + 1) * ld + t], v[j + 1], acc2); }
                if (j < d) acc = fma(Am[(size_t)j * ld + t], v[j], acc);
                y[t] = acc + acc2;
                if (y2) {
                    acc = 0.0; acc2 = 0.0; j = 0;
                    for (; j + 1 < d; j += 2) { acc = fma(Am[(size_t)t * ld + j], v[j], acc); acc2 = fma(Am[(size_t)t * ld + j + 1], v[j + 1], acc2); }
                    if (j < d) acc = fma(Am[(size_t)t * ld + j], v[j], acc);
                    y2[t] = acc + acc2;
                }
            }
        } else if (i < d) {
            double acc = 0.0;
            double acc2 = 0.0;
            int j = 0;
            if (t < 64) {
                for (; j + 1 < d; j += 2) { acc = fma(Am[(size_t)j * ld + i], v[j], acc); acc2 = fma(Am[(size_t)(j + 1) * ld + i], v[j + 1], acc2); }
                if (j < d) acc = fma(Am[(size_t)j * ld + i], v[j], acc);
                y[i] = acc + acc2;
            } else if (y2) {   // A^T v with the same coalesced column access: threads 64.. walk the other half of the columns
                for (; j + 1 < d; j += 2) { acc = fma(Am[(size_t)i * ld + j], v[j], acc); acc2 = fma(Am[(size_t)i * ld + j + 1], v[j + 1], acc2); }
                if (j < d) acc = fma(Am[(size_t)i * ld + j], v[j], acc);
                y2[i] = acc + acc2;
            }
        }
        rc_sync();
    };
    MC metric;
    // log det G the way the reference takes it — 2 sum log diag(chol G) (core/log_det.hpp:35), NaN when G is not numerically
    // positive definite, which the accept rule then treats like the reference does — for the metric `m` prepared at xm; W is
    // used as scratch and holds G again on return
    auto logdet_chol = [&](const MC& m, const double* xm) -> double {
        rc_cholesky_inplace<WIDE>(W, d, ld);
        const double s = rc_block_sum(t < d ? 2.0 * log(W[(size_t)t * ld + t]) : 0.0, red);
        rc_sync();
        m.template build<WIDE>(xm, d, W, ld);
        rc_sync();
        return s;
    };
    // out = (eps F)/2 at position y with momentum q, metric inverse Am evaluated at the point the metric object was prepared for
    auto mntm_update = [&](const double* y, const double* q, const double* Am, const MC& m, const double* xm, double* out) {
        target_at(y, false, true);
        gemv2(Am, q, uv, upv);
        m.template contract<WIDE>(xm, d, Am, ld, uv, upv, scr, red, cv);
        if (t < d) out[t] = (eps * fma(0.5, cv[t], -gv[t])) * 0.5;
        rc_sync();
    };
    auto store_global = [&](double* G) {   // W (ld) -> global (d), before it is inverted in place
        if (wide) {
            if (t < d)
                for (int j = 0; j < d; ++j) G[(size_t)j * d + t] = W[(size_t)j * ld + t];
            return;
        }
        const int i = t & 63;
        if (i < d)
            for (int j = t >> 6; j < d; j += 2) G[(size_t)j * d + i] = W[(size_t)j * ld + i];
    };
    auto load_global = [&](const double* G) {   // global (d) -> W (ld)
        if (wide) {
            if (t < d)
                for (int j = 0; j < d; ++j) W[(size_t)j * ld + t] = G[(size_t)j * d + t];
            return;
        }
        const int i = t & 63;
        if (i < d)
            for (int j = t >> 6; j < d; j += 2) W[(size_t)j * ld + i] = G[(size_t)j * d + i];
    };

    // ---- set-up: pre-loop normals (value unused), metric / inverse / energy at the initial point ----
    normals_to(-1, zv);
    if (t < d) xprev[t] = a.x0[(a.broadcast_x0 ? 0 : chain * d) + t];
    rc_sync();
    metric.prepare(xprev, d);
    metric.template build<WIDE>(xprev, d, W, ld);
    rc_sync();
    store_global(Gacc);
    double logdet_prev = logdet_chol(metric, xprev);
    rc_invert<WIDE>(W, d, ld, piv, rcbuf, rtperm);
    for (int k = t; k < d * ld; k += RC_THREADS) Ainv0[k] = W[k];
    MC metric_prev = metric;   // the start-of-trajectory metric (Q17) of every draw until an accept replaces it
    target_at(xprev, true, false);
    double prev_U = (a.cons_term - sc_lp) + 0.5 * logdet_prev;
    rc_sync();

    int n_acc = 0;
    const int n_total = (int)(a.n_burnin + a.n_keep), n_burnin = (int)a.n_burnin;
    double* out_row = a.draws + chain * a.n_keep * d;
    double* out_lp = a.logp ? a.logp + chain * a.n_keep : nullptr;

    for (int it = 0; it < n_total; ++it) {
        normals_to(it, zv);
        if (warp == 0) {
            const double u = rng.uniform(a.rng, it, 0);
            if (lane == 0) sc_u = u;
        }
        // p = chol(G_prev) z with the Eigen matrixLLT storage quirk when asked for (Q8); K0 = p.(G_prev^-1 p)/2
        load_global(Gacc);
        rc_sync();
        rc_cholesky_inplace<WIDE>(W, d, ld);
        if (t < d) {
            double acc = 0.0;
            const int jmax = (a.chol_mode == MCMCB200_CHOL_EIGEN_LLT) ? d : t + 1;
            for (int j = 0; j < jmax; ++j) acc = fma(W[(size_t)j * ld + t], zv[j], acc);
            pv[t] = acc;
            xcur[t] = xprev[t];
        }
        rc_sync();
        gemv2(Ainv0, pv, tv, nullptr);
        const double prev_K = 0.5 * rc_block_sum(t < d ? pv[t] * tv[t] : 0.0, red);
        double logdet_new = logdet_prev;
        MC metric_new = metric_prev;
        bool have_new = false;

        for (int s = 0; s < a.n_leap; ++s) {
            if (t < d) qv[t] = pv[t];
            rc_sync();
            for (int kk = 0; kk < a.n_fp; ++kk) {   // momentum half step, fixed point, start-of-trajectory metric (Q16/Q17)
                mntm_update(xcur, qv, Ainv0, metric_prev, xprev, wv);
                if (t < d) qv[t] = pv[t] + wv[t];
                rc_sync();
            }
            if (t < d) { pv[t] = qv[t]; wv[t] = xcur[t]; }
            rc_sync();
            for (int kk = 0; kk < a.n_fp; ++kk) {   // position step, fixed point: w = x + (eps/2)(G_prev^-1 + G(w)^-1) p
                MC mw;
                mw.prepare(wv, d);
                mw.template build<WIDE>(wv, d, W, ld);
                rc_sync();
                rc_invert<WIDE>(W, d, ld, piv, rcbuf, rtperm);
                if (t < d) {
                    double acc = 0.0;
                    for (int j = 0; j < d; ++j) acc = fma(Ainv0[(size_t)j * ld + t] + W[(size_t)j * ld + t], heps * pv[j], acc);
                    tv[t] = xcur[t] + acc;
                }
                rc_sync();
                if (t < d) wv[t] = tv[t];
                rc_sync();
            }
            if (t < d) xcur[t] = wv[t];
            rc_sync();
            metric_new.prepare(xcur, d);
            metric_new.template build<WIDE>(xcur, d, W, ld);
            rc_sync();
            if (s + 1 == a.n_leap) {   // only the end point's metric can become the next draw's G_prev and enters the energy
                store_global(Gnew);
                logdet_new = logdet_chol(metric_new, xcur);
            }
            rc_invert<WIDE>(W, d, ld, piv, rcbuf, rtperm);
            have_new = true;
            mntm_update(xcur, pv, W, metric_new, xcur, wv);
            if (t < d) pv[t] = pv[t] + wv[t];
            rc_sync();
        }
        if (!have_new) {   // n_leap == 0: the "new" metric is the current one
            for (int k = t; k < d * ld; k += RC_THREADS) W[k] = Ainv0[k];
            for (int k = t; k < d * d; k += RC_THREADS) Gnew[k] = Gacc[k];
            rc_sync();
        }
        target_at(xcur, true, false);
        double prop_U = (a.cons_term - sc_lp) + 0.5 * logdet_new;
        if (!isfinite(prop_U)) prop_U = CUDART_INF;
        gemv2(W, pv, tv, nullptr);
        const double prop_K = 0.5 * rc_block_sum(t < d ? pv[t] * tv[t] : 0.0, red);
        const double comp = fmin(0.01, -(prop_U + prop_K) + (prev_U + prev_K));   // src/rmhmc.cpp:250 (min(0.01, NaN) = 0.01: a NaN energy accepts)
        const bool acc = sc_u < exp(comp);
        if (acc) {   // src/rmhmc.cpp:254-261
            if (t < d) xprev[t] = xcur[t];
            for (int k = t; k < d * ld; k += RC_THREADS) Ainv0[k] = W[k];
            prev_U = prop_U;
            logdet_prev = logdet_new;
            metric_prev = metric_new;
            double* tp = Gacc; Gacc = Gnew; Gnew = tp;
        }
        rc_sync();
        if (it >= n_burnin) {
            if (t < d) out_row[t] = xprev[t];
            out_row += d;
            if (out_lp) {
                if (t == 0) *out_lp = -(prev_U - a.cons_term);
                ++out_lp;
            }
            n_acc += acc ? 1 : 0;
        }
    }
    if (t == 0 && a.n_accept) a.n_accept[chain] = n_acc;
}

long long rmhmc_cta_work_doubles(int d) { return 2ll * d * d + (long long)d * ((d + 2) & ~1); }

bool rmhmc_cta_applicable(int target_id, int metric_id, int d, bool strict, bool has_bounds)
{
    if (strict || has_bounds || d < 2 || d > RC_MAXD_WIDE) return false;
    return target_id == MCMCB200_TARGET_FUNNEL && metric_id >= 0 && metric_id <= 2;
}

template <class T, class MC> static int launch_cta(const RmhmcLaunch& a)
{
    const int d = a.d, ld = (d + 2) & ~1, dp = (d + 1) & ~1;
    const size_t smem = ((size_t)d * ld + (size_t)16 * dp) * sizeof(double);
    {
        const char* e = std::getenv("MCMCB200_RMHMC_REGTILE");
        const int v = (e && e[0] == '0') ? 1 : 0;
        MCMCB200_CUDA_TRY(cudaMemcpyToSymbolAsync(rc_force_shared_gj, &v, sizeof(int), 0, cudaMemcpyHostToDevice, a.stream));

    }
    if (a.rng.mode == RNG_PHILOX) rc_build_rng_tab<<<1, 256, 0, a.stream>>>();
    auto launch = [&](auto kern) -> int {
        // static (tables, buffers: ~18 KB) + dynamic shared memory together exceed the 48 KB default long before the dynamic part alone does
        if (smem > 24 * 1024) MCMCB200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)a.n_chains, RC_THREADS, smem, a.stream>>>(a);
        MCMCB200_CUDA_TRY(cudaGetLastError());
        return MCMCB200_OK;
    };
    const char* ew = std::getenv("MCMCB200_RMHMC_WIDE");
    const bool wide = d > RC_MAXD || (ew && ew[0] == '1');
    if (a.rng.mode == RNG_PHILOX) return wide ? launch(rmhmc_cta_kernel<T, MC, RNG_PHILOX, true>) : launch(rmhmc_cta_kernel<T, MC, RNG_PHILOX, false>);
    return wide ? launch(rmhmc_cta_kernel<T, MC, RNG_TAPE, true>) : launch(rmhmc_cta_kernel<T, MC, RNG_TAPE, false>);
}

int launch_rmhmc_cta(const RmhmcLaunch& a)
{
    if (a.broadcast_x0 && a.n_chains > 0 && false) return MCMCB200_ERR_UNSUPPORTED;
    if (a.target_id == MCMCB200_TARGET_FUNNEL) return a.metric_id == 2 ? launch_cta<Funnel, FunnelSoftabsCta>(a) : launch_cta<Funnel, FunnelFisherCta>(a);
    set_error("rmhmc (CTA kernel): target %d has no contraction-form metric", a.target_id);
    return MCMCB200_ERR_UNSUPPORTED;
}

}  // namespace mcmcb200
