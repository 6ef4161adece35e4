// In-kernel random variates for the sampler kernels.
//
// PHILOX (production): Philox4x32-10, key = (seed_lo, seed_hi), counter = (index, draw+1, chain, stream).
//   stream 0, index q: one 128-bit block (r0,r1,r2,r3) -> the Box-Muller pair for elements (2q, 2q+1):
//       k1 = (r0:r1) >> 12  (52 bits)    u1 = (k1 + 1/2) 2^-52  in (0,1)     R = sqrt(-2 ln u1)
//       k2 = (r2:r3) >> 12  (52 bits)    phi = 2 pi (k2 + 1/2) 2^-52   in (0, 2 pi)
//       z[2q] = R cos(phi)               z[2q+1] = R sin(phi)
//     the 24 unused bits (r1 & 0xfff, r3 & 0xfff) of blocks q=0 and q=1 form the 48-bit integer
//       s = (r1_q0 & 0xfff) << 36 | (r3_q0 & 0xfff) << 24 | (r1_q1 & 0xfff) << 12 | (r3_q1 & 0xfff)
//     and uniform #0 of the draw is (s + 1/2) 2^-48 (blocks 0 and 1 are always generated).
//   stream 1, index k >= 1: uniform #k of the draw = (((r0:r1) >> 12) + 1/2) 2^-52   (NUTS only).
//   draw = -1 is the pre-loop draw of NUTS / RM-HMC.
//   This replaces bmo::stats::rnorm_vec_inplace / runif (include/BaseMatrixOps/include/stats/rnorm.hpp:120-128,
//   runif.hpp:93-99), whose std::mt19937_64 stream is inherently serial.  ln, sqrt, sin and cos are evaluated
//   straight from the integer bits: 256-bin table + degree-7 log1p for ln, MUFU seed + one cubic step for sqrt,
//   256-bin (cos, sin) table + a small-angle rotation for the angle — all accurate to a few ulp; the oracle
//   restates the formulas above with libm and agrees to <= 1e-14.
// TAPE (parity): a flat per-chain stream of doubles consumed in order — the reference's own variates,
//   replayed on the host (host_tape.cpp) or recorded by the caller.
#pragma once

#include "rng_args.h"
#include "warp.cuh"

namespace mcmcb200
{

constexpr int LOG_TAB_BITS = 8;
constexpr int LOG_TAB_SIZE = 1 << LOG_TAB_BITS;  // double2 entries (1/c_i, -2 ln c_i)
constexpr int ANG_TAB_BITS = 8;
constexpr int ANG_TAB_SIZE = 1 << ANG_TAB_BITS;  // double2 entries (cos, sin) of the bin-centre angles
constexpr int RNG_TAB_DOUBLE2 = LOG_TAB_SIZE + ANG_TAB_SIZE;  // 8 KB of shared memory per CTA
constexpr double LN2 = 0.69314718055994530942;

// Round keys come precomputed in the kernel-parameter (constant) bank.
__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, const RngArgs& a, unsigned (&out)[4])
{
    constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const unsigned hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        c0 = hi1 ^ c1 ^ a.rk[2 * r];
        c1 = lo1;
        c2 = hi0 ^ c3 ^ a.rk[2 * r + 1];
        c3 = lo0;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Polynomial coefficients live in constant memory (loaded once per use into a uniform register and shared by
// the two Box-Muller pairs that are evaluated in lock-step).
static __constant__ double RNG_C[16] = {
    // q(r) = -2 log1p(r)/r, Horner from the r^6 term down     [0..6]
    -2.0 / 7.0, 2.0 / 6.0, -2.0 / 5.0, 2.0 / 4.0, -2.0 / 3.0, 1.0, -2.0,
    // [7] -2 ln 2   [8] 2^52 + 1076   [9] 2 pi 2^-52   [10] 2^52 + 2^43
    -2.0 * 0.69314718055994530942, 4503599627370496.0 + 1076.0, 6.283185307179586476925 * 2.220446049250313e-16,
    4503599627370496.0 + 8796093022208.0,
    // sin(t)/t - 1 ~ t^2 (S0 + t^2 (S1 + t^2 S2)),  cos(t) - 1 ~ t^2 (C0 + t^2 (C1 + t^2 C2)) for |t| <= pi/256   [11..15]
    -1.0 / 6.0, 1.0 / 120.0, -0.5, 1.0 / 24.0, -1.0 / 720.0};

// Shared-memory tables, built once per CTA by its own threads:
//   tab[i], i < 256:       (A_i, B_i) with A_i ~ 1/c_i, B_i = 2 ln A_i, c_i the centre of mantissa bin i of [1,2)
//                          (anchored to exactly 1 and 2 in the first / last bin so ln u stays accurate as u -> 1);
//   tab[256 + j], j < 256: (cos, sin) of the centre of angle bin j, (j + 1/2) 2 pi / 256.
__device__ __forceinline__ void build_rng_tables(double2* tab)
{
    for (int i = threadIdx.x; i < LOG_TAB_SIZE; i += blockDim.x) {
        double A, B2;
        if (i == 0) {
            A = 1.0;
            B2 = 0.0;
        } else if (i == LOG_TAB_SIZE - 1) {
            A = 0.5;
            B2 = -2.0 * LN2;  // -2 ln 2 with the same rounded constant the kernel multiplies by
        } else {
            const double c = 1.0 + (i + 0.5) * (1.0 / LOG_TAB_SIZE);
            A = 1.0 / c;
            B2 = 2.0 * log(A);
        }
        tab[i] = make_double2(A, B2);
    }
    for (int j = threadIdx.x; j < ANG_TAB_SIZE; j += blockDim.x) {
        double sn, cs;
        sincospi((2 * j + 1) * (1.0 / ANG_TAB_SIZE), &sn, &cs);  // angle / pi = (j + 1/2) * 2 / 256
        tab[LOG_TAB_SIZE + j] = make_double2(cs, sn);
    }
}

// One Box-Muller pair from a Philox block, staged so that two pairs can be advanced in lock-step.
struct BmPair {
    double m, ed, rr, q, t2;   // log part
    double2 le;                 // log table entry
    double2 cssn;               // angle table entry
    double dl, z, sd, cd;       // small-angle part
    __device__ __forceinline__ void setup(const unsigned (&r)[4], const double2* __restrict__ tab)
    {
        // u1 = n 2^-53, n = 2 k1 + 1 (odd, < 2^53), k1 = (r0:r1) >> 12
        const unsigned n_hi = r[0] >> 11;
        const unsigned n_lo = __funnelshift_r(r[1], r[0], 11) | 1u;
        const double nd = static_cast<double>((static_cast<unsigned long long>(n_hi) << 32) | n_lo);  // exact
        const unsigned h = static_cast<unsigned>(__double2hiint(nd));
        m = __hiloint2double((h & 0x000fffffu) | 0x3ff00000u, __double2loint(nd));  // mantissa in [1,2)
        le = tab[(h >> (20 - LOG_TAB_BITS)) & (LOG_TAB_SIZE - 1)];
        ed = __hiloint2double(0x43300000, static_cast<int>(h >> 20));              // 2^52 + biased exponent
        // angle: k2 = (r2:r3) >> 12; top 8 bits pick the table bin, the low 44 bits g give the offset from its centre
        const unsigned k2_hi = r[2] >> 12, k2_lo = __funnelshift_r(r[3], r[2], 12);
        cssn = tab[LOG_TAB_SIZE + (k2_hi >> (20 - ANG_TAB_BITS))];
        dl = __hiloint2double(0x43300000 | (k2_hi & 0xfffu), k2_lo);               // 2^52 + g, g < 2^44
    }
};

// Evaluate NP (1 or 2) staged pairs in lock-step: every constant is fetched once for all of them.
template <int NP> __device__ __forceinline__ void bm_eval(BmPair (&b)[NP], double (&z0)[NP], double (&z1)[NP])
{
    double L[NP], R[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        b[i].ed = b[i].ed - RNG_C[8];                     // exponent - 1023 - 53
        b[i].rr = fma(b[i].m, b[i].le.x, -1.0);           // m / c_i - 1, |rr| <= 2^-8
        b[i].dl = b[i].dl - RNG_C[10];                    // g - 2^43, exact
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) b[i].q = RNG_C[0];
#pragma unroll
    for (int c = 1; c <= 6; ++c) {
#pragma unroll
        for (int i = 0; i < NP; ++i) b[i].q = fma(b[i].q, b[i].rr, RNG_C[c]);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        b[i].t2 = fma(b[i].ed, RNG_C[7], b[i].le.y);       // -2 (E ln2 + ln c_i)
        b[i].dl = b[i].dl * RNG_C[9] + 0.5 * RNG_C[9];     // delta = 2 pi (g - 2^43 + 1/2) 2^-52, |delta| <= pi/256
        L[i] = fma(b[i].rr, b[i].q, b[i].t2);              // -2 ln u1
        b[i].z = b[i].dl * b[i].dl;
    }
    // R = sqrt(L): MUFU.RSQ64H seed, one cubically convergent step (relative error ~ e^3 <= 2^-60), no slow path:
    // L is a normal number in [2^-52, 75).
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(L[i]));
        const double t = L[i] * y;
        const double e = fma(-t, y, 1.0);
        const double pe = fma(e, 0.375, 0.5);
        const double ye = y * e;
        y = fma(ye, pe, y);
        R[i] = L[i] * y;
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        b[i].sd = fma(b[i].z, RNG_C[12], RNG_C[11]);        // sin(d)/d - 1 = z (S0 + z S1) (|d| <= 0.0123: next term 2e-18)
        b[i].cd = fma(b[i].z, RNG_C[15], RNG_C[14]);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        b[i].cd = fma(b[i].z, b[i].cd, RNG_C[13]);          // cos(d) - 1 = z (C0 + z (C1 + z C2))
        b[i].sd = b[i].sd * b[i].z;
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const double cdm = b[i].cd * b[i].z;                // cos(d) - 1
        const double sdv = fma(b[i].sd, b[i].dl, b[i].dl);  // sin(d)
        const double C = b[i].cssn.x, S = b[i].cssn.y;
        // cos(a + d) = C + (C cdm - S sdv),  sin(a + d) = S + (S cdm + C sdv)
        const double cph = C + fma(C, cdm, -S * sdv);
        const double sph = S + fma(S, cdm, C * sdv);
        z0[i] = R[i] * cph;
        z1[i] = R[i] * sph;
    }
}

// per-chain RNG cursor; MODE is RNG_PHILOX or RNG_TAPE (compile time)
template <int MODE> struct ChainRng {
    unsigned chain;       // global chain id (Philox counter word 2)
    unsigned spare;       // Philox: 24 spare bits of this lane's block m = 0 of the current draw
    const double* tape;   // tape mode: this chain's stream
    long long cursor;

    __device__ __forceinline__ void init(const RngArgs& a, long long local_chain, long long global_chain)
    {
        chain = static_cast<unsigned>(global_chain);
        spare = 0;
        tape = (MODE == RNG_TAPE) ? a.tape + local_chain * a.tape_stride : nullptr;
        cursor = 0;
    }

    // d standard normals into the lane-striped vector z (FT: d == 32*EPL, no padding slots).
    // When several warps share a chain, each warp generates its own segment: q_base = first Box-Muller pair of the
    // segment, seg_off = first element of the segment, d_total = n_dim of the chain (d is the segment's length).
    template <int EPL, bool FT>
    __device__ __forceinline__ void normals(const RngArgs& a, long long draw, int d, int lane, const double2* __restrict__ tab,
                                            double (&z)[EPL], int q_base = 0, int seg_off = 0, int d_total = -1)
    {
        if (MODE == RNG_PHILOX) {
            constexpr int NPAIR = EPL / 2;
#pragma unroll
            for (int m0 = 0; m0 < NPAIR; m0 += 2) {
                constexpr int dummy = 0;
                (void)dummy;
                if (m0 + 1 < NPAIR) {
                    BmPair b[2];
                    double z0[2], z1[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int q = (m0 + i) * 32 + lane;
                        unsigned r[4];
                        philox4x32_10(static_cast<unsigned>(q_base + q), static_cast<unsigned>(draw + 1), chain, 0u, a, r);
                        if (m0 + i == 0) spare = ((r[1] & 0xfffu) << 12) | (r[3] & 0xfffu);
                        b[i].setup(r, tab);
                    }
                    bm_eval<2>(b, z0, z1);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int q = (m0 + i) * 32 + lane;
                        z[2 * (m0 + i)] = (FT || 2 * q < d) ? z0[i] : 0.0;
                        z[2 * (m0 + i) + 1] = (FT || 2 * q + 1 < d) ? z1[i] : 0.0;
                    }
                } else {
                    BmPair b[1];
                    double z0[1], z1[1];
                    const int q = m0 * 32 + lane;
                    unsigned r[4];
                    philox4x32_10(static_cast<unsigned>(q_base + q), static_cast<unsigned>(draw + 1), chain, 0u, a, r);
                    if (m0 == 0) spare = ((r[1] & 0xfffu) << 12) | (r[3] & 0xfffu);
                    b[0].setup(r, tab);
                    bm_eval<1>(b, z0, z1);
                    z[2 * m0] = (FT || 2 * q < d) ? z0[0] : 0.0;
                    z[2 * m0 + 1] = (FT || 2 * q + 1 < d) ? z1[0] : 0.0;
                }
            }
        } else {
            const double* t = tape + cursor + seg_off;
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                const int j = elem_index(lane, k);
                z[k] = (j < d) ? t[j] : 0.0;
            }
            cursor += (d_total < 0) ? d : d_total;
        }
    }

    // k-th uniform of the draw (warp-uniform result); k == 0 must follow normals() of the same draw
    __device__ __forceinline__ double uniform(const RngArgs& a, long long draw, int k)
    {
        if (MODE == RNG_PHILOX) {
            if (k == 0) {
                const unsigned s0 = __shfl_sync(FULL, spare, 0), s1 = __shfl_sync(FULL, spare, 1);
                // 2^52 + (s0 << 24 | s1), then exact subtraction
                const double sd = __hiloint2double(0x43300000 | (s0 >> 8), (s0 << 24) | s1) - 4503599627370496.0;
                return fma(sd, 3.5527136788005009e-15, 1.7763568394002505e-15);  // (s + 1/2) 2^-48
            }
            unsigned r[4];
            philox4x32_10(static_cast<unsigned>(k), static_cast<unsigned>(draw + 1), chain, 1u, a, r);
            const unsigned hi = r[0] >> 12, lo = __funnelshift_r(r[1], r[0], 12);
            const double kd = __hiloint2double(0x43300000 | hi, lo) - 4503599627370496.0;
            return fma(kd, 2.220446049250313e-16, 1.1102230246251565e-16);  // (k + 1/2) 2^-52
        }
        return tape[cursor++];
    }
};

}  // namespace mcmcb200
