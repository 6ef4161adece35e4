// In-kernel random variates for the sampler kernels.
//
// PHILOX (production): Philox4x32-10, key = (seed_lo, seed_hi), counter = (index, draw+1, chain, stream).
//   stream 0, index q: one 128-bit block (r0,r1,r2,r3) -> the Box-Muller pair for elements (2q, 2q+1):
//       k1 = (r0:r1) >> 12  (52 bits)    u1 = (k1 + 1/2) 2^-52  in (0,1)     R = sqrt(-2 ln u1)
//       k2 = (r2:r3) >> 12  (52 bits)    f = k2 mod 2^50, b0 = bit 50, b1 = bit 51 of k2
//       phi = (pi/2) (f + 1/2) 2^-50 in (0, pi/2)
//       z[2q] = (-1)^b0 R cos(phi)       z[2q+1] = (-1)^b1 R sin(phi)
//     the 24 unused bits (r1 & 0xfff, r3 & 0xfff) of blocks q=0 and q=1 form the 48-bit integer
//       s = (r1_q0 & 0xfff) << 36 | (r3_q0 & 0xfff) << 24 | (r1_q1 & 0xfff) << 12 | (r3_q1 & 0xfff)
//     and uniform #0 of the draw is (s + 1/2) 2^-48 (blocks 0 and 1 are always generated).
//   stream 1, index k >= 1: uniform #k of the draw = (((r0:r1) >> 12) + 1/2) 2^-52   (NUTS only).
//   draw = -1 is the pre-loop draw of NUTS / RM-HMC.
//   This replaces bmo::stats::rnorm_vec_inplace / runif (include/BaseMatrixOps/include/stats/rnorm.hpp:120-128,
//   runif.hpp:93-99), whose std::mt19937_64 stream is inherently serial.  ln, sin and cos are evaluated by
//   range-reduced polynomials straight from the integer bits (table-driven log, fdlibm-style kernels on
//   |theta| <= pi/4 followed by an exact pi/4 rotation), accurate to ~1 ulp — the oracle restates the
//   formulas above with libm and agrees to <= 1e-14.
// TAPE (parity): a flat per-chain stream of doubles consumed in order — the reference's own variates,
//   replayed on the host (host_tape.cpp) or recorded by the caller.
#pragma once

#include "rng_args.h"
#include "warp.cuh"

namespace mcmcb200
{

constexpr int LOG_TAB_BITS = 8;
constexpr int LOG_TAB_SIZE = 1 << LOG_TAB_BITS;  // double2 entries: 4 KB of shared memory per CTA
constexpr double LN2 = 0.69314718055994530942;

__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1,
                                              unsigned (&out)[4])
{
    constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const unsigned hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Table for ln(m), m in [1,2): entry i = (A_i, B_i) with A_i ~ 1/c_i, B_i = -2 * (-ln A_i) = 2 ln A_i, c_i the centre of
// bin i (anchored to exactly 1 and 2 in the first / last bin so that ln u stays relatively accurate as u -> 1).
// Built once per CTA by its own threads.
__device__ __forceinline__ void build_log_table(double2* tab)
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
            B2 = 2.0 * log(A);  // -2 ln(1/A)
        }
        tab[i] = make_double2(A, B2);
    }
}

// (r0..r3) -> two independent N(0,1) variates, per the definition in the header comment.
__device__ __forceinline__ void normal_pair_from_bits(const unsigned (&r)[4], const double2* __restrict__ tab, double& z0, double& z1)
{
    // ---- L = -2 ln u1, u1 = n 2^-53, n = 2 k1 + 1 (odd, < 2^53) ----
    const unsigned n_hi = r[0] >> 11;                               // bits 52..32 of n
    const unsigned n_lo = __funnelshift_r(r[1], r[0], 11) | 1u;     // bits 31..0 of n  (= (k1 << 1) | 1)
    const double nd = static_cast<double>((static_cast<unsigned long long>(n_hi) << 32) | n_lo);  // exact
    const unsigned h = static_cast<unsigned>(__double2hiint(nd));
    const double m = __hiloint2double((h & 0x000fffffu) | 0x3ff00000u, __double2loint(nd));      // mantissa in [1,2)
    const double2 e = tab[(h >> (20 - LOG_TAB_BITS)) & (LOG_TAB_SIZE - 1)];
    // exponent of nd as a double, minus 53:  (h >> 20) - 1023 - 53
    const double ed = __hiloint2double(0x43300000, static_cast<int>(h >> 20)) - (4503599627370496.0 + 1076.0);
    const double rr = fma(m, e.x, -1.0);                            // m / c_i - 1, |rr| <= 2^-8
    double q = 2.0 / 7.0;                                           // -2 log1p(rr) = rr * q(rr)
    q = fma(q, rr, -2.0 / 6.0);
    q = fma(q, rr, 2.0 / 5.0);
    q = fma(q, rr, -2.0 / 4.0);
    q = fma(q, rr, 2.0 / 3.0);
    q = fma(q, rr, -1.0);
    q = fma(q, rr, 2.0);
    q = -q;
    const double t2 = fma(ed, -2.0 * LN2, e.y);                     // -2 (E ln2 + ln c_i)
    const double L = fma(rr, q, t2);
    const double R2 = sqrt(0.5 * L);                                // R / sqrt(2)

    // ---- angle: theta = phi - pi/4 = 2 pi (f - 2^49 + 1/2) 2^-52, |theta| <= pi/4 ----
    const unsigned k2_hi = r[2] >> 12;                              // 20 bits: b1 b0 f[49:32]
    const unsigned k2_lo = __funnelshift_r(r[3], r[2], 12);
    const double fd = __hiloint2double(0x43300000 | (k2_hi & 0x3ffffu), k2_lo);   // 2^52 + f
    const double sI = fd - (4503599627370496.0 + 562949953421312.0);              // f - 2^49, exact
    constexpr double C2PI = 6.283185307179586476925 * 2.220446049250313e-16;       // 2 pi 2^-52
    const double th = fma(sI, C2PI, 0.5 * C2PI);
    const double z = th * th;
    double ps = 1.58969099521155010221e-10;
    ps = fma(ps, z, -2.50507602534068634195e-08);
    ps = fma(ps, z, 2.75573137070700676789e-06);
    ps = fma(ps, z, -1.98412698298579493134e-04);
    ps = fma(ps, z, 8.33333333332248946124e-03);
    ps = fma(ps, z, -1.66666666666666324348e-01);
    const double sn = fma(th * z, ps, th);
    double pc = -1.13596475577881948265e-11;
    pc = fma(pc, z, 2.08757232129817482790e-09);
    pc = fma(pc, z, -2.75573143513906633035e-07);
    pc = fma(pc, z, 2.48015872894767294178e-05);
    pc = fma(pc, z, -1.38888888888741095749e-03);
    pc = fma(pc, z, 4.16666666666666019037e-02);
    const double cs = fma(z * z, pc, fma(z, -0.5, 1.0));
    // rotate by pi/4: sqrt(2) cos(phi) = cs - sn, sqrt(2) sin(phi) = cs + sn; signs from bits 50, 51 of k2
    const double a0 = R2 * (cs - sn), a1 = R2 * (cs + sn);
    z0 = __hiloint2double(__double2hiint(a0) ^ static_cast<int>((k2_hi << 13) & 0x80000000u), __double2loint(a0));
    z1 = __hiloint2double(__double2hiint(a1) ^ static_cast<int>((k2_hi << 12) & 0x80000000u), __double2loint(a1));
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

    // d standard normals into the lane-striped vector z
    template <int EPL>
    __device__ __forceinline__ void normals(const RngArgs& a, long long draw, int d, int lane, const double2* __restrict__ tab,
                                            double (&z)[EPL])
    {
        if (MODE == RNG_PHILOX) {
#pragma unroll
            for (int m = 0; m < EPL / 2; ++m) {
                const int q = m * 32 + lane;
                unsigned r[4];
                philox4x32_10(static_cast<unsigned>(q), static_cast<unsigned>(draw + 1), chain, 0u, a.k0, a.k1, r);
                if (m == 0) spare = ((r[1] & 0xfffu) << 12) | (r[3] & 0xfffu);
                double z0, z1;
                normal_pair_from_bits(r, tab, z0, z1);
                z[2 * m] = (2 * q < d) ? z0 : 0.0;
                z[2 * m + 1] = (2 * q + 1 < d) ? z1 : 0.0;
            }
        } else {
            const double* t = tape + cursor;
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                const int j = elem_index(lane, k);
                z[k] = (j < d) ? t[j] : 0.0;
            }
            cursor += d;
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
            philox4x32_10(static_cast<unsigned>(k), static_cast<unsigned>(draw + 1), chain, 1u, a.k0, a.k1, r);
            const unsigned hi = r[0] >> 12, lo = __funnelshift_r(r[1], r[0], 12);
            const double kd = __hiloint2double(0x43300000 | hi, lo) - 4503599627370496.0;
            return fma(kd, 2.220446049250313e-16, 1.1102230246251565e-16);  // (k + 1/2) 2^-52
        }
        return tape[cursor++];
    }
};

}  // namespace mcmcb200
