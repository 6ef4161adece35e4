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
//   straight from the integer bits: 512-bin table + degree-5 log1p for ln, MUFU seed + one cubic step for sqrt,
//   512-bin (cos, sin) table + a small-angle rotation for the angle — all accurate to a few ulp; the oracle
//   restates the formulas above with libm and agrees to <= 1e-14.
// TAPE (parity): a flat per-chain stream of doubles consumed in order — the reference's own variates,
//   replayed on the host (host_tape.cpp) or recorded by the caller.
#pragma once

#include <type_traits>

#include "rng_args.h"
#include "warp.cuh"

namespace mcmcb200
{

constexpr int LOG_TAB_BITS = 9;
constexpr int LOG_TAB_SIZE = 1 << LOG_TAB_BITS;  // double2 entries (1/c_i, -2 ln c_i)
constexpr int ANG_TAB_BITS = 9;
constexpr int ANG_TAB_SIZE = 1 << ANG_TAB_BITS;  // double2 entries (cos, sin) of the bin-centre angles
constexpr int RNG_TAB_DOUBLE2 = LOG_TAB_SIZE + ANG_TAB_SIZE;  // 16 KB of shared memory per CTA
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
    // q(r) = -2 log1p(r)/r, Horner from the r^4 term down     [0..4]      (|r| <= 2^-9: truncation r^6/3 <= 2e-17)
    -2.0 / 5.0, 2.0 / 4.0, -2.0 / 3.0, 1.0, -2.0,
    // [5] -2 ln 2   [6] 2^52 + 1076   [7] 2^-52   [8] -(2^52 + 2^42) 2^-52 = -(1 + 2^-10)
    -2.0 * 0.69314718055994530942, 4503599627370496.0 + 1076.0, 2.220446049250313e-16, -(1.0 + 0.0009765625),
    // with f in turns, |f| <= 2^-10, w = f^2:
    //   sin(2 pi f) = f (2 pi + w (S0 + w S1)),  S0 = -(2 pi)^3/6, S1 = (2 pi)^5/120     (next term 7e-20)     [9..11]
    //   cos(2 pi f) - 1 = w (C0 + w C1),         C0 = -(2 pi)^2/2, C1 = (2 pi)^4/24      (next term 7e-17)     [12..13]
    6.283185307179586476925, -41.341702240399760233968, 81.605249276075054203397,
    -19.739208802178717237669, 64.939394022668291490944,
    // sqrt refinement: [14] 0.375  [15] 0.5
    0.375, 0.5};

// Shared-memory tables, built once per CTA by its own threads:
//   tab[i], i < 512:       (A_i, B_i) with A_i ~ 1/c_i, B_i = 2 ln A_i, c_i the centre of mantissa bin i of [1,2)
//                          (anchored to exactly 1 and 2 in the first / last bin so ln u stays accurate as u -> 1);
//   tab[512 + j], j < 512: (cos, sin) of the angle 2 pi ((2j + 1)/1024 + 2^-53): the centre of angle bin j plus the
//                          half-step of the (k2 + 1/2) grid, so that the in-bin offset (g - 2^42) 2^-52 is an exact
//                          double obtained with ONE fma from the raw bits.
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
        // angle / pi = (2j + 1)/512 + 2^-52, exactly representable (< 2, 53 significant bits)
        sincospi((2 * j + 1) * (1.0 / ANG_TAB_SIZE) + 2.220446049250313e-16, &sn, &cs);
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
        // angle: k2 = (r2:r3) >> 12; top 9 bits pick the table bin, the low 43 bits g give the offset from its centre
        const unsigned k2_hi = r[2] >> 12, k2_lo = __funnelshift_r(r[3], r[2], 12);
        cssn = tab[LOG_TAB_SIZE + (k2_hi >> (20 - ANG_TAB_BITS))];
        dl = __hiloint2double(0x43300000 | (k2_hi & ((1u << (20 - ANG_TAB_BITS)) - 1u)), k2_lo);   // 2^52 + g, g < 2^43
    }
};

// Evaluate NP (1 or 2) staged pairs in lock-step: every constant is fetched once for all of them.
// Per pair 26 fp64 instructions: ln 8, sqrt 5, angle 7, rotation + scaling 6.  Lout (optional) receives
// L = -2 ln u1 = z0^2 + z1^2 up to rounding, so a caller with M = I gets the kinetic energy for free.
template <int NP> __device__ __forceinline__ void bm_eval(BmPair (&b)[NP], double (&z0)[NP], double (&z1)[NP], double* Lout = nullptr)
{
    double L[NP], R[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        b[i].ed = b[i].ed - RNG_C[6];                     // exponent - 1023 - 53
        b[i].rr = fma(b[i].m, b[i].le.x, -1.0);           // m / c_i - 1, |rr| <= 2^-9
        b[i].dl = fma(b[i].dl, 2.220446049250313e-16, -(1.0 + 0.0009765625));       // f = (g - 2^42) 2^-52 turns, exact, |f| <= 2^-10
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) b[i].q = RNG_C[0];
#pragma unroll
    for (int i = 0; i < NP; ++i) b[i].q = fma(b[i].q, b[i].rr, 0.5);          // 0.5, 1.0, -2.0: DFMA immediates
#pragma unroll
    for (int i = 0; i < NP; ++i) b[i].q = fma(b[i].q, b[i].rr, RNG_C[2]);
#pragma unroll
    for (int i = 0; i < NP; ++i) b[i].q = fma(b[i].q, b[i].rr, 1.0);
#pragma unroll
    for (int i = 0; i < NP; ++i) b[i].q = fma(b[i].q, b[i].rr, -2.0);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        b[i].t2 = fma(b[i].ed, RNG_C[5], b[i].le.y);       // -2 (E ln2 + ln c_i)
        L[i] = fma(b[i].rr, b[i].q, b[i].t2);              // -2 ln u1
        b[i].z = b[i].dl * b[i].dl;
        if (Lout) Lout[i] = L[i];
    }
    // R = sqrt(L) = t + t e (1/2 + 3/8 e), t = L y, e = 1 - t y, y = MUFU.RSQ64H seed: one cubically convergent step
    // (relative error ~ e^3 <= 2^-60), no slow path: L is a normal number in [2^-52, 75).
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(L[i]));
        const double t = L[i] * y;
        const double e = fma(-t, y, 1.0);
        const double pe = fma(e, 0.375, 0.5);
        const double te = t * e;
        R[i] = fma(te, pe, t);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        b[i].sd = fma(b[i].z, RNG_C[11], RNG_C[10]);        // S0 + w S1
        b[i].cd = fma(b[i].z, RNG_C[13], RNG_C[12]);        // C0 + w C1
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        b[i].sd = fma(b[i].sd, b[i].z, RNG_C[9]);           // 2 pi + w (S0 + w S1)
        b[i].cd = b[i].cd * b[i].z;                         // cos(d) - 1
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const double sdv = b[i].sd * b[i].dl;               // sin(d)
        const double RC = R[i] * b[i].cssn.x, RS = R[i] * b[i].cssn.y;
        // R cos(a + d) = RC + RC cdm - RS sdv,  R sin(a + d) = RS + RS cdm + RC sdv
        z0[i] = fma(RC, b[i].cd, fma(-RS, sdv, RC));
        z1[i] = fma(RS, b[i].cd, fma(RC, sdv, RS));
    }
}

template <int I, int N, class F> __device__ __forceinline__ void static_for(F&& f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>());
        static_for<I + 1, N>(f);
    }
}

// The same NP Box-Muller pairs as ChainRng::normals (identical arithmetic, identical results), cut into 20 units of
// work that a caller interleaves by hand with other straight-line code (hmc_pipe_kernel: one slice per leapfrog step),
// so that the integer Philox rounds and the short polynomial chains of the NEXT draw sit between the dependent DFMAs
// of the CURRENT trajectory in program order.  Units 0-9: Philox round r of all pairs; 10-11: bits -> staged operands
// (I2F, table loads; first / second half of the pairs), spare bits; 12-15: ln; 16: sqrt; 17-18: small-angle
// polynomials; 19: rotation + scaling.  slice<S, NS>() runs the units whose cumulative issue cost (UNIT_COST, in
// dispatch cycles: fp64 = 2, everything else = 1) starts inside the S-th of NS equal shares.
template <int NP> struct BmPipe {
    static constexpr int N_UNITS = 20;
    static constexpr int UNIT_COST[N_UNITS] = {8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 17, 17, 12, 8, 12, 10, 22, 8, 8, 24};
    static constexpr int cost_before(int u)
    {
        int c = 0;
        for (int i = 0; i < u; ++i) c += UNIT_COST[i];
        return c;
    }
    unsigned c[NP][4];
    BmPair b[NP];
    double L[NP], R[NP];
    unsigned spare;

    __device__ __forceinline__ void begin(int lane, long long draw, unsigned chain)
    {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            c[i][0] = static_cast<unsigned>(i * 32 + lane);
            c[i][1] = static_cast<unsigned>(draw + 1);
            c[i][2] = chain;
            c[i][3] = 0u;
        }
    }

    // z: lane-striped normals (2*NP slots); lsum: sum over pairs of -2 ln u1
    template <int U> __device__ __forceinline__ void unit(const RngArgs& a, const double2* __restrict__ tab, double (&z)[2 * NP], double& lsum)
    {
        constexpr int H = (NP + 1) / 2;   // pairs staged by unit 10; the rest by unit 11
        if constexpr (U < 10) {
            constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const unsigned hi0 = __umulhi(M0, c[i][0]), lo0 = M0 * c[i][0];
                const unsigned hi1 = __umulhi(M1, c[i][2]), lo1 = M1 * c[i][2];
                c[i][0] = hi1 ^ c[i][1] ^ a.rk[2 * U];
                c[i][1] = lo1;
                c[i][2] = hi0 ^ c[i][3] ^ a.rk[2 * U + 1];
                c[i][3] = lo0;
            }
        } else if constexpr (U == 10) {
            spare = ((c[0][1] & 0xfffu) << 12) | (c[0][3] & 0xfffu);
#pragma unroll
            for (int i = 0; i < H; ++i) b[i].setup(c[i], tab);
        } else if constexpr (U == 11) {
#pragma unroll
            for (int i = H; i < NP; ++i) b[i].setup(c[i], tab);
        } else if constexpr (U == 12) {
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                b[i].ed = b[i].ed - RNG_C[6];
                b[i].rr = fma(b[i].m, b[i].le.x, -1.0);
                b[i].dl = fma(b[i].dl, 2.220446049250313e-16, -(1.0 + 0.0009765625));
            }
        } else if constexpr (U == 13) {
#pragma unroll
            for (int i = 0; i < NP; ++i) b[i].q = fma(RNG_C[0], b[i].rr, 0.5);
#pragma unroll
            for (int i = 0; i < NP; ++i) b[i].q = fma(b[i].q, b[i].rr, RNG_C[2]);
        } else if constexpr (U == 14) {
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                b[i].q = fma(b[i].q, b[i].rr, 1.0);
                b[i].t2 = fma(b[i].ed, RNG_C[5], b[i].le.y);
                b[i].z = b[i].dl * b[i].dl;
            }
        } else if constexpr (U == 15) {
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                b[i].q = fma(b[i].q, b[i].rr, -2.0);
                L[i] = fma(b[i].rr, b[i].q, b[i].t2);
                lsum = (i == 0) ? L[0] : lsum + L[i];
            }
        } else if constexpr (U == 16) {
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                double y;
                asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(L[i]));
                const double t = L[i] * y;
                const double e = fma(-t, y, 1.0);
                const double pe = fma(e, 0.375, 0.5);
                const double te = t * e;
                R[i] = fma(te, pe, t);
            }
        } else if constexpr (U == 17) {
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                b[i].sd = fma(b[i].z, RNG_C[11], RNG_C[10]);
                b[i].cd = fma(b[i].z, RNG_C[13], RNG_C[12]);
            }
        } else if constexpr (U == 18) {
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                b[i].sd = fma(b[i].sd, b[i].z, RNG_C[9]);
                b[i].cd = b[i].cd * b[i].z;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const double sdv = b[i].sd * b[i].dl;
                const double RC = R[i] * b[i].cssn.x, RS = R[i] * b[i].cssn.y;
                z[2 * i] = fma(RC, b[i].cd, fma(-RS, sdv, RC));
                z[2 * i + 1] = fma(RS, b[i].cd, fma(RC, sdv, RS));
            }
        }
    }

    // slice S of NS: the units whose cumulative cost starts in [S, S+1) * total / NS
    template <int S, int NS> __device__ __forceinline__ void slice(const RngArgs& a, const double2* __restrict__ tab, double (&z)[2 * NP], double& lsum)
    {
        constexpr int total = cost_before(N_UNITS);
        static_for<0, N_UNITS>([&](auto u) {
            constexpr int U = decltype(u)::value;
            constexpr int at = cost_before(U) * NS;
            if constexpr (at >= S * total && at < (S + 1) * total) this->template unit<U>(a, tab, z, lsum);
        });
    }

    // uniform #0 of the draw from the spare bits of blocks 0 and 1 (valid after unit 10); same definition as ChainRng::uniform
    __device__ __forceinline__ double uniform0() const
    {
        const unsigned s0 = __shfl_sync(FULL, spare, 0), s1 = __shfl_sync(FULL, spare, 1);
        const double sd = __hiloint2double(0x43300000 | (s0 >> 8), (s0 << 24) | s1) - 4503599627370496.0;
        return fma(sd, 3.5527136788005009e-15, 1.7763568394002505e-15);
    }
};

// per-chain RNG cursor; MODE is RNG_PHILOX or RNG_TAPE (compile time)
template <int MODE> struct ChainRng {
    unsigned chain;       // global chain id (Philox counter word 2)
    unsigned spare;       // Philox: 24 spare bits of this lane's block m = 0 of the current draw
    const double* tape;   // tape mode: this chain's stream
    long long cursor;
    long long limit;      // tape mode: doubles available to this chain (reads past it return a constant and raise a.err_flag)

    __device__ __forceinline__ void init(const RngArgs& a, long long local_chain, long long global_chain)
    {
        chain = static_cast<unsigned>(global_chain);
        spare = 0;
        tape = (MODE == RNG_TAPE) ? a.tape + local_chain * a.tape_stride : nullptr;
        cursor = 0;
        limit = a.tape_stride;
    }

    // d standard normals into the lane-striped vector z (FT: d == 32*EPL, no padding slots).
    // When several warps share a chain, each warp generates its own segment: q_base = first Box-Muller pair of the
    // segment, seg_off = first element of the segment, d_total = n_dim of the chain (d is the segment's length).
    template <int EPL, bool FT>
    __device__ __forceinline__ void normals(const RngArgs& a, long long draw, int d, int lane, const double2* __restrict__ tab,
                                            double (&z)[EPL], int q_base = 0, int seg_off = 0, int d_total = -1, double* lsum = nullptr)
    {
        // lsum (Philox, full tiles only): receives sum over this lane's pairs of L = -2 ln u1 (= sum z^2 up to rounding)
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
                    double Lp[2];
                    bm_eval<2>(b, z0, z1, lsum ? Lp : nullptr);
                    if (lsum) *lsum = (m0 == 0) ? Lp[0] + Lp[1] : *lsum + (Lp[0] + Lp[1]);
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
                    double Lp[1];
                    bm_eval<1>(b, z0, z1, lsum ? Lp : nullptr);
                    if (lsum) *lsum = (m0 == 0) ? Lp[0] : *lsum + Lp[0];
                    z[2 * m0] = (FT || 2 * q < d) ? z0[0] : 0.0;
                    z[2 * m0 + 1] = (FT || 2 * q + 1 < d) ? z1[0] : 0.0;
                }
            }
        } else {
            const long long adv = (d_total < 0) ? d : d_total;
            if (cursor + adv > limit) {   // tape exhausted (only reachable with data-dependent consumption: host checks static counts)
                if (a.err_flag) *a.err_flag = 1;
#pragma unroll
                for (int k = 0; k < EPL; ++k) z[k] = 0.0;
            } else {
                const double* t = tape + cursor + seg_off;
#pragma unroll
                for (int k = 0; k < EPL; ++k) {
                    const int j = elem_index(lane, k);
                    z[k] = (j < d) ? t[j] : 0.0;
                }
            }
            cursor += adv;
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
        if (cursor >= limit) {
            if (a.err_flag) *a.err_flag = 1;
            ++cursor;
            return 0.5;
        }
        return tape[cursor++];
    }

    // Random access to the k-th uniform of the draw without moving the cursor (NUTS resolves its theta' selection lazily,
    // out of order).  Philox is counter-based anyway; in tape mode the uniforms of a draw sit at ubase + k, ubase = the
    // cursor right after the draw's normals.  The caller sets cursor = ubase + (uniforms consumed) at the end of the draw.
    __device__ __forceinline__ double uniform_at(const RngArgs& a, long long draw, int k, long long ubase)
    {
        if (MODE == RNG_PHILOX) return uniform(a, draw, k);
        const long long pos = ubase + k;
        if (pos >= limit) {
            if (a.err_flag) *a.err_flag = 1;
            return 0.5;
        }
        return tape[pos];
    }
};

}  // namespace mcmcb200
