// In-kernel random variates for the sampler kernels.
//
// PHILOX (production): Philox4x32-10, key = (seed_lo, seed_hi), counter =
//   (index, draw+1, chain, stream).  stream 0: Box-Muller pair q -> N(0,1) for elements
//   (2q, 2q+1); stream 1: the k-th U(0,1) of the draw.  draw = -1 is the pre-loop draw of
//   NUTS / RM-HMC.  Replaces bmo::stats::rnorm_vec_inplace / runif
//   (include/BaseMatrixOps/include/stats/rnorm.hpp:120-128, runif.hpp:93-99), whose
//   std::mt19937_64 stream is inherently serial.
// TAPE (parity): a flat per-chain stream of doubles consumed in order — the reference's
//   own variates, replayed on the host (host_tape.cpp) or recorded by the caller.
#pragma once

#include "rng_args.h"
#include "warp.cuh"

namespace mcmcb200
{

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

// 53 random bits -> (k + 1/2) 2^-53 in the open interval (0,1)
__device__ __forceinline__ double u53_open(unsigned hi, unsigned lo)
{
    const unsigned long long k = ((static_cast<unsigned long long>(hi) << 32) | lo) >> 11;
    return (static_cast<double>(k) + 0.5) * 1.1102230246251565404e-16;
}

// per-chain RNG cursor
struct ChainRng {
    unsigned chain;             // global chain id (Philox counter word 2)
    const double* tape;         // this chain's tape
    long long cursor;

    __device__ __forceinline__ void init(const RngArgs& a, long long local_chain, long long global_chain)
    {
        chain = static_cast<unsigned>(global_chain);
        tape = (a.mode == RNG_TAPE) ? a.tape + local_chain * a.tape_stride : nullptr;
        cursor = 0;
    }

    // d standard normals into the lane-striped vector z
    template <int EPL> __device__ __forceinline__ void normals(const RngArgs& a, long long draw, int d, int lane, double (&z)[EPL])
    {
        if (a.mode == RNG_PHILOX) {
#pragma unroll
            for (int m = 0; m < EPL / 2; ++m) {
                const int q = m * 32 + lane;
                unsigned r[4];
                philox4x32_10(static_cast<unsigned>(q), static_cast<unsigned>(draw + 1), chain, 0u, a.k0, a.k1, r);
                const double u1 = u53_open(r[0], r[1]), u2 = u53_open(r[2], r[3]);
                const double rad = sqrt(-2.0 * log(u1));
                double s, c;
                sincospi(2.0 * u2, &s, &c);
                z[2 * m] = (2 * q < d) ? rad * c : 0.0;
                z[2 * m + 1] = (2 * q + 1 < d) ? rad * s : 0.0;
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

    // k-th uniform of the draw (warp-uniform result)
    __device__ __forceinline__ double uniform(const RngArgs& a, long long draw, int k)
    {
        if (a.mode == RNG_PHILOX) {
            unsigned r[4];
            philox4x32_10(static_cast<unsigned>(k), static_cast<unsigned>(draw + 1), chain, 1u, a.k0, a.k1, r);
            return u53_open(r[0], r[1]);
        }
        return tape[cursor++];
    }
};

}  // namespace mcmcb200
