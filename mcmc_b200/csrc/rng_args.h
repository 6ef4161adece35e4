// Plain (host-visible) description of the RNG source a kernel launch uses; see rng.cuh.
#pragma once

namespace mcmcb200
{

enum { RNG_PHILOX = 0, RNG_TAPE = 1 };

struct RngArgs {
    int mode;
    unsigned k0, k1;        // Philox key = (seed_lo, seed_hi)
    unsigned rk[20];        // the 10 round keys (k0 + r*W0, k1 + r*W1): read straight from the constant bank by LOP3
    const double* tape;     // [n_chains][tape_stride]
    long long tape_stride;
    int* err_flag;          // device int, set to 1 by a chain whose tape cursor would pass tape_stride (checked after the run)
};

inline void rng_set_key(RngArgs& a, unsigned long long seed)
{
    a.k0 = (unsigned)(seed & 0xffffffffull);
    a.k1 = (unsigned)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        a.rk[2 * r] = a.k0 + (unsigned)r * 0x9E3779B9u;
        a.rk[2 * r + 1] = a.k1 + (unsigned)r * 0xBB67AE85u;
    }
}

}  // namespace mcmcb200
