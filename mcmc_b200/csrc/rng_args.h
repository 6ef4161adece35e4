// Plain (host-visible) description of the RNG source a kernel launch uses; see rng.cuh.
#pragma once

namespace mcmcb200
{

enum { RNG_PHILOX = 0, RNG_TAPE = 1 };

struct RngArgs {
    int mode;
    unsigned k0, k1;        // Philox key = (seed_lo, seed_hi)
    const double* tape;     // [n_chains][tape_stride]
    long long tape_stride;
};

}  // namespace mcmcb200
