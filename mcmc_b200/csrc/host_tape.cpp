// Host side of MCMCB200_RNG_MT19937_TAPE: replay the reference's random stream.
//
// The reference seeds one std::mt19937_64 per call (src/hmc.cpp:68, rand_engine_t =
// std::mt19937_64, include/misc/mcmc_options.hpp:101) and draws
//   * normals with bmo::stats::rnorm: a NEW std::normal_distribution<fp_t> for every variate
//     (include/BaseMatrixOps/include/stats/rnorm.hpp:57-59, vector form :120-128), so the
//     distribution's cached second variate is always thrown away;
//   * uniforms with bmo::stats::runif: a_adj = nextafter(0, 1), then
//     std::uniform_real_distribution<fp_t>(a_adj, 1) (stats/runif.hpp:58-63).
// That stream is serial and standard-library specific, so it cannot be generated on the
// GPU bit-for-bit; for drop-in parity the library generates, per chain, exactly the variates
// the reference would consume and the kernels read them in order:
//   [n_pre_normals normals]  then per draw  [d normals][1 uniform]
// (n_pre_normals = d for NUTS / RM-HMC; 0 for HMC / MALA).
#include <cmath>
#include <random>
#include <thread>
#include <vector>

#include "engine.h"

namespace mcmcb200
{

void host_mt19937_tape(uint64_t seed, long long n_pre_normals, long long n_draws, int d, double* out)
{
    std::mt19937_64 engine(seed);
    long long pos = 0;
    auto one_normal = [&]() {
        std::normal_distribution<double> dist(0.0, 1.0);
        return 0.0 + 1.0 * dist(engine);
    };
    const double lo = std::nextafter(0.0, 1.0);
    for (long long i = 0; i < n_pre_normals; ++i) out[pos++] = one_normal();
    for (long long t = 0; t < n_draws; ++t) {
        for (int j = 0; j < d; ++j) out[pos++] = one_normal();
        std::uniform_real_distribution<double> ud(lo, 1.0);
        out[pos++] = ud(engine);
    }
}

// mcmc::de's stream (src/de.cpp:92-99: the master engine seeds ONE per-thread engine through generate_seed_value,
// include/stats/seed_values.hpp:26-30; single-threaded member loop), in consumption order: n_pop*d initial uniforms, then
// per generation and member: c1 and c2 (rind with its rejection loops, stats/rind.hpp:34-37; the ACCEPTED indices are
// recorded, as doubles), d proposal uniforms in (-b, b), z.
static double bmo_runif(std::mt19937_64& eng, double a, double b)
{
    const double a_adj = std::nextafter(a, b);   // stats/runif.hpp:58-63
    std::uniform_real_distribution<double> ud(a_adj, b);
    return ud(eng);
}
void host_de_tape(uint64_t seed, long long n_pop, int d, long long n_gen, double par_b, double* out)
{
    std::mt19937_64 master(seed), eng;
    const double u0 = bmo_runif(master, 0.0, 1.0);
    eng.seed(static_cast<size_t>((u0 + 0 + 1) * 1000));   // generate_seed_value(0, 1, rand_engine)
    long long pos = 0;
    for (long long k = 0; k < n_pop * d; ++k) out[pos++] = bmo_runif(eng, 0.0, 1.0);
    for (long long g = 0; g < n_gen; ++g)
        for (long long i = 0; i < n_pop; ++i) {
            long long c1, c2;
            do { c1 = (long long)static_cast<size_t>(bmo_runif(eng, 0.0, double(n_pop - 1) + 1.0)); } while (c1 == i);
            do { c2 = (long long)static_cast<size_t>(bmo_runif(eng, 0.0, double(n_pop - 1) + 1.0)); } while (c2 == i || c2 == c1);
            out[pos++] = (double)c1;
            out[pos++] = (double)c2;
            for (int j = 0; j < d; ++j) out[pos++] = bmo_runif(eng, -par_b, par_b);
            out[pos++] = bmo_runif(eng, 0.0, 1.0);
        }
}

struct HostMtStreams {
    std::vector<std::mt19937_64> eng;
};

HostMtStreams* host_mt_streams_create(uint64_t seed0, long long n_chains)
{
    HostMtStreams* s = new HostMtStreams;
    s->eng.reserve((size_t)n_chains);
    for (long long c = 0; c < n_chains; ++c) s->eng.emplace_back(seed0 + (uint64_t)c);
    return s;
}
void host_mt_streams_destroy(HostMtStreams* s) { delete s; }

template <class F> static void over_chains(long long C, F&& f)
{
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 16) nt = 16;
    if ((long long)nt > C) nt = (unsigned)C;
    if (nt <= 1) { for (long long c = 0; c < C; ++c) f(c); return; }
    std::vector<std::thread> th;
    for (unsigned ti = 0; ti < nt; ++ti)
        th.emplace_back([&, ti]() { for (long long c = ti; c < C; c += nt) f(c); });
    for (auto& t : th) t.join();
}

void host_mt_streams_fill(HostMtStreams* s, long long n_normals, long long pool, double* out)
{
    const long long stride = n_normals + pool;
    const double lo = std::nextafter(0.0, 1.0);
    over_chains((long long)s->eng.size(), [&](long long c) {
        std::mt19937_64& e = s->eng[(size_t)c];
        double* o = out + c * stride;
        for (long long i = 0; i < n_normals; ++i) {
            std::normal_distribution<double> dist(0.0, 1.0);   // a fresh distribution per variate (stats/rnorm.hpp:57-59)
            o[i] = 0.0 + 1.0 * dist(e);
        }
        std::mt19937_64 look = e;   // the uniforms the chain MAY consume next; the engine itself advances by the count used
        for (long long i = 0; i < pool; ++i) {
            std::uniform_real_distribution<double> ud(lo, 1.0);
            o[n_normals + i] = ud(look);
        }
    });
}

void host_mt_streams_advance(HostMtStreams* s, const long long* used)
{
    // bmo::stats::runif draws one std::uniform_real_distribution<double> value = one raw 64-bit output of mt19937_64
    // (generate_canonical<double, 53> needs a single call of a 64-bit engine)
    over_chains((long long)s->eng.size(), [&](long long c) { s->eng[(size_t)c].discard((unsigned long long)used[c]); });
}

}  // namespace mcmcb200
