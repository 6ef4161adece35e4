// TEST INFRASTRUCTURE ONLY (oracle/).  Thin extern "C" driver around the
// UNMODIFIED reference samplers.  It is compiled together with
// /root/reference/src/{hmc,mala,nuts,rmhmc,rwmh,de}.cpp (where they lie; nothing is copied
// into this repo) against the stand-in Eigen header — see oracle/Makefile —
// into oracle/_ref/libmcmc_ref_{strict,fast}.so.
//
//   strict : -O2 -ffp-contract=off                    -> parity comparator
//   fast   : the reference's release flags (configure:196-214, with a portable
//            -march) + -fopenmp                       -> CPU baseline timing
//
// Entry points call mcmc::hmc / mala / nuts / rmhmc exactly as a user would
// (examples/eigen/hmc_normal.cpp:108 etc.), with the callbacks of
// host_targets.hpp.  The many-chain harness is the OpenMP loop described in
// BASELINE.md §3 (one mcmc::X call per chain, seed = seed_base + chain).
#ifndef MCMC_ENABLE_EIGEN_WRAPPERS
#define MCMC_ENABLE_EIGEN_WRAPPERS
#endif
#include "mcmc.hpp"

#include <chrono>
#include <cstdint>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "host_targets.hpp"

namespace
{

struct tgt_ctx_t {
    int target_id;
    const double* data;
    int d;
    int metric_id;
};

double log_kernel_cb(const mcmc::ColVec_t& vals, mcmc::ColVec_t* grad_out, void* ctx_v)
{
    const tgt_ctx_t* ctx = static_cast<const tgt_ctx_t*>(ctx_v);
    if (grad_out) {
        grad_out->resize(ctx->d);
        return otgt::value_and_grad(ctx->target_id, ctx->data, vals.data(), grad_out->data(), ctx->d, otgt::SUM_SEQ);
    }
    return otgt::value_and_grad(ctx->target_id, ctx->data, vals.data(), nullptr, ctx->d, otgt::SUM_SEQ);
}

// value-only callback of mcmc::rwmh (include/mcmc/rwmh.hpp:46)
double log_kernel_val_cb(const mcmc::ColVec_t& vals, void* ctx_v)
{
    const tgt_ctx_t* ctx = static_cast<const tgt_ctx_t*>(ctx_v);
    return otgt::value_and_grad(ctx->target_id, ctx->data, vals.data(), nullptr, ctx->d, otgt::SUM_SEQ);
}

mcmc::Mat_t tensor_cb(const mcmc::ColVec_t& vals, mcmc::Cube_t* deriv_out, void* ctx_v)
{
    const tgt_ctx_t* ctx = static_cast<const tgt_ctx_t*>(ctx_v);
    const int d = ctx->d;
    mcmc::Mat_t G(d, d);
    if (deriv_out) {
        std::vector<double> dG(size_t(d) * d * d);
        otgt::metric(ctx->target_id, ctx->metric_id, ctx->data, vals.data(), d, G.data(), dG.data());
        deriv_out->setZero(d, d, d);
        for (int i = 0; i < d; ++i)
            for (int k = 0; k < d * d; ++k) deriv_out->mat(i).data()[k] = dG[size_t(i) * d * d + k];
    } else {
        otgt::metric(ctx->target_id, ctx->metric_id, ctx->data, vals.data(), d, G.data(), nullptr);
    }
    return G;
}

void fill_precond(mcmc::Mat_t& m, const double* precond_colmajor, int d)
{
    if (!precond_colmajor) return;
    m.resize(d, d);
    for (int k = 0; k < d * d; ++k) m.data()[k] = precond_colmajor[k];
}

// draws (n_keep x d, column-major Mat_t, Q23) -> row-major [n_keep][d]
void copy_draws(const mcmc::Mat_t& draws, double* out, long n_keep, int d)
{
    if (!out) return;
    for (long t = 0; t < n_keep; ++t)
        for (int j = 0; j < d; ++j) out[t * d + j] = draws(t, j);
}

}  // namespace

extern "C" {

// sampler ids shared by the entry points below
enum { REF_HMC = 0, REF_MALA = 1, REF_NUTS = 2, REF_RMHMC = 3, REF_RWMH = 4 };

struct ref_settings_t {
    long n_burnin, n_keep;
    long n_leap_steps;     // hmc, rmhmc
    double step_size;      // all (nuts: eps_bar_0; rwmh: par_scale)
    const double* precond; // d*d column-major or null (hmc, mala, nuts; rwmh: cov_mat)
    long n_fp_steps;       // rmhmc
    long n_adapt_draws;    // nuts
    double target_accept_rate, gamma_val, t0_val, kappa_val;  // nuts
    long max_tree_depth;   // nuts
    int use_nuts_defaults; // 1 -> keep the struct defaults for the five nuts fields above
    int vals_bound;        // algo_settings_t::vals_bound
    const double* lower;   // d entries (+-inf = unbounded side) when vals_bound
    const double* upper;
    int metric_id;         // rmhmc: metric registered with the target (0 = default)
};

int ref_run_chain(int sampler, int target_id, const double* tdata, int d, const double* x0, const ref_settings_t* st,
                  unsigned long seed, double* draws_out, long* n_accept)
{
    tgt_ctx_t ctx = {target_id, tdata, d, st->metric_id};
    mcmc::ColVec_t init(d);
    for (int j = 0; j < d; ++j) init(j) = x0[j];

    mcmc::algo_settings_t s;
    s.rng_seed_value = seed;
    if (st->vals_bound) {
        s.vals_bound = true;
        s.lower_bounds.resize(d);
        s.upper_bounds.resize(d);
        for (int j = 0; j < d; ++j) { s.lower_bounds(j) = st->lower[j]; s.upper_bounds(j) = st->upper[j]; }
    }
    mcmc::Mat_t draws;
    bool ok = false;
    long acc = 0;

    switch (sampler) {
    case REF_HMC:
        s.hmc_settings.n_burnin_draws = size_t(st->n_burnin);
        s.hmc_settings.n_keep_draws = size_t(st->n_keep);
        s.hmc_settings.n_leap_steps = size_t(st->n_leap_steps);
        s.hmc_settings.step_size = st->step_size;
        fill_precond(s.hmc_settings.precond_mat, st->precond, d);
        ok = mcmc::hmc(init, log_kernel_cb, draws, &ctx, s);
        acc = long(s.hmc_settings.n_accept_draws);
        break;
    case REF_MALA:
        s.mala_settings.n_burnin_draws = size_t(st->n_burnin);
        s.mala_settings.n_keep_draws = size_t(st->n_keep);
        s.mala_settings.step_size = st->step_size;
        fill_precond(s.mala_settings.precond_mat, st->precond, d);
        ok = mcmc::mala(init, log_kernel_cb, draws, &ctx, s);
        acc = long(s.mala_settings.n_accept_draws);
        break;
    case REF_NUTS:
        s.nuts_settings.n_burnin_draws = size_t(st->n_burnin);
        s.nuts_settings.n_keep_draws = size_t(st->n_keep);
        s.nuts_settings.step_size = st->step_size;
        if (!st->use_nuts_defaults) {
            s.nuts_settings.n_adapt_draws = size_t(st->n_adapt_draws);
            s.nuts_settings.target_accept_rate = st->target_accept_rate;
            s.nuts_settings.gamma_val = st->gamma_val;
            s.nuts_settings.t0_val = st->t0_val;
            s.nuts_settings.kappa_val = st->kappa_val;
            s.nuts_settings.max_tree_depth = size_t(st->max_tree_depth);
        }
        fill_precond(s.nuts_settings.precond_mat, st->precond, d);
        ok = mcmc::nuts(init, log_kernel_cb, draws, &ctx, s);
        acc = long(s.nuts_settings.n_accept_draws);
        break;
    case REF_RMHMC:
        s.rmhmc_settings.n_burnin_draws = size_t(st->n_burnin);
        s.rmhmc_settings.n_keep_draws = size_t(st->n_keep);
        s.rmhmc_settings.n_leap_steps = size_t(st->n_leap_steps);
        s.rmhmc_settings.step_size = st->step_size;
        s.rmhmc_settings.n_fp_steps = size_t(st->n_fp_steps);
        ok = mcmc::rmhmc(init, log_kernel_cb, tensor_cb, draws, &ctx, &ctx, s);
        acc = long(s.rmhmc_settings.n_accept_draws);
        break;
    case REF_RWMH:
        s.rwmh_settings.n_burnin_draws = size_t(st->n_burnin);
        s.rwmh_settings.n_keep_draws = size_t(st->n_keep);
        s.rwmh_settings.par_scale = st->step_size;
        fill_precond(s.rwmh_settings.cov_mat, st->precond, d);
        ok = mcmc::rwmh(init, log_kernel_val_cb, draws, &ctx, s);
        acc = long(s.rwmh_settings.n_accept_draws);
        break;
    default:
        return -1;
    }
    copy_draws(draws, draws_out, st->n_keep, d);
    if (n_accept) *n_accept = acc;
    return ok ? 0 : 1;
}

// Many chains: OpenMP loop over chains, one reference call per chain
// (seed = seed_base + c, x0s chain-major [C][d], draws_out [C][n_keep][d] or null).
// Returns wall seconds spent in the loop via *elapsed_s.
int ref_run_chains(int sampler, int target_id, const double* tdata, int d, long n_chains, const double* x0s,
                   const ref_settings_t* st, unsigned long seed_base, double* draws_out, long* n_accept,
                   int n_threads, double* elapsed_s)
{
    int bad = 0;
    const auto t0 = std::chrono::steady_clock::now();
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic) reduction(+ : bad)
#endif
    for (long c = 0; c < n_chains; ++c) {
        long acc = 0;
        double* out = draws_out ? draws_out + size_t(c) * size_t(st->n_keep) * size_t(d) : nullptr;
        std::vector<double> scratch;
        if (!out) {
            scratch.resize(size_t(st->n_keep) * size_t(d));
            out = scratch.data();
        }
        bad += ref_run_chain(sampler, target_id, tdata, d, x0s + size_t(c) * size_t(d), st, seed_base + (unsigned long)c,
                             out, &acc);
        if (n_accept) n_accept[c] = acc;
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
    return bad;
}

// mcmc::de (src/de.cpp:30-271): one population of n_pop members, single-threaded (omp_n_threads = 1: the member loop updates X
// in place, so its sequential semantics are the deterministic ones).  draws_out: [n_keep][n_pop][d] (Cube_t: one n_pop x d
// matrix per kept generation).  init_lb / init_ub: d entries or null (defaults initial_vals -+ 0.5, src/de.cpp:70-71).
struct ref_de_settings_t {
    long n_pop, n_burnin, n_keep;
    int jumps;
    double par_b, par_gamma_jump;
    const double* init_lb;
    const double* init_ub;
    int vals_bound;
    const double* lower;
    const double* upper;
};

int ref_run_de(int target_id, const double* tdata, int d, const double* x0, const ref_de_settings_t* st, unsigned long seed,
               double* draws_out, long* n_accept)
{
    tgt_ctx_t ctx = {target_id, tdata, d, 0};
    mcmc::ColVec_t init(d);
    for (int j = 0; j < d; ++j) init(j) = x0[j];
    mcmc::algo_settings_t s;
    s.rng_seed_value = seed;
    if (st->vals_bound) {
        s.vals_bound = true;
        s.lower_bounds.resize(d);
        s.upper_bounds.resize(d);
        for (int j = 0; j < d; ++j) { s.lower_bounds(j) = st->lower[j]; s.upper_bounds(j) = st->upper[j]; }
    }
    s.de_settings.n_pop = size_t(st->n_pop);
    s.de_settings.n_burnin_draws = size_t(st->n_burnin);
    s.de_settings.n_keep_draws = size_t(st->n_keep);
    s.de_settings.jumps = st->jumps != 0;
    s.de_settings.par_b = st->par_b;
    s.de_settings.par_gamma_jump = st->par_gamma_jump;
    s.de_settings.omp_n_threads = 1;
    if (st->init_lb) { s.de_settings.initial_lb.resize(d); for (int j = 0; j < d; ++j) s.de_settings.initial_lb(j) = st->init_lb[j]; }
    if (st->init_ub) { s.de_settings.initial_ub.resize(d); for (int j = 0; j < d; ++j) s.de_settings.initial_ub(j) = st->init_ub[j]; }
    mcmc::Cube_t draws;
    const bool ok = mcmc::de(init, log_kernel_val_cb, draws, &ctx, s);
    if (draws_out)
        for (long t = 0; t < st->n_keep; ++t)
            for (long i = 0; i < st->n_pop; ++i)
                for (int j = 0; j < d; ++j) draws_out[(size_t(t) * st->n_pop + i) * d + j] = draws.mat(t)(i, j);
    if (n_accept) *n_accept = long(s.de_settings.n_accept_draws);
    return ok ? 0 : 1;
}

int ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// The BMO RNG primitives themselves: n normals then m uniforms from engine(seed).
void ref_rng_stream(unsigned long seed, long n_norm, long n_unif, double* out)
{
    mcmc::rand_engine_t eng(seed);
    for (long i = 0; i < n_norm; ++i) out[i] = bmo::stats::rnorm<double>(eng);
    for (long i = 0; i < n_unif; ++i) out[n_norm + i] = bmo::stats::runif<double>(eng);
}

}  // extern "C"
