// C-ABI layer of the engine (include/mcmc_b200.h): argument checking, one-time host
// linear algebra on the preconditioner (what src/hmc.cpp:57-59 does with BMO_MATOPS_INV /
// BMO_MATOPS_CHOL_LOWER), host<->device marshalling, reference-stream tape generation,
// kernel dispatch and timing.  No CPU sampling path exists here: if CUDA is unavailable
// every run call fails with MCMCB200_ERR_CUDA.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "engine.h"

namespace mcmcb200
{

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int epl_for_dim(int d)
{
    if (d <= 0) return 0;
    if (d <= 64) return 2;
    if (d <= 128) return 4;
    if (d <= 256) return 8;
    if (d <= 512) return 16;
    return 0;
}

// ---- grow-only device scratch, per (device, slot) -----------------------------------------
enum Slot { SLOT_TDATA = 0, SLOT_LB, SLOT_UB, SLOT_X0, SLOT_DRAWS, SLOT_LOGP, SLOT_NACC, SLOT_TAPE, SLOT_MAT_A, SLOT_MAT_B, SLOT_MAT_C,
            SLOT_STEP, SLOT_NLF, SLOT_WORK, SLOT_EVAL_X, SLOT_EVAL_V, SLOT_EVAL_G, SLOT_COUNT };
constexpr int MAX_DEVICES = 16;
// The scratch belongs to the calling HOST THREAD: like the reference's samplers (no globals, src/hmc.cpp) the run calls
// are re-entrant — several host threads may sample at the same time, on the same or on different devices, each with its
// own buffers.  A thread's buffers are freed when it exits (or by mcmcb200_release_workspace()).
struct Buf { void* p = nullptr; size_t bytes = 0; };
struct Pool {
    Buf b[MAX_DEVICES][SLOT_COUNT];
    ~Pool()
    {
        for (int d = 0; d < MAX_DEVICES; ++d)
            for (int s = 0; s < SLOT_COUNT; ++s)
                if (b[d][s].p) cudaFree(b[d][s].p);   // errors at process teardown (runtime already unloading) are harmless
    }
};
static thread_local Pool g_pool;

static int pool_get(int dev, Slot s, size_t bytes, void** out)
{
    Buf& b = g_pool.b[dev][s];
    if (bytes == 0) bytes = 8;
    if (b.bytes < bytes) {
        if (b.p) cudaFree(b.p);
        b.p = nullptr;
        b.bytes = 0;
        MCMCB200_CUDA_TRY(cudaMalloc(&b.p, bytes));
        b.bytes = bytes;
    }
    *out = b.p;
    return MCMCB200_OK;
}

struct DeviceScope {
    int prev = -1, dev = -1;
    bool changed = false;
    int enter(int want)
    {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n <= 0) {
            set_error("no usable CUDA device (%s); mcmc_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
            return MCMCB200_ERR_CUDA;
        }
        MCMCB200_CUDA_TRY(cudaGetDevice(&prev));
        dev = (want < 0) ? prev : want;
        if (dev >= n || dev >= MAX_DEVICES) {
            set_error("device %d out of range (%d visible)", dev, n);
            return MCMCB200_ERR_INVALID_ARG;
        }
        if (dev != prev) {
            MCMCB200_CUDA_TRY(cudaSetDevice(dev));
            changed = true;
        }
        return MCMCB200_OK;
    }
    ~DeviceScope()
    {
        if (changed) cudaSetDevice(prev);
    }
};

static int64_t target_data_len(int target_id, int d)
{
    if (d <= 0) return -1;
    switch (target_id) {
    case MCMCB200_TARGET_ISO_GAUSS: return 0;
    case MCMCB200_TARGET_DIAG_GAUSS: return d;
    case MCMCB200_TARGET_DENSE_GAUSS: return (int64_t)d * d;
    case MCMCB200_TARGET_LINREG: return (int64_t)d * d + d;
    case MCMCB200_TARGET_NORMAL_MODEL: return d == 2 ? 3 : -1;
    case MCMCB200_TARGET_FUNNEL: return d >= 2 ? 0 : -1;
    default: return -1;
    }
}

// Everything a run needs on the device, resolved from the host-facing structs.
struct Staged {
    DeviceScope scope;
    cudaStream_t stream = nullptr;
    CommonLaunch c{};
    double* draws_host = nullptr;
    double* logp_host = nullptr;
    long long n_keep = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ~Staged()
    {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
    }
};

static int upload(int dev, Slot slot, const double* host, size_t n, cudaStream_t st, const double** dev_out)
{
    void* p = nullptr;
    int rc = pool_get(dev, slot, n * sizeof(double), &p);
    if (rc) return rc;
    if (n) MCMCB200_CUDA_TRY(cudaMemcpyAsync(p, host, n * sizeof(double), cudaMemcpyHostToDevice, st));
    *dev_out = static_cast<const double*>(p);
    return MCMCB200_OK;
}

// tape_per_chain > 0: doubles of reference-stream tape each chain needs in MT19937 mode (0 = sampler cannot use it)
static int stage_common(Staged& s, const mcmcb200_problem_t* pr, const mcmcb200_rng_t* rng, int arith, long long n_burnin,
                        long long n_keep, long long n_pre_normals, bool mt_tape_supported, mcmcb200_output_t* out, int max_dim = 32 * MAX_EPL)
{
    if (!pr || !rng || !out) { set_error("null argument"); return MCMCB200_ERR_INVALID_ARG; }
    if (pr->n_chains <= 0 || pr->n_dim <= 0) { set_error("n_chains and n_dim must be positive"); return MCMCB200_ERR_INVALID_ARG; }
    if (n_burnin < 0 || n_keep < 0 || n_burnin + n_keep > 0x7ffffff0ll) { set_error("draw counts out of range"); return MCMCB200_ERR_INVALID_ARG; }
    if (!pr->initial_vals) { set_error("initial_vals is null"); return MCMCB200_ERR_INVALID_ARG; }
    if (n_keep > 0 && !out->draws_out) { set_error("draws_out is null"); return MCMCB200_ERR_INVALID_ARG; }
    const int d = pr->n_dim;
    const int64_t need = target_data_len(pr->target_id, d);
    if (need < 0) {
        set_error("unknown target id %d (or n_dim=%d invalid for it)", pr->target_id, d);
        return MCMCB200_ERR_UNKNOWN_TARGET;
    }
    if (pr->target_data_len < need || (need > 0 && !pr->target_data)) {
        set_error("target %d needs %lld doubles of data, got %lld", pr->target_id, (long long)need, (long long)pr->target_data_len);
        return MCMCB200_ERR_INVALID_ARG;
    }
    if (d > max_dim) {
        set_error("n_dim=%d exceeds what this sampler/target combination supports (max %d)", d, max_dim);
        return MCMCB200_ERR_UNSUPPORTED;
    }
    int rc = s.scope.enter(pr->device);
    if (rc) return rc;
    const int dev = s.scope.dev;
    s.stream = static_cast<cudaStream_t>(pr->stream);
    cudaStream_t st = s.stream;
    CommonLaunch& c = s.c;
    c.n_chains = pr->n_chains;
    c.d = d;
    c.target_id = pr->target_id;
    c.chain_offset = pr->chain_offset;
    c.stream = st;
    c.strict = (arith == MCMCB200_ARITH_STRICT);
    c.broadcast_x0 = pr->broadcast_initial ? 1 : 0;
    s.n_keep = n_keep;

    if ((rc = upload(dev, SLOT_TDATA, pr->target_data, (size_t)need, st, &c.tdata))) return rc;
    c.lb = c.ub = nullptr;
    if (pr->vals_bound) {
        if (!pr->lower_bounds || !pr->upper_bounds) { set_error("vals_bound is set but lower_bounds / upper_bounds is null"); return MCMCB200_ERR_INVALID_ARG; }
        for (int j = 0; j < d; ++j)
            if (pr->lower_bounds[j] != pr->lower_bounds[j] || pr->upper_bounds[j] != pr->upper_bounds[j] || !(pr->lower_bounds[j] < pr->upper_bounds[j])) {
                set_error("bounds of element %d are not ordered (lower < upper required)", j);
                return MCMCB200_ERR_INVALID_ARG;
            }
        if ((rc = upload(dev, SLOT_LB, pr->lower_bounds, (size_t)d, st, &c.lb))) return rc;
        if ((rc = upload(dev, SLOT_UB, pr->upper_bounds, (size_t)d, st, &c.ub))) return rc;
    }

    const size_t n_x0 = (size_t)(c.broadcast_x0 ? 1 : pr->n_chains) * d;
    if (pr->initial_mem == MCMCB200_MEM_DEVICE) c.x0 = pr->initial_vals;
    else if ((rc = upload(dev, SLOT_X0, pr->initial_vals, n_x0, st, &c.x0))) return rc;

    // RNG
    const long long n_total = n_burnin + n_keep;
    rng_set_key(c.rng, rng->seed);
    c.rng.tape = nullptr;
    c.rng.tape_stride = 0;
    if (rng->mode == MCMCB200_RNG_PHILOX) {
        c.rng.mode = RNG_PHILOX;
    } else if (rng->mode == MCMCB200_RNG_MT19937_TAPE) {
        if (!mt_tape_supported) {
            set_error("MT19937 tape mode needs a static variate count per draw; this sampler consumes a data-dependent number "
                      "(use PHILOX or USER_TAPE)");
            return MCMCB200_ERR_UNSUPPORTED;
        }
        c.rng.mode = RNG_TAPE;
        const long long stride = n_pre_normals + n_total * (d + 1);
        std::vector<double> tape((size_t)pr->n_chains * (size_t)stride);
        const long long C = pr->n_chains;
        unsigned nt = std::thread::hardware_concurrency();
        if (nt == 0) nt = 1;
        if ((long long)nt > C) nt = (unsigned)C;
        std::vector<std::thread> th;
        for (unsigned ti = 0; ti < nt; ++ti)
            th.emplace_back([&, ti]() {
                for (long long ch = ti; ch < C; ch += nt)
                    host_mt19937_tape(rng->seed + (uint64_t)(pr->chain_offset + ch), n_pre_normals, n_total, d,
                                      tape.data() + (size_t)ch * (size_t)stride);
            });
        for (auto& t : th) t.join();
        if ((rc = upload(dev, SLOT_TAPE, tape.data(), tape.size(), st, &c.rng.tape))) return rc;
        MCMCB200_CUDA_TRY(cudaStreamSynchronize(st));  // `tape` goes out of scope
        c.rng.tape_stride = stride;
    } else if (rng->mode == MCMCB200_RNG_USER_TAPE) {
        if (!rng->tape || rng->tape_stride <= 0) { set_error("USER_TAPE needs tape and tape_stride"); return MCMCB200_ERR_INVALID_ARG; }
        c.rng.mode = RNG_TAPE;
        c.rng.tape_stride = rng->tape_stride;
        if (rng->tape_mem == MCMCB200_MEM_DEVICE) c.rng.tape = rng->tape;
        else if ((rc = upload(dev, SLOT_TAPE, rng->tape, (size_t)pr->n_chains * (size_t)rng->tape_stride, st, &c.rng.tape))) return rc;
    } else {
        set_error("unknown rng mode %d", rng->mode);
        return MCMCB200_ERR_INVALID_ARG;
    }

    // outputs
    const size_t n_draws = (size_t)pr->n_chains * (size_t)n_keep * d;
    if (out->draws_mem == MCMCB200_MEM_DEVICE) {
        c.draws = out->draws_out;
        c.logp = out->logp_out;
    } else {
        void* p = nullptr;
        if ((rc = pool_get(dev, SLOT_DRAWS, n_draws * sizeof(double), &p))) return rc;
        c.draws = static_cast<double*>(p);
        s.draws_host = out->draws_out;
        if (out->logp_out) {
            if ((rc = pool_get(dev, SLOT_LOGP, (size_t)pr->n_chains * n_keep * sizeof(double), &p))) return rc;
            c.logp = static_cast<double*>(p);
            s.logp_host = out->logp_out;
        } else {
            c.logp = nullptr;
        }
    }
    void* p = nullptr;
    if ((rc = pool_get(dev, SLOT_NACC, (size_t)pr->n_chains * sizeof(long long), &p))) return rc;
    c.n_accept = static_cast<long long*>(p);
    MCMCB200_CUDA_TRY(cudaEventCreate(&s.ev0));
    MCMCB200_CUDA_TRY(cudaEventCreate(&s.ev1));
    return MCMCB200_OK;
}

static int finish_common(Staged& s, mcmcb200_output_t* out, int launches)
{
    cudaStream_t st = s.stream;
    const CommonLaunch& c = s.c;
    if (s.draws_host)
        MCMCB200_CUDA_TRY(cudaMemcpyAsync(s.draws_host, c.draws, (size_t)c.n_chains * s.n_keep * c.d * sizeof(double),
                                          cudaMemcpyDeviceToHost, st));
    if (s.logp_host)
        MCMCB200_CUDA_TRY(cudaMemcpyAsync(s.logp_host, c.logp, (size_t)c.n_chains * s.n_keep * sizeof(double),
                                          cudaMemcpyDeviceToHost, st));
    if (out->n_accept_draws) {
        static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
        MCMCB200_CUDA_TRY(cudaMemcpyAsync(out->n_accept_draws, c.n_accept, (size_t)c.n_chains * sizeof(int64_t),
                                          cudaMemcpyDeviceToHost, st));
    }
    MCMCB200_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    MCMCB200_CUDA_TRY(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
    out->kernel_ms = ms;
    out->kernel_launches = launches;
    return MCMCB200_OK;
}

// M -> (sqrt factor, inverse [, M itself]) on the device; all column-major.  null precond -> nulls (M = I).
static int stage_precond(int dev, cudaStream_t st, const double* precond, int d, int chol_mode, const double** S_dev,
                         const double** Minv_dev, const double** M_dev)
{
    *S_dev = nullptr;
    if (Minv_dev) *Minv_dev = nullptr;
    if (M_dev) *M_dev = nullptr;
    if (!precond) return MCMCB200_OK;
    const size_t nn = (size_t)d * d;
    std::vector<double> S(nn), Minv(nn);
    if (!host_cholesky_colmajor(precond, d, chol_mode, S.data())) {
        set_error("precond_mat is not positive definite");
        return MCMCB200_ERR_INVALID_ARG;
    }
    int rc;
    if ((rc = upload(dev, SLOT_MAT_A, S.data(), nn, st, S_dev))) return rc;
    if (Minv_dev) {
        if (!host_inverse_colmajor(precond, d, Minv.data())) {
            set_error("precond_mat is singular");
            return MCMCB200_ERR_INVALID_ARG;
        }
        if ((rc = upload(dev, SLOT_MAT_B, Minv.data(), nn, st, Minv_dev))) return rc;
    }
    if (M_dev && (rc = upload(dev, SLOT_MAT_C, precond, nn, st, M_dev))) return rc;
    MCMCB200_CUDA_TRY(cudaStreamSynchronize(st));  // host vectors go out of scope
    return MCMCB200_OK;
}

}  // namespace mcmcb200

using namespace mcmcb200;

extern "C" {

void mcmcb200_hmc_settings_default(mcmcb200_hmc_settings_t* s)
{
    std::memset(s, 0, sizeof(*s));
    s->n_burnin_draws = 1000;  // mcmc_structs.hpp:68-74
    s->n_keep_draws = 1000;
    s->n_leap_steps = 1;
    s->step_size = 1.0;
    s->chol_mode = MCMCB200_CHOL_EIGEN_LLT;
    s->arith = MCMCB200_ARITH_FAST;
}
void mcmcb200_mala_settings_default(mcmcb200_mala_settings_t* s)
{
    std::memset(s, 0, sizeof(*s));
    s->n_burnin_draws = 1000;  // mcmc_structs.hpp:125-130
    s->n_keep_draws = 1000;
    s->step_size = 1.0;
    s->chol_mode = MCMCB200_CHOL_EIGEN_LLT;
    s->arith = MCMCB200_ARITH_FAST;
}
void mcmcb200_nuts_settings_default(mcmcb200_nuts_settings_t* s)
{
    std::memset(s, 0, sizeof(*s));
    s->n_burnin_draws = 1000;  // mcmc_structs.hpp:84-97
    s->n_keep_draws = 1000;
    s->n_adapt_draws = 1000;
    s->target_accept_rate = 0.55;
    s->max_tree_depth = 10;
    s->step_size = 1.0;
    s->gamma_val = 0.05;
    s->t0_val = 10;
    s->kappa_val = 0.75;
    s->chol_mode = MCMCB200_CHOL_EIGEN_LLT;
    s->arith = MCMCB200_ARITH_FAST;
}
void mcmcb200_rmhmc_settings_default(mcmcb200_rmhmc_settings_t* s)
{
    std::memset(s, 0, sizeof(*s));
    s->n_burnin_draws = 1000;  // mcmc_structs.hpp:107-116
    s->n_keep_draws = 1000;
    s->n_leap_steps = 1;
    s->step_size = 1.0;
    s->n_fp_steps = 5;
    s->metric_id = 0;
    s->chol_mode = MCMCB200_CHOL_EIGEN_LLT;
    s->arith = MCMCB200_ARITH_FAST;
}

void mcmcb200_rwmh_settings_default(mcmcb200_rwmh_settings_t* s)
{
    std::memset(s, 0, sizeof(*s));
    s->n_burnin_draws = 1000;  // mcmc_structs.hpp:138-149
    s->n_keep_draws = 1000;
    s->par_scale = 1.0;
    s->chol_mode = MCMCB200_CHOL_EIGEN_LLT;
    s->arith = MCMCB200_ARITH_FAST;
}

int mcmcb200_hmc_run(const mcmcb200_problem_t* pr, const mcmcb200_rng_t* rng, const mcmcb200_hmc_settings_t* st,
                     mcmcb200_output_t* out)
{
    if (!st) { set_error("null settings"); return MCMCB200_ERR_INVALID_ARG; }
    if (st->n_leap_steps < 0 || st->n_leap_steps > 0x7fffffff) { set_error("bad n_leap_steps"); return MCMCB200_ERR_INVALID_ARG; }
    Staged s;
    const bool wide = pr && hmc_wide_supported(pr->target_id, pr->n_dim, st->precond_mat != nullptr) && !pr->vals_bound;
    int rc = stage_common(s, pr, rng, st->arith, st->n_burnin_draws, st->n_keep_draws, 0, true, out, wide ? 2048 : 32 * MAX_EPL);
    if (rc) return rc;
    HmcLaunch a;
    static_cast<CommonLaunch&>(a) = s.c;
    a.n_burnin = st->n_burnin_draws;
    a.n_keep = st->n_keep_draws;
    a.n_leap = (int)st->n_leap_steps;
    a.eps = st->step_size;
    if ((rc = stage_precond(s.scope.dev, s.stream, st->precond_mat, pr->n_dim, st->chol_mode, &a.S_cm, &a.Minv_cm, nullptr))) return rc;
    MCMCB200_CUDA_TRY(cudaEventRecord(s.ev0, s.stream));
    if ((rc = wide ? launch_hmc_wide(a) : launch_hmc(a))) return rc;
    MCMCB200_CUDA_TRY(cudaEventRecord(s.ev1, s.stream));
    if (out->n_leapfrog_out)
        for (long long c = 0; c < pr->n_chains; ++c) out->n_leapfrog_out[c] = (st->n_burnin_draws + st->n_keep_draws) * st->n_leap_steps;
    if (out->step_size_out)
        for (long long c = 0; c < pr->n_chains; ++c) out->step_size_out[c] = st->step_size;
    return finish_common(s, out, 1);
}

int mcmcb200_mala_run(const mcmcb200_problem_t* pr, const mcmcb200_rng_t* rng, const mcmcb200_mala_settings_t* st,
                      mcmcb200_output_t* out)
{
    if (!st) { set_error("null settings"); return MCMCB200_ERR_INVALID_ARG; }
    Staged s;
    // dense quadratic targets with M = I run chain-batched (one fp64 tensor-core GEMM per draw for all chains) when the
    // dimension is beyond the register-resident kernels or there are enough chains to fill GEMM tiles
    const bool wide_ok = pr && mala_wide_supported(pr->target_id, pr->n_dim, st->precond_mat != nullptr) && !pr->broadcast_initial && !pr->vals_bound;
    const bool use_wide = wide_ok && (pr->n_dim > 32 * MAX_EPL || pr->n_chains >= 256);
    int rc = stage_common(s, pr, rng, st->arith, st->n_burnin_draws, st->n_keep_draws, 0, true, out, use_wide ? 2048 : 32 * MAX_EPL);
    if (rc) return rc;
    MalaLaunch a;
    static_cast<CommonLaunch&>(a) = s.c;
    a.n_burnin = st->n_burnin_draws;
    a.n_keep = st->n_keep_draws;
    a.eps = st->step_size;
    a.M_cm = a.S_cm = a.SigInv_cm = nullptr;
    if (st->precond_mat) {
        // Sigma = eps^2 M is the proposal covariance (mala.ipp:63-64); its inverse replaces the two
        // per-draw O(d^3) dmvnorm factorizations (the log-dets cancel exactly).
        const int d = pr->n_dim;
        const size_t nn = (size_t)d * d;
        std::vector<double> Sigma(nn), SigInv(nn), S(nn);
        const double e2 = st->step_size * st->step_size;
        for (size_t k = 0; k < nn; ++k) Sigma[k] = st->precond_mat[k] * e2;
        if (!host_cholesky_colmajor(st->precond_mat, d, st->chol_mode, S.data()) || !host_inverse_colmajor(Sigma.data(), d, SigInv.data())) {
            set_error("precond_mat is not positive definite");
            return MCMCB200_ERR_INVALID_ARG;
        }
        if ((rc = upload(s.scope.dev, SLOT_MAT_A, S.data(), nn, s.stream, &a.S_cm))) return rc;
        if ((rc = upload(s.scope.dev, SLOT_MAT_B, SigInv.data(), nn, s.stream, &a.SigInv_cm))) return rc;
        if ((rc = upload(s.scope.dev, SLOT_MAT_C, st->precond_mat, nn, s.stream, &a.M_cm))) return rc;
        MCMCB200_CUDA_TRY(cudaStreamSynchronize(s.stream));
    }
    int launches = 1;
    if (use_wide) {
        void* wp = nullptr;
        if ((rc = pool_get(s.scope.dev, SLOT_WORK, (size_t)mala_wide_work_doubles(pr->n_chains, pr->n_dim) * sizeof(double), &wp))) return rc;
        MCMCB200_CUDA_TRY(cudaEventRecord(s.ev0, s.stream));
        if ((rc = launch_mala_wide(a, static_cast<double*>(wp), &launches))) return rc;
    } else {
        MCMCB200_CUDA_TRY(cudaEventRecord(s.ev0, s.stream));
        if ((rc = launch_mala(a))) return rc;
    }
    MCMCB200_CUDA_TRY(cudaEventRecord(s.ev1, s.stream));
    if (out->n_leapfrog_out)
        for (long long c = 0; c < pr->n_chains; ++c) out->n_leapfrog_out[c] = 0;
    if (out->step_size_out)
        for (long long c = 0; c < pr->n_chains; ++c) out->step_size_out[c] = st->step_size;
    return finish_common(s, out, launches);
}

int mcmcb200_rwmh_run(const mcmcb200_problem_t* pr, const mcmcb200_rng_t* rng, const mcmcb200_rwmh_settings_t* st,
                      mcmcb200_output_t* out)
{
    if (!st) { set_error("null settings"); return MCMCB200_ERR_INVALID_ARG; }
    Staged s;
    int rc = stage_common(s, pr, rng, st->arith, st->n_burnin_draws, st->n_keep_draws, 0, true, out);
    if (rc) return rc;
    RwmhLaunch a;
    static_cast<CommonLaunch&>(a) = s.c;
    a.n_burnin = st->n_burnin_draws;
    a.n_keep = st->n_keep_draws;
    a.par_scale = st->par_scale;
    a.S_cm = nullptr;
    if (st->cov_mat) {
        // cov_mcmc_chol = par_scale * chol(cov_mat), materialised once (src/rwmh.cpp:116)
        const int d = pr->n_dim;
        const size_t nn = (size_t)d * d;
        std::vector<double> S(nn);
        if (!host_cholesky_colmajor(st->cov_mat, d, st->chol_mode, S.data())) {
            set_error("cov_mat is not positive definite");
            return MCMCB200_ERR_INVALID_ARG;
        }
        for (size_t k = 0; k < nn; ++k) S[k] = st->par_scale * S[k];
        if ((rc = upload(s.scope.dev, SLOT_MAT_A, S.data(), nn, s.stream, &a.S_cm))) return rc;
        MCMCB200_CUDA_TRY(cudaStreamSynchronize(s.stream));
    }
    MCMCB200_CUDA_TRY(cudaEventRecord(s.ev0, s.stream));
    if ((rc = launch_rwmh(a))) return rc;
    MCMCB200_CUDA_TRY(cudaEventRecord(s.ev1, s.stream));
    if (out->n_leapfrog_out)
        for (long long c = 0; c < pr->n_chains; ++c) out->n_leapfrog_out[c] = 0;
    if (out->step_size_out)
        for (long long c = 0; c < pr->n_chains; ++c) out->step_size_out[c] = st->par_scale;
    return finish_common(s, out, 1);
}

int mcmcb200_nuts_run(const mcmcb200_problem_t* pr, const mcmcb200_rng_t* rng, const mcmcb200_nuts_settings_t* st,
                      mcmcb200_output_t* out)
{
    if (!st) { set_error("null settings"); return MCMCB200_ERR_INVALID_ARG; }
    if (st->max_tree_depth < 0 || st->max_tree_depth > 20) { set_error("max_tree_depth must be in [0,20]"); return MCMCB200_ERR_INVALID_ARG; }
    Staged s;
    int rc = stage_common(s, pr, rng, st->arith, st->n_burnin_draws, st->n_keep_draws, pr ? pr->n_dim : 0, false, out);
    if (rc) return rc;
    NutsLaunch a;
    static_cast<CommonLaunch&>(a) = s.c;
    a.n_burnin = st->n_burnin_draws;
    a.n_keep = st->n_keep_draws;
    const long long n_total = a.n_burnin + a.n_keep;
    a.n_adapt = (st->n_adapt_draws <= n_total) ? st->n_adapt_draws : n_total;  // src/nuts.cpp:54
    a.max_depth = (int)st->max_tree_depth;
    a.eps_bar0 = st->step_size;
    a.delta = st->target_accept_rate;
    a.gamma = st->gamma_val;
    a.t0 = st->t0_val;
    a.kappa = st->kappa_val;
    if ((rc = stage_precond(s.scope.dev, s.stream, st->precond_mat, pr->n_dim, st->chol_mode, &a.S_cm, &a.Minv_cm, nullptr))) return rc;
    void* p = nullptr;
    if ((rc = pool_get(s.scope.dev, SLOT_STEP, (size_t)pr->n_chains * sizeof(double), &p))) return rc;
    a.step_out = static_cast<double*>(p);
    if ((rc = pool_get(s.scope.dev, SLOT_NLF, (size_t)pr->n_chains * sizeof(long long), &p))) return rc;
    a.n_leapfrog = static_cast<long long*>(p);
    a.work_stride = nuts_work_doubles_per_chain(pr->n_dim, a.max_depth);
    if ((rc = pool_get(s.scope.dev, SLOT_WORK, (size_t)pr->n_chains * (size_t)a.work_stride * sizeof(double), &p))) return rc;
    a.work = static_cast<double*>(p);
    // dense targets with enough chains to fill the GPU run 8 chains per CTA with cooperative gradients (nuts.cu);
    // MCMCB200_NUTS_COOP=0/1 forces the choice (tests compare the two kernels bit for bit)
    // by default only for the target whose cooperative path is covered by the GPU parity tests (dense_gauss, the C4 target);
    // linreg has the same kernel instantiated and can be switched on with the environment variable
    a.coop = pr->n_chains >= 64 && pr->target_id == MCMCB200_TARGET_DENSE_GAUSS;
    if (const char* e = std::getenv("MCMCB200_NUTS_COOP")) a.coop = (e[0] == '1');
    a.coop_batch = 6;   // measured on B200 (C4 shape, 1184 chains x 40 draws): 2 -> 456 ms, 4 -> 332, 6 -> 314, 8 -> 332
    if (const char* e = std::getenv("MCMCB200_NUTS_BATCH")) a.coop_batch = std::atoi(e) > 0 ? std::atoi(e) : 1;
    MCMCB200_CUDA_TRY(cudaEventRecord(s.ev0, s.stream));
    if ((rc = launch_nuts(a))) return rc;
    MCMCB200_CUDA_TRY(cudaEventRecord(s.ev1, s.stream));
    if (out->step_size_out)
        MCMCB200_CUDA_TRY(cudaMemcpyAsync(out->step_size_out, a.step_out, (size_t)pr->n_chains * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    if (out->n_leapfrog_out)
        MCMCB200_CUDA_TRY(cudaMemcpyAsync(out->n_leapfrog_out, a.n_leapfrog, (size_t)pr->n_chains * sizeof(int64_t), cudaMemcpyDeviceToHost, s.stream));
    return finish_common(s, out, 1);
}

int mcmcb200_rmhmc_run(const mcmcb200_problem_t* pr, const mcmcb200_rng_t* rng, const mcmcb200_rmhmc_settings_t* st,
                       mcmcb200_output_t* out)
{
    if (!st) { set_error("null settings"); return MCMCB200_ERR_INVALID_ARG; }
    Staged s;
    int rc = stage_common(s, pr, rng, st->arith, st->n_burnin_draws, st->n_keep_draws, pr ? pr->n_dim : 0, true, out);
    if (rc) return rc;
    RmhmcLaunch a;
    static_cast<CommonLaunch&>(a) = s.c;
    a.n_burnin = st->n_burnin_draws;
    a.n_keep = st->n_keep_draws;
    a.n_leap = (int)st->n_leap_steps;
    a.n_fp = (int)st->n_fp_steps;
    a.eps = st->step_size;
    a.chol_mode = st->chol_mode;
    a.cons_term = (double)(0.5 * (double)(size_t)pr->n_dim * 1.83787706640934548356L);
    a.metric_id = st->metric_id;
    a.work = nullptr;
    a.work_stride = 0;
    // the 2-parameter Normal model runs thread-per-chain in registers (rmhmc.cu); everything else — and the Normal model
    // too when MCMCB200_RMHMC_GENERAL=1 (tests compare the two kernels) — runs warp-per-chain with the metric algebra
    // in a per-chain scratch area (rmhmc_general.cu)
    bool general = pr->target_id != MCMCB200_TARGET_NORMAL_MODEL;
    if (const char* e = std::getenv("MCMCB200_RMHMC_GENERAL")) general = general || e[0] == '1';
    if (general) {
        if (!rmhmc_general_supported(pr->target_id, st->metric_id, pr->n_dim)) {
            set_error("rmhmc: target %d has no registered metric %d for n_dim=%d (general kernel: n_dim <= 64)", pr->target_id, st->metric_id, pr->n_dim);
            return MCMCB200_ERR_UNSUPPORTED;
        }
        a.work_stride = rmhmc_general_work_doubles(pr->n_dim);
        void* wp = nullptr;
        if ((rc = pool_get(s.scope.dev, SLOT_WORK, (size_t)pr->n_chains * (size_t)a.work_stride * sizeof(double), &wp))) return rc;
        a.work = static_cast<double*>(wp);
    }
    MCMCB200_CUDA_TRY(cudaEventRecord(s.ev0, s.stream));
    if ((rc = general ? launch_rmhmc_general(a) : launch_rmhmc(a))) return rc;
    MCMCB200_CUDA_TRY(cudaEventRecord(s.ev1, s.stream));
    if (out->n_leapfrog_out)
        for (long long c = 0; c < pr->n_chains; ++c) out->n_leapfrog_out[c] = (st->n_burnin_draws + st->n_keep_draws) * st->n_leap_steps;
    if (out->step_size_out)
        for (long long c = 0; c < pr->n_chains; ++c) out->step_size_out[c] = st->step_size;
    return finish_common(s, out, 1);
}

int mcmcb200_target_lookup(const char* name)
{
    if (!name) return -1;
    static const struct { const char* n; int id; } tbl[] = {
        {"iso_gauss", MCMCB200_TARGET_ISO_GAUSS},     {"diag_gauss", MCMCB200_TARGET_DIAG_GAUSS},
        {"dense_gauss", MCMCB200_TARGET_DENSE_GAUSS}, {"linreg", MCMCB200_TARGET_LINREG},
        {"normal_model", MCMCB200_TARGET_NORMAL_MODEL}, {"funnel", MCMCB200_TARGET_FUNNEL}};
    for (const auto& e : tbl)
        if (std::strcmp(e.n, name) == 0) return e.id;
    return -1;
}

int64_t mcmcb200_target_data_len(int target_id, int32_t n_dim) { return target_data_len(target_id, n_dim); }

int mcmcb200_target_eval(int target_id, const double* target_data, int64_t target_data_len_, int32_t n_dim, int64_t n_points,
                         const double* x, double* value_out, double* grad_out, int32_t arith)
{
    const int64_t need = target_data_len(target_id, n_dim);
    if (need < 0) { set_error("unknown target id %d", target_id); return MCMCB200_ERR_UNKNOWN_TARGET; }
    if (target_data_len_ < need || !x || !value_out || n_points <= 0) { set_error("bad arguments"); return MCMCB200_ERR_INVALID_ARG; }
    if (epl_for_dim(n_dim) == 0) { set_error("n_dim=%d unsupported", n_dim); return MCMCB200_ERR_UNSUPPORTED; }
    DeviceScope sc;
    int rc = sc.enter(-1);
    if (rc) return rc;
    EvalLaunch a{};
    a.target_id = target_id;
    a.d = n_dim;
    a.n_points = n_points;
    a.strict = (arith == MCMCB200_ARITH_STRICT);
    a.stream = nullptr;
    if ((rc = upload(sc.dev, SLOT_TDATA, target_data, (size_t)need, nullptr, &a.tdata))) return rc;
    if ((rc = upload(sc.dev, SLOT_EVAL_X, x, (size_t)n_points * n_dim, nullptr, &a.x))) return rc;
    void* p = nullptr;
    if ((rc = pool_get(sc.dev, SLOT_EVAL_V, (size_t)n_points * sizeof(double), &p))) return rc;
    a.value = static_cast<double*>(p);
    a.grad = nullptr;
    if (grad_out) {
        if ((rc = pool_get(sc.dev, SLOT_EVAL_G, (size_t)n_points * n_dim * sizeof(double), &p))) return rc;
        a.grad = static_cast<double*>(p);
    }
    if ((rc = launch_target_eval(a))) return rc;
    MCMCB200_CUDA_TRY(cudaMemcpy(value_out, a.value, (size_t)n_points * sizeof(double), cudaMemcpyDeviceToHost));
    if (grad_out) MCMCB200_CUDA_TRY(cudaMemcpy(grad_out, a.grad, (size_t)n_points * n_dim * sizeof(double), cudaMemcpyDeviceToHost));
    return MCMCB200_OK;
}

int mcmcb200_mt19937_tape(uint64_t seed, int64_t n_pre_normals, int64_t n_draws, int32_t n_dim, double* tape_out)
{
    if (!tape_out || n_pre_normals < 0 || n_draws < 0 || n_dim <= 0) { set_error("bad arguments"); return MCMCB200_ERR_INVALID_ARG; }
    host_mt19937_tape(seed, n_pre_normals, n_draws, n_dim, tape_out);
    return MCMCB200_OK;
}

int mcmcb200_philox_stream(uint64_t seed, int64_t chain, int64_t draw, int32_t n_dim, int32_t n_unif, double* out)
{
    if (!out || n_dim <= 0 || n_unif < 0 || epl_for_dim(n_dim) == 0) { set_error("bad arguments"); return MCMCB200_ERR_INVALID_ARG; }
    DeviceScope sc;
    int rc = sc.enter(-1);
    if (rc) return rc;
    void* p = nullptr;
    if ((rc = pool_get(sc.dev, SLOT_EVAL_V, (size_t)(n_dim + n_unif) * sizeof(double), &p))) return rc;
    if ((rc = launch_philox_stream(seed, chain, draw, n_dim, n_unif, static_cast<double*>(p), nullptr)))
        return rc;
    MCMCB200_CUDA_TRY(cudaMemcpy(out, p, (size_t)(n_dim + n_unif) * sizeof(double), cudaMemcpyDeviceToHost));
    return MCMCB200_OK;
}

const char* mcmcb200_last_error(void) { return g_err; }

int mcmcb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

void mcmcb200_release_workspace(void)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= MAX_DEVICES) return;
    for (int s = 0; s < SLOT_COUNT; ++s) {
        if (g_pool.b[dev][s].p) cudaFree(g_pool.b[dev][s].p);
        g_pool.b[dev][s] = Buf();
    }
}

}  // extern "C"
