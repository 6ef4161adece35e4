// C-ABI layer of the engine (include/mcmc_b200.h): argument checking, one-time host
// linear algebra on the preconditioner (what src/hmc.cpp:57-59 does with BMO_MATOPS_INV /
// BMO_MATOPS_CHOL_LOWER), host<->device marshalling, reference-stream tape generation,
// kernel dispatch and timing.  No CPU sampling path exists here: if CUDA is unavailable
// every run call fails with MCMCB200_ERR_CUDA.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <mutex>
#include <string>
#include <new>
#include <stdexcept>
#include <thread>
#include <vector>

#include "engine.h"
#include "transpose.h"
#include "rmhmc_cta.h"
#include "hmc_batched.h"
#include "nuts_batched.h"
#include "hmc_duo.h"
#include "hmc_half.h"

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
            SLOT_STEP, SLOT_NLF, SLOT_WORK, SLOT_EVAL_X, SLOT_EVAL_V, SLOT_EVAL_G, SLOT_ERR, SLOT_AUX, SLOT_DRAWS_T, SLOT_COUNT };
constexpr int MAX_DEVICES = 16;
// Device scratch is a per-device cache of buffers shared by all host threads; a run call LEASES the buffers it needs and
// hands them back when it returns.  Like the reference's samplers (no globals, src/hmc.cpp) the run calls stay re-entrant:
// concurrent calls — on the same or on different devices — never share a buffer, while consecutive calls (also from
// short-lived threads, e.g. the per-device threads of mcmc_b200.hpp's b200.devices) reuse the allocations.  Idle buffers
// beyond MCMCB200_POOL_CAP_MB (default 16384) per device are freed largest-first; mcmcb200_release_workspace() frees all
// idle buffers of the current device.
struct CacheBuf { void* p; size_t bytes; bool in_use; };
static std::mutex g_pool_mu;
static std::vector<CacheBuf>& pool_of(int dev)
{
    static std::vector<CacheBuf>* pools = new std::vector<CacheBuf>[MAX_DEVICES];   // never destroyed: the CUDA runtime may be gone at exit
    return pools[dev];
}
static size_t pool_cap_bytes()
{
    static const size_t cap = [] {
        const char* e = std::getenv("MCMCB200_POOL_CAP_MB");
        const long long mb = e ? std::atoll(e) : 16384;
        return (size_t)(mb < 0 ? 0 : mb) << 20;
    }();
    return cap;
}
static void pool_trim_locked(int dev, size_t cap)
{
    std::vector<CacheBuf>& v = pool_of(dev);
    for (;;) {
        size_t idle = 0, worst = v.size();
        for (size_t i = 0; i < v.size(); ++i)
            if (!v[i].in_use) {
                idle += v[i].bytes;
                if (worst == v.size() || v[i].bytes > v[worst].bytes) worst = i;
            }
        if (idle <= cap || worst == v.size()) return;
        cudaFree(v[worst].p);
        v.erase(v.begin() + (long)worst);
    }
}
struct Lease {
    int dev = -1;
    cudaStream_t stream = nullptr;
    bool settled = false;            // the stream was synchronized after the last use of the buffers
    void* held[SLOT_COUNT] = {};
    size_t held_bytes[SLOT_COUNT] = {};
    void release_slot_locked(int s)
    {
        if (!held[s]) return;
        for (CacheBuf& b : pool_of(dev))
            if (b.p == held[s]) b.in_use = false;
        held[s] = nullptr;
        held_bytes[s] = 0;
    }
    ~Lease()
    {
        if (dev < 0) return;
        if (!settled) cudaStreamSynchronize(stream);   // error paths: kernels may still be using the buffers
        std::lock_guard<std::mutex> g(g_pool_mu);
        for (int s = 0; s < SLOT_COUNT; ++s) release_slot_locked(s);
        pool_trim_locked(dev, pool_cap_bytes());
    }
};

static int pool_get(Lease& L, int dev, Slot s, size_t bytes, void** out)
{
    if (bytes == 0) bytes = 8;
    L.dev = dev;
    if (L.held[s] && L.held_bytes[s] >= bytes) { *out = L.held[s]; return MCMCB200_OK; }
    std::lock_guard<std::mutex> g(g_pool_mu);
    L.release_slot_locked(s);
    std::vector<CacheBuf>& v = pool_of(dev);
    size_t best = v.size();
    for (size_t i = 0; i < v.size(); ++i)   // best fit among the idle buffers, but never waste more than 2x
        if (!v[i].in_use && v[i].bytes >= bytes && v[i].bytes <= 2 * bytes + (1u << 20) && (best == v.size() || v[i].bytes < v[best].bytes)) best = i;
    if (best == v.size()) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaErrorMemoryAllocation) {   // give the idle cache back and retry once
            cudaGetLastError();
            pool_trim_locked(dev, 0);
            e = cudaMalloc(&p, bytes);
        }
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            return (e == cudaErrorMemoryAllocation) ? MCMCB200_ERR_OOM : MCMCB200_ERR_CUDA;
        }
        v.push_back(CacheBuf{p, bytes, false});
        best = v.size() - 1;
    }
    v[best].in_use = true;
    L.held[s] = v[best].p;
    L.held_bytes[s] = v[best].bytes;
    *out = v[best].p;
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

// ---- user-registered targets (mcmcb200_register_target) ---------------------------------------------------------
struct UserTarget { std::string name; mcmcb200_user_target_t vt; };
static std::mutex g_user_mu;
static std::vector<UserTarget>& user_targets()
{
    static std::vector<UserTarget>* v = new std::vector<UserTarget>;   // never destroyed: user libraries may unload after us
    return *v;
}
static bool user_target_get(int target_id, mcmcb200_user_target_t* out)
{
    std::lock_guard<std::mutex> g(g_user_mu);
    const int k = target_id - MCMCB200_USER_TARGET_BASE;
    if (k < 0 || k >= (int)user_targets().size()) return false;
    *out = user_targets()[(size_t)k].vt;
    return true;
}
bool user_target_has(int kind, int target_id)
{
    mcmcb200_user_target_t vt;
    return kind >= 0 && kind < USER_LAUNCH_COUNT && user_target_get(target_id, &vt) && vt.launch[kind] != nullptr;
}
int user_target_launch(int kind, int target_id, const void* launch_struct)
{
    mcmcb200_user_target_t vt;
    if (kind < 0 || kind >= USER_LAUNCH_COUNT || !user_target_get(target_id, &vt) || !vt.launch[kind]) {
        set_error("target %d has no launcher for this sampler", target_id);
        return MCMCB200_ERR_UNKNOWN_TARGET;
    }
    return vt.launch[kind](launch_struct);
}

static int64_t target_data_len(int target_id, int d)
{
    if (d <= 0) return -1;
    if (target_id >= MCMCB200_USER_TARGET_BASE) {
        mcmcb200_user_target_t vt;
        if (!user_target_get(target_id, &vt)) return -1;
        return vt.data_len ? vt.data_len(d) : 0;
    }
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
    DeviceScope scope;   // declared first: destroyed last, so the lease is returned while the device is still current
    Lease lease;
    cudaStream_t stream = nullptr;
    CommonLaunch c{};
    double* draws_host = nullptr;
    double* logp_host = nullptr;
    long long n_keep = 0;
    bool deferred_mt_tape = false;
    double* draws_T = nullptr;   // MCMCB200_LAYOUT_COLMAJOR: device buffer receiving the transposed draws (what is handed back)
    float transpose_ms = 0.f;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ~Staged()
    {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
    }
};

static int upload(Lease& L, int dev, Slot slot, const double* host, size_t n, cudaStream_t st, const double** dev_out)
{
    void* p = nullptr;
    int rc = pool_get(L, dev, slot, n * sizeof(double), &p);
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
    s.lease.stream = s.stream;
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

    // built-in targets have a fixed blob size; a user's functor declares a minimum and receives everything the caller passed
    const size_t n_tdata = (pr->target_id >= MCMCB200_USER_TARGET_BASE) ? (size_t)pr->target_data_len : (size_t)need;
    if ((rc = upload(s.lease, dev, SLOT_TDATA, pr->target_data, n_tdata, st, &c.tdata))) return rc;
    c.lb = c.ub = nullptr;
    if (pr->vals_bound) {
        if (!pr->lower_bounds || !pr->upper_bounds) { set_error("vals_bound is set but lower_bounds / upper_bounds is null"); return MCMCB200_ERR_INVALID_ARG; }
        for (int j = 0; j < d; ++j)
            if (pr->lower_bounds[j] != pr->lower_bounds[j] || pr->upper_bounds[j] != pr->upper_bounds[j] || !(pr->lower_bounds[j] < pr->upper_bounds[j])) {
                set_error("bounds of element %d are not ordered (lower < upper required)", j);
                return MCMCB200_ERR_INVALID_ARG;
            }
        if ((rc = upload(s.lease, dev, SLOT_LB, pr->lower_bounds, (size_t)d, st, &c.lb))) return rc;
        if ((rc = upload(s.lease, dev, SLOT_UB, pr->upper_bounds, (size_t)d, st, &c.ub))) return rc;
    }

    const size_t n_x0 = (size_t)(c.broadcast_x0 ? 1 : pr->n_chains) * d;
    if (pr->initial_mem == MCMCB200_MEM_DEVICE) c.x0 = pr->initial_vals;
    else if ((rc = upload(s.lease, dev, SLOT_X0, pr->initial_vals, n_x0, st, &c.x0))) return rc;

    // RNG
    const long long n_total = n_burnin + n_keep;
    rng_set_key(c.rng, rng->seed);
    c.rng.tape = nullptr;
    c.rng.tape_stride = 0;
    if (rng->mode == MCMCB200_RNG_PHILOX) {
        c.rng.mode = RNG_PHILOX;
    } else if (rng->mode == MCMCB200_RNG_MT19937_TAPE) {
        c.rng.mode = RNG_TAPE;
        if (!mt_tape_supported) {
            // data-dependent variate count (NUTS): the caller drives the kernel draw by draw and supplies each segment's
            // slice of the reference stream itself (nuts_run_impl)
            s.deferred_mt_tape = true;
        } else {
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
        if ((rc = upload(s.lease, dev, SLOT_TAPE, tape.data(), tape.size(), st, &c.rng.tape))) return rc;
        MCMCB200_CUDA_TRY(cudaStreamSynchronize(st));  // `tape` goes out of scope
        c.rng.tape_stride = stride;
        }
    } else if (rng->mode == MCMCB200_RNG_USER_TAPE) {
        if (!rng->tape || rng->tape_stride <= 0) { set_error("USER_TAPE needs tape and tape_stride"); return MCMCB200_ERR_INVALID_ARG; }
        // samplers with a static variate count per draw: the whole stream must be there (the kernels read it unchecked);
        // NUTS consumes a data-dependent number of uniforms: its kernel checks every read against tape_stride and the run
        // fails with MCMCB200_ERR_INVALID_ARG when a chain runs off its tape
        if (mt_tape_supported && rng->tape_stride < n_pre_normals + n_total * (d + 1)) {
            set_error("USER_TAPE: tape_stride %lld is shorter than the %lld variates a chain consumes (%lld pre-loop normals + %lld draws x (n_dim + 1))",
                      (long long)rng->tape_stride, (long long)(n_pre_normals + n_total * (d + 1)), (long long)n_pre_normals, (long long)n_total);
            return MCMCB200_ERR_INVALID_ARG;
        }
        c.rng.mode = RNG_TAPE;
        c.rng.tape_stride = rng->tape_stride;
        if (rng->tape_mem == MCMCB200_MEM_DEVICE) c.rng.tape = rng->tape;
        else if ((rc = upload(s.lease, dev, SLOT_TAPE, rng->tape, (size_t)pr->n_chains * (size_t)rng->tape_stride, st, &c.rng.tape))) return rc;
    } else {
        set_error("unknown rng mode %d", rng->mode);
        return MCMCB200_ERR_INVALID_ARG;
    }

    {
        void* ep = nullptr;
        if ((rc = pool_get(s.lease, dev, SLOT_ERR, sizeof(int), &ep))) return rc;
        MCMCB200_CUDA_TRY(cudaMemsetAsync(ep, 0, sizeof(int), st));
        c.rng.err_flag = static_cast<int*>(ep);
    }
    // outputs
    const size_t n_draws = (size_t)pr->n_chains * (size_t)n_keep * d;
    if (out->draws_layout != MCMCB200_LAYOUT_CHAIN_ROWS && out->draws_layout != MCMCB200_LAYOUT_COLMAJOR) { set_error("unknown draws_layout %d", out->draws_layout); return MCMCB200_ERR_INVALID_ARG; }
    const bool colmajor = out->draws_layout == MCMCB200_LAYOUT_COLMAJOR && n_keep > 0;
    if (colmajor) {
        // the kernels write chain-major rows; the reference's Mat_t layout is produced by one HBM-bound transpose on the
        // device (finish_common), so the host receives bytes it can use as they are
        void* p = nullptr;
        if ((rc = pool_get(s.lease, dev, SLOT_DRAWS, n_draws * sizeof(double), &p))) return rc;
        c.draws = static_cast<double*>(p);
        if (out->draws_mem == MCMCB200_MEM_DEVICE) {
            s.draws_T = out->draws_out;
            c.logp = out->logp_out;
        } else {
            if ((rc = pool_get(s.lease, dev, SLOT_DRAWS_T, n_draws * sizeof(double), &p))) return rc;
            s.draws_T = static_cast<double*>(p);
            s.draws_host = out->draws_out;
        }
    }
    if (!colmajor && out->draws_mem == MCMCB200_MEM_DEVICE) {
        c.draws = out->draws_out;
        c.logp = out->logp_out;
    } else if (out->draws_mem != MCMCB200_MEM_DEVICE) {
        void* p = nullptr;
        if (!colmajor) {
            if ((rc = pool_get(s.lease, dev, SLOT_DRAWS, n_draws * sizeof(double), &p))) return rc;
            c.draws = static_cast<double*>(p);
            s.draws_host = out->draws_out;
        }
        if (out->logp_out) {
            if ((rc = pool_get(s.lease, dev, SLOT_LOGP, (size_t)pr->n_chains * n_keep * sizeof(double), &p))) return rc;
            c.logp = static_cast<double*>(p);
            s.logp_host = out->logp_out;
        } else {
            c.logp = nullptr;
        }
    }
    void* p = nullptr;
    if ((rc = pool_get(s.lease, dev, SLOT_NACC, (size_t)pr->n_chains * sizeof(long long), &p))) return rc;
    c.n_accept = static_cast<long long*>(p);
    MCMCB200_CUDA_TRY(cudaEventCreate(&s.ev0));
    MCMCB200_CUDA_TRY(cudaEventCreate(&s.ev1));
    return MCMCB200_OK;
}

static int finish_common(Staged& s, mcmcb200_output_t* out, int launches)
{
    cudaStream_t st = s.stream;
    const CommonLaunch& c = s.c;
    if (s.draws_T) {
        int rc = launch_transpose_draws(c.draws, s.draws_T, c.n_chains, s.n_keep, c.d, st);
        if (rc) return rc;
    }
    if (s.draws_host)
        MCMCB200_CUDA_TRY(cudaMemcpyAsync(s.draws_host, s.draws_T ? s.draws_T : c.draws, (size_t)c.n_chains * s.n_keep * c.d * sizeof(double),
                                          cudaMemcpyDeviceToHost, st));
    if (s.logp_host)
        MCMCB200_CUDA_TRY(cudaMemcpyAsync(s.logp_host, c.logp, (size_t)c.n_chains * s.n_keep * sizeof(double),
                                          cudaMemcpyDeviceToHost, st));
    if (out->n_accept_draws) {
        static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
        MCMCB200_CUDA_TRY(cudaMemcpyAsync(out->n_accept_draws, c.n_accept, (size_t)c.n_chains * sizeof(int64_t),
                                          cudaMemcpyDeviceToHost, st));
    }
    int dev_err = 0;
    if (c.rng.err_flag) MCMCB200_CUDA_TRY(cudaMemcpyAsync(&dev_err, c.rng.err_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    MCMCB200_CUDA_TRY(cudaStreamSynchronize(st));
    if (dev_err) {
        set_error("a chain consumed more variates than its tape holds (tape_stride too short for this run)");
        return MCMCB200_ERR_INVALID_ARG;
    }
    s.lease.settled = true;
    float ms = 0.f;
    MCMCB200_CUDA_TRY(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
    out->kernel_ms = ms;
    out->kernel_launches = launches;
    return MCMCB200_OK;
}

// M -> (sqrt factor, inverse [, M itself]) on the device; all column-major.  null precond -> nulls (M = I).
static int stage_precond(Lease& lease, int dev, cudaStream_t st, const double* precond, int d, int chol_mode, const double** S_dev,
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
    if ((rc = upload(lease, dev, SLOT_MAT_A, S.data(), nn, st, S_dev))) return rc;
    if (Minv_dev) {
        if (!host_inverse_colmajor(precond, d, Minv.data())) {
            set_error("precond_mat is singular");
            return MCMCB200_ERR_INVALID_ARG;
        }
        if ((rc = upload(lease, dev, SLOT_MAT_B, Minv.data(), nn, st, Minv_dev))) return rc;
    }
    if (M_dev && (rc = upload(lease, dev, SLOT_MAT_C, precond, nn, st, M_dev))) return rc;
    MCMCB200_CUDA_TRY(cudaStreamSynchronize(st));  // host vectors go out of scope
    return MCMCB200_OK;
}

}  // namespace mcmcb200

// The C entry points never let a C++ exception escape (std::bad_alloc from a host staging vector, e.g. a reference-stream
// tape of several GB): they return MCMCB200_ERR_OOM / MCMCB200_ERR_CUDA with the text in mcmcb200_last_error().
template <class F> static int guarded(F&& f)
{
    try {
        return f();
    } catch (const std::bad_alloc&) {
        ::mcmcb200::set_error("host memory allocation failed");
        return MCMCB200_ERR_OOM;
    } catch (const std::exception& e) {
        ::mcmcb200::set_error("internal error: %s", e.what());
        return MCMCB200_ERR_CUDA;
    }
}
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

static int hmc_run_impl(const mcmcb200_problem_t* pr, const mcmcb200_rng_t* rng, const mcmcb200_hmc_settings_t* st,
                     mcmcb200_output_t* out)
{
    if (!st) { set_error("null settings"); return MCMCB200_ERR_INVALID_ARG; }
    if (st->n_leap_steps < 0 || st->n_leap_steps > 0x7fffffff) { set_error("bad n_leap_steps"); return MCMCB200_ERR_INVALID_ARG; }
    Staged s;
    const bool wide = pr && hmc_wide_supported(pr->target_id, pr->n_dim, st->precond_mat != nullptr) && !pr->vals_bound;
    // a dense mass matrix and / or a dense quadratic target with many chains (or n_dim beyond the register-resident kernels):
    // chain-batched path, every d x d product as one fp64 tensor-core GEMM over all chains (hmc_batched.cu); FAST arithmetic
    const bool batched = pr && !wide && !pr->broadcast_initial &&
                         hmc_batched_supported(pr->target_id, pr->n_dim, st->precond_mat != nullptr, st->arith == MCMCB200_ARITH_STRICT, pr->vals_bound != 0, pr->n_chains);
    int rc = stage_common(s, pr, rng, st->arith, st->n_burnin_draws, st->n_keep_draws, 0, true, out, (wide || batched) ? 2048 : 32 * MAX_EPL);
    if (rc) return rc;
    HmcLaunch a;
    static_cast<CommonLaunch&>(a) = s.c;
    a.n_burnin = st->n_burnin_draws;
    a.n_keep = st->n_keep_draws;
    a.n_leap = (int)st->n_leap_steps;
    a.eps = st->step_size;
    if ((rc = stage_precond(s.lease, s.scope.dev, s.stream, st->precond_mat, pr->n_dim, st->chol_mode, &a.S_cm, &a.Minv_cm, nullptr))) return rc;
    int launches = 1;
    if (batched) {
        void* wp = nullptr;
        if ((rc = pool_get(s.lease, s.scope.dev, SLOT_WORK, (size_t)hmc_batched_work_doubles(pr->n_chains, pr->n_dim) * sizeof(double), &wp))) return rc;
        MCMCB200_CUDA_TRY(cudaEventRecord(s.ev0, s.stream));
        if ((rc = launch_hmc_batched(a, static_cast<double*>(wp), &launches))) return rc;
    } else {
        MCMCB200_CUDA_TRY(cudaEventRecord(s.ev0, s.stream));
        // few chains (strong-scaling shards): two warps per chain, variates of draw t + 1 generated under the trajectory of draw t
        if (!wide && hmc_half_supported(a)) { if ((rc = launch_hmc_half(a))) return rc; }   // n_dim <= 32: two chains per warp
        else if (!wide && hmc_duo_supported(a)) { if ((rc = launch_hmc_duo(a))) return rc; }
        else if ((rc = wide ? launch_hmc_wide(a) : launch_hmc(a))) return rc;
    }
    MCMCB200_CUDA_TRY(cudaEventRecord(s.ev1, s.stream));
    if (out->n_leapfrog_out)
        for (long long c = 0; c < pr->n_chains; ++c) out->n_leapfrog_out[c] = (st->n_burnin_draws + st->n_keep_draws) * st->n_leap_steps;
    if (out->step_size_out)
        for (long long c = 0; c < pr->n_chains; ++c) out->step_size_out[c] = st->step_size;
    return finish_common(s, out, launches);
}

static int mala_run_impl(const mcmcb200_problem_t* pr, const mcmcb200_rng_t* rng, const mcmcb200_mala_settings_t* st,
                      mcmcb200_output_t* out)
{
    if (!st) { set_error("null settings"); return MCMCB200_ERR_INVALID_ARG; }
    if (pr && pr->vals_bound && st->precond_mat) {
        // The reference then evaluates both proposal densities with the covariance eps^2 J(proposal) M (mala.ipp:55-56), which is
        // not symmetric: dmvnorm's LLT log-det reads its lower triangle and turns NaN whenever that is not positive definite, and
        // min(0.01, NaN) = 0.01 accepts the proposal.  Reproducing that needs an O(d^3) factorisation per draw and chain only to
        // detect the NaN; the device path refuses the combination instead of sampling something else.  (HMC and NUTS take
        // bounds together with a precond_mat: there the mass matrix and the Jacobian do not mix.)
        set_error("mala: vals_bound together with precond_mat is not supported on the device path");
        return MCMCB200_ERR_UNSUPPORTED;
    }
    Staged s;
    // dense quadratic targets with M = I run chain-batched (one fp64 tensor-core GEMM per draw for all chains) when the
    // dimension is beyond the register-resident kernels or there are enough chains to fill GEMM tiles
    const bool wide_ok = pr && mala_wide_supported(pr->target_id, pr->n_dim, st->precond_mat != nullptr) && !pr->broadcast_initial && !pr->vals_bound;
    // (FAST arithmetic only: the GEMM path has no un-contracted operation order.  The choice depends on the chain count
    // of THIS call, so a FAST run sharded into pieces of < 256 chains takes the warp kernel: both paths are within the FAST
    // tolerance of the reference, but not bit-identical to each other; STRICT always takes the warp kernel.)
    const bool use_wide = wide_ok && st->arith == MCMCB200_ARITH_FAST && (pr->n_dim > 32 * MAX_EPL || pr->n_chains >= 256);
    if (wide_ok && !use_wide && pr->n_dim > 32 * MAX_EPL) {
        set_error("mala: n_dim=%d > %d runs on the chain-batched tensor-core path, which has no STRICT arithmetic", pr->n_dim, 32 * MAX_EPL);
        return MCMCB200_ERR_UNSUPPORTED;
    }
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
        // with box constraints the proposal covariance is eps^2 J(proposal) M, different at every draw: the kernel then needs
        // M^-1 itself (Sigma^-1 r = M^-1 (r / (J eps^2)), mala.cu)
        if (!host_cholesky_colmajor(st->precond_mat, d, st->chol_mode, S.data()) ||
            !host_inverse_colmajor(pr->vals_bound ? st->precond_mat : Sigma.data(), d, SigInv.data())) {
            set_error("precond_mat is not positive definite");
            return MCMCB200_ERR_INVALID_ARG;
        }
        if ((rc = upload(s.lease, s.scope.dev, SLOT_MAT_A, S.data(), nn, s.stream, &a.S_cm))) return rc;
        if ((rc = upload(s.lease, s.scope.dev, SLOT_MAT_B, SigInv.data(), nn, s.stream, &a.SigInv_cm))) return rc;
        if ((rc = upload(s.lease, s.scope.dev, SLOT_MAT_C, st->precond_mat, nn, s.stream, &a.M_cm))) return rc;
        MCMCB200_CUDA_TRY(cudaStreamSynchronize(s.stream));
    }
    int launches = 1;
    if (use_wide) {
        void* wp = nullptr;
        if ((rc = pool_get(s.lease, s.scope.dev, SLOT_WORK, (size_t)mala_wide_work_doubles(pr->n_chains, pr->n_dim) * sizeof(double), &wp))) return rc;
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

static int rwmh_run_impl(const mcmcb200_problem_t* pr, const mcmcb200_rng_t* rng, const mcmcb200_rwmh_settings_t* st,
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
        if ((rc = upload(s.lease, s.scope.dev, SLOT_MAT_A, S.data(), nn, s.stream, &a.S_cm))) return rc;
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

void mcmcb200_de_settings_default(mcmcb200_de_settings_t* s)
{
    std::memset(s, 0, sizeof(*s));
    s->n_burnin_draws = 1000;  // mcmc_structs.hpp:44-62
    s->n_keep_draws = 1000;
    s->n_pop = 100;
    s->jumps = 0;
    s->par_b = 1e-4;
    s->par_gamma_jump = 2.0;
    s->arith = MCMCB200_ARITH_FAST;
}

static int de_run_impl(const mcmcb200_problem_t* pr, const mcmcb200_rng_t* rng, const mcmcb200_de_settings_t* st, mcmcb200_output_t* out)
{
    if (!st) { set_error("null settings"); return MCMCB200_ERR_INVALID_ARG; }
    if (st->n_pop < 3 || st->n_pop > 1000000) { set_error("de: n_pop must be in [3, 1e6] (a proposal needs two other members)"); return MCMCB200_ERR_INVALID_ARG; }
    if (out && out->logp_out) { set_error("de: logp_out is not available"); return MCMCB200_ERR_UNSUPPORTED; }
    if (pr && pr->broadcast_initial) { set_error("de: broadcast_initial is not supported"); return MCMCB200_ERR_UNSUPPORTED; }
    if (pr && pr->initial_mem == MCMCB200_MEM_DEVICE) { set_error("de: initial_vals must be host memory (the sampling box is derived from it on the host)"); return MCMCB200_ERR_UNSUPPORTED; }
    Staged s;
    const long long n_pop = st->n_pop;
    // stage_common sizes draws_out as n_chains * n_keep * n_dim: a "kept draw" of a population is its whole n_pop x n_dim matrix
    if (st->n_keep_draws < 0 || st->n_burnin_draws < 0 || st->n_keep_draws * n_pop > 0x7ffffff0ll) { set_error("draw counts out of range"); return MCMCB200_ERR_INVALID_ARG; }
    int rc = stage_common(s, pr, rng, st->arith, 0, st->n_keep_draws * n_pop, 0, false, out);
    if (rc) return rc;
    const int d = pr->n_dim;
    const long long C = pr->n_chains, n_total = st->n_burnin_draws + st->n_keep_draws;
    if (n_total > 0x7ffffff0ll / n_pop) { set_error("draw counts out of range"); return MCMCB200_ERR_INVALID_ARG; }
    DeLaunch a;
    static_cast<CommonLaunch&>(a) = s.c;
    a.n_burnin = st->n_burnin_draws;
    a.n_keep = st->n_keep_draws;
    a.n_pop = (int)n_pop;
    a.jumps = st->jumps ? 1 : 0;
    a.par_b = st->par_b;
    a.gamma = 2.38 / std::sqrt(2.0 * (double)(size_t)d);   // src/de.cpp:59 (settings.par_gamma is never read)
    a.gamma_jump = st->par_gamma_jump;
    // sampling box of the initial population, per population (src/de.cpp:70-71), clamped to the hard bounds
    // (sampling_bounds_check, include/misc/bounds_check.hpp:27-57)
    std::vector<double> lo((size_t)C * d), hi((size_t)C * d);
    for (long long c = 0; c < C; ++c)
        for (int j = 0; j < d; ++j) {
            double l = st->initial_lb ? st->initial_lb[j] : pr->initial_vals[c * d + j] + (-0.5);
            double u = st->initial_ub ? st->initial_ub[j] : pr->initial_vals[c * d + j] + 0.5;
            if (pr->vals_bound) {
                const bool has_l = std::isfinite(pr->lower_bounds[j]), has_u = std::isfinite(pr->upper_bounds[j]);
                if (has_l) l = std::max(pr->lower_bounds[j], l);
                if (has_u) u = std::min(pr->upper_bounds[j], u);
            }
            lo[(size_t)c * d + j] = l;
            hi[(size_t)c * d + j] = u;
        }
    if ((rc = upload(s.lease, s.scope.dev, SLOT_MAT_A, lo.data(), lo.size(), s.stream, &a.init_lb))) return rc;
    if ((rc = upload(s.lease, s.scope.dev, SLOT_MAT_B, hi.data(), hi.size(), s.stream, &a.init_ub))) return rc;
    a.init_per_pop = 1;
    const long long stride = n_pop * d + n_total * n_pop * (d + 3);
    std::vector<double> tape;
    if (s.deferred_mt_tape) {   // MCMCB200_RNG_MT19937_TAPE: the reference's own stream, one std::mt19937_64 per population
        tape.resize((size_t)C * (size_t)stride);
        unsigned nt = std::thread::hardware_concurrency();
        if (nt == 0) nt = 1;
        if ((long long)nt > C) nt = (unsigned)C;
        std::vector<std::thread> th;
        for (unsigned ti = 0; ti < nt; ++ti)
            th.emplace_back([&, ti]() {
                for (long long c = ti; c < C; c += nt)
                    host_de_tape(rng->seed + (uint64_t)(pr->chain_offset + c), n_pop, d, n_total, st->par_b, tape.data() + (size_t)c * (size_t)stride);
            });
        for (auto& t : th) t.join();
        if ((rc = upload(s.lease, s.scope.dev, SLOT_TAPE, tape.data(), tape.size(), s.stream, &a.rng.tape))) return rc;
        a.rng.tape_stride = stride;
    } else if (rng->mode == MCMCB200_RNG_USER_TAPE && rng->tape_stride < stride) {
        set_error("de: USER_TAPE tape_stride %lld is shorter than the %lld variates a population consumes", (long long)rng->tape_stride, (long long)stride);
        return MCMCB200_ERR_INVALID_ARG;
    }
    MCMCB200_CUDA_TRY(cudaStreamSynchronize(s.stream));   // lo / hi / tape go out of scope
    const int dp = (d + 1) & ~1;
    void* wp = nullptr;
    if ((rc = pool_get(s.lease, s.scope.dev, SLOT_WORK, (size_t)C * (size_t)n_pop * dp * sizeof(double), &wp))) return rc;
    a.work = static_cast<double*>(wp);
    MCMCB200_CUDA_TRY(cudaEventRecord(s.ev0, s.stream));
    if ((rc = launch_de(a))) return rc;
    MCMCB200_CUDA_TRY(cudaEventRecord(s.ev1, s.stream));
    if (out->n_leapfrog_out)
        for (long long c = 0; c < C; ++c) out->n_leapfrog_out[c] = 0;
    if (out->step_size_out)
        for (long long c = 0; c < C; ++c) out->step_size_out[c] = a.gamma;
    return finish_common(s, out, 1);
}

static int nuts_run_impl(const mcmcb200_problem_t* pr, const mcmcb200_rng_t* rng, const mcmcb200_nuts_settings_t* st,
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
    if ((rc = stage_precond(s.lease, s.scope.dev, s.stream, st->precond_mat, pr->n_dim, st->chol_mode, &a.S_cm, &a.Minv_cm, nullptr))) return rc;
    void* p = nullptr;
    if ((rc = pool_get(s.lease, s.scope.dev, SLOT_STEP, (size_t)pr->n_chains * sizeof(double), &p))) return rc;
    a.step_out = static_cast<double*>(p);
    if ((rc = pool_get(s.lease, s.scope.dev, SLOT_NLF, (size_t)pr->n_chains * sizeof(long long), &p))) return rc;
    a.n_leapfrog = static_cast<long long*>(p);
    a.work_stride = nuts_work_doubles_per_chain(pr->n_dim, a.max_depth);
    // dense quadratic target, many chains, FAST arithmetic: chain-batched rounds — one fp64 tensor-core GEMM per round for the
    // gradient products of ALL chains + a resumable per-chain state machine (nuts_batched.cu); everything else: nuts.cu
    const bool batched = !s.deferred_mt_tape && pr->target_id < MCMCB200_USER_TARGET_BASE &&
                         nuts_batched_supported(pr->target_id, pr->n_dim, a.S_cm != nullptr, a.strict, a.lb != nullptr, pr->n_chains, a.max_depth);
    const size_t work_doubles = batched ? (size_t)nuts_batched_work_doubles(pr->n_chains, pr->n_dim, a.max_depth)
                                        : (size_t)pr->n_chains * (size_t)a.work_stride;
    if ((rc = pool_get(s.lease, s.scope.dev, SLOT_WORK, work_doubles * sizeof(double), &p))) return rc;
    a.work = static_cast<double*>(p);
    // dense targets with enough chains to fill the GPU run 8 chains per CTA with cooperative gradients (nuts.cu);
    // MCMCB200_NUTS_COOP=0/1 forces the choice (tests compare the two kernels bit for bit)
    // by default only for the target whose cooperative path is covered by the GPU parity tests (dense_gauss, the C4 target);
    // linreg has the same kernel instantiated and can be switched on with the environment variable
    a.coop = pr->n_chains >= 64 && pr->target_id == MCMCB200_TARGET_DENSE_GAUSS;
    if (const char* e = std::getenv("MCMCB200_NUTS_COOP")) a.coop = (e[0] == '1');
    a.coop_batch = 6;   // measured on B200 (C4 shape, 1184 chains x 40 draws): 2 -> 456 ms, 4 -> 332, 6 -> 314, 8 -> 332
    if (const char* e = std::getenv("MCMCB200_NUTS_BATCH")) a.coop_batch = std::atoi(e) > 0 ? std::atoi(e) : 1;
    a.coop_prefetch = true;
    if (const char* e = std::getenv("MCMCB200_NUTS_PREFETCH")) a.coop_prefetch = (e[0] != '0');
    a.coop_dmma = true;
    if (const char* e = std::getenv("MCMCB200_NUTS_DMMA")) a.coop_dmma = (e[0] != '0');
    a.t_begin = 0;
    a.t_end = n_total;
    a.save_state = false;
    a.tape_used = nullptr;
    int launches = 1;
    MCMCB200_CUDA_TRY(cudaEventRecord(s.ev0, s.stream));
    if (batched) {
        if ((rc = launch_nuts_batched(a, a.work, &launches, nullptr))) return rc;
    } else if (!s.deferred_mt_tape) {
        if ((rc = launch_nuts(a))) return rc;
    } else {
        // Reference-stream mode (MCMCB200_RNG_MT19937_TAPE).  The reference draws, per iteration, n_dim normals and then a
        // data-dependent number of uniforms from ONE serial std::mt19937_64 (src/nuts.cpp:199-206,233,261, nuts.ipp:214),
        // so the stream cannot be laid out in advance.  The kernel is driven draw by draw: the host generates the draw's
        // normals plus a look-ahead pool of the uniforms the tree MAY consume (at most 2^D + D for max_tree_depth D: one
        // slice variable, per doubling a direction, its 2^j - 1 merges and the accept test), the kernel reports how many it
        // used, and the chain's engine is advanced by exactly that count — each bmo::stats::runif is one raw 64-bit draw.
        const int D = a.max_depth;
        if (D > 16) { set_error("nuts: MT19937 reference-stream mode supports max_tree_depth <= 16 (look-ahead pool of 2^D uniforms)"); return MCMCB200_ERR_UNSUPPORTED; }
        const long long pool = (1ll << D) + D + 1;
        const long long C = pr->n_chains;
        const int d = pr->n_dim;
        HostMtStreams* ms = host_mt_streams_create(rng->seed + (uint64_t)pr->chain_offset, C);
        struct Guard { HostMtStreams* p; ~Guard() { host_mt_streams_destroy(p); } } guard{ms};
        std::vector<double> tape_h((size_t)C * (size_t)(2 * d + pool));
        std::vector<long long> used_h((size_t)C);
        void* tp = nullptr;
        if ((rc = pool_get(s.lease, s.scope.dev, SLOT_TAPE, tape_h.size() * sizeof(double), &tp))) return rc;
        void* up = nullptr;
        if ((rc = pool_get(s.lease, s.scope.dev, SLOT_AUX, (size_t)C * sizeof(long long), &up))) return rc;
        a.rng.tape = static_cast<const double*>(tp);
        a.tape_used = static_cast<long long*>(up);
        a.save_state = true;
        launches = 0;
        const long long n_seg = n_total > 0 ? n_total : 1;   // with no draws at all one launch still performs the set-up
        for (long long t = 0; t < n_seg; ++t) {
            const long long n_norm = (t == 0 ? 2ll : 1ll) * d;   // the pre-loop momentum draw precedes draw 0
            const long long stride = n_norm + pool;
            host_mt_streams_fill(ms, (n_total > 0) ? n_norm : d, pool, tape_h.data());
            MCMCB200_CUDA_TRY(cudaMemcpyAsync(tp, tape_h.data(), (size_t)C * (size_t)((n_total > 0 ? n_norm : d) + pool) * sizeof(double), cudaMemcpyHostToDevice, s.stream));
            a.rng.tape_stride = (n_total > 0) ? stride : d + pool;
            a.t_begin = t;
            a.t_end = (n_total > 0) ? t + 1 : 0;
            if ((rc = launch_nuts(a))) return rc;
            ++launches;
            MCMCB200_CUDA_TRY(cudaMemcpyAsync(used_h.data(), up, (size_t)C * sizeof(long long), cudaMemcpyDeviceToHost, s.stream));
            MCMCB200_CUDA_TRY(cudaStreamSynchronize(s.stream));
            for (long long c = 0; c < C; ++c) used_h[(size_t)c] -= (n_total > 0) ? n_norm : d;
            host_mt_streams_advance(ms, used_h.data());
        }
    }
    MCMCB200_CUDA_TRY(cudaEventRecord(s.ev1, s.stream));
    if (out->step_size_out)
        MCMCB200_CUDA_TRY(cudaMemcpyAsync(out->step_size_out, a.step_out, (size_t)pr->n_chains * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    if (out->n_leapfrog_out)
        MCMCB200_CUDA_TRY(cudaMemcpyAsync(out->n_leapfrog_out, a.n_leapfrog, (size_t)pr->n_chains * sizeof(int64_t), cudaMemcpyDeviceToHost, s.stream));
    return finish_common(s, out, launches);
}

static int rmhmc_run_impl(const mcmcb200_problem_t* pr, const mcmcb200_rng_t* rng, const mcmcb200_rmhmc_settings_t* st,
                       mcmcb200_output_t* out)
{
    if (!st) { set_error("null settings"); return MCMCB200_ERR_INVALID_ARG; }
    if (st->n_leap_steps < 0 || st->n_leap_steps > 0x7fffffff || st->n_fp_steps < 0 || st->n_fp_steps > 0x7fffffff) {
        set_error("bad n_leap_steps / n_fp_steps");
        return MCMCB200_ERR_INVALID_ARG;
    }
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
    // FAST arithmetic with a metric registered in contraction form (the funnel metrics): one CTA per chain, all metric algebra
    // in shared memory, no derivative cube (rmhmc_cta.cu).  MCMCB200_RMHMC_CTA=0 forces the cube kernel (tests compare them).
    bool cta = general && rmhmc_cta_applicable(pr->target_id, st->metric_id, pr->n_dim, st->arith == MCMCB200_ARITH_STRICT, pr->vals_bound != 0);
    if (const char* e = std::getenv("MCMCB200_RMHMC_CTA")) cta = cta && e[0] != '0';
    if (cta) {
        a.work_stride = rmhmc_cta_work_doubles(pr->n_dim);
        void* wp = nullptr;
        if ((rc = pool_get(s.lease, s.scope.dev, SLOT_WORK, (size_t)pr->n_chains * (size_t)a.work_stride * sizeof(double), &wp))) return rc;
        a.work = static_cast<double*>(wp);
    } else if (general) {
        const bool user_metric = pr->target_id >= MCMCB200_USER_TARGET_BASE;
        if (user_metric ? !(user_target_has(USER_LAUNCH_RMHMC, pr->target_id) && pr->n_dim <= 64)
                        : !rmhmc_general_supported(pr->target_id, st->metric_id, pr->n_dim)) {
            set_error("rmhmc: target %d has no registered metric %d for n_dim=%d (cube kernel: n_dim <= 64; FAST arithmetic with a contraction-form metric: n_dim <= 128)", pr->target_id, st->metric_id, pr->n_dim);
            return MCMCB200_ERR_UNSUPPORTED;
        }
        a.work_stride = rmhmc_general_work_doubles(pr->n_dim);
        void* wp = nullptr;
        if ((rc = pool_get(s.lease, s.scope.dev, SLOT_WORK, (size_t)pr->n_chains * (size_t)a.work_stride * sizeof(double), &wp))) return rc;
        a.work = static_cast<double*>(wp);
    }
    MCMCB200_CUDA_TRY(cudaEventRecord(s.ev0, s.stream));
    if (cta) {
        if ((rc = launch_rmhmc_cta(a))) return rc;
    } else if (general && pr->target_id >= MCMCB200_USER_TARGET_BASE) {
        RmhmcLaunch b = a;
        b.target_id = MCMCB200_TARGET_USER;
        if ((rc = user_target_launch(USER_LAUNCH_RMHMC, pr->target_id, &b))) return rc;
    } else if ((rc = general ? launch_rmhmc_general(a) : launch_rmhmc(a))) return rc;
    MCMCB200_CUDA_TRY(cudaEventRecord(s.ev1, s.stream));
    if (out->n_leapfrog_out)
        for (long long c = 0; c < pr->n_chains; ++c) out->n_leapfrog_out[c] = (st->n_burnin_draws + st->n_keep_draws) * st->n_leap_steps;
    if (out->step_size_out)
        for (long long c = 0; c < pr->n_chains; ++c) out->step_size_out[c] = st->step_size;
    return finish_common(s, out, 1);
}

#define MCMCB200_ENTRY(name)                                                                                             \
    int mcmcb200_##name##_run(const mcmcb200_problem_t* pr, const mcmcb200_rng_t* rng, const mcmcb200_##name##_settings_t* st, \
                              mcmcb200_output_t* out)                                                                    \
    {                                                                                                                    \
        return guarded([&] { return name##_run_impl(pr, rng, st, out); });                                               \
    }
MCMCB200_ENTRY(hmc)
MCMCB200_ENTRY(mala)
MCMCB200_ENTRY(rwmh)
MCMCB200_ENTRY(nuts)
MCMCB200_ENTRY(rmhmc)
MCMCB200_ENTRY(de)
#undef MCMCB200_ENTRY

int mcmcb200_target_lookup(const char* name)
{
    if (!name) return -1;
    static const struct { const char* n; int id; } tbl[] = {
        {"iso_gauss", MCMCB200_TARGET_ISO_GAUSS},     {"diag_gauss", MCMCB200_TARGET_DIAG_GAUSS},
        {"dense_gauss", MCMCB200_TARGET_DENSE_GAUSS}, {"linreg", MCMCB200_TARGET_LINREG},
        {"normal_model", MCMCB200_TARGET_NORMAL_MODEL}, {"funnel", MCMCB200_TARGET_FUNNEL}};
    for (const auto& e : tbl)
        if (std::strcmp(e.n, name) == 0) return e.id;
    std::lock_guard<std::mutex> g(g_user_mu);
    for (size_t k = 0; k < user_targets().size(); ++k)
        if (user_targets()[k].name == name) return MCMCB200_USER_TARGET_BASE + (int)k;
    return -1;
}

int mcmcb200_register_target(const char* name, const mcmcb200_user_target_t* table)
{
    if (!name || !*name || !table) { set_error("register_target: null argument"); return -MCMCB200_ERR_INVALID_ARG; }
    if (table->abi_version != MCMCB200_USER_ABI) {
        set_error("register_target(%s): built against user ABI %u, this library has %u — rebuild the target library against this mcmc_b200", name,
                  table->abi_version, MCMCB200_USER_ABI);
        return -MCMCB200_ERR_UNSUPPORTED;
    }
    if (mcmcb200_target_lookup(name) >= 0) { set_error("register_target: the name %s is taken", name); return -MCMCB200_ERR_INVALID_ARG; }
    std::lock_guard<std::mutex> g(g_user_mu);
    user_targets().push_back(UserTarget{name, *table});
    return MCMCB200_USER_TARGET_BASE + (int)user_targets().size() - 1;
}

int64_t mcmcb200_target_data_len(int target_id, int32_t n_dim) { return target_data_len(target_id, n_dim); }

int mcmcb200_metric_lookup(const char* name, int* target_id_out, int* metric_id_out)
{
    static const struct { const char* n; int target, metric; } tbl[] = {
        {"normal_fisher", MCMCB200_TARGET_NORMAL_MODEL, 0}, {"funnel_fisher", MCMCB200_TARGET_FUNNEL, 1}, {"funnel_softabs", MCMCB200_TARGET_FUNNEL, 2}};
    if (name)
        for (const auto& e : tbl)
            if (std::strcmp(e.n, name) == 0) {
                if (target_id_out) *target_id_out = e.target;
                if (metric_id_out) *metric_id_out = e.metric;
                return MCMCB200_OK;
            }
    set_error("unknown metric %s", name ? name : "(null)");
    return MCMCB200_ERR_UNKNOWN_TARGET;
}

int mcmcb200_target_eval(int target_id, const double* target_data, int64_t target_data_len_, int32_t n_dim, int64_t n_points,
                         const double* x, double* value_out, double* grad_out, int32_t arith)
{
    const int64_t need = target_data_len(target_id, n_dim);
    if (need < 0) { set_error("unknown target id %d", target_id); return MCMCB200_ERR_UNKNOWN_TARGET; }
    if (target_data_len_ < need || !x || !value_out || n_points <= 0) { set_error("bad arguments"); return MCMCB200_ERR_INVALID_ARG; }
    if (epl_for_dim(n_dim) == 0) { set_error("n_dim=%d unsupported", n_dim); return MCMCB200_ERR_UNSUPPORTED; }
    DeviceScope sc;
    Lease lease;
    int rc = sc.enter(-1);
    if (rc) return rc;
    EvalLaunch a{};
    a.target_id = target_id;
    a.d = n_dim;
    a.n_points = n_points;
    a.strict = (arith == MCMCB200_ARITH_STRICT);
    a.stream = nullptr;
    if ((rc = upload(lease, sc.dev, SLOT_TDATA, target_data, target_id >= MCMCB200_USER_TARGET_BASE ? (size_t)target_data_len_ : (size_t)need, nullptr, &a.tdata))) return rc;
    if ((rc = upload(lease, sc.dev, SLOT_EVAL_X, x, (size_t)n_points * n_dim, nullptr, &a.x))) return rc;
    void* p = nullptr;
    if ((rc = pool_get(lease, sc.dev, SLOT_EVAL_V, (size_t)n_points * sizeof(double), &p))) return rc;
    a.value = static_cast<double*>(p);
    a.grad = nullptr;
    if (grad_out) {
        if ((rc = pool_get(lease, sc.dev, SLOT_EVAL_G, (size_t)n_points * n_dim * sizeof(double), &p))) return rc;
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

int mcmcb200_de_tape(uint64_t seed, int64_t n_pop, int32_t n_dim, int64_t n_gen, double par_b, double* tape_out)
{
    if (!tape_out || n_pop < 3 || n_gen < 0 || n_dim <= 0) { set_error("bad arguments"); return MCMCB200_ERR_INVALID_ARG; }
    host_de_tape(seed, n_pop, n_dim, n_gen, par_b, tape_out);
    return MCMCB200_OK;
}

int mcmcb200_philox_stream(uint64_t seed, int64_t chain, int64_t draw, int32_t n_dim, int32_t n_unif, double* out)
{
    if (!out || n_dim <= 0 || n_unif < 0 || epl_for_dim(n_dim) == 0) { set_error("bad arguments"); return MCMCB200_ERR_INVALID_ARG; }
    DeviceScope sc;
    Lease lease;
    int rc = sc.enter(-1);
    if (rc) return rc;
    void* p = nullptr;
    if ((rc = pool_get(lease, sc.dev, SLOT_EVAL_V, (size_t)(n_dim + n_unif) * sizeof(double), &p))) return rc;
    if ((rc = launch_philox_stream(seed, chain, draw, n_dim, n_unif, static_cast<double*>(p), nullptr)))
        return rc;
    MCMCB200_CUDA_TRY(cudaMemcpy(out, p, (size_t)(n_dim + n_unif) * sizeof(double), cudaMemcpyDeviceToHost));
    return MCMCB200_OK;
}

// Page-locked host buffers.  Pinning 4 GB costs about a second, so freed buffers are kept (at most 4, reused best-fit when a
// request is no less than half their size) until mcmcb200_release_workspace(): a caller that samples repeatedly — the C++
// wrapper stages every many-chain draws_out through one — pays for the pinning once.
struct PinnedBuf { void* p; size_t bytes; };
static std::mutex g_pinned_mu;
static std::vector<PinnedBuf>& pinned_cache()
{
    static std::vector<PinnedBuf>* v = new std::vector<PinnedBuf>;
    return *v;
}
static std::vector<PinnedBuf>& pinned_live()
{
    static std::vector<PinnedBuf>* v = new std::vector<PinnedBuf>;
    return *v;
}
void* mcmcb200_host_alloc(size_t bytes)
{
    if (bytes == 0) bytes = 8;
    {
        std::lock_guard<std::mutex> g(g_pinned_mu);
        std::vector<PinnedBuf>& c = pinned_cache();
        size_t best = c.size();
        for (size_t i = 0; i < c.size(); ++i)
            if (c[i].bytes >= bytes && c[i].bytes <= 2 * bytes && (best == c.size() || c[i].bytes < c[best].bytes)) best = i;
        if (best != c.size()) {
            PinnedBuf b = c[best];
            c.erase(c.begin() + (long)best);
            pinned_live().push_back(b);
            return b.p;
        }
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaHostAlloc(%zu bytes) failed", bytes);
        return nullptr;
    }
    std::lock_guard<std::mutex> g(g_pinned_mu);
    pinned_live().push_back(PinnedBuf{p, bytes});
    return p;
}
void mcmcb200_host_free(void* p)
{
    if (!p) return;
    PinnedBuf b{p, 0};
    void* evict = nullptr;
    {
        std::lock_guard<std::mutex> g(g_pinned_mu);
        std::vector<PinnedBuf>& l = pinned_live();
        for (size_t i = 0; i < l.size(); ++i)
            if (l[i].p == p) { b = l[i]; l.erase(l.begin() + (long)i); break; }
        if (b.bytes == 0) { evict = p; }   // not ours to cache (unknown size): just free it
        else {
            std::vector<PinnedBuf>& c = pinned_cache();
            c.push_back(b);
            if (c.size() > 4) {   // keep the four largest
                size_t smallest = 0;
                for (size_t i = 1; i < c.size(); ++i)
                    if (c[i].bytes < c[smallest].bytes) smallest = i;
                evict = c[smallest].p;
                c.erase(c.begin() + (long)smallest);
            }
        }
    }
    if (evict) cudaFreeHost(evict);
}

int mcmcb200_fp64_peak(int32_t device, double* tflops_out)
{
    if (!tflops_out) { set_error("null argument"); return MCMCB200_ERR_INVALID_ARG; }
    DeviceScope sc;
    Lease lease;
    int rc = sc.enter(device);
    if (rc) return rc;
    void* p = nullptr;
    if ((rc = pool_get(lease, sc.dev, SLOT_EVAL_V, 64, &p))) return rc;
    rc = launch_fp64_peak(static_cast<double*>(p), nullptr, tflops_out);
    lease.settled = true;
    return rc;
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
    {
        std::lock_guard<std::mutex> g(g_pool_mu);
        pool_trim_locked(dev, 0);
    }
    std::vector<PinnedBuf> drop;
    {
        std::lock_guard<std::mutex> g(g_pinned_mu);
        drop.swap(pinned_cache());
    }
    for (const PinnedBuf& b : drop) cudaFreeHost(b.p);
}

}  // extern "C"
