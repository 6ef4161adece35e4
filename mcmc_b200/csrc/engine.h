// Internal host-side declarations shared by the C-ABI layer (engine.cu) and the
// per-sampler kernel translation units (hmc.cu, mala.cu, nuts.cu, rmhmc.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mcmc_b200.h"
#include "rng_args.h"

namespace mcmcb200
{

constexpr int WARPS_PER_BLOCK = 4;   // one warp per chain, 4 chains per CTA
constexpr int MCMCB200_TARGET_USER = 1000;   // id under which a user's translation unit instantiates its own functor
// A user's translation unit may cap the instantiated tile sizes (MCMCB200_USER_MAX_EPL = 2, 4, 8 or 16 elements per lane,
// i.e. n_dim <= 32 * that) to keep its compile time down; the library itself builds all of them.
#ifndef MCMCB200_USER_MAX_EPL
#define MCMCB200_USER_MAX_EPL 16
#endif
#define MCMCB200_EPL_CASE(n, expr)                                          \
    case n:                                                                 \
        if constexpr (n <= MCMCB200_USER_MAX_EPL && n <= target_max_epl<T>::value) { return expr; }          \
        else { set_error("n_dim needs %d elements per lane; this target / build stops at %d", n,          \
                         MCMCB200_USER_MAX_EPL < target_max_epl<T>::value ? MCMCB200_USER_MAX_EPL : target_max_epl<T>::value); return MCMCB200_ERR_UNSUPPORTED; }
constexpr int MAX_EPL = 16;          // n_dim <= 64*MAX_EPL/2 = 512 in the register-resident kernels

// everything below is DEVICE memory unless noted
struct CommonLaunch {
    long long n_chains;
    int d;
    int target_id;
    const double* tdata;
    const double* x0;
    int broadcast_x0;
    long long chain_offset;
    RngArgs rng;
    double* draws;          // [n_chains][n_keep][d]
    double* logp;           // [n_chains][n_keep] or null
    long long* n_accept;    // [n_chains]
    cudaStream_t stream;
    bool strict;
    const double* lb;       // box constraints (vals_bound): lower / upper bounds [d] on the device, or null
    const double* ub;
};

struct HmcLaunch : CommonLaunch {
    long long n_burnin, n_keep;
    int n_leap;
    double eps;
    const double* S_cm;     // sqrt factor, column-major, or null for M = I
    const double* Minv_cm;  // M^-1, column-major (symmetric), or null
};

struct MalaLaunch : CommonLaunch {
    long long n_burnin, n_keep;
    double eps;
    const double* M_cm;       // M (drift), column-major, or null
    const double* S_cm;       // sqrt factor (noise)
    const double* SigInv_cm;  // (eps^2 M)^-1
};

struct RwmhLaunch : CommonLaunch {
    long long n_burnin, n_keep;
    double par_scale;
    const double* S_cm;       // par_scale * chol(cov_mat), column-major, or null for cov = I
};

struct NutsLaunch : CommonLaunch {
    long long n_burnin, n_keep, n_adapt;
    int max_depth;
    double eps_bar0, delta, gamma, t0, kappa;
    const double* S_cm;
    const double* Minv_cm;
    double* step_out;          // [n_chains] or null
    long long* n_leapfrog;     // [n_chains] or null
    double* work;              // per-chain scratch for the memoised tree states
    long long work_stride;     // doubles per chain
    bool coop;                 // dense targets: 8 chains per CTA with cooperative gradients (nuts.cu)
    int coop_batch;            // requests that must be pending before busy warps attend a cooperative round
    bool coop_prefetch;        // cooperative kernel: request the next trajectory state's product before the tree logic of the current one
    bool coop_dmma;            // FAST arithmetic: cooperative products on the fp64 tensor cores (DMMA); MCMCB200_NUTS_DMMA=0 keeps the scalar body
    // segmented runs (reference-stream mode drives the kernel draw by draw from the host): this launch performs draws
    // [t_begin, t_end); with t_begin > 0 the chain state is reloaded from the work area, with save_state it is parked there
    long long t_begin, t_end;
    bool save_state;
    long long* tape_used;      // [n_chains] or null: tape doubles consumed by this launch (tape mode)
};

struct RmhmcLaunch : CommonLaunch {
    long long n_burnin, n_keep;
    int n_leap, n_fp;
    double eps;
    int chol_mode;
    double cons_term;  // 0.5 * n_dim * log(2 pi), computed in long double then narrowed (src/rmhmc.cpp:188)
    int metric_id;     // which of the target's registered metrics (0 = default)
    double* work;      // general kernel: per-chain scratch for the d x d matrices and the two derivative cubes
    long long work_stride;
};

struct DeLaunch : CommonLaunch {   // n_chains = number of independent populations
    long long n_burnin, n_keep;
    int n_pop;
    int jumps;
    double par_b, gamma, gamma_jump;
    const double* init_lb;   // sampling box of the initial population: [n_chains][d] (init_per_pop) or [d]
    const double* init_ub;
    int init_per_pop;
    double* work;            // [n_chains][n_pop][dp] population matrices
};

struct EvalLaunch {
    int target_id;
    const double* tdata;
    int d;
    long long n_points;
    const double* x;
    double* value;
    double* grad;  // or null
    bool strict;
    cudaStream_t stream;
};

// each returns a cudaError_t-like status: 0 ok, MCMCB200_ERR_* otherwise (message via set_error)
int launch_hmc(const HmcLaunch& a);
// 4 warps per chain, 512 < n_dim <= 2048, separable targets, M = I (hmc_wide.cu)
bool hmc_wide_supported(int target_id, int d, bool has_precond);
int launch_hmc_wide(const HmcLaunch& a);
int launch_mala(const MalaLaunch& a);
// chain-batched path (mala_wide.cu): dense quadratic targets, M = I, n_dim <= 2048
bool mala_wide_supported(int target_id, int d, bool has_precond);
long long mala_wide_work_doubles(long long n_chains, int d);
int launch_mala_wide(const MalaLaunch& a, double* work, int* launches);
int launch_nuts(const NutsLaunch& a);
int launch_rwmh(const RwmhLaunch& a);
int launch_de(const DeLaunch& a);
int launch_rmhmc(const RmhmcLaunch& a);
// warp-per-chain kernel for general n_dim <= 64 (rmhmc_general.cu)
bool rmhmc_general_supported(int target_id, int metric_id, int d);
long long rmhmc_general_work_doubles(int d);
int launch_rmhmc_general(const RmhmcLaunch& a);
int launch_target_eval(const EvalLaunch& a);
int launch_philox_stream(unsigned long long seed, long long chain, long long draw, int d, int n_unif, double* out_dev,
                         cudaStream_t stream);
long long nuts_work_doubles_per_chain(int d, int max_depth);
int launch_fp64_peak(double* scratch_dev, cudaStream_t stream, double* tflops_out);

void set_error(const char* fmt, ...);
// user-registered targets (ids >= MCMCB200_USER_TARGET_BASE, engine.cu): kind = index into mcmcb200_user_target_t's launchers
enum { USER_LAUNCH_HMC = 0, USER_LAUNCH_MALA, USER_LAUNCH_NUTS, USER_LAUNCH_RWMH, USER_LAUNCH_DE, USER_LAUNCH_EVAL, USER_LAUNCH_RMHMC, USER_LAUNCH_COUNT };
int user_target_launch(int kind, int target_id, const void* launch_struct);
bool user_target_has(int kind, int target_id);
int epl_for_dim(int d);  // 2,4,8,16 or 0 if unsupported

// host helpers (host_linalg.cpp, host_tape.cpp)
bool host_inverse_colmajor(const double* A, int n, double* inv);
bool host_cholesky_colmajor(const double* A, int n, int chol_mode, double* L);
void host_mt19937_tape(uint64_t seed, long long n_pre_normals, long long n_draws, int d, double* out);
// the variates one reference DE population consumes (src/de.cpp:92-99,118-190): out[n_pop*d + n_gen*n_pop*(d+3)]
void host_de_tape(uint64_t seed, long long n_pop, int d, long long n_gen, double par_b, double* out);
// reference stream for samplers with a data-dependent uniform count (NUTS): one std::mt19937_64 per chain, advanced draw by draw
struct HostMtStreams;
HostMtStreams* host_mt_streams_create(uint64_t seed0, long long n_chains);
void host_mt_streams_destroy(HostMtStreams*);
// per chain: n_normals normals consumed from the engine (bmo rnorm semantics), then `pool` uniforms generated from a COPY of
// the engine (bmo runif semantics: each is exactly one raw 64-bit draw); out is [n_chains][n_normals + pool]
void host_mt_streams_fill(HostMtStreams*, long long n_normals, long long pool, double* out);
// advance chain c's engine by used[c] uniforms
void host_mt_streams_advance(HostMtStreams*, const long long* used);

}  // namespace mcmcb200

#define MCMCB200_CUDA_TRY(expr)                                                             \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ::mcmcb200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (_e == cudaErrorMemoryAllocation) ? MCMCB200_ERR_OOM : MCMCB200_ERR_CUDA; \
        }                                                                                   \
    } while (0)
