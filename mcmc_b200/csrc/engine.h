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
int launch_rmhmc(const RmhmcLaunch& a);
// warp-per-chain kernel for general n_dim <= 64 (rmhmc_general.cu)
bool rmhmc_general_supported(int target_id, int metric_id, int d);
long long rmhmc_general_work_doubles(int d);
int launch_rmhmc_general(const RmhmcLaunch& a);
int launch_target_eval(const EvalLaunch& a);
int launch_philox_stream(unsigned long long seed, long long chain, long long draw, int d, int n_unif, double* out_dev,
                         cudaStream_t stream);
long long nuts_work_doubles_per_chain(int d, int max_depth);

void set_error(const char* fmt, ...);
int epl_for_dim(int d);  // 2,4,8,16 or 0 if unsupported

// host helpers (host_linalg.cpp, host_tape.cpp)
bool host_inverse_colmajor(const double* A, int n, double* inv);
bool host_cholesky_colmajor(const double* A, int n, int chol_mode, double* L);
void host_mt19937_tape(uint64_t seed, long long n_pre_normals, long long n_draws, int d, double* out);

}  // namespace mcmcb200

#define MCMCB200_CUDA_TRY(expr)                                                             \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ::mcmcb200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (_e == cudaErrorMemoryAllocation) ? MCMCB200_ERR_OOM : MCMCB200_ERR_CUDA; \
        }                                                                                   \
    } while (0)
