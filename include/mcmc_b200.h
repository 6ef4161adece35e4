/*
 * mcmc_b200 — C ABI of the B200-native many-chain HMC / MALA / NUTS / RM-HMC engine.
 *
 * This header is the drop-in boundary for the hot path of kthohr/mcmc
 * (MCMCLib 2.1.0).  Each entry point replaces one reference interface; the
 * reference is a C++ library without an FFI layer, so "what the reference's FFI
 * would bind" is its public C++ API flattened to plain pointers and sizes:
 *
 *   mcmcb200_hmc_run    <- bool mcmc::hmc  (initial_vals, target_log_kernel, draws_out, target_data, settings)
 *                          include/mcmc/hmc.hpp:43-72,  src/hmc.cpp:30-254
 *   mcmcb200_mala_run   <- bool mcmc::mala (...)   include/mcmc/mala.hpp:43-73,  src/mala.cpp:30-235
 *   mcmcb200_nuts_run   <- bool mcmc::nuts (...)   include/mcmc/nuts.hpp:43-72,  src/nuts.cpp:30-359
 *   mcmcb200_rmhmc_run  <- bool mcmc::rmhmc(..., tensor_fn, ..., tensor_data, settings)
 *                          include/mcmc/rmhmc.hpp:47-86, src/rmhmc.cpp:30-325
 *   mcmcb200_rwmh_run   <- bool mcmc::rwmh (initial_vals, target_log_kernel (value only), draws_out, target_data, settings)
 *                          include/mcmc/rwmh.hpp:43-72, src/rwmh.cpp:30-199   (the gradient-free sibling)
 *   mcmcb200_de_run     <- bool mcmc::de   (initial_vals, target_log_kernel (value only), Cube_t& draws_out, target_data, settings)
 *                          include/mcmc/de.hpp:43-72, src/de.cpp:30-271   (the population sampler)
 *   mcmcb200_*_settings <- hmc_/mala_/nuts_/rmhmc_/rwmh_/de_settings_t + algo_settings_t
 *                          include/misc/mcmc_structs.hpp:66-134,151-184 (same field names and defaults)
 *
 * Differences forced by the device boundary:
 *   - the std::function log-kernel callback cannot cross to the GPU; callers pick a
 *     REGISTERED __device__ functor by id (mcmcb200_target_t) and hand over its data
 *     blob, which the library copies to the device;
 *   - one call runs MANY independent chains (the reference runs one); chain c uses
 *     rng seed  seed + chain_offset + c  in MT19937 mode, and Philox counter word
 *     chain_offset + c in Philox mode, so results do not depend on how chains are
 *     sharded across GPUs;
 *   - draws_out is chain-major: [n_chains][n_keep_draws][n_dim] doubles (each chain a
 *     row-major n_keep x n_dim matrix).  include/mcmc_b200.hpp converts to the
 *     reference's column-major Mat_t (n_keep x n_dim).
 *
 * The reference returns `true` unconditionally (src/hmc.cpp:207,226) and never
 * throws; these functions return MCMCB200_OK (0) on success and a non-zero code on
 * CUDA / argument errors (mcmcb200_last_error() gives the text).  There is no CPU
 * fallback: without a usable CUDA device every run call fails with
 * MCMCB200_ERR_CUDA.
 */
#ifndef MCMC_B200_H
#define MCMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCMCB200_VERSION_MAJOR 0
#define MCMCB200_VERSION_MINOR 1

/* return codes */
enum {
    MCMCB200_OK = 0,
    MCMCB200_ERR_INVALID_ARG = 1,
    MCMCB200_ERR_UNKNOWN_TARGET = 2,
    MCMCB200_ERR_UNSUPPORTED = 3, /* valid request outside the compiled kernel set (e.g. n_dim too large) */
    MCMCB200_ERR_CUDA = 4,
    MCMCB200_ERR_OOM = 5
};

/* registered __device__ log-density functors (mcmc_b200/csrc/targets.cuh) */
typedef enum {
    MCMCB200_TARGET_ISO_GAUSS = 0,   /* log pi = -1/2 |x|^2 ; no data                                   */
    MCMCB200_TARGET_DIAG_GAUSS = 1,  /* log pi = -1/2 sum_i w_i x_i^2 ; data = w[n_dim]                  */
    MCMCB200_TARGET_DENSE_GAUSS = 2, /* log pi = -1/2 x' P x ; data = P[n_dim^2], symmetric             */
    MCMCB200_TARGET_LINREG = 3,      /* log pi = -1/2 t' A t + b' t ; data = A[n_dim^2] (sym), b[n_dim]  */
    MCMCB200_TARGET_NORMAL_MODEL = 4, /* Normal(mu, sigma) likelihood of the examples/eigen/..._normal.cpp programs on
                                        sufficient statistics; n_dim = 2, data = {n, xbar, sum (x-xbar)^2};
                                        carries the Fisher metric used by mcmcb200_rmhmc_run */
    MCMCB200_TARGET_FUNNEL = 5       /* Neal's funnel (BASELINE config 5): x[0] = v ~ N(0, 3^2), x[i] | v ~ N(0, e^v);
                                        n_dim >= 2, no data; metrics for mcmcb200_rmhmc_run (n_dim <= 64):
                                        1 = "funnel_fisher" diag(1/9 + (n_dim-1)/2, e^-v, ..., e^-v) (default),
                                        2 = "funnel_softabs" SoftAbs of the Hessian, alpha = 1e6 (closed form) */
} mcmcb200_target_t;

typedef enum {
    MCMCB200_RNG_PHILOX = 0,       /* counter-based Philox4x32-10 generated in-kernel (production mode) */
    MCMCB200_RNG_MT19937_TAPE = 1, /* reference stream: the library replays std::mt19937_64(seed + chain) through
                                      the reference's BaseMatrixOps rnorm/runif semantics on the host and the
                                      kernel consumes that tape (drop-in parity mode; HMC/MALA/RM-HMC/RWMH) */
    MCMCB200_RNG_USER_TAPE = 2     /* caller-supplied stream of doubles per chain, consumed in order */
} mcmcb200_rng_mode_t;

typedef enum { MCMCB200_MEM_HOST = 0, MCMCB200_MEM_DEVICE = 1 } mcmcb200_mem_t;

typedef enum {
    MCMCB200_ARITH_FAST = 0,  /* fused multiply-add, merged half-kicks: <= 1e-10 from the reference */
    MCMCB200_ARITH_STRICT = 1 /* the reference's operation order, no contraction: element-wise bit-exact */
} mcmcb200_arith_t;

typedef enum {
    MCMCB200_CHOL_LOWER = 0,    /* true lower factor (Armadillo backend, core/cholesky.hpp:31)            */
    MCMCB200_CHOL_EIGEN_LLT = 1 /* Eigen backend's matrixLLT() full storage (core/cholesky.hpp:37, Q8)   */
} mcmcb200_chol_t;

/* What to sample from and where the chains start. */
typedef struct mcmcb200_problem {
    int64_t n_chains;
    int32_t n_dim;
    int32_t target_id;          /* mcmcb200_target_t */
    const double* target_data;  /* HOST pointer, target_data_len doubles (may be NULL when len == 0) */
    int64_t target_data_len;
    const double* initial_vals; /* [n_chains][n_dim] chain-major, or [n_dim] when broadcast_initial != 0 */
    int32_t initial_mem;        /* mcmcb200_mem_t */
    int32_t broadcast_initial;
    int64_t chain_offset;       /* global index of chain 0 of this call (multi-GPU sharding) */
    int32_t device;             /* CUDA device ordinal; -1 = current device */
    int32_t vals_bound;         /* algo_settings_t::vals_bound (mcmc_structs.hpp:159): box constraints on          */
    void* stream;               /* cudaStream_t to launch on (NULL = default stream) */
    const double* lower_bounds; /* HOST, n_dim entries, -inf = no lower bound (mcmc_structs.hpp:161-162,            */
    const double* upper_bounds; /* HOST, n_dim entries, +inf = no upper bound  determine_bounds_type.hpp:39-50)     */
} mcmcb200_problem_t;

typedef struct mcmcb200_rng {
    int32_t mode;  /* mcmcb200_rng_mode_t */
    int32_t tape_mem; /* for USER_TAPE */
    uint64_t seed; /* algo_settings_t::rng_seed_value (mcmc_structs.hpp:155) */
    const double* tape; /* USER_TAPE: [n_chains][tape_stride] */
    int64_t tape_stride;
} mcmcb200_rng_t;

/* hmc_settings_t (mcmc_structs.hpp:66-78) */
typedef struct mcmcb200_hmc_settings {
    int64_t n_burnin_draws; /* default 1000 */
    int64_t n_keep_draws;   /* default 1000 */
    int64_t n_leap_steps;   /* default 1    */
    double step_size;       /* default 1.0  */
    const double* precond_mat; /* HOST, n_dim^2 column-major mass matrix M, or NULL -> identity (src/hmc.cpp:57) */
    int32_t chol_mode;      /* mcmcb200_chol_t */
    int32_t arith;          /* mcmcb200_arith_t */
} mcmcb200_hmc_settings_t;

/* mala_settings_t (mcmc_structs.hpp:123-134) */
typedef struct mcmcb200_mala_settings {
    int64_t n_burnin_draws;
    int64_t n_keep_draws;
    double step_size;
    const double* precond_mat;
    int32_t chol_mode;
    int32_t arith;
} mcmcb200_mala_settings_t;

/* nuts_settings_t (mcmc_structs.hpp:82-101) */
typedef struct mcmcb200_nuts_settings {
    int64_t n_burnin_draws;
    int64_t n_keep_draws;
    int64_t n_adapt_draws;     /* default 1000 */
    double target_accept_rate; /* default 0.55 */
    int64_t max_tree_depth;    /* default 10   */
    double step_size;          /* eps_bar_0, default 1.0 */
    double gamma_val;          /* default 0.05 */
    double t0_val;             /* default 10   */
    double kappa_val;          /* default 0.75 */
    const double* precond_mat;
    int32_t chol_mode;
    int32_t arith;
} mcmcb200_nuts_settings_t;

/* rmhmc_settings_t (mcmc_structs.hpp:105-119); precond_mat is never read by the reference (Q18) */
typedef struct mcmcb200_rmhmc_settings {
    int64_t n_burnin_draws;
    int64_t n_keep_draws;
    int64_t n_leap_steps;
    double step_size;
    int64_t n_fp_steps; /* default 5 */
    int32_t chol_mode;
    int32_t arith;
    int32_t metric_id;  /* which of the target's registered metrics plays the reference's tensor_fn; 0 = its default */
    int32_t reserved0;
} mcmcb200_rmhmc_settings_t;

/* rwmh_settings_t (mcmc_structs.hpp:138-149) */
typedef struct mcmcb200_rwmh_settings {
    int64_t n_burnin_draws; /* default 1000 */
    int64_t n_keep_draws;   /* default 1000 */
    double par_scale;       /* default 1.0  */
    const double* cov_mat;  /* HOST, n_dim^2 column-major proposal covariance, or NULL -> identity (src/rwmh.cpp:57) */
    int32_t chol_mode;      /* mcmcb200_chol_t */
    int32_t arith;          /* mcmcb200_arith_t */
} mcmcb200_rwmh_settings_t;

/* de_settings_t (mcmc_structs.hpp:44-62).  par_gamma is kept out: the reference never reads it (src/de.cpp:58-59 uses
   2.38 / sqrt(2 n_vals)).  One call runs problem.n_chains independent POPULATIONS of n_pop members each;
   problem.initial_vals is [n_chains][n_dim] (the reference's initial_vals of each population);
   draws_out is [n_chains][n_keep_draws][n_pop][n_dim] (per population the reference's Cube_t: n_keep matrices of
   n_pop x n_vals, row-major here); n_accept_draws[p] counts accepted member updates after burn-in (src/de.cpp:198).
   MCMCB200_RNG_USER_TAPE stride per population: n_pop*n_dim + (n_burnin+n_keep)*n_pop*(n_dim+3), see csrc/de.cu. */
typedef struct mcmcb200_de_settings {
    int64_t n_burnin_draws; /* default 1000 */
    int64_t n_keep_draws;   /* default 1000 */
    int64_t n_pop;          /* default 100; >= 3 */
    int32_t jumps;          /* default 0 */
    int32_t arith;          /* mcmcb200_arith_t */
    double par_b;           /* default 1e-4 */
    double par_gamma_jump;  /* default 2.0 */
    const double* initial_lb; /* HOST [n_dim] or NULL -> initial_vals - 0.5 (src/de.cpp:70) */
    const double* initial_ub; /* HOST [n_dim] or NULL -> initial_vals + 0.5 (src/de.cpp:71) */
} mcmcb200_de_settings_t;

typedef enum {
    MCMCB200_LAYOUT_CHAIN_ROWS = 0, /* draws_out[chain][t][j]: each chain a row-major n_keep x n_dim matrix (default)          */
    MCMCB200_LAYOUT_COLMAJOR = 1    /* draws_out[chain][j][t]: each chain the reference's column-major n_keep x n_dim Mat_t
                                       (SURVEY Q23), transposed on the device before it is handed back                     */
} mcmcb200_layout_t;

typedef struct mcmcb200_output {
    double* draws_out;       /* [n_chains][n_keep_draws][n_dim] (see draws_layout) */
    int32_t draws_mem;       /* mcmcb200_mem_t */
    int32_t draws_layout;    /* mcmcb200_layout_t */
    int64_t* n_accept_draws; /* HOST, [n_chains]: post-burn-in acceptances per chain (src/hmc.cpp:196-199); may be NULL */
    double* logp_out;        /* optional, same memory space as draws_out: log pi of each kept draw, [n_chains][n_keep] */
    double* step_size_out;   /* optional HOST [n_chains]: NUTS step size after the last draw */
    int64_t* n_leapfrog_out; /* optional HOST [n_chains]: leapfrog steps actually computed */
    /* filled by the library */
    float kernel_ms;         /* device time of the sampling kernel(s), CUDA events on `stream` */
    int32_t kernel_launches; /* sampling-kernel launches issued by this call */
} mcmcb200_output_t;

/* defaults identical to the reference structs */
void mcmcb200_hmc_settings_default(mcmcb200_hmc_settings_t* s);
void mcmcb200_mala_settings_default(mcmcb200_mala_settings_t* s);
void mcmcb200_nuts_settings_default(mcmcb200_nuts_settings_t* s);
void mcmcb200_rmhmc_settings_default(mcmcb200_rmhmc_settings_t* s);
void mcmcb200_rwmh_settings_default(mcmcb200_rwmh_settings_t* s);
void mcmcb200_de_settings_default(mcmcb200_de_settings_t* s);

int mcmcb200_hmc_run(const mcmcb200_problem_t* problem, const mcmcb200_rng_t* rng,
                     const mcmcb200_hmc_settings_t* settings, mcmcb200_output_t* out);
int mcmcb200_mala_run(const mcmcb200_problem_t* problem, const mcmcb200_rng_t* rng,
                      const mcmcb200_mala_settings_t* settings, mcmcb200_output_t* out);
int mcmcb200_nuts_run(const mcmcb200_problem_t* problem, const mcmcb200_rng_t* rng,
                      const mcmcb200_nuts_settings_t* settings, mcmcb200_output_t* out);
int mcmcb200_rmhmc_run(const mcmcb200_problem_t* problem, const mcmcb200_rng_t* rng,
                       const mcmcb200_rmhmc_settings_t* settings, mcmcb200_output_t* out);

int mcmcb200_rwmh_run(const mcmcb200_problem_t* problem, const mcmcb200_rng_t* rng,
                      const mcmcb200_rwmh_settings_t* settings, mcmcb200_output_t* out);

int mcmcb200_de_run(const mcmcb200_problem_t* problem, const mcmcb200_rng_t* rng,
                    const mcmcb200_de_settings_t* settings, mcmcb200_output_t* out);

/* Target registry: id by name ("iso_gauss", "diag_gauss", "dense_gauss", "linreg", "normal_model", "funnel"), -1 if unknown;
   number of doubles the target's data blob must hold for a given n_dim (-1 if unknown / n_dim invalid). */
int mcmcb200_target_lookup(const char* name);
int64_t mcmcb200_target_data_len(int target_id, int32_t n_dim);
/* User-defined targets: the reference's log-kernel is an arbitrary callback (include/mcmc/hmc.hpp:43-58); on the device
   it is a __device__ functor compiled by the USER (include/mcmc_b200_device.cuh shows how: one struct + one macro in the
   user's own .cu, built into the user's shared library, no rebuild of libmcmc_b200.so).  The macro fills this table with
   launchers of the sampler kernels instantiated for that functor and calls mcmcb200_register_target() when the user's
   library is loaded; the returned id (>= MCMCB200_USER_TARGET_BASE) is used like a built-in one.  A NULL launcher means
   "sampler not instantiated" (the run call then fails with MCMCB200_ERR_UNKNOWN_TARGET). */
#define MCMCB200_USER_TARGET_BASE 64
#define MCMCB200_USER_ABI 2u
typedef struct mcmcb200_user_target {
    uint32_t abi_version;                   /* MCMCB200_USER_ABI: layout of the internal launch structs */
    int64_t (*data_len)(int32_t n_dim);      /* minimum doubles of target data for n_dim; < 0 = n_dim not valid for this target */
    int (*launch[8])(const void* launch_struct); /* hmc, mala, nuts, rwmh, de, target_eval, rmhmc (user metric), reserved */
} mcmcb200_user_target_t;
int mcmcb200_register_target(const char* name, const mcmcb200_user_target_t* table); /* id, or -(error code) */

/* Metric registry (the reference's tensor_fn, include/mcmc/rmhmc.hpp:51): "normal_fisher" (normal_model, id 0),
   "funnel_fisher" (funnel, id 1), "funnel_softabs" (funnel, id 2) -> the target it belongs to and the metric_id for
   mcmcb200_rmhmc_settings_t.  MCMCB200_ERR_UNKNOWN_TARGET if the name is not registered. */
int mcmcb200_metric_lookup(const char* name, int* target_id_out, int* metric_id_out);

/* Evaluate a registered functor on the device: x is HOST [n_points][n_dim]; value_out HOST [n_points];
   grad_out HOST [n_points][n_dim] or NULL.  Used to check device functors against host callbacks. */
int mcmcb200_target_eval(int target_id, const double* target_data, int64_t target_data_len, int32_t n_dim,
                         int64_t n_points, const double* x, double* value_out, double* grad_out, int32_t arith);

/* Host side of MCMCB200_RNG_MT19937_TAPE: the variates one reference chain consumes from
   std::mt19937_64(seed) — n_pre_normals, then per draw n_dim normals followed by 1 uniform —
   written to tape_out[(n_pre_normals + n_draws*(n_dim+1))].  Exposed so tests can compare it with the oracle. */
int mcmcb200_mt19937_tape(uint64_t seed, int64_t n_pre_normals, int64_t n_draws, int32_t n_dim, double* tape_out);

/* Host side of MCMCB200_RNG_MT19937_TAPE for mcmc::de: the variates one reference population consumes from
   std::mt19937_64(seed) (src/de.cpp:92-99,118-190), tape_out[n_pop*n_dim + n_gen*n_pop*(n_dim+3)]. */
int mcmcb200_de_tape(uint64_t seed, int64_t n_pop, int32_t n_dim, int64_t n_gen, double par_b, double* tape_out);

/* Raw device Philox stream (for tests): normals of draw `draw` for `chain`, then n_unif uniforms. HOST out. */
int mcmcb200_philox_stream(uint64_t seed, int64_t chain, int64_t draw, int32_t n_dim, int32_t n_unif, double* out);

/* Multi-GPU assembly of draws_out: one process (or host thread) per GPU; rank 0 creates a 128-byte id, every
   rank calls comm_init with it, then allgather_draws puts each rank's chain-major block local_dev[chains_per_rank[rank]]
   [n_keep][n_dim] into full_dev[sum chains][n_keep][n_dim] (rank order) on every rank, on `stream`, over NCCL / NVLink.
   NCCL is bound at run time (libnccl.so.2); MCMCB200_ERR_UNSUPPORTED if it is not installed.  All pointers are DEVICE. */
int mcmcb200_comm_unique_id(void* id_out, size_t id_bytes);
int mcmcb200_comm_init(const void* id, size_t id_bytes, int32_t world_size, int32_t rank, int32_t device, void** comm_out);
int mcmcb200_comm_destroy(void* comm);
int mcmcb200_allgather_draws(void* comm, const double* local_dev, const int64_t* chains_per_rank, int64_t n_keep, int32_t n_dim,
                             double* full_dev, void* stream);

/* Page-locked host memory for draws_out (cudaHostAlloc / cudaFreeHost): a D2H copy into it runs at the PCIe rate, a copy
   into pageable memory at a fraction of it.  NULL on failure. */
void* mcmcb200_host_alloc(size_t bytes);
void mcmcb200_host_free(void* p);

/* Measured fp64 FMA peak of the device (TFLOP/s, dependent-free DFMA loop on every SM): the roofline denominator of the
   compute-bound kernels, taken in the same process (bench.py). */
int mcmcb200_fp64_peak(int32_t device, double* tflops_out);

const char* mcmcb200_last_error(void);
int mcmcb200_device_count(void);
/* releases cached device scratch buffers of the calling thread's current device */
void mcmcb200_release_workspace(void);

#ifdef __cplusplus
}
#endif

#endif /* MCMC_B200_H */
