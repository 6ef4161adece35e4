// mcmc_b200_device.cuh — write your own log-density (and RM-HMC metric) as a __device__ functor, in YOUR .cu, and run
// the library's sampler kernels on it without touching or rebuilding libmcmc_b200.so.
//
// This is the device-side replacement for the reference's callback arguments
//   std::function<fp_t (const ColVec_t& vals_inp, ColVec_t* grad_out, void* target_data)> target_log_kernel
//       (/root/reference/include/mcmc/hmc.hpp:43-58; contract: examples/eigen/hmc_normal.cpp:44-76 — grad_out may be null,
//        the return value is log pi)
//   std::function<Mat_t (const ColVec_t& vals_inp, Cube_t* tensor_deriv_out, void* tensor_data)> tensor_fn
//       (include/mcmc/rmhmc.hpp:47-66)
// A std::function cannot run on the GPU; a functor compiled into the kernels can.  Usage (examples/user_target/):
//
//     #define MCMCB200_USER_TARGET_TAG normal_raw            // a C identifier, unique per target library
//     #define MCMCB200_USER_MAX_EPL 2                        // optional: largest tile to instantiate (see "Tile sizes")
//     #include "mcmc_b200_device.cuh"                        // the warp-level building blocks a functor may use
//     struct NormalRaw {                                      // the functor concept of mcmc_b200/csrc/targets.cuh
//         static constexpr bool needs_scratch = false, dense_matrix = false, separable = false, per_element_data = false;
//         template <int EPL, bool STRICT, bool WANT_VALUE, bool WANT_GRAD, bool REDUCE = true, class Ctx = mcmcb200::WarpCtx>
//         static __device__ __forceinline__ double eval(const double* data, const Ctx& w, const double (&x)[EPL], double (&g)[EPL]);
//     };
//     static int64_t normal_raw_data_len(int32_t n_dim) { return n_dim == 2 ? 1 : -1; }
//     #define MCMCB200_USER_FUNCTOR NormalRaw
//     #define MCMCB200_USER_TARGET_NAME "normal_raw"
//     #define MCMCB200_USER_DATA_LEN normal_raw_data_len
//     #include "mcmc_b200_register.cuh"                      // instantiates the sampler kernels + registers at load time
//
// and build it as a shared library next to libmcmc_b200.so:
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -shared -Xcompiler -fPIC \
//          -I<repo>/include -I<repo>/mcmc_b200/csrc my_target.cu -L<repo>/mcmc_b200 -lmcmc_b200 -o libmy_target.so
// Loading libmy_target.so (dlopen / linking it) registers "normal_raw"; mcmcb200_target_lookup("normal_raw") /
// mcmc::device_kernel("normal_raw") then returns its id, and every mcmcb200_*_run call accepts it.
//
// The functor contract (see targets.cuh for the built-in ones):
//   * x and g are LANE-STRIPED over the 32 lanes of the chain's warp: element j lives on lane (j % 64) / 2, slot
//     2 (j / 64) + (j % 2) (mcmcb200::elem_index(lane, k) gives j for slot k); slots with j >= w.d hold 0 and must stay 0 in g;
//   * every lane of the warp calls eval together (warp collectives such as mcmcb200::warp_sum are allowed);
//   * WANT_VALUE: return log pi(x) — the warp-uniform total when REDUCE, else this lane's partial sum (the caller adds the
//     lanes up); WANT_GRAD: write d log pi / dx into g;
//   * STRICT asks for un-contracted IEEE arithmetic in the reference's operation order (use mcmcb200::Ar<STRICT>), FAST may fuse;
//   * data is the target's blob in global memory (what the caller passed as target_data).
// Tile sizes: n_dim <= 64 needs EPL 2, <= 128 EPL 4, <= 256 EPL 8, <= 512 EPL 16; MCMCB200_USER_MAX_EPL caps what is
// compiled.  Samplers: define MCMCB200_USER_NO_NUTS / _NO_MALA / _NO_RWMH / _NO_DE / _NO_HMC before the include to skip a
// sampler's kernels (compile time); a skipped sampler reports MCMCB200_ERR_UNKNOWN_TARGET for this target.
// RM-HMC: define MCMCB200_USER_METRIC_TYPE to a struct with
//     template <bool STRICT> static __device__ void eval(const double* data, int d, int lane, const double* xs, double* G, double* dG);
// (G: d x d column-major; dG: d matrices of d x d or null; xs: the position, every lane may read any element; buffers are
// zero-initialised once, write your fixed sparsity pattern only — see FunnelSoftabsMetric in csrc/rmhmc_general.cu); n_dim <= 64.
#pragma once

#ifndef MCMCB200_USER_TARGET_TAG
#error "define MCMCB200_USER_TARGET_TAG (a C identifier unique to this target library) before including mcmc_b200_device.cuh"
#endif

#include "mcmc_b200.h"

// The kernel sources refer to the user's functor through this name; mcmc_b200_register.cuh makes it a typedef of
// MCMCB200_USER_FUNCTOR inside namespace mcmcb200.
#define MCMCB200_USER_TARGET_TYPE UserFunctor

#include "../mcmc_b200/csrc/engine.h"
#include "../mcmc_b200/csrc/warp.cuh"
#include "../mcmc_b200/csrc/rng.cuh"
#include "../mcmc_b200/csrc/targets.cuh"
#include "../mcmc_b200/csrc/box.cuh"
