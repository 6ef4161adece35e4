// Chain-batched HMC (hmc_batched.cu): dense mass matrix and / or dense quadratic target, every d x d product as one fp64
// tensor-core GEMM over all chains; FAST arithmetic, n_dim even and <= 2048.
#pragma once
#include "engine.h"

namespace mcmcb200
{
bool hmc_batched_supported(int target_id, int d, bool has_precond, bool strict, bool has_bounds, long long n_chains);
long long hmc_batched_work_doubles(long long n_chains, int d);
int launch_hmc_batched(const HmcLaunch& a, double* work, int* launches);
}
