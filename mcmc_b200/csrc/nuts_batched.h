// Chain-batched NUTS (nuts_batched.cu): dense quadratic targets, M = I, FAST arithmetic — every chain's next gradient product
// of a step is part of ONE fp64 tensor-core GEMM over all chains; the tree logic runs as a resumable per-chain state machine.
#pragma once
#include "engine.h"

namespace mcmcb200
{
bool nuts_batched_supported(int target_id, int d, bool has_precond, bool strict, bool has_bounds, long long n_chains, int max_depth);
long long nuts_batched_work_doubles(long long n_chains, int d, int max_depth);
// steps_out (optional): number of lock-step rounds (GEMM + step-kernel pairs) the run took
int launch_nuts_batched(const NutsLaunch& a, double* work, int* launches, long long* steps_out);
}
