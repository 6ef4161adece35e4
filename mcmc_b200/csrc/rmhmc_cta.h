// CTA-per-chain RM-HMC kernel (rmhmc_cta.cu): FAST arithmetic, metrics registered in contraction form, n_dim <= 64.
#pragma once
#include "engine.h"

namespace mcmcb200
{
bool rmhmc_cta_applicable(int target_id, int metric_id, int d, bool strict, bool has_bounds);
long long rmhmc_cta_work_doubles(int d);
int launch_rmhmc_cta(const RmhmcLaunch& a);
}
