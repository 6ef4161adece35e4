// HMC for n_dim <= 32 (hmc_half.cu): two chains per warp, 16 lanes each.
#pragma once
#include "engine.h"

namespace mcmcb200
{
bool hmc_half_supported(const HmcLaunch& a);
int launch_hmc_half(const HmcLaunch& a);
}
