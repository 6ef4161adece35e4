// Low-occupancy HMC (hmc_duo.cu): two warps per chain — one generates the next draw's variates while the other runs the
// trajectory — for calls with fewer chains than the GPU has warp schedulers to fill (the strong-scaling shards).
#pragma once
#include "engine.h"

namespace mcmcb200
{
bool hmc_duo_supported(const HmcLaunch& a);
int launch_hmc_duo(const HmcLaunch& a);
}
