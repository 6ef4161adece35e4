// mcmc_b200_register.cuh — second half of the user-target recipe of mcmc_b200_device.cuh: include it AFTER the functor
// (and, for RM-HMC, the metric) is defined and MCMCB200_USER_FUNCTOR / MCMCB200_USER_TARGET_NAME / MCMCB200_USER_DATA_LEN
// are set.  It pulls the library's sampler kernel templates into this translation unit, instantiated for the user's
// functor only, and registers their launchers with libmcmc_b200.so when the user's library is loaded
// (mcmcb200_register_target, include/mcmc_b200.h).
#pragma once

#if !defined(MCMCB200_USER_FUNCTOR) || !defined(MCMCB200_USER_TARGET_NAME) || !defined(MCMCB200_USER_DATA_LEN)
#error "define MCMCB200_USER_FUNCTOR, MCMCB200_USER_TARGET_NAME and MCMCB200_USER_DATA_LEN before including mcmc_b200_register.cuh"
#endif

namespace mcmcb200
{
typedef MCMCB200_USER_FUNCTOR UserFunctor;
#ifdef MCMCB200_USER_METRIC_TYPE
typedef MCMCB200_USER_METRIC_TYPE UserMetric;
#endif
}

#ifndef MCMCB200_USER_NO_HMC
#include "../mcmc_b200/csrc/hmc.cu"
#endif
#ifndef MCMCB200_USER_NO_MALA
#include "../mcmc_b200/csrc/mala.cu"
#endif
#ifndef MCMCB200_USER_NO_NUTS
#include "../mcmc_b200/csrc/nuts.cu"
#endif
#ifndef MCMCB200_USER_NO_RWMH
#include "../mcmc_b200/csrc/rwmh.cu"
#endif
#ifndef MCMCB200_USER_NO_DE
#include "../mcmc_b200/csrc/de.cu"
#endif
#include "../mcmc_b200/csrc/util_kernels.cu"
#ifdef MCMCB200_USER_METRIC_TYPE
#include "../mcmc_b200/csrc/rmhmc_general.cu"
#endif

namespace
{
struct McmcB200UserRegistration {
    int id;
    McmcB200UserRegistration()
    {
        using namespace mcmcb200;
        mcmcb200_user_target_t t;
        t.abi_version = MCMCB200_USER_ABI;
        t.data_len = MCMCB200_USER_DATA_LEN;
        for (auto& l : t.launch) l = nullptr;
#ifndef MCMCB200_USER_NO_HMC
        t.launch[USER_LAUNCH_HMC] = [](const void* p) { return MCMCB200_SLICED(launch_hmc)(*static_cast<const HmcLaunch*>(p)); };
#endif
#ifndef MCMCB200_USER_NO_MALA
        t.launch[USER_LAUNCH_MALA] = [](const void* p) { return MCMCB200_SLICED(launch_mala)(*static_cast<const MalaLaunch*>(p)); };
#endif
#ifndef MCMCB200_USER_NO_NUTS
        t.launch[USER_LAUNCH_NUTS] = [](const void* p) { return MCMCB200_SLICED(launch_nuts)(*static_cast<const NutsLaunch*>(p)); };
#endif
#ifndef MCMCB200_USER_NO_RWMH
        t.launch[USER_LAUNCH_RWMH] = [](const void* p) { return MCMCB200_SLICED(launch_rwmh)(*static_cast<const RwmhLaunch*>(p)); };
#endif
#ifndef MCMCB200_USER_NO_DE
        t.launch[USER_LAUNCH_DE] = [](const void* p) { return MCMCB200_SLICED(launch_de)(*static_cast<const DeLaunch*>(p)); };
#endif
        t.launch[USER_LAUNCH_EVAL] = [](const void* p) { return MCMCB200_SLICED(launch_target_eval)(*static_cast<const EvalLaunch*>(p)); };
#ifdef MCMCB200_USER_METRIC_TYPE
        t.launch[USER_LAUNCH_RMHMC] = [](const void* p) { return launch_tm<UserFunctor, UserMetric>(*static_cast<const RmhmcLaunch*>(p)); };
#endif
        id = mcmcb200_register_target(MCMCB200_USER_TARGET_NAME, &t);
    }
};
static McmcB200UserRegistration mcmcb200_user_registration_instance;
}  // namespace
