// By-target dispatch for the samplers whose kernels are compiled one translation unit per registered target
// (hmc.cu, mala.cu, nuts.cu, rwmh.cu, de.cu with -DMCMCB200_TARGET_SLICE=k, see mcmc_b200/build.py): parallel compilation only,
// no behavioural content.
#include "engine.h"

namespace mcmcb200
{

#ifdef MCMCB200_FAST_BUILD   // developer build: only target 0 is compiled
#define MCMCB200_SLICE_OR_MISSING(call) \
    (set_error("this developer build (MCMCB200_FAST_BUILD) only contains the iso_gauss target"), MCMCB200_ERR_UNKNOWN_TARGET)
#else
#define MCMCB200_SLICE_OR_MISSING(call) call
#endif

#define DECL(k)                                        \
    int launch_hmc_slice##k(const HmcLaunch& a);       \
    int launch_mala_slice##k(const MalaLaunch& a);     \
    int launch_nuts_slice##k(const NutsLaunch& a);     \
    int launch_rwmh_slice##k(const RwmhLaunch& a);     \
    int launch_de_slice##k(const DeLaunch& a);
DECL(0)
#ifndef MCMCB200_FAST_BUILD
DECL(1) DECL(2) DECL(3) DECL(4) DECL(5)
#endif
#undef DECL

#define DISPATCH(fn, what, KIND)                                             \
    switch (a.target_id) {                                                   \
    case 0: return fn##_slice0(a);                                           \
    case 1: return MCMCB200_SLICE_OR_MISSING(fn##_slice1(a));                                         \
    case 2: return MCMCB200_SLICE_OR_MISSING(fn##_slice2(a));                                         \
    case 3: return MCMCB200_SLICE_OR_MISSING(fn##_slice3(a));                                         \
    case 4: return MCMCB200_SLICE_OR_MISSING(fn##_slice4(a));                                         \
    case 5: return MCMCB200_SLICE_OR_MISSING(fn##_slice5(a));                                         \
    default:                                                                 \
        if (user_target_has(KIND, a.target_id)) {   /* registered by a user's library (mcmc_b200_device.cuh) */ \
            auto b = a;                                                      \
            b.target_id = MCMCB200_TARGET_USER;                              \
            return user_target_launch(KIND, a.target_id, &b);                \
        }                                                                    \
        set_error(what ": unknown target id %d (or this sampler was not instantiated for it)", a.target_id); \
        return MCMCB200_ERR_UNKNOWN_TARGET;                                  \
    }

int launch_hmc(const HmcLaunch& a) { DISPATCH(launch_hmc, "hmc", USER_LAUNCH_HMC) }
int launch_mala(const MalaLaunch& a) { DISPATCH(launch_mala, "mala", USER_LAUNCH_MALA) }
int launch_nuts(const NutsLaunch& a) { DISPATCH(launch_nuts, "nuts", USER_LAUNCH_NUTS) }
int launch_rwmh(const RwmhLaunch& a) { DISPATCH(launch_rwmh, "rwmh", USER_LAUNCH_RWMH) }
int launch_de(const DeLaunch& a) { DISPATCH(launch_de, "de", USER_LAUNCH_DE) }

}  // namespace mcmcb200
