#include "engine.h"
namespace mcmcb200 {
int launch_nuts(const NutsLaunch&) { set_error("nuts kernel not built yet"); return MCMCB200_ERR_UNSUPPORTED; }
}
namespace mcmcb200 { long long nuts_work_doubles_per_chain(int, int) { return 1; } }
