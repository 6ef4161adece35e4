#include "engine.h"
namespace mcmcb200 {
int launch_rmhmc(const RmhmcLaunch&) { set_error("rmhmc kernel not built yet"); return MCMCB200_ERR_UNSUPPORTED; }
}
