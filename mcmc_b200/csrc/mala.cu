#include "engine.h"
namespace mcmcb200 {
int launch_mala(const MalaLaunch&) { set_error("mala kernel not built yet"); return MCMCB200_ERR_UNSUPPORTED; }
}
