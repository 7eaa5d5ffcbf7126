// placeholder, replaced below
#include "ops.h"
std::vector<MapHandle> join_mono_batch(Context &, const std::vector<MapHandle> &, const std::vector<MapHandle> &)
{
    throw LsfmError(LSFM_ERR_ARG, "mono join not built yet");
}
