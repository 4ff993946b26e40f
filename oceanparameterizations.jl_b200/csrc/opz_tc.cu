// opz_tc.cu — tcgen05 tensor-core path (placeholder until the kernel lands).
#include "opz_internal.h"
namespace opz {
int tc_model_prepare(opz_model*) { set_error("tensor-core path not built yet"); return OPZ_EUNSUP; }
void tc_model_release(opz_model*) {}
int tc_rollout(opz_ctx*, const opz_model*, const RolloutCall&, int) { set_error("tensor-core path not built yet"); return OPZ_EUNSUP; }
}
