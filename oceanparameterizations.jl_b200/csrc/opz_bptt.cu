// opz_bptt.cu — discrete-adjoint (BPTT) kernels (placeholder until the kernel lands).
#include "opz_internal.h"
namespace opz {
int fp32_loss_grad(opz_ctx*, const opz_model*, const LossGradCall&) { set_error("BPTT path not built yet"); return OPZ_EUNSUP; }
}
