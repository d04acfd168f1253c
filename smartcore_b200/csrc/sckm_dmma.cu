// sckm_dmma.cu -- K1: FP64 DMMA tile kernel (placeholder until the tile kernel lands)
#include "sckm_common.cuh"
namespace sckm {
bool dmma_supported(const sckm_dataset*, uint64_t) { return false; }
int launch_assign_dmma(sckm_dataset* ds, uint64_t) { return fail(ds->ctx, SCKM_ERR_STATE, "DMMA kernel not built"); }
}
