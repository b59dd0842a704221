// Kernel translation unit: FV field expressions, uniform-level flux schemes, linear combination, initial condition.
#include "launch_impl.cuh"

#define SMR_FV_DIM(D)                                                     \
    SMR_INST_BATCH(smr_item_fv, smr::FvOp<D, false>)                      \
    SMR_INST_BATCH(smr_item_fv, smr::FvOp<D, true>)                       \
    SMR_INST_BATCH(smr_item_fvstrip, smr::FvStripOp<D, false>)            \
    SMR_INST_BATCH(smr_item_fvstrip, smr::FvStripOp<D, true>)             \
    SMR_INST_BATCH(smr_item_fv, smr::FluxLinHomOp<D>)                     \
    SMR_INST_BATCH(smr_item_fvstrip, smr::FluxLinHomStripOp<D>)           \
    SMR_INST_BATCH(smr_item_fv, smr::InitBallOp<D>)

SMR_FV_DIM(1)
SMR_FV_DIM(2)
SMR_FV_DIM(3)
SMR_INST_BATCH(smr_item_fv, smr::LinCombOp)
