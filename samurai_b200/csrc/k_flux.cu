// Kernel translation unit: flux-based schemes across level jumps (gather form).
#include "launch_impl.cuh"

SMR_INST_BATCH(smr_item_flux, smr::FluxGenOp<1, 0>)
SMR_INST_BATCH(smr_item_flux, smr::FluxGenOp<2, 0>)
SMR_INST_BATCH(smr_item_flux, smr::FluxGenOp<3, 0>)
SMR_INST_BATCH(smr_item_flux, smr::FluxGenOp<1, 1>)
SMR_INST_BATCH(smr_item_flux, smr::FluxGenOp<2, 1>)
SMR_INST_BATCH(smr_item_flux, smr::FluxGenOp<3, 1>)
SMR_INST_BATCH(smr_item_fluxw, smr::FluxWenoOp<1, 0, 1>)
SMR_INST_BATCH(smr_item_fluxw, smr::FluxWenoOp<2, 0, 1>)
SMR_INST_BATCH(smr_item_fluxw, smr::FluxWenoOp<3, 0, 1>)
SMR_INST_BATCH(smr_item_fluxw, smr::FluxWenoOp<1, 1, 1>)
SMR_INST_BATCH(smr_item_fluxw, smr::FluxWenoOp<2, 1, 1>)
SMR_INST_BATCH(smr_item_fluxw, smr::FluxWenoOp<3, 1, 1>)
SMR_INST_BATCH(smr_item_fluxw, smr::FluxWenoOp<2, 1, 2>)
SMR_INST_BATCH(smr_item_fluxw, smr::FluxWenoOp<3, 1, 3>)
SMR_INST_BATCH(smr_item_flux, smr::FluxVecOp<2>)
SMR_INST_BATCH(smr_item_flux, smr::FluxVecOp<3>)
