// Kernel translation unit: per-sweep launches of the multiresolution operators (projection, prediction, detail,
// criteria, keep propagation, copy) and the per-level ghost phase.
#include "launch_impl.cuh"

#define SMR_MR_DIM(D)                                                     \
    SMR_INST_BATCH(smr_item_proj, smr::ProjOp<D>)                         \
    SMR_INST_BATCH(smr_item_pred, smr::PredOp<D, 0>)                      \
    SMR_INST_BATCH(smr_item_pred, smr::PredOp<D, 1>)                      \
    SMR_INST_BATCH(smr_item_detail, smr::DetailOp<D, 0>)                  \
    SMR_INST_BATCH(smr_item_detail, smr::DetailOp<D, 1>)                  \
    SMR_INST_BATCH(smr_item_tag, smr::CriteriaOp<D>)                      \
    SMR_INST_BATCH(smr_item_tag, smr::MaximumOp<D, true>)

SMR_MR_DIM(1)
SMR_MR_DIM(2)
SMR_MR_DIM(3)
SMR_INST_BATCH(smr_item_fv, smr::AbsMaxOp)
SMR_INST_BATCH(smr_item_fv, smr::KeepLeavesOp)
SMR_INST_BATCH(smr_item_fv, smr::TagsChangeOp)
SMR_INST_RECORDS(smr_item_fv, smr::KeepLeavesOp)
SMR_INST_RECORDS(smr_item_fv, smr::TagsChangeOp)
SMR_INST_BATCH(smr_item_copy, smr::CopyOp)
SMR_INST_BATCH(smr_item_copy, smr::TagOrOp)

namespace smr
{
    cudaError_t launch_ghost_phase_kernel(int dim, int grid, cudaStream_t st, const BcView& bc, int bc_ctas, const BatchView<smr_item_proj>& pv, double* f)
    {
        switch (dim)
        {
            case 1:
                ghost_phase_kernel<1><<<grid, SMR_CTA_THREADS, 0, st>>>(bc, bc_ctas, pv, f);
                break;
            case 2:
                ghost_phase_kernel<2><<<grid, SMR_CTA_THREADS, 0, st>>>(bc, bc_ctas, pv, f);
                break;
            default:
                ghost_phase_kernel<3><<<grid, SMR_CTA_THREADS, 0, st>>>(bc, bc_ctas, pv, f);
                break;
        }
        return cudaGetLastError();
    }
} // namespace smr
