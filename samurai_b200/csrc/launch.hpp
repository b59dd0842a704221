// Host-callable launchers of the kernels in kernels.cuh.  The kernels are compiled in several translation units
// (k_fv.cu, k_flux.cu, k_mr.cu and one k_wf.cu object per <dim, prediction radius>) so the library builds in parallel;
// capi.cu only sees these declarations and never instantiates a kernel itself.
#pragma once
#include "derive.cuh"
#include "kernels.cuh"

namespace smr
{
    // batch_kernel<Item, Op><<<grid, SMR_CTA_THREADS, 0, st>>>(v, op); explicit instantiations live in the k_*.cu files
    template <class Item, class Op>
    cudaError_t launch_batch(int grid, cudaStream_t st, const BatchView<Item>& v, const Op& op);

    // record_kernel<Item, Op><<<ceil(n_items / 8), 256>>>: one warp per record
    template <class Item, class Op>
    cudaError_t launch_records(cudaStream_t st, const Item* items, int n_items, const Op& op);

    cudaError_t launch_ghost_phase_kernel(int dim, int grid, cudaStream_t st, const BcView& bc, int bc_ctas, const BatchView<smr_item_proj>& pv, double* f);

    // records from seeds + CSR (derive.cuh, k_derive.cu)
    cudaError_t launch_derive(int grid, cudaStream_t st, const DeriveArgs& a);

    // fused wavefront, one instantiation per (dim, radius)
    template <int DIM, int RADIUS>
    cudaError_t wf_launch_inst(WfArgs& a, int grid, size_t smem, cudaStream_t st);
    template <int DIM, int RADIUS>
    int wf_occupancy_inst(size_t smem);

    // every kernel translation unit owns a copy of the __constant__ peer table: they register a setter here
    using PeerSetter = cudaError_t (*)(const PeerTable&);
    void register_peer_setter(PeerSetter f);
    cudaError_t set_peer_table(const PeerTable& t);
} // namespace smr
