// Definitions behind launch.hpp, included by the kernel translation units only.
#pragma once
#include "launch.hpp"

namespace smr
{
    template <class Item, class Op>
    cudaError_t launch_batch(int grid, cudaStream_t st, const BatchView<Item>& v, const Op& op)
    {
        batch_kernel<Item, Op><<<grid, SMR_CTA_THREADS, 0, st>>>(v, op);
        return cudaGetLastError();
    }

    template <class Item, class Op>
    cudaError_t launch_records(cudaStream_t st, const Item* items, int n_items, const Op& op)
    {
        record_kernel<Item, Op><<<(n_items + SMR_CTA_THREADS / 32 - 1) / (SMR_CTA_THREADS / 32), SMR_CTA_THREADS, 0, st>>>(items, n_items, op);
        return cudaGetLastError();
    }

    template <int DIM, int RADIUS>
    cudaError_t wf_launch_inst(WfArgs& a, int grid, size_t smem, cudaStream_t st)
    {
        void* args[] = {&a};
        return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&wavefront_kernel<DIM, RADIUS>), dim3(static_cast<unsigned>(grid)),
                                           dim3(SMR_CTA_THREADS), args, smem, st);
    }

    template <int DIM, int RADIUS>
    int wf_occupancy_inst(size_t smem)
    {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wavefront_kernel<DIM, RADIUS>, SMR_CTA_THREADS, smem) != cudaSuccess)
        {
            return -1;
        }
        return per_sm;
    }

    namespace
    {
        struct PeerRegistrar
        {
            PeerRegistrar()
            {
                register_peer_setter([](const PeerTable& t) { return cudaMemcpyToSymbol(g_peers, &t, sizeof(t)); });
            }
        } s_peer_registrar;
    }
} // namespace smr

#define SMR_INST_RECORDS(Item, ...) template cudaError_t smr::launch_records<Item, __VA_ARGS__>(cudaStream_t, const Item*, int, const __VA_ARGS__&);
#define SMR_INST_BATCH(Item, ...) template cudaError_t smr::launch_batch<Item, __VA_ARGS__>(int, cudaStream_t, const smr::BatchView<Item>&, const __VA_ARGS__&);
