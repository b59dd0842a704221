// Kernel translation unit: device-side derivation of the index records from seeds + the reference sub-mesh CSR.
#define SMR_DERIVE_KERNEL
#include "derive.cuh"

namespace smr
{
    cudaError_t launch_derive(int grid, cudaStream_t st, const DeriveArgs& a)
    {
        derive_kernel<<<grid, SMR_CTA_THREADS, 0, st>>>(a);
        return cudaGetLastError();
    }
} // namespace smr
