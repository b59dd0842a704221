// Kernel translation unit: the fused level wavefront, one object file per (WF_DIM, WF_RADIUS) (see __graft_entry__.build)
#include "launch_impl.cuh"

#if !defined(WF_DIM) || !defined(WF_RADIUS)
#error "compile with -DWF_DIM=1|2|3 -DWF_RADIUS=0|1"
#endif

template cudaError_t smr::wf_launch_inst<WF_DIM, WF_RADIUS>(smr::WfArgs&, int, size_t, cudaStream_t);
template int smr::wf_occupancy_inst<WF_DIM, WF_RADIUS>(size_t);
