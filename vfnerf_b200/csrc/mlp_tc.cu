// Tensor-core (tcgen05 / TMEM) MLP path -- placeholder until the fused kernels land.
#include "mlp_tc.cuh"

namespace vfn {

#define TC_UNBUILT() do { set_error("tensor-core precision modes are not built yet; use VFNERF_PREC_FP32"); return 3; } while (0)

int tc_carve(char*, int64_t&, const vfnerf_render_cfg&, const vfnerf_mlp_desc&, const vfnerf_mlp_desc&, int64_t, int, TcPlan&) { TC_UNBUILT(); }
int tc_prepare(const vfnerf_render_cfg&, const vfnerf_mlp_desc&, const float*, const vfnerf_mlp_desc&, const float*, TcPlan&, cudaStream_t) { TC_UNBUILT(); }
int tc_vf_forward(const vfnerf_render_cfg&, const TcPlan&, const float*, int64_t, float*, int64_t, int, const GridSpec*, int, cudaStream_t) { TC_UNBUILT(); }
int tc_rn_forward(const vfnerf_render_cfg&, const TcPlan&, const float*, int64_t, int64_t, float*, int, cudaStream_t) { TC_UNBUILT(); }
int64_t tc_vf_workspace_bytes(const vfnerf_mlp_desc&, int64_t, int, int, int) { set_error("tensor-core precision modes are not built yet"); return -1; }
int tc_vf_query(const vfnerf_mlp_desc&, const float*, int, int, float, int, const float*, int64_t, float*, int64_t, int, void*, int64_t, cudaStream_t) { TC_UNBUILT(); }
int tc_vf_grid_query(const vfnerf_mlp_desc&, const float*, int, int, float, int, int, int64_t, int64_t, const GridSpec&, float*, void*, int64_t, cudaStream_t) { TC_UNBUILT(); }

}  // namespace vfn
