// Tensor-core (tcgen05 / TMEM) path of the two MLPs: interface used by api.cu.
#pragma once
#include "common.cuh"

namespace vfn {

struct TcPlan {
  void* vf_pack = nullptr;     // packed bf16 weight images + folded affine of the VF net
  void* rn_pack = nullptr;     // same for the colour net
  void* stash = nullptr;       // activation stash for the backward
  int64_t vf_pack_bytes = 0, rn_pack_bytes = 0, stash_bytes = 0;
};

// carve the tensor-core buffers out of the workspace (base may be NULL when only sizing)
int tc_carve(char* base, int64_t& off, const vfnerf_render_cfg& cfg, const vfnerf_mlp_desc& vf,
             const vfnerf_mlp_desc& rn, int64_t n_points, int keep, TcPlan& plan);
// fold BatchNorm + convert/tile the weights of both nets into their shared-memory images
int tc_prepare(const vfnerf_render_cfg& cfg, const vfnerf_mlp_desc& vf, const float* vf_arena,
               const vfnerf_mlp_desc& rn, const float* rn_arena, TcPlan& plan, cudaStream_t s);
// VF MLP on n points: positional encoding fused in the prologue; writes the first n_out_cols columns
// of [v, feat] to out (row stride out_ld)
int tc_vf_forward(const vfnerf_render_cfg& cfg, const TcPlan& plan, const float* points, int64_t n,
                  float* out, int64_t out_ld, int n_out_cols, const GridSpec* grid, int keep,
                  cudaStream_t s);
int tc_rn_forward(const vfnerf_render_cfg& cfg, const TcPlan& plan, const float* cin, int64_t cin_ld,
                  int64_t n, float* colors, int keep, cudaStream_t s);
int64_t tc_vf_workspace_bytes(const vfnerf_mlp_desc& vf, int64_t n_points, int multires, int keep,
                              int precision);
int tc_vf_query(const vfnerf_mlp_desc& vf, const float* vf_arena, int multires, int skip_layer,
                float bn_eps, int precision, const float* points, int64_t n, float* out, int64_t out_ld,
                int n_out_cols, void* workspace, int64_t workspace_bytes, cudaStream_t s);
int tc_vf_grid_query(const vfnerf_mlp_desc& vf, const float* vf_arena, int multires, int skip_layer,
                     float bn_eps, int precision, int res, int64_t i0, int64_t n, const GridSpec& gs,
                     float* out, void* workspace, int64_t workspace_bytes, cudaStream_t s);

}  // namespace vfn
