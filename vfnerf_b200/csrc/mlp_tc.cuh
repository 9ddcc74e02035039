// Tensor-core (tcgen05 / TMEM) path of the two MLPs: interface used by api.cu.
#pragma once
#include "common.cuh"

namespace vfn {

constexpr int kTcMaxSteps = 16;

// One GEMM step of the fused chain: acc[128 x N] = A[:, a_col0 : a_col0+K] * Wimg^T, then an epilogue.
struct TcStep {
  int K;          // multiple of 16
  int a_col0;     // first activation-tile column consumed
  int N;          // UMMA N (multiple of 16, <= 256)
  int n_valid;    // real output channels (<= N)
  int chunk_k;    // K columns per pipeline chunk (multiple of 16, N*chunk_k*2 <= 16 KiB)
  int n_chunks;
  int epi;        // TcEpi
  int aff_off;    // offset (floats) of this step's shift vector in the affine table
  long long w_off;  // byte offset of this step's weight image in the pack buffer
  // pack-time description of the source weights
  int net;        // 0 = VF net, 1 = colour net
  int layer;      // source Linear
  int row0;       // first source row
  int colmap;     // 0: identity (zero padded), 1: [W | W] duplicated for a hi/lo split input of width `dup_w`,
                  // 2: colour-net input permutation (features first, then the `small` leading columns)
  int dup_w;      // padded width of one copy (colmap 1) / number of leading small columns (colmap 2)
  float post_scale;  // folded 1/sqrt(2) of the skip connection (applies to scale and shift)
};

enum TcEpi { TC_EPI_RELU = 0, TC_EPI_RELU_SKIPFILL = 1, TC_EPI_V = 2, TC_EPI_FEAT = 3, TC_EPI_RGB = 4 };
enum TcMode { TC_MODE_V_ONLY = 0, TC_MODE_VF_FULL = 1, TC_MODE_RENDER = 2 };

struct TcProgram {
  int n_steps;
  int act_cols;     // activation tile width in bf16 columns (256, or 304 with the colour-net aux region)
  int n_stages;
  int emb_w;        // 3 + 6*multires
  int emb_pad;      // emb_w rounded up to 16
  int multires, multires_view;
  int small_w;      // 3 + (3 + 6*multires_view) + 3
  TcStep s[kTcMaxSteps];
};

struct TcPlan {
  uint8_t* wpack = nullptr;   // weight images of every step of the RENDER program (VF steps are shared by all modes)
  float* affine = nullptr;    // shift vectors, 256 floats per step
  int64_t wpack_bytes = 0;
  TcProgram render{}, vf_full{}, v_only{};
};

// carve the tensor-core buffers out of the workspace (base may be NULL when only sizing)
int tc_carve(char* base, int64_t& off, int multires, int multires_view, int skip_layer,
             const vfnerf_mlp_desc& vf, const vfnerf_mlp_desc* rn, TcPlan& plan);
// fold BatchNorm + convert/tile the weights of both nets into their shared-memory images
int tc_prepare(const vfnerf_mlp_desc& vf, const float* vf_arena, const vfnerf_mlp_desc* rn,
               const float* rn_arena, float bn_eps, const TcPlan& plan, cudaStream_t s);
// One fused launch.  points [n,3] (or generated from `grid` when non-null); ray_dirs [n/samples_per_ray, 3]
// (RENDER mode only).  out_v [n, v_ld] receives the 3 vector outputs; out_feat [n, feat_ld] the features
// (VF_FULL); colors [n,3] (RENDER).
int tc_forward(const TcPlan& plan, int mode, const float* points, const GridSpec* grid, int grid_res,
               int64_t grid_i0, int64_t n, const float* ray_dirs, int samples_per_ray, float* out_v,
               int64_t v_ld, float* out_feat, int64_t feat_ld, float* colors, cudaStream_t s);

}  // namespace vfn
