// Tensor-core (tcgen05 / TMEM) path of the two MLPs: interface used by api.cu.
#pragma once
#include "common.cuh"

namespace vfn {

constexpr int kTcMaxSteps = 16;
constexpr int kTcMaxSegs = 3;

// Activation-tile column map (bf16 columns of the 128-row A operand, K-slab layout, tc_common.cuh):
//   [0,256)    main: current layer input / output (step 0 reads the hi/lo embedding from [0, 2*emb_pad))
//   [256,304)  aux:  colour-net small inputs (normal written by the V step, point + view embedding by the prologue warps)
//   [304,352)  skip: positional encoding / sqrt(2) for the skip layer               (written by the prologue)
//   [352,368)  ones: [1, 1, 0, ...] constant -- multiplies the (hi, lo) bias row of every weight image,
//                    so the folded BatchNorm shift is added by the tensor core, not by the epilogue
//   [368,464)  emb0: bf16 hi | lo split of the positional encoding, the A operand of layer 0 (written by the
//                    prologue warps, so it never collides with the main columns)
// aux column order: [n(3), 0 x5 | p(3), embed(view dir), 0...]: the V step only rewrites the first 16-byte unit.
constexpr int kColAux = 256, kColSkip = 304, kColOnes = 352, kColEmb0 = 368, kActCols = 464;

// One GEMM step of the fused chain: acc[128 x N] = sum over segments A[:, col0 : col0+k] * Wimg^T, then an epilogue.
struct TcStep {
  int N;          // UMMA N (multiple of 16, <= 256)
  int n_valid;    // real output channels (<= N)
  int n_seg;
  int seg_col0[kTcMaxSegs];   // first activation-tile column of the segment
  int seg_k[kTcMaxSegs];      // columns (multiple of 16)
  int K;          // sum of seg_k == columns of the weight image
  int chunk_k;    // K columns per pipeline chunk (multiple of 16, N*chunk_k*2 <= 32 KiB); chunks never straddle segments
  int epi;        // TcEpi
  int fresh_mask; // readiness barriers (see mlp_tc.cu) whose phase this step's MMAs must wait for
  int pre_wait_mask;  // barriers waited on before the first MMA (accumulator hand-off only)
  long long w_off;    // byte offset of this step's weight image in the pack buffer
  // pack-time description of the source weights
  int net;        // 0 = VF net, 1 = colour net
  int layer;      // source Linear
  int row0;       // first source row
  int colmap;     // 0 identity | 1 hi/lo duplicated embedding | 2 colour-net input permutation | 3 skip layer
  int src_split;  // colmap 2: number of leading small columns; colmap 3: columns fed by the previous layer
  float post_scale;  // folded 1/sqrt(2) of the skip connection (applies to scale and shift)
};

enum TcEpi { TC_EPI_RELU = 0, TC_EPI_V = 2, TC_EPI_FEAT = 3, TC_EPI_RGB = 4 };
enum TcMode { TC_MODE_V_ONLY = 0, TC_MODE_VF_FULL = 1, TC_MODE_RENDER = 2 };

struct TcProgram {
  int n_steps;
  int render;       // colour steps follow (V step writes the aux columns, FEAT step writes the main columns)
  int emb_w;        // 3 + 6*multires
  int emb_pad;      // emb_w rounded up to 16
  int multires, multires_view;
  int small_w;      // 3 + (3 + 6*multires_view) + 3
  int skip_step;    // index of the step that consumes the skip columns (-1: none)
  int aux_step;     // index of the step that consumes the aux columns (-1: none)
  TcStep s[kTcMaxSteps];
};

struct TcPlan {
  uint8_t* wpack = nullptr;   // weight images of every step of the RENDER program (VF steps are shared by all modes)
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
