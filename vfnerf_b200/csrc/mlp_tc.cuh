// Tensor-core (tcgen05 / TMEM) path of the two MLPs: interface used by api.cu.
#pragma once
#include "common.cuh"

namespace vfn {

constexpr int kTcMaxSteps = 16;
constexpr int kTcMaxSegs = 3;
constexpr int kTcMaxChunks = 20;   // pipeline chunks per step (split-precision steps: hi + lo chunk per K range + bias)
constexpr int kTcTableSteps = 13;  // steps of the longest program (render(): 8 + 1 + 4; its dgrad chain: 5 + 8) -- rows of the
                                   // chunk table in the kernel parameters (TcParams::ctab)

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
// Split-precision (VFNERF_PREC_BF16X3) tile: every VF activation exists twice, as its bf16 rounding ("hi", main columns)
// and as the bf16 rounding of the remainder ("lo", columns kX3ColLo + c); a VF product is three MMAs
//   acc += A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T          (the A_lo W_lo term is below fp32 accumulation noise)
// so the chain carries ~16 significant bits instead of 8.  No skip region: the skip layer reads the (hi | lo) layer-0
// embedding with its weights scaled by 1/sqrt(2).  The colour net stays plain bf16 (its error is 2e-4).
//   [0,256) main hi | [256,272) ones | [272,368) emb0 hi|lo | [368,624) main lo;  aux = [368,416), ALIASING the first 48 lo
//   columns: the lo copies are dead once the last hidden VF layer's MMAs have completed (inside render() the feature step
//   reads only the hi copy, and that layer's epilogue writes no lo copy), which is exactly when the prologue warps write the
//   colour net's small inputs (normal, point, view embedding); the next tile's first epilogue rewrites the lo columns long
//   after the colour net's first layer has read them.  The 12 KiB this saves are the fourth slot of the weight ring.
constexpr int kX3ColOnes = 256, kX3ColEmb0 = 272, kX3ColLo = 368, kX3ColAux = kX3ColLo, kX3ActCols = 624;
// fp16 + fp8 remainders (VFNERF_PREC_FP16F8): the same tile, but a VF product is
//   acc += A16 W16^T  +  e5m2(A - A16) e4m3(W)^T  +  e5m2(2^-12 A) e4m3(2^12 (W - W16))^T
// with A16 / W16 the fp16 roundings: one 16-bit MMA (K = 16) plus two 8-bit MMAs (kind::f8f6f4, K = 32 at the same issue
// cost: profiles/probe_f8.py) per 16 K columns instead of three 16-bit ones -- 2 tensor-core units per product, not 3.
// The remainders are 2^-12 of the product, so the 3-4 mantissa bits of the 8-bit formats leave ~2^-15 (measured against
// the reference goldens: normals 8e-4 vs 2.7e-4 for bf16x3 and 1.4e-1 for bf16).  The activation remainders (2^-12 of
// O(1) activations) fit e5m2's range unscaled and the folded weights (O(0.1)) e4m3's; the weight remainders (2^-12 of the
// weights) need the power-of-two scale, undone on the activation copy.  The lo columns [368,624) hold the two 8-bit copies
// of the main columns: e5m2 remainders in slabs 0..15 (16 columns per 16-byte unit), e5m2 scaled values in 16..31.
// The ones-columns hold 2.0 -- the same bit pattern in bf16 and fp16 -- and the packed bias rows are halved.
constexpr float kF8ScaleHi = 4096.f;

// One GEMM step of the fused chain: acc[128 x N] = sum over segments A[:, col0 : col0+k] * Wimg^T, then an epilogue.
struct TcStep {
  int N;          // UMMA N (multiple of 16, <= 256)
  int n_valid;    // real output channels (<= N)
  int n_seg;
  int seg_col0[kTcMaxSegs];   // first activation-tile column of the segment
  int seg_k[kTcMaxSegs];      // columns (multiple of 16)
  int seg_lo[kTcMaxSegs];     // split-precision image: activation-tile column distance from the hi to the lo copy of
                              // this segment's A operand (0: plain segment).  The weight image then holds, per K chunk,
                              // the bf16 weights W_hi followed by the remainders W_lo
  float seg_wscale[kTcMaxSegs];   // extra factor on this segment's weights (skip-layer embedding columns: 1/sqrt(2))
  int seg_bar[kTcMaxSegs];    // readiness barriers of this segment's columns given explicitly (0: derived from the columns) --
                              // the split-precision tile's aux columns alias lo columns
  int seg_f8[kTcMaxSegs];     // fp16 + fp8-remainder segment (VFNERF_PREC_FP16F8, below): the "lo" half of every chunk of the
                              // weight image holds two 8-bit images instead of one 16-bit image
  int a_f16;      // this step's 16-bit operands (activations, weights, bias pair) are fp16, not bf16
  int use_lo;     // this program issues the lo products of the split segments (0: only A_hi W_hi^T, e.g. the feature step
                  // inside render(), whose output is rounded to bf16 for the colour net anyway)
  int out_lo;     // epilogue also writes the lo copy of its output (the next step is a split-precision step)
  int dot;        // 1 / 2: the 3-wide output layer that reads this step's activation -- the VF vector (1: tanh, to out_v
                  // and, in render(), to the aux columns) or the colour (2: sigmoid, to colors) -- is evaluated by the
                  // prologue warps as fp32 CUDA-core dot products of this step's fp32 accumulator row with the fp32
                  // weight rows: an M = 256 pair MMA costs 128 cycles whatever its N, so a 3-channel tensor-core step is
                  // as expensive as a 256-channel one (round 1: 13 % of the tensor time for 6 of 3 334 channels)
  int dot_guard_next;   // this step reuses the accumulator buffer of the PREVIOUS tile's last dot step: from the second
                        // tile on, the issuer waits for the prologue warps' "row consumed" barrier first (same-tile
                        // reuse is a pre_wait_mask bit)
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
  // training support
  int no_bias;       // 1: no ones-column bias segment (backward steps)
  int stash_out;     // activation-stash tensor that receives this step's epilogue output (-1: none)
  int mask_src;      // backward: stash tensor whose sign pattern (ReLU) or value (tanh) gates this step's output
};

enum TcEpi { TC_EPI_RELU = 0, TC_EPI_FEAT = 3,
             TC_EPI_BWD_RELU = 5,    // dX * (Y > 0)            (Y = stashed forward activation)
             TC_EPI_BWD_TANH = 6 };  // dX * (1 - Y^2)          (Y = stashed tanh features)
enum TcMode { TC_MODE_V_ONLY = 0, TC_MODE_VF_FULL = 1, TC_MODE_RENDER = 2, TC_MODE_RENDER_STASH = 3, TC_MODE_BWD = 4,
              TC_MODE_V_ONLY_STASH = 5, TC_MODE_VF_FULL_STASH = 6,   // VF-only module call kept for a backward
              TC_MODE_VF_BWD = 7 };                                   // dgrad chain of the VF net alone

// Activation stash (training): every tensor is bf16 in "tile-major K-slab" order, i.e. the exact shared-memory
// image of a 128-point tile, tile after tile:  [tile][slab = 8 channels][row = point in tile][8 channels].
// The same bytes serve as a K-major operand (dgrad: points x channels) and, with the MN-major descriptor bits,
// as the transposed operand of the weight-gradient GEMM (channels x points) -- no transpose pass exists.
constexpr int kTcMaxStash = 32;
struct TcStash {
  int n_y;                          // number of forward activation tensors (VF hidden layers, features, colour hidden layers)
  int idx_emb0, idx_skip, idx_aux;  // prologue-written side inputs
  int idx_d0;                       // first gradient tensor: D_i = idx_d0 + i mirrors Y_i
  int idx_dcolu, idx_dvu;           // 16-channel (hi, lo) units of d(colour pre-sigmoid) / d(vector pre-tanh), written by the
                                    // dgrad kernel's prologue: the A operands of the two 3-row weight-gradient GEMMs
  int n_tensors;
  int slabs[kTcMaxStash];           // 8-channel slabs per tile
  long long off[kTcMaxStash];       // byte offset of the tensor inside the stash buffer
  long long gate_off[kTcMaxStash];  // ReLU gates of hidden-layer tensor i as bits: [tile][64-column group][row] x 8 bytes
  long long bytes;
};

struct TcProgram {
  int n_steps;
  int render;       // colour steps follow (V step writes the aux columns, FEAT step writes the main columns)
  int emb_w;        // 3 + 6*multires
  int emb_pad;      // emb_w rounded up to 16
  int multires, multires_view;
  int small_w;      // 3 + (3 + 6*multires_view) + 3
  int skip_step;    // index of the step that consumes the skip columns (-1: none)
  int aux_step;     // index of the step that consumes the aux columns (-1: none)
  int bwd;          // 1: backward (dgrad) program of render() -- different prologue, no bias segments; 2: of the VF net alone
  int x3;           // split-precision tile layout (kX3Col*), 16 KiB ring slots
  int f8;           // fp16 + fp8-remainder variant of the split-precision tile (implies x3)
  int emb0_last_step;   // the emb0 columns may be rewritten for the next tile once this step's MMAs have completed
  int dot_step[2];      // steps with TcStep::dot, in order (-1: none)
  int col_aux, col_skip, col_ones, col_emb0, col_lo;   // activation-tile layout of this program
  TcStep s[kTcMaxSteps];
};

// fp32 rows of the two 3-wide output layers (TcStep::dot), appended to the weight images:
// [vector rows 3 x 256 | colour rows 3 x 256 | vector bias 3, colour bias 3, 0, 0]
constexpr int kTcDotFloats = 2 * 3 * 256 + 8;
struct TcPlan {
  uint8_t* wpack = nullptr;   // weight images of every step of the RENDER program (VF steps are shared by all modes)
  int64_t wpack_bytes = 0;
  int64_t dot_off = 0;        // byte offset of the kTcDotFloats block inside wpack
  TcProgram render{}, vf_full{}, v_only{};
  // training (keep_for_backward): stash layout, the dgrad program with its transposed weight images, scratch
  TcProgram bwd{}, bwd_vf{};
  TcStash stash{};
  uint8_t* stash_buf = nullptr;
  uint8_t* wpack_bwd = nullptr;
  int64_t wpack_bwd_bytes = 0, wpack_bwd_vf_bytes = 0;
  float* gbuf = nullptr;       // fp32 weight-gradient scratch: one [256 x 320] block per layer + column sums
  int64_t gbuf_floats = 0;
  float* d3 = nullptr;         // [P,3] x 2: d(colour pre-sigmoid), d(vector pre-tanh)
};

// carve the tensor-core buffers out of the workspace (base may be NULL when only sizing)
int tc_carve(char* base, int64_t& off, int multires, int multires_view, int skip_layer,
             const vfnerf_mlp_desc& vf, const vfnerf_mlp_desc* rn, TcPlan& plan, int64_t n_points = 0, int keep = 0,
             int x3 = 0);
// fold BatchNorm + convert/tile the weights of both nets into their shared-memory images
int tc_prepare(const vfnerf_mlp_desc& vf, const float* vf_arena, const vfnerf_mlp_desc* rn,
               const float* rn_arena, float bn_eps, const TcPlan& plan, cudaStream_t s);
// One fused launch.  points [n,3] (or generated from `grid` when non-null); ray_dirs [n/samples_per_ray, 3]
// (RENDER mode only).  out_v [n, v_ld] receives the 3 vector outputs; out_feat [n, feat_ld] the features
// (VF_FULL); colors [n,3] (RENDER).
int tc_forward(const TcPlan& plan, int mode, const float* points, const GridSpec* grid, int grid_res,
               int64_t grid_i0, int64_t n, const float* ray_dirs, int samples_per_ray, float* out_v,
               int64_t v_ld, float* out_feat, int64_t feat_ld, float* colors, cudaStream_t s,
               int64_t stash_tile0 = 0);   // *_STASH modes: first 128-point tile of the stash this launch writes

// Training backward of the RENDER program on the tensor cores.  d_colors [n,3] = dL/d colours, d_v [n,3] = dL/d VF
// vectors (both fp32, already including every upstream term); colors / normals are the forward outputs.  Writes
// (not accumulates) the gradients of every Linear / BatchNorm parameter into the two gradient arenas.
int tc_backward(const TcPlan& plan, const vfnerf_mlp_desc& vf, const float* vf_arena, const vfnerf_mlp_desc& rn,
                const float* rn_arena, float bn_eps, int64_t n, const float* colors, const float* normals,
                const float* d_colors, const float* d_v, float* vf_grad, float* rn_grad, cudaStream_t s);

// Training backward of the VF-only module call (VF_FULL / V_ONLY program kept with *_STASH): out [n, out_ld] are the
// forward outputs (vector | features), d_out [n, d_ld] their gradients (first n_out_cols columns, 3 or all).
int tc_backward_vf(const TcPlan& plan, const vfnerf_mlp_desc& vf, const float* vf_arena, float bn_eps, int64_t n,
                   const float* out, int64_t out_ld, const float* d_out, int64_t d_ld, int n_out_cols, float* vf_grad,
                   int accumulate, cudaStream_t s);

// test support (host only, no launch): the host-built chunk records of one program (0 render, 1 vf_full, 2 v_only, 3 dgrad
// of render(), 4 dgrad of the VF net): records [n_steps][kTcMaxChunks][4], n_chunks [n_steps], step_facts [n_steps][6]
int tc_debug_chunk_table(const TcPlan& plan, int program, uint32_t* records, int* n_chunks, int* step_facts, int* n_steps);
// test support: convert one stash tensor (or, tensor == 1000, the two [n,3] output-layer gradients) to row-major fp32
int tc_debug_stash_read(const TcPlan& plan, int tensor, int64_t n, float* out, int* n_cols, cudaStream_t s);

}  // namespace vfn
