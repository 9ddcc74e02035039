// Fused tensor-core MLP chain (precisions bf16, bf16x3, fp16f8): the VF MLP (positional encoding -> 9 Linear+BN+ReLU
// layers with skip -> tanh) and, in RENDER mode, the colour MLP (5 layers -> sigmoid) evaluated for a
// tile of 128 points entirely on chip.  SURVEY.md §8 rows a3 + a8.  DESIGN.md §4 is the long description.
//
// One CTA per SM, 480 threads, CTAs paired into clusters of two (tcgen05 cta_group::2): each CTA owns one
// 128-point tile (its A operand and its TMEM accumulators) but only HALF of every weight chunk (N/2 output
// channels); one MMA instruction of the leader drives both SMs' tensor cores with M = 256.  Per SM this halves
// the shared-memory bytes moved per MMA (the limiter of the 1-CTA version) and the L2 traffic.
//   warp 0        weight producer: one lane streams this CTA's half of each pre-tiled weight chunk from L2 into the
//                 shared-memory ring (3 x 32 KiB, split-precision tile: 4 x 16 KiB) with cp.async.bulk + mbarriers (UBLKCP)
//   warp 1        leader CTA: MMA issuer -- the whole warp runs the loop on the uniform datapath (chunk records from the
//                 kernel parameters), one elected lane issues tcgen05.mma.cta_group::2 (M=256, N<=256; bf16 / fp16 K=16,
//                 8-bit K=32 -> fp32 in TMEM); tcgen05.commit (multicast) releases ring slots and publishes the
//                 accumulators in both CTAs (SASS UTCHMMA / UTCQMMA)
//   warps 2..9    epilogue: tcgen05.ld the fp32 accumulator (one row per thread, 32 columns at a time), ReLU / tanh,
//                 conversion (bf16; bf16 hi + lo; fp16 + two 8-bit copies), and the NEXT layer's A operand written
//                 straight into the activation tile in the K-slab UMMA layout -- activations never leave the SM.
//                 The folded BatchNorm scale lives in the weights and the shift is a (hi, lo) bias row multiplied
//                 by a constant ones-column of A, so the epilogue has no per-channel operand.
//   warps 10..13  prologue, one tile ahead: positional encodings, colour-net side inputs; the two 3-wide output layers
//                 (VF vector, colour) as fp32 CUDA-core dot products on the accumulator row
//   warp 14       training: bulk-copies finished column groups to the activation stash; inference: the peer CTA's relay
//                 lane ("my half of the chunk has landed" -> the leader's `full` barrier)
// Overlap inside the CTA: two 256-column TMEM accumulators (step s uses buffer s&1) and one "ready" mbarrier per
// 64-column group of the activation tile; the MMAs of layer l+1 start on K chunk g as soon as group g of layer l's
// output is in shared memory.  Weights: 1.1 MB (VF) + 0.6 MB (colour) per 16-bit copy, L2 resident, re-streamed per tile.
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include "mlp_tc.cuh"
#include "tc_common.cuh"

namespace vfn {
using namespace tc;

constexpr int kTileM = 128;
// weight-ring depth: the split-precision tile's 16 KiB slots are refilled in ~1.1 k cycles (MMA completion -> `empty` -> bulk
// copy from L2 -> relay) while a slot's four MMAs take 512: three slots bound the chunk rate at ~540 cycles, four do not
// activation-tile layout and ring-slot size of the two tile formats (mlp_tc.cuh)
template <bool kX3> struct Lay;
template <> struct Lay<false> {
  static constexpr int aux = kColAux, skip = kColSkip, ones = kColOnes, emb0 = kColEmb0, lo = 0, cols = kActCols;
  static constexpr int stage_bytes = 32768, stages = 3;
};
template <> struct Lay<true> {
  static constexpr int aux = kX3ColAux, skip = -1, ones = kX3ColOnes, emb0 = kX3ColEmb0, lo = kX3ColLo, cols = kX3ActCols;
  static constexpr int stage_bytes = 16384, stages = 4;
};
template <bool kX3> constexpr size_t tc_smem_bytes() {
  return (size_t)Lay<kX3>::cols * kTileM * 2 + (size_t)Lay<kX3>::stages * Lay<kX3>::stage_bytes + 512 + 128 +
         (size_t)2 * 3 * 256 * 4;
}
static_assert(tc_smem_bytes<true>() <= 227 * 1024 && tc_smem_bytes<false>() <= 227 * 1024, "activation tile + ring exceed shared memory");
constexpr int kTcThreads = 480;      // warp 0 producer, warp 1 MMA/relay, warps 2..9 epilogue, warps 10..13 prologue,
                                     // warp 14: activation-stash store lane (training forward)
constexpr int kAccCols = 256;
constexpr float kInvSqrt2 = 0.70710678118654752f;
// readiness barriers (leader CTA): 0..3 = 64-column groups of the main region (one per K chunk; group g is written by
// the epilogue warps of half g%2), 4 = the normal in the aux region (half-0 epilogue warps); 5 = skip region, 6 = emb0
// region, 7 = point/view part of the aux region (prologue warps).  Each expects 4 warps x 2 CTAs = 8 arrivals.
constexpr int kGroups = 9, kBarAux = 4, kBarSkip = 5, kBarEmb0 = 6, kBarAuxStatic = 7,
              kBarDot = 8;   // 8: the prologue warps have read the accumulator row of a TcStep::dot step (both CTAs)

// One pipeline chunk of a step (= one weight-ring slot), as the producer and the MMA issuer see it:
//   x = A start-address increment (16-byte units), y = K columns | readiness-barrier mask << 16,
//   z = bytes of this CTA's half of the weight chunk | (its offset inside the step's half image / 16) << 16,
//   w = bit 0: last chunk of the step; bit 1: fp8 remainder chunk; bits 8..: split-precision "hi" chunk -- distance
//       (16-byte units) from the hi to the lo copy of the A columns: the chunk's MMAs are issued a second time on the lo copy
struct TcChunk { uint32_t x, y, z, w; };

struct TcParams {
  TcProgram prog;
  // the chunk records of every step, built on the host (tc_chunk_records): the MMA-issuing warp reads them from the
  // parameter constant bank with a warp-uniform index, so every descriptor word it forms lives in uniform registers and
  // a tcgen05.mma costs it ~4 uniform-datapath instructions (from a shared-memory table + shuffles the compiler wrapped
  // every MMA in ELECT / 8 x R2UR.BROADCAST: profiles/r02_forward_kernel_experiments.md, section 4)
  TcChunk ctab[kTcTableSteps * kTcMaxChunks + 1];   // + 1: the issuer reads one record ahead
  int cnum[kTcTableSteps];                          // chunks per step (relay lane)
  const uint8_t* wpack;
  const float* points;
  int use_grid;
  GridSpec grid;
  int grid_res;
  long long grid_i0;
  long long n_points;
  const float* ray_dirs;
  int samples_per_ray;
  float* out_v; long long v_ld;
  float* out_feat; long long feat_ld;
  float* colors;
  // training: activation stash (forward writes Y tensors, backward reads them and writes D tensors)
  uint8_t* stash;
  TcStash sinfo;
  long long dot_off;       // fp32 rows of the 3-wide output layers inside wpack (forward programs)
  const float* dcol_pre;   // backward input [n,3]: dL/d(colour pre-sigmoid)
  const float* dv_pre;     // backward input [n,3]: dL/d(vector pre-tanh)
  int dbg;              // experiments: 1 = MMA issuer ignores A-readiness, 2 = epilogue skips TMEM loads / math / stores
  long long* dbg_buf;   // VFNERF_TC_DBG=64: cycle counters of CTA 0 (MMA thread [0..2], epilogue warp 2 [8..12], warp 3 [16..20])
};

// ---------------------------------------------------------------------------------------------
// weight packing: fp32 Linear weights (+ folded BatchNorm scale, + the bias row) -> bf16 K-slab images
// ---------------------------------------------------------------------------------------------
__global__ void tc_pack_kernel(TcProgram prog, vfnerf_mlp_desc vf, const float* __restrict__ vf_arena,
                               vfnerf_mlp_desc rn, const float* __restrict__ rn_arena, float eps,
                               uint8_t* __restrict__ wpack) {
  const TcStep st = prog.s[blockIdx.y];
  const vfnerf_mlp_desc& d = st.net == 0 ? vf : rn;
  const float* arena = st.net == 0 ? vf_arena : rn_arena;
  const int l = st.layer, in_dim = d.in_dim[l];
  const bool bn = d.gamma_off[l] >= 0;
  const int E = prog.emb_w, Epad = prog.emb_pad;
  const int total = st.N * st.K;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int n = e / st.K, k = e - n * st.K;
    // locate (segment, chunk, hi or lo part, column in segment) and the byte offset inside the image.
    // each CTA of the pair streams one half image: output channels [half*N/2, (half+1)*N/2).  Inside a half the
    // segments follow each other; a segment is a sequence of K chunks; a split-precision segment stores, per chunk,
    // the bf16 weights W_hi [N/2 x kc] and right after them the remainders W_lo [N/2 x kc]
    const int nh = st.N >> 1, half = n / nh, nn = n - half * nh;
    int sg = 0, q = k;
    int64_t base = st.w_off + (int64_t)half * nh * st.K * 2;
    for (;;) {
      const int wimg = st.seg_k[sg] * (st.seg_lo[sg] ? 2 : 1);
      if (q < wimg) break;
      q -= wimg; base += (int64_t)nh * wimg * 2; ++sg;
    }
    const bool split = st.seg_lo[sg] != 0;
    const int span = st.chunk_k * (split ? 2 : 1);
    const int ci = q / span, r = q - ci * span;
    const int kc = min(st.chunk_k, st.seg_k[sg] - ci * st.chunk_k);
    const bool is_lo = r >= kc;
    const int kk = r - (is_lo ? kc : 0);
    const int kin = ci * st.chunk_k + kk;
    const int64_t off = base + (int64_t)ci * span * nh * 2 + (is_lo ? (int64_t)kc * nh * 2 : 0) +
                        (int64_t)(kk / 8) * nh * 16 + nn * 16 + (kk & 7) * 2;
    float w = 0.f;
    if (n < st.n_valid) {
      const int row = st.colmap >= 10 ? 0 : st.row0 + n;
      float sc = 1.f;
      if (bn && st.colmap < 10) sc = arena[d.gamma_off[l] + row] / sqrtf(arena[d.var_off[l] + row] + eps);
      if (sg == st.n_seg - 1 && !st.no_bias) {
        // bias row: folded shift as a bf16 (hi, lo) pair in columns 0 and 1
        if (kin < 2) {
          const float b = arena[d.b_off[l] + row];
          float sh = b;
          if (bn) sh = arena[d.beta_off[l] + row] + (b - arena[d.mean_off[l] + row]) * sc;
          sh *= st.post_scale;
          if (prog.f8) sh *= 0.5f;                    // the ones-columns of this tile hold 2.0 (mlp_tc.cuh)
          const float hi = st.a_f16 ? __half2float(__float2half_rn(sh)) : __bfloat162float(__float2bfloat16(sh));
          w = kin == 0 ? hi : sh - hi;
        }
      } else {
        int src = -1;
        if (st.colmap >= 10) {
          // backward image: MMA "output channel" n = forward input column, MMA "K" index = forward output channel
          int n_out = -1;
          if (st.colmap == 11 || (st.colmap == 12 && sg == 1)) n_out = kin < 3 ? kin : (kin < 6 ? kin - 3 : -1);   // (hi, lo) unit
          else n_out = st.row0 + kin;
          const int k_in = st.src_split + n;
          w = 0.f;
          if (n_out >= 0 && n_out < d.out_dim[l] && k_in < in_dim) {
            float sc2 = 1.f;
            if (bn) sc2 = arena[d.gamma_off[l] + n_out] / sqrtf(arena[d.var_off[l] + n_out] + eps);
            w = arena[d.w_off[l] + (int64_t)n_out * in_dim + k_in] * sc2 * st.post_scale;
          }
          src = -2;
        }
        else if (st.colmap == 0) src = kin;
        else if (st.colmap == 1) { src = kin < Epad ? kin : kin - Epad; if (src >= E) src = -1; }
        else if (st.colmap == 2) {
          // colour-net layer 0: main segment = features; aux segment = [n(3), 0 x5, p(3), embed(view), 0...]
          const int ev = st.src_split - 6;                    // width of the view-direction embedding
          if (sg == 0) src = st.src_split + kin;
          else if (kin < 3) src = 3 + ev + kin;               // normals (reference columns 3+ev ..)
          else if (kin >= 8 && kin < 11) src = kin - 8;       // point
          else if (kin >= 11 && kin < 11 + ev) src = 3 + (kin - 11);
          else src = -1;
        }
        else src = sg == 0 ? (kin < st.src_split ? kin : -1) : (kin < E ? st.src_split + kin : -1);
        if (src >= 0 && src < in_dim) w = arena[d.w_off[l] + (int64_t)row * in_dim + src] * sc * st.post_scale;
      }
      w *= st.seg_wscale[sg];
    }
    const bool f16 = st.a_f16 != 0;
    if (split && st.seg_f8[sg]) {
      // fp16 + fp8-remainder segment: the hi half is fp16(W); the lo half holds, per 16-column unit, e4m3(W) in its
      // first kc * N/2 bytes (multiplies the e5m2 remainders of the activations) and e4m3(2^12 (W - fp16(W))) in the next
      // kc * N/2 (multiplies the e5m2 copies of the activations)
      const float hi = __half2float(__float2half_rn(w));
      if (!is_lo) {
        *reinterpret_cast<__half*>(wpack + off) = __float2half_rn(w);
      } else {
        const int64_t lo_base = base + (int64_t)ci * span * nh * 2 + (int64_t)kc * nh * 2;
        const int64_t o8 = lo_base + (int64_t)(kk / 16) * nh * 16 + nn * 16 + (kk & 15);
        wpack[o8] = (uint8_t)__nv_cvt_float_to_fp8(w, __NV_SATFINITE, __NV_E4M3);
        wpack[o8 + (int64_t)kc * nh] = (uint8_t)__nv_cvt_float_to_fp8((w - hi) * kF8ScaleHi, __NV_SATFINITE, __NV_E4M3);
      }
      continue;
    }
    if (split) {
      const float hi = f16 ? __half2float(__float2half_rn(w)) : __bfloat162float(__float2bfloat16(w));
      if (is_lo) w -= hi;
    }
    if (f16) *reinterpret_cast<__half*>(wpack + off) = __float2half_rn(w);
    else *reinterpret_cast<__nv_bfloat16*>(wpack + off) = __float2bfloat16(w);
  }
}

// fp32 rows + biases of the 3-wide output layers (vector rows of the VF net's last layer, the colour net's last layer),
// with a BatchNorm folded in should the layer have one: [v rows 3 x 256 | colour rows 3 x 256 | bias v 3, bias c 3, 0 0]
__global__ void tc_pack_dot_kernel(vfnerf_mlp_desc vf, const float* __restrict__ vf_arena, vfnerf_mlp_desc rn,
                                   const float* __restrict__ rn_arena, int have_rn, float eps, float* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kTcDotFloats; i += gridDim.x * blockDim.x) {
    const bool is_bias = i >= 2 * 3 * 256;
    const int net = is_bias ? (i - 2 * 3 * 256) / 3 : i / 768;
    const int row = is_bias ? (i - 2 * 3 * 256) % 3 : (i % 768) / 256;
    const int col = i % 256;
    float v = 0.f;
    if (net < 2 && (net == 0 || have_rn)) {
      const vfnerf_mlp_desc& d = net == 0 ? vf : rn;
      const float* arena = net == 0 ? vf_arena : rn_arena;
      const int l = d.n_layers - 1, in_dim = d.in_dim[l];
      float sc = 1.f, sh = arena[d.b_off[l] + row];
      if (d.gamma_off[l] >= 0) {
        sc = arena[d.gamma_off[l] + row] / sqrtf(arena[d.var_off[l] + row] + eps);
        sh = arena[d.beta_off[l] + row] + (sh - arena[d.mean_off[l] + row]) * sc;
      }
      v = is_bias ? sh : (col < in_dim ? arena[d.w_off[l] + (int64_t)row * in_dim + col] * sc : 0.f);
    }
    out[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// device helpers of the epilogue warps
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t tanh_bf16x2(uint32_t x) {
  uint32_t y;
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
// fp32 pair -> bf16x2 with the ReLU folded into the conversion (SASS F2FP.RELU.BF16.F32.PACK_AB): one instruction per
// two outputs instead of a conversion plus a max.bf16x2
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t x) {
  uint32_t y;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(y) : "r"(x), "r"(0u));
  return y;
}

// 8 consecutive columns of one row -> one 16-byte store into slab `slab` of the activation tile
__device__ __forceinline__ void store_slab_f(uint8_t* s_act, int slab, int row, const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(s_act + slab * (kTileM * 16) + row * 16) = u;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void store_slab_h(uint8_t* s_act, int slab, int row, const float* f) {     // fp16 variant
  uint4 u;
  u.x = pack_f16x2(f[0], f[1]); u.y = pack_f16x2(f[2], f[3]);
  u.z = pack_f16x2(f[4], f[5]); u.w = pack_f16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(s_act + slab * (kTileM * 16) + row * 16) = u;
}
__device__ __forceinline__ void store_slab_u(uint8_t* s_act, int slab, int row, uint32_t a, uint32_t b, uint32_t c,
                                             uint32_t d) {
  *reinterpret_cast<uint4*>(s_act + slab * (kTileM * 16) + row * 16) = make_uint4(a, b, c, d);
}

// positional encoding of embedder.py:11-37 into e[0 .. 3+6*L).  sin/cos of the base frequency come from
// sincosf; the octaves above it use the double-angle recurrence (error doubles per octave: <= 4e-6 at 2^5,
// far below the bf16 quantisation this path feeds).  Fully unrolled so e[] stays in registers.
// kExact (split-precision path): every octave is its own sincosf of the exactly scaled argument, like the reference's
// torch.sin(x * 2^k) -- the recurrence's 4e-6 would be a visible share of that path's 1e-3 budget.
constexpr int kMaxRes = 7;
template <bool kExact = false>
__device__ __forceinline__ void embed3(const float* p, int L, float* e) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    e[c] = p[c];
    float sv, cv;
    sincosf(p[c], &sv, &cv);
#pragma unroll
    for (int k = 0; k < kMaxRes; ++k) {
      if (k < L) {
        e[3 + 6 * k + c] = sv;
        e[6 + 6 * k + c] = cv;
        if (kExact) {
          sincosf(p[c] * (float)(2 << k), &sv, &cv);
        } else {
          const float s2 = 2.f * sv * cv, c2 = cv * cv - sv * sv;
          sv = s2; cv = c2;
        }
      }
    }
  }
}

__device__ __forceinline__ void load_point(const TcParams& p, long long pi, bool valid, float* pt) {
  pt[0] = pt[1] = pt[2] = 0.f;
  if (!valid) return;
  if (p.use_grid) {
    const long long g = p.grid_i0 + pi;
    const long long res = p.grid_res;
    const long long idx[3] = {(g / res / res) % res, (g / res) % res, g % res};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = __fadd_rn(__fmul_rn((float)idx[c], p.grid.voxel), p.grid.origin[c]);
      v = __fadd_rn(v, p.grid.translation[c]);
      pt[c] = __fadd_rn(v, p.grid.centroid[c]);
    }
  } else {
    pt[0] = p.points[3 * pi]; pt[1] = p.points[3 * pi + 1]; pt[2] = p.points[3 * pi + 2];
  }
}

// readiness barrier that guards activation-tile column `col` (-1: constant region, nothing to wait for)
// 16-byte unit (slab, row) of tile `tile` of stash tensor t
__device__ __forceinline__ uint4* stash_unit(const TcParams& p, int t, long long tile, int slab, int row) {
  return reinterpret_cast<uint4*>(p.stash + p.sinfo.off[t] + tile * ((long long)p.sinfo.slabs[t] * (kTileM * 16)) +
                                  (long long)slab * (kTileM * 16) + row * 16);
}

__device__ __forceinline__ uint2* gate_unit(const TcParams& p, int t, long long tile, int group, int row) {
  return reinterpret_cast<uint2*>(p.stash + p.sinfo.gate_off[t] + ((tile * 4 + group) * kTileM + row) * 8);
}

// readiness barriers (bit mask) that guard the 64-column chunk starting at activation-tile column `col`
__host__ __device__ inline uint32_t col_barriers(bool x3, int col) {
  if (col < kColAux) return 1u << (col >> 6);
  if (x3) {
    if (col < kX3ColEmb0) return 0u;              // ones (the aux columns alias lo columns: TcStep::seg_bar names their barriers)
    if (col < kX3ColLo) return 1u << kBarEmb0;
    return 1u << ((col - kX3ColLo) >> 6);     // the lo copy of a group is published by the same arrival as its hi copy
  }
  if (col < kColSkip) return (1u << kBarAux) | (1u << kBarAuxStatic);
  if (col < kColOnes) return 1u << kBarSkip;
  if (col < kColEmb0) return 0u;
  return 1u << kBarEmb0;
}

// The chunk records of step `si` (at most kTcMaxChunks): run by the host for TcParams::ctab (MMA issuer) and by one
// thread per step inside the kernel for the shared-memory copy the weight producer and the relay lane walk.
__host__ __device__ inline int tc_chunk_records(const TcProgram& prog, int si, TcChunk* out) {
  const TcStep& st = prog.s[si];
  const bool x3 = prog.x3 != 0, f8 = prog.f8 != 0;
  constexpr uint32_t kSlabUnits = (kTileM * 16u) >> 4;     // 16-byte units per K-slab of the activation tile
  uint32_t seen = 0, src16 = 0;
  int nc = 0;
  for (int sg = 0; sg < st.n_seg; ++sg) {
    const int lo = st.seg_lo[sg];
    for (int k0 = 0; k0 < st.seg_k[sg]; k0 += st.chunk_k) {
      const int kc = (st.chunk_k < st.seg_k[sg] - k0) ? st.chunk_k : st.seg_k[sg] - k0;
      const int col = st.seg_col0[sg] + k0;
      uint32_t need = 0;
      if (st.seg_bar[sg]) need = (uint32_t)st.seg_bar[sg];
      else
        for (int cc = col; cc < col + kc; cc += 64) need |= col_barriers(x3, cc);
      need &= (uint32_t)st.fresh_mask & ~seen;
      seen |= need;
      const uint32_t a_off = (uint32_t)(col >> 3) * kSlabUnits, bytes = (uint32_t)((st.N >> 1) * kc * 2);
      const bool f8seg = f8 && st.seg_f8[sg] != 0;
      const uint32_t dual = (lo && st.use_lo && !f8seg) ? ((uint32_t)(lo >> 3) * kSlabUnits) << 8 : 0u;
      out[nc++] = TcChunk{a_off, (uint32_t)kc | (need << 16), bytes | (src16 << 16), dual};
      src16 += bytes >> 4;
      if (lo) {              // the image holds W_lo right after W_hi whether or not this program uses it
        if (st.use_lo) {
          // fp8 remainder chunk (w bit 1): x = the e5m2 remainders of these columns in the lo region (16 columns per unit)
          const uint32_t a8 = ((uint32_t)((st.seg_col0[sg] + lo) >> 3) + (uint32_t)(k0 >> 4)) * kSlabUnits;
          out[nc++] = f8seg ? TcChunk{a8, (uint32_t)kc, bytes | (src16 << 16), 2u}
                            : TcChunk{a_off, (uint32_t)kc, bytes | (src16 << 16), 0u};
        }
        src16 += bytes >> 4;
      }
    }
  }
  out[nc - 1].w |= 1u;
  return nc;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma2_bf16_split(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                 uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
// Warp-collective variants: executed by all 32 lanes in uniform control flow, one elected lane issues.
__device__ __forceinline__ void umma2_bf16_split_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                   uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      ".reg .b64 da, db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_f8_split_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                 uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      ".reg .b64 da, db;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit_u32_w(uint32_t bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t"
      "}" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// completion of all previously issued MMAs -> the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// 32-bit shared-address forms for the single-lane issue loop (no generic->shared conversion per call)
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma2_commit_u32(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// In-kernel cycle counters / per-step timeline (VFNERF_TC_DBG=64) exist only in builds with -DVFNERF_TC_PROFILE: the
// single MMA-issuing lane has no slack -- even predicated-off clock reads and one more run-time switch in its loop cost
// several per cent (profiles/r01_forward_kernel_experiments.md) -- so the product build compiles the hooks out.
#ifdef VFNERF_TC_PROFILE
constexpr bool kTcProfile = true;
#define TCK(acc_) do { if (prof) { long long t1_ = clock64(); acc_ += t1_ - t0; t0 = t1_; } } while (0)
#else
constexpr bool kTcProfile = false;
#define TCK(acc_) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------------
// activation-tile stores of the split-precision epilogue (profile builds: VFNERF_TC_DBG bit 1 drops them -- garbage results --
// to see what the epilogue's shared-memory writes cost the MMAs that run beside it)
#define EPI_STORE(...) do { if (!(kTcProfile && (kdbg & 2))) store_slab_u(__VA_ARGS__); } while (0)
template <bool kBwd, bool kStash, bool kX3, bool kF8 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
mlp_tc_kernel(const __grid_constant__ TcParams p) {
  static_assert(!kX3 || (!kBwd && !kStash), "the split-precision tile is built for the forward-only programs");
  static_assert(!kF8 || kX3, "the fp16 + fp8-remainder chain uses the split-precision tile");
  using L = Lay<kX3>;
  // epilogue order of the split-precision tile (see the activation-store loop): all eight warps on one 64-column group
  // at a time.  -DVFN_X3_SERIAL_GROUPS=0 builds the plain tile's order (two groups per warp half) for A/B timing.
#ifndef VFN_X3_SERIAL_GROUPS
#define VFN_X3_SERIAL_GROUPS 1
#endif
  constexpr bool kSerialGroups = kX3 && VFN_X3_SERIAL_GROUPS;
  static_assert(!kF8 || kSerialGroups, "the fp8 hand-off writes 32 columns per warp and iteration");
  // weight ring.  The stream from L2 is latency-bound (a slot is busy for its fill latency + its wait + its MMAs, whatever
  // its size; profiles/r02_forward_kernel_experiments.md), so the fp16 + fp8 chain, whose MMAs are short, runs the same 48 KiB
  // as six 8 KiB slots (32 K columns per chunk) instead of three 16 KiB ones
  constexpr int kStageBytes = L::stage_bytes;
  constexpr int kStages = L::stages;
  const int kdbg = kTcProfile ? p.dbg : 0;   // experiment switches (VFNERF_TC_DBG) exist in profile builds only
  extern __shared__ __align__(1024) uint8_t smem[];
  // launch-phase stamps of CTA 0 (profile builds, VFNERF_TC_DBG=64): entry, set-up done, first MMA, last step committed, exit
  const bool stamp = kTcProfile && p.dbg_buf && blockIdx.x == 0 && threadIdx.x == 32;
  if (stamp) p.dbg_buf[240] = clock64();
  const TcProgram& prog = p.prog;
  uint8_t* s_act = smem;
  uint8_t* s_stage = smem + L::cols * (kTileM * 2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + kStages * kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* fullp = bars + 2 * kStages;      // leader only: "the peer's half of the chunk has landed"
  uint64_t* acc_full = bars + 3 * kStages;
  // one "accumulator complete" barrier per TMEM accumulator buffer: buffer b is committed once per two steps, and the
  // MMAs of step g+2 depend (through the column-group barriers) on every epilogue warp having finished step g+1, hence
  // having passed its wait for step g -- the issuer can never complete a phase twice before a slow warp has seen it
  uint64_t* grp = acc_full + 2;                 // leader only: A-operand readiness
  uint64_t* reg_free = grp + kGroups;           // [0] emb0, [1] skip, [2] aux-static: their reader MMAs have completed
  // training forward: the bf16 activations an epilogue warp writes into the activation tile ARE the stash image of
  // that tile (same K-slab bytes), so they are not stored to global memory by the epilogue warps: one lane of warp 14
  // copies each finished 64-column group with cp.async.bulk (shared -> global) and tells the epilogue when the group's
  // columns may be overwritten.  st_ready[g]: the 4 warps owning group g have written + fenced it; st_done[g]: the bulk
  // copy has finished reading it.
  uint64_t* st_ready = reg_free + 3;
  uint64_t* st_done = st_ready + 4;
  // "the accumulator of a TcStep::dot step is complete", for the prologue warps.  They cannot share acc_full: a waiter
  // that does not observe EVERY phase of an mbarrier cannot tell the phase it wants from an older one of equal parity.
  uint64_t* dot_full = st_done + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dot_full + 1);
  // (the per-chunk facts of every step -- TcChunk records -- live in the kernel parameters, TcParams::ctab: the producer,
  //  the relay lane and the MMA issuer read them from the constant bank with a warp-uniform index; rounds 1 and 2 built a
  //  shared-memory copy here, 11 k cycles of set-up per launch)
  // fp32 rows (and biases) of the two 3-wide output layers the epilogue warps evaluate on the CUDA cores (TcStep::dot)
  float* s_dotb = reinterpret_cast<float*>(bars + 64);                        // [8]: vector bias 3, colour bias 3
  float* s_dotw = s_dotb + 32;                                                // [2][3][256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    // leader: a ring slot is full when its own bulk copy has landed AND the peer's relay has arrived
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], rank == 0 ? 2 : 1); mbar_init(&empty[i], 1); mbar_init(&fullp[i], 1); }
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    // 4 warps (one half of the epilogue warps, or the prologue warps) x 2 CTAs; the split-precision tile's column groups
    // are written by all 8 epilogue warps
    for (int i = 0; i < kGroups; ++i) mbar_init(&grp[i], (kSerialGroups && i < 4) ? 16 : 8);
    for (int i = 0; i < 3; ++i) mbar_init(&reg_free[i], 1);
    for (int i = 0; i < 4; ++i) { mbar_init(&st_ready[i], 4); mbar_init(&st_done[i], 1); }
    mbar_init(dot_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (!kBwd) {
    const float* dw = reinterpret_cast<const float*>(p.wpack + p.dot_off);
    for (int i = threadIdx.x; i < 2 * 3 * 256; i += kTcThreads) s_dotw[i] = __ldg(dw + i);
    if (threadIdx.x < 8) s_dotb[threadIdx.x] = __ldg(dw + 2 * 3 * 256 + threadIdx.x);
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kTileM) {
    // constant ones-columns [1, 1, 0, ...] that pick up the bias row of every weight image
    const int r = threadIdx.x - 64;
    store_slab_u(s_act, L::ones / 8, r, kF8 ? 0x40004000u : 0x3F803F80u, 0u, 0u, 0u);   // kF8: 2.0 in bf16 AND fp16
    store_slab_u(s_act, L::ones / 8 + 1, r, 0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  tc_fence_before_sync();
  cluster_sync_all();                       // barriers of BOTH CTAs are initialised before any remote arrive / multicast
  tc_fence_after_sync();
  if (stamp) p.dbg_buf[241] = clock64();
  const uint32_t tmem = *tmem_slot;
  const long long num_tiles = (p.n_points + kTileM - 1) / kTileM;
  // tiles are handed out in pairs: cluster c processes tile pairs c, c + n_clusters, ...; CTA `rank` takes tile 2*pair + rank
  const long long num_pairs = (num_tiles + 1) / 2;
  const long long pair0 = blockIdx.x >> 1, pair_step = gridDim.x >> 1;
  // called by all 32 lanes of an epilogue warp after each lane's fence.proxy.async: one arrival per warp
  auto arrive_grp = [&](int g) {
    __syncwarp();
    if (lane == 0) {
      if (rank == 0) mbar_arrive(&grp[g]); else mbar_arrive_remote(&grp[g], 0);
    }
  };
  // the same with the stash hand-shake of the training kernels folded in: ONE warp barrier, the arrival the MMA issuer is
  // waiting for first, then "this group may be copied out" for warp 14
  auto publish = [&](int g, bool to_issuer, uint64_t* st_ready_bar) {
    __syncwarp();
    if (lane == 0) {
      if (to_issuer) { if (rank == 0) mbar_arrive(&grp[g]); else mbar_arrive_remote(&grp[g], 0); }
      if (st_ready_bar) mbar_arrive(st_ready_bar);
    }
  };

  if (warp == 0) {
    // ===================== weight producer =====================
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (long long pair = pair0; pair < num_pairs; pair += pair_step) {
        for (int si = 0; si < prog.n_steps; ++si) {
          const TcStep& st = prog.s[si];
          const uint8_t* src = p.wpack + st.w_off + (int64_t)rank * (st.N >> 1) * st.K * 2;
          // records from the parameter constant bank (host-built, TcParams::ctab), the next one fetched before the wait
          int cidx = si * kTcMaxChunks;
          TcChunk c = p.ctab[cidx];
          for (;;) {
            const TcChunk c_next = p.ctab[cidx + 1];
            const uint32_t bytes = c.z & 0xFFFFu;
            mbar_wait(&empty[stage], phase ^ 1);
            if (kdbg & 4) {
              // experiment: no weight traffic at all (results are garbage) -- isolates the L2 -> shared-memory streaming
              mbar_arrive(&full[stage]);
            } else {
              mbar_arrive_expect_tx(&full[stage], bytes);
              bulk_g2s(s_stage + stage * kStageBytes, src + (size_t)(c.z >> 16) * 16, bytes, &full[stage]);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            if (c.w & 1u) break;
            c = c_next; ++cidx;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== leader CTA: MMA issuer =====================
    // The WHOLE warp runs this loop (waits, chunk-table reads, descriptor arithmetic) in uniform control flow and only the
    // tcgen05 instructions themselves are issued by lane 0: the compiler then keeps the descriptor words in uniform
    // registers.  Run by lane 0 alone, every tcgen05.mma was wrapped in an ELECT / 7 x R2UR.BROADCAST waterfall (17
    // instructions of scalar work per 128-cycle MMA).  Values that come from shared memory are made uniform with a shuffle.
    if (rank == 0) {
      const uint32_t tmem_w = __shfl_sync(0xffffffffu, tmem, 0);
      int stage = 0, phase = 0;
      uint32_t grp_par = 0, gstep = 0;
      long long t_grp = 0, t_full = 0, t_issue = 0, t0 = 0;
      const bool prof = kTcProfile && p.dbg_buf && blockIdx.x == 0;
      if (prof) t0 = clock64();
      const uint32_t act_base = smem_u32(s_act), stage_base = smem_u32(s_stage);
      const uint32_t grp_u32 = smem_u32(grp), full_u32 = smem_u32(full), empty_u32 = smem_u32(empty);
      int tile_no = 0;
      for (long long pair = pair0; pair < num_pairs; pair += pair_step, ++tile_no) {
        const bool tl = prof && tile_no == 2;
        for (int si = 0; si < prog.n_steps; ++si, ++gstep) {
          const TcStep& st = prog.s[si];
          bool first_mma = true;
          const uint32_t idesc = (kF8 && st.a_f16) ? make_idesc_f16(2 * kTileM, st.N) : make_idesc_bf16(2 * kTileM, st.N);
          // 8-bit remainder products: e5m2 activations (remainders, then scaled copies) x e4m3 weights
          const uint32_t idesc8a = make_idesc_f8(2 * kTileM, st.N, 1, 0), idesc8b = idesc8a;
          // descriptor halves: only the start-address field of the low words changes between MMAs
          const uint32_t desc_hi = (128u >> 4) | (1u << 14);                   // SBO = 128 B, version 1
          const uint32_t a_lo0 = ((act_base >> 4) & 0x3FFF) | (((kTileM * 16u) >> 4) << 16);
          const uint32_t b_lo0 = ((stage_base >> 4) & 0x3FFF) | ((((uint32_t)(st.N >> 1) * 16u) >> 4) << 16);
          const uint32_t b_kstep = (uint32_t)st.N;                             // two K-slabs of the HALF weight chunk, in 16-byte units
          const uint32_t acc = tmem_w + (gstep & 1) * kAccCols;
          const uint32_t fresh = (kdbg & 1) ? 0u : 0xFFFFu;   // per-chunk masks are pre-filtered in the chunk table
          uint32_t accumulate = 0;
          TCK(t_issue);
          const int pre_wait = st.pre_wait_mask | ((st.dot_guard_next && tile_no > 0) ? (1 << kBarDot) : 0);
          for (int g = 0; g < kGroups; ++g) {
            if (pre_wait & (1 << g)) {
              mbar_wait_cluster(&grp[g], (grp_par >> g) & 1u);
              grp_par ^= (1u << g);
            }
          }
          // chunk records come from the parameter constant bank with a warp-uniform index (TcParams::ctab): no
          // shared-memory read, no shuffle, and everything derived from them stays in uniform registers
          int cidx = si * kTcMaxChunks;
          int ci = 0;                              // chunk number inside the step (profile builds: per-chunk stamps)
          TcChunk ck = p.ctab[cidx];
          for (;;) {
            const TcChunk nxt = p.ctab[cidx + 1];          // next chunk's facts arrive while this chunk's MMAs issue
            // the A columns of this chunk must have been (re)written: wait for their readiness barrier(s) (at most two)
            uint32_t need = (ck.y >> 16) & fresh;
            // tcgen05.fence::after_thread_sync orders this thread's MMAs after what OTHER threads did to the activation tile
            // and the accumulator (their stores + tcgen05.ld) before they arrived on a group barrier: needed after a
            // group / hand-off wait, not after the weight ring's "slot full" (a bulk copy completing on an mbarrier)
            const bool synced = need != 0 || accumulate == 0;
            if (need) {
              int b = __ffs(need) - 1;
              mbar_wait_u32(grp_u32 + 8u * b, (grp_par >> b) & 1u);
              grp_par ^= (1u << b);
              need &= need - 1;
              if (need) {
                b = __ffs(need) - 1;
                mbar_wait_u32(grp_u32 + 8u * b, (grp_par >> b) & 1u);
                grp_par ^= (1u << b);
              }
            }
            TCK(t_grp);
            if (tl && si == 2 && ci < 10) p.dbg_buf[168 + 4 * ci + 0] = clock64();
            mbar_wait_u32(full_u32 + 8u * stage, phase);   // both halves of the weight chunk have landed
            TCK(t_full);
            if (tl && si == 2 && ci < 10) p.dbg_buf[168 + 4 * ci + 1] = clock64();
            if (synced) tc_fence_after_sync();
            const uint32_t a_lo = a_lo0 + ck.x;
            const uint32_t b_lo = b_lo0 + (uint32_t)stage * (kStageBytes >> 4);
            constexpr uint32_t a_kstep = 2u * ((kTileM * 16u) >> 4);
            const int kc = (int)(ck.y & 0xFFFFu);
            if (tl && first_mma) { p.dbg_buf[64 + si * 8 + 0] = clock64(); first_mma = false; }
            if (prof && lane == 0 && gstep == 0 && ci == 0) p.dbg_buf[242] = clock64();
            if (kF8 && (ck.w & 2u)) {
              // fp8 remainder chunk: K = 32 per instruction, two 16-column units of A and of B each.  First half of the
              // slot: e4m3(W) against the e5m2 remainders; second half: e4m3(2^12 W_lo) against the e5m2 copies,
              // which sit 16 slabs after the remainders
              const uint32_t b2 = b_lo + ((b_kstep * (uint32_t)kc) >> 5);   // (N / 2) * kc bytes, from the register b_kstep lives in
              const uint32_t a2 = a_lo + 16u * ((kTileM * 16u) >> 4);
              uint32_t j = 0;
              if (kTcProfile && (kdbg & 32)) {
                // experiment: 8-bit MMAs not issued at all (garbage results): what a step costs without their tensor time
              } else if (kc == 64) {
                umma2_f8_split_w(acc, a_lo, desc_hi, b_lo, desc_hi, idesc8a, 1u);
                umma2_f8_split_w(acc, a_lo + a_kstep, desc_hi, b_lo + b_kstep, desc_hi, idesc8a, 1u);
                umma2_f8_split_w(acc, a2, desc_hi, b2, desc_hi, idesc8b, 1u);
                umma2_f8_split_w(acc, a2 + a_kstep, desc_hi, b2 + b_kstep, desc_hi, idesc8b, 1u);
              } else {
              for (int kk = 0; kk < kc; kk += 32, ++j)
                umma2_f8_split_w(acc, a_lo + j * a_kstep, desc_hi, b_lo + j * b_kstep, desc_hi, idesc8a, 1u);
              j = 0;
              for (int kk = 0; kk < kc; kk += 32, ++j)
                umma2_f8_split_w(acc, a2 + j * a_kstep, desc_hi, b2 + j * b_kstep, desc_hi, idesc8b, 1u);
              }
              if (tl && si == 2 && ci < 10) p.dbg_buf[168 + 4 * ci + 3] = clock64();
              umma2_commit_u32_w(empty_u32 + 8u * stage);
            } else if constexpr (kX3) {
              // 16 KiB ring slots: at most 64 K columns per chunk.  A split-precision "hi" chunk multiplies W_hi with
              // the hi AND the lo copy of the A columns (the weights are fetched once for both products); the "lo"
              // chunk that follows multiplies W_lo with the hi copy.
              const uint32_t lo_off = ck.w >> 8;
              if (kc == 64) {
                umma2_bf16_split_w(acc, a_lo, desc_hi, b_lo, desc_hi, idesc, accumulate);
#pragma unroll
                for (uint32_t j = 1; j < 4; ++j)
                  umma2_bf16_split_w(acc, a_lo + j * a_kstep, desc_hi, b_lo + j * b_kstep, desc_hi, idesc, 1u);
                if (lo_off) {
#pragma unroll
                  for (uint32_t j = 0; j < 4; ++j)
                    umma2_bf16_split_w(acc, a_lo + lo_off + j * a_kstep, desc_hi, b_lo + j * b_kstep, desc_hi, idesc, 1u);
                }
              } else {
                uint32_t al = a_lo, bl = b_lo, ac = accumulate;
                for (int kk = 0; kk < kc; kk += 16) {
                  umma2_bf16_split_w(acc, al, desc_hi, bl, desc_hi, idesc, ac);
                  ac = 1u; al += a_kstep; bl += b_kstep;
                }
                if (lo_off) {
                  al = a_lo + lo_off; bl = b_lo;
                  for (int kk = 0; kk < kc; kk += 16) {
                    umma2_bf16_split_w(acc, al, desc_hi, bl, desc_hi, idesc, 1u);
                    al += a_kstep; bl += b_kstep;
                  }
                }
              }
              if (tl && si == 2 && ci < 10) p.dbg_buf[168 + 4 * ci + 3] = clock64();
              umma2_commit_u32_w(empty_u32 + 8u * stage);
            } else {
              if (kc == 128) {
                // steady state: eight MMAs with constant descriptor increments (the issue thread must stay well
                // under 128 cycles of scalar work per MMA, profiles/run_umma_bench.py)
                umma2_bf16_split_w(acc, a_lo, desc_hi, b_lo, desc_hi, idesc, accumulate);
#pragma unroll
                for (uint32_t j = 1; j < 8; ++j)
                  umma2_bf16_split_w(acc, a_lo + j * a_kstep, desc_hi, b_lo + j * b_kstep, desc_hi, idesc, 1u);
              } else {
                uint32_t al = a_lo, bl = b_lo, ac = accumulate;
                for (int kk = 0; kk < kc; kk += 16) {
                  umma2_bf16_split_w(acc, al, desc_hi, bl, desc_hi, idesc, ac);
                  ac = 1u; al += a_kstep; bl += b_kstep;
                }
              }
              if (tl && si == 2 && ci < 10) p.dbg_buf[168 + 4 * ci + 3] = clock64();
              umma2_commit_u32_w(empty_u32 + 8u * stage);    // ring slot (in both CTAs) reusable once these MMAs have read it
            }
            accumulate = 1;
            if (tl && si == 2 && ci < 10) p.dbg_buf[168 + 4 * ci + 2] = clock64();
            if (kTcProfile) ++ci;
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            if (ck.w & 1u) break;
            ck = nxt; ++cidx;
          }
          {
            umma2_commit_u32_w(smem_u32(&acc_full[gstep & 1]));      // accumulators complete -> epilogue warps of both CTAs
            if (st.dot) umma2_commit_u32_w(smem_u32(dot_full));       // ... and -> the prologue warps (3-wide output layer)
            // side regions whose only reader was this step may now be rewritten for the next tile (prologue warps)
            // [0]: layer-0 operand region (forward) / the whole main region once the VF-only dgrad tile is finished
            if (si == prog.emb0_last_step) umma2_commit_u32_w(smem_u32(&reg_free[0]));
            if (si == prog.skip_step) umma2_commit_u32_w(smem_u32(&reg_free[1]));
            if (si == prog.aux_step) umma2_commit_u32_w(smem_u32(&reg_free[2]));
          }
          if (tl) p.dbg_buf[64 + si * 8 + 1] = clock64();
        }
      }
      TCK(t_issue);
      if (prof && lane == 0) p.dbg_buf[243] = clock64();
      if (prof) { p.dbg_buf[0] = t_grp; p.dbg_buf[1] = t_full; p.dbg_buf[2] = t_issue; }
    }
    // ===================== peer CTA: relay lane =====================
    else if (kStash && lane == 0) {   // (inference variants: the relay runs in warp 14, see there)
      int stage = 0, phase = 0;
      for (long long pair = pair0; pair < num_pairs; pair += pair_step) {
        for (int si = 0; si < prog.n_steps; ++si) {
          const int nc = p.cnum[si];
          for (int c = 0; c < nc; ++c) {
            mbar_wait(&full[stage], phase);           // my half of the chunk is in my shared memory
            mbar_arrive_remote(&full[stage], 0);      // second arrival on the leader's "slot full" barrier
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 14) {
    // ===================== activation-stash store lane (training forward) =====================
    if constexpr (!kStash) {
      // inference variants: this warp has no stash to store, so its lane 0 is the peer CTA's relay lane.  (Side effect
      // worth knowing: with this branch empty, ptxas stops treating the issuer warp's loop as converged and wraps every
      // tcgen05.mma in ELECT + R2UR.BROADCAST again.)
      if (lane == 0 && rank == 1) {
      int stage = 0, phase = 0;
      for (long long pair = pair0; pair < num_pairs; pair += pair_step) {
        for (int si = 0; si < prog.n_steps; ++si) {
          const int nc = p.cnum[si];
          for (int c = 0; c < nc; ++c) {
            mbar_wait(&full[stage], phase);           // my half of the chunk is in my shared memory
            mbar_arrive_remote(&full[stage], 0);      // second arrival on the leader's "slot full" barrier
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
          }
    }
    if constexpr (kStash) {
      if (lane == 0) {
        uint32_t su = 0;
        for (long long pair = pair0; pair < num_pairs; pair += pair_step) {
          const long long tile = 2 * pair + (long long)rank;
          for (int si = 0; si < prog.n_steps; ++si) {
            const TcStep& st = prog.s[si];
            // forward: hidden-layer / feature outputs that become the next A operand; backward (render() dgrad chain):
            // every step but the last hands its gated gradient on through the activation tile
            const bool via_tile = kBwd ? (prog.bwd == 1 && si + 1 < prog.n_steps && st.stash_out >= 0)
                                       : ((st.epi == TC_EPI_RELU || (st.epi == TC_EPI_FEAT && prog.render != 0)) && st.stash_out >= 0);
            if (!via_tile) continue;
            for (int bg = 0; bg < 4; ++bg) {
              const int c0 = bg * 64;
              mbar_wait(&st_ready[bg], su & 1);
              if (c0 < st.N && tile < num_tiles && !(kdbg & 8)) {
                const uint32_t bytes = (uint32_t)(min(64, st.N - c0) >> 3) * (kTileM * 16);
                bulk_s2g(stash_unit(p, st.stash_out, tile, c0 >> 3, 0), s_act + (size_t)(c0 >> 3) * (kTileM * 16), bytes);
                bulk_wait_read_all();
              }
              mbar_arrive(&st_done[bg]);
            }
            ++su;
          }
        }
        bulk_wait_all();
      }
    }
  } else if (warp >= 10) {
    // ===================== prologue warps (one row per thread) =====================
    // They build, one tile ahead, every A-operand region that depends only on the inputs: the bf16 hi|lo positional
    // encoding of layer 0 (emb0 columns), the encoding / sqrt(2) for the skip layer, and the point + view-direction
    // part of the colour net's small inputs.  A region is rewritten only after the MMAs that read it for the
    // previous tile have completed (reg_free barriers, committed by the MMA issuer).
    const int row = threadIdx.x - 320;
    const int E = prog.emb_w, Epad = prog.emb_pad, nsl = Epad >> 3;
    auto tile_of = [&](long long pair) { return 2 * pair + (long long)rank; };
    auto arrive_pro = [&](int b) {
      __syncwarp();
      if (lane == 0) { if (rank == 0) mbar_arrive(&grp[b]); else mbar_arrive_remote(&grp[b], 0); }
    };
    int n = 0;
    if (kBwd) {
      // backward: the two 3-wide gradient inputs enter as bf16 (hi, lo) pairs in 16-column units:
      //   d(colour pre-sigmoid) -> aux columns 0..15 (A operand of the first step), d(vector pre-tanh) -> skip columns 0..15
      //   VF-only dgrad (prog.bwd == 2): no colour unit; the main region starts as d(feature pre-tanh), which
      //   tc_backward_vf left in the gradient stash as a bf16 tile image -- copied here, one row per thread
      const bool with_rn = prog.aux_step >= 0;
      for (long long pair = pair0; pair < num_pairs; pair += pair_step, ++n) {
        const long long tile = tile_of(pair);
        const long long pi = tile * kTileM + row;
        const bool valid = pi < p.n_points;
        float u[16], w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { u[j] = 0.f; w[j] = 0.f; }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float a = with_rn ? p.dcol_pre[3 * pi + j] : 0.f, b = p.dv_pre[3 * pi + j];
            u[j] = __bfloat162float(__float2bfloat16(a)); u[3 + j] = a - u[j];
            w[j] = __bfloat162float(__float2bfloat16(b)); w[3 + j] = b - w[j];
          }
        }
        if (tile < num_tiles) {
          // the same units are the A operands of the 3-row weight-gradient GEMMs of the two output layers
#pragma unroll
          for (int sl = 0; sl < 2; ++sl) {
            const float* a8 = u + 8 * sl;
            const float* b8 = w + 8 * sl;
            if (with_rn)
              *stash_unit(p, p.sinfo.idx_dcolu, tile, sl, row) =
                  make_uint4(pack_bf16x2(a8[0], a8[1]), pack_bf16x2(a8[2], a8[3]), pack_bf16x2(a8[4], a8[5]), pack_bf16x2(a8[6], a8[7]));
            *stash_unit(p, p.sinfo.idx_dvu, tile, sl, row) =
                make_uint4(pack_bf16x2(b8[0], b8[1]), pack_bf16x2(b8[2], b8[3]), pack_bf16x2(b8[4], b8[5]), pack_bf16x2(b8[6], b8[7]));
          }
        }
        if (with_rn) {
          if (n > 0) mbar_wait(&reg_free[2], (n - 1) & 1);
          store_slab_f(s_act, L::aux / 8, row, u);
          store_slab_f(s_act, L::aux / 8 + 1, row, u + 8);
          fence_proxy_async_smem();
          arrive_pro(kBarAuxStatic);
        } else {
          if (n > 0) mbar_wait(&reg_free[0], (n - 1) & 1);
          const int t_feat = p.sinfo.idx_d0 + p.prog.s[0].mask_src + 1;      // gradient twin of the feature tensor
          const bool t_ok = tile < num_tiles;
#pragma unroll 4
          for (int sl = 0; sl < 32; ++sl) {
            const uint4 v = t_ok ? __ldg(stash_unit(p, t_feat, tile, sl, row)) : make_uint4(0u, 0u, 0u, 0u);
            store_slab_u(s_act, sl, row, v.x, v.y, v.z, v.w);
          }
          fence_proxy_async_smem();
          arrive_pro(0); arrive_pro(1); arrive_pro(2); arrive_pro(3);
        }
        if (n > 0) mbar_wait(&reg_free[1], (n - 1) & 1);
        store_slab_f(s_act, L::skip / 8, row, w);
        store_slab_f(s_act, L::skip / 8 + 1, row, w + 8);
        fence_proxy_async_smem();
        arrive_pro(kBarSkip);
      }
    } else {
      // ---- inputs of tile `pair` (the n-th tile of this CTA), part A: emb0 (+ skip) columns
      auto prep_a = [&](long long pair, int n) {
        const long long tile = tile_of(pair);
        const long long pi = tile * kTileM + row;
        const bool valid = pi < p.n_points;
        const bool st_on = kStash && tile < num_tiles;
        float pt[3], emb[48];
        load_point(p, pi, valid, pt);
#pragma unroll
        for (int i = 0; i < 48; ++i) emb[i] = 0.f;
        embed3<kX3>(pt, prog.multires, emb);
#pragma unroll
        for (int i = 0; i < 48; ++i)
          if (i >= E) emb[i] = 0.f;
        // ---- emb0: hi slabs then lo slabs
        if (n > 0) mbar_wait(&reg_free[0], (n - 1) & 1);
#pragma unroll
        for (int sl = 0; sl < 6; ++sl) {
          if (sl < nsl) {
            float hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float v = emb[sl * 8 + j];
              hi[j] = kF8 ? __half2float(__float2half_rn(v)) : __bfloat162float(__float2bfloat16(v));
              lo[j] = v - hi[j];
            }
            if (kF8) {
              store_slab_h(s_act, L::emb0 / 8 + sl, row, hi);
              store_slab_h(s_act, L::emb0 / 8 + nsl + sl, row, lo);
            } else {
              store_slab_f(s_act, L::emb0 / 8 + sl, row, hi);
              store_slab_f(s_act, L::emb0 / 8 + nsl + sl, row, lo);
            }
            if (st_on) {
              *stash_unit(p, p.sinfo.idx_emb0, tile, sl, row) =
                  make_uint4(pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]), pack_bf16x2(hi[4], hi[5]), pack_bf16x2(hi[6], hi[7]));
              *stash_unit(p, p.sinfo.idx_emb0, tile, nsl + sl, row) =
                  make_uint4(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]), pack_bf16x2(lo[4], lo[5]), pack_bf16x2(lo[6], lo[7]));
            }
          }
        }
        fence_proxy_async_smem();
        arrive_pro(kBarEmb0);
        // ---- skip region: embedding / sqrt(2)   (split-precision tile: the skip layer reads the emb0 columns instead)
        if (!kX3 && prog.skip_step >= 0) {
          if (n > 0) mbar_wait(&reg_free[1], (n - 1) & 1);
#pragma unroll
          for (int i = 0; i < 48; ++i) emb[i] *= kInvSqrt2;
#pragma unroll
          for (int sl = 0; sl < 6; ++sl) {
            store_slab_f(s_act, L::skip / 8 + sl, row, emb + 8 * sl);
            if (st_on) {
              const float* e8 = emb + 8 * sl;
              *stash_unit(p, p.sinfo.idx_skip, tile, sl, row) =
                  make_uint4(pack_bf16x2(e8[0], e8[1]), pack_bf16x2(e8[2], e8[3]), pack_bf16x2(e8[4], e8[5]), pack_bf16x2(e8[6], e8[7]));
            }
          }
          fence_proxy_async_smem();
          arrive_pro(kBarSkip);
        }
      };
      // ---- part B: aux region, columns 8..47: [p(3), embed(view dir)(3+6*Lv), 0...]  (columns 0..7: the normal, below)
      auto prep_b = [&](long long pair, int n) {
        const long long tile = tile_of(pair);
        const long long pi = tile * kTileM + row;
        const bool valid = pi < p.n_points;
        const bool st_on = kStash && tile < num_tiles;
        float pt[3], a[40], d[3] = {0.f, 0.f, 0.f};
        load_point(p, pi, valid, pt);
#pragma unroll
        for (int j = 0; j < 40; ++j) a[j] = 0.f;
        if (valid) {
          const long long r = pi / p.samples_per_ray;
          d[0] = __ldg(p.ray_dirs + 3 * r); d[1] = __ldg(p.ray_dirs + 3 * r + 1); d[2] = __ldg(p.ray_dirs + 3 * r + 2);
        }
        embed3<kX3>(d, prog.multires_view, a + 3);
        const int ev = 3 + 6 * prog.multires_view;
#pragma unroll
        for (int j = 0; j < 40; ++j) {
          if (j < 3) a[j] = pt[j];
          if (j >= 3 + ev) a[j] = 0.f;
        }
        if (n > 0) mbar_wait(&reg_free[2], (n - 1) & 1);
#pragma unroll
        for (int sl = 0; sl < 5; ++sl) {
          store_slab_f(s_act, L::aux / 8 + 1 + sl, row, a + 8 * sl);
          if (st_on) {
            const float* e8 = a + 8 * sl;
            *stash_unit(p, p.sinfo.idx_aux, tile, 1 + sl, row) =
                make_uint4(pack_bf16x2(e8[0], e8[1]), pack_bf16x2(e8[2], e8[3]), pack_bf16x2(e8[4], e8[5]), pack_bf16x2(e8[6], e8[7]));
          }
        }
        fence_proxy_async_smem();
        arrive_pro(kBarAuxStatic);
      };
      // ---- the 3-wide output layer that reads step si's activation (TcStep::dot), on the CUDA cores: this thread owns
      // TMEM lane (warp % 4) * 32 + lane = one point; it reads the step's fp32 accumulator row (256 columns), applies the
      // ReLU and takes the three dot products with the fp32 weight rows (broadcast reads from shared memory).  The
      // epilogue warps do the hand-off of the same accumulator concurrently; the MMA issuer does not reuse the
      // accumulator buffer before this warp has arrived on kBarDot.
      const int qd = warp & 3, rowd = qd * 32 + lane;
      uint32_t n_dots = 0;
      auto dot_step = [&](int si, uint32_t gstep, long long pair) {
        const int dot = prog.s[si].dot;
        const long long tile = tile_of(pair);
        const long long pi = tile * kTileM + rowd;
        const bool valid = pi < p.n_points;
        mbar_wait(dot_full, n_dots & 1);
        ++n_dots;
        tc_fence_after_sync();
        const uint32_t acc = tmem + (gstep & 1) * kAccCols + ((uint32_t)(qd * 32) << 16);
        const float* wd = s_dotw + (dot - 1) * 768;
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, e0 = 0.f, e1 = 0.f, e2 = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(acc + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 w0 = *reinterpret_cast<const float4*>(wd + c0 + 4 * c4);
            const float4 w1 = *reinterpret_cast<const float4*>(wd + 256 + c0 + 4 * c4);
            const float4 w2 = *reinterpret_cast<const float4*>(wd + 512 + c0 + 4 * c4);
            const float y0 = fmaxf(__uint_as_float(v[4 * c4]), 0.f), y1 = fmaxf(__uint_as_float(v[4 * c4 + 1]), 0.f);
            const float y2 = fmaxf(__uint_as_float(v[4 * c4 + 2]), 0.f), y3 = fmaxf(__uint_as_float(v[4 * c4 + 3]), 0.f);
            d0 = fmaf(y0, w0.x, d0); e0 = fmaf(y1, w0.y, e0); d0 = fmaf(y2, w0.z, d0); e0 = fmaf(y3, w0.w, e0);
            d1 = fmaf(y0, w1.x, d1); e1 = fmaf(y1, w1.y, e1); d1 = fmaf(y2, w1.z, d1); e1 = fmaf(y3, w1.w, e1);
            d2 = fmaf(y0, w2.x, d2); e2 = fmaf(y1, w2.y, e2); d2 = fmaf(y2, w2.z, d2); e2 = fmaf(y3, w2.w, e2);
          }
        }
        tc_fence_before_sync();
        // accumulator row consumed: the buffer may be overwritten (matched by the issuer's wait at TcProgram::dot_guard)
        { __syncwarp(); if (lane == 0) { if (rank == 0) mbar_arrive(&grp[kBarDot]); else mbar_arrive_remote(&grp[kBarDot], 0); } }
        const float* bias = s_dotb + 3 * (dot - 1);
        const float r0 = d0 + e0 + bias[0], r1 = d1 + e1 + bias[1], r2 = d2 + e2 + bias[2];
        if (dot == 1) {
          const float nv[3] = {tanhf(r0), tanhf(r1), tanhf(r2)};
          if (valid) {
#pragma unroll
            for (int j = 0; j < 3; ++j) p.out_v[pi * p.v_ld + j] = nv[j];
          }
          if (prog.render) {
            // the normal is the first 16-byte unit of the colour net's small inputs (aux columns 0..7); its previous
            // reader (the colour net's first layer of the previous tile) completed long ago
            const float a[8] = {nv[0], nv[1], nv[2], 0.f, 0.f, 0.f, 0.f, 0.f};
            store_slab_f(s_act, L::aux / 8, rowd, a);
            if (kStash && tile < num_tiles)
              *stash_unit(p, p.sinfo.idx_aux, tile, 0, rowd) = make_uint4(pack_bf16x2(nv[0], nv[1]), pack_bf16x2(nv[2], 0.f), 0u, 0u);
            fence_proxy_async_smem();
            arrive_pro(kBarAux);           // matched by the colour net's first step
          }
        } else if (valid) {
          p.colors[3 * pi] = 1.f / (1.f + expf(-r0));
          p.colors[3 * pi + 1] = 1.f / (1.f + expf(-r1));
          p.colors[3 * pi + 2] = 1.f / (1.f + expf(-r2));
        }
      };
      // per tile: [inputs A of the NEXT tile] -> vector dot -> [inputs B of the next tile] -> colour dot, so the next
      // tile's operands are in place long before its first MMA and no region is rewritten while it is still being read
      const int d_first = prog.dot_step[0], d_second = prog.dot_step[1];
      // split-precision tile: the aux columns alias lo columns that are live until the last hidden VF layer's MMAs have
      // completed, so part B is written for the CURRENT tile right after the vector dot step (which waits for exactly that)
      constexpr bool kAuxLate = kX3;
      if (pair0 < num_pairs) { prep_a(pair0, 0); if (!kAuxLate && prog.aux_step >= 0) prep_b(pair0, 0); }
      for (long long pair = pair0; pair < num_pairs; pair += pair_step, ++n) {
        const long long next = pair + pair_step;
        const uint32_t g0 = (uint32_t)n * (uint32_t)prog.n_steps;
        if (next < num_pairs) prep_a(next, n + 1);
        if (d_first >= 0) dot_step(d_first, g0 + d_first, pair);
        if (kAuxLate) { if (prog.aux_step >= 0) prep_b(pair, n); }
        else if (next < num_pairs && prog.aux_step >= 0) prep_b(next, n + 1);
        if (d_second >= 0) dot_step(d_second, g0 + d_second, pair);
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access (hardware: warp % 4)
    const int h = (warp - 2) >> 2;          // column-group parity: this warp owns groups g with g % 2 == h
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const bool render = prog.render != 0;
    uint32_t gstep = 0;
    uint32_t su = 0;          // stashed steps whose output left through warp 14 so far (st_ready / st_done phases)
    long long t_acc = 0, t_ld = 0, t_math = 0, t_sig = 0, t_other = 0, t0 = 0;
    const bool prof = kTcProfile && p.dbg_buf && blockIdx.x == 0 && lane == 0 && (warp == 2 || warp == 3);
    auto tile_of = [&](long long pair) { return 2 * pair + (long long)rank; };
    if (prof) t0 = clock64();

    int tile_no = 0;
    for (long long pair = pair0; pair < num_pairs; pair += pair_step, ++tile_no) {
      const long long tile = tile_of(pair);
      const long long pi = tile * kTileM + row;
      const bool valid = pi < p.n_points;
      const bool tl = prof && tile_no == 2;
      for (int si = 0; si < prog.n_steps; ++si, ++gstep) {
        const TcStep& st = prog.s[si];
        const uint32_t acc = tmem + (gstep & 1) * kAccCols + lane_off;
        if constexpr (kBwd) {
          // dgrad step: acc = dL/d(this layer's input); gate it with the stashed forward activation Y of that input
          // (ReLU: Y > 0; tanh: 1 - Y^2), hand it on as the next step's A operand and stash it for the wgrad GEMM.
          // ReLU gates were stored by the forward as one bit per element (8 bytes per row and 64 columns); they are
          // fetched BEFORE waiting for the accumulator, so their latency hides behind this step's MMAs.
          const bool is_tanh = st.epi == TC_EPI_BWD_TANH;
          const int stN = st.N;
          const bool t_ok = tile < num_tiles;
          // the last step's output (dL/d pre-activation of layer 0) has no consumer in this launch: it is only stashed,
          // and it must not arrive on the column-group barriers (every arrival set is matched by exactly one wait of
          // the MMA issuer; an unmatched one would shift the barrier phases of the next tile)
          const bool hand_on = si + 1 < prog.n_steps;
          // render() dgrad chain: handed-on gradients are stashed by warp 14 from the activation tile (see st_ready)
          const bool st_tile = prog.bwd == 1 && hand_on && st.stash_out >= 0;
          uint32_t gbits[2][2] = {{0u, 0u}, {0u, 0u}};
          if (!is_tanh && t_ok) {
#pragma unroll
            for (int it = 0; it < 2; ++it) {
              const int bg = 2 * it + h;
              if (bg * 64 < stN) {
                const uint2 g2 = __ldg(gate_unit(p, st.mask_src, tile, bg, row));
                gbits[it][0] = g2.x; gbits[it][1] = g2.y;
              }
            }
          }
          TCK(t_other);
          mbar_wait(&acc_full[gstep & 1], (gstep >> 1) & 1);
          TCK(t_acc);
          tc_fence_after_sync();
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int bg = 2 * it + h, c0 = bg * 64;
            if (c0 >= stN) {
              if (st_tile) mbar_wait(&st_done[bg], (su & 1) ^ 1);
              publish(bg, hand_on, st_tile ? &st_ready[bg] : nullptr);
              continue;
            }
            const bool second = c0 + 32 < stN;
            // both 32-column halves of the group are loaded together (like the forward epilogue): one wait, not two
            uint32_t vab[2][32];
            tmem_ld32(acc + c0, vab[0]);
            if (second) tmem_ld32(acc + c0 + 32, vab[1]);
            // the bulk copy of this group's previous contents must have read them before they are overwritten: checked
            // while the accumulator load is in flight
            if (st_tile) mbar_wait(&st_done[bg], (su & 1) ^ 1);
            tmem_ld_wait();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              if (half == 1 && !second) break;
              const int cb = c0 + 32 * half;
              const uint32_t* va = vab[half];
              uint4 yt[4];
              if (is_tanh) {
#pragma unroll
                for (int sl = 0; sl < 4; ++sl)
                  yt[sl] = t_ok ? __ldg(stash_unit(p, st.mask_src, tile, (cb >> 3) + sl, row)) : make_uint4(0, 0, 0, 0);
              }
              const uint32_t bits = gbits[it][half];
#pragma unroll
              for (int sl = 0; sl < 4; ++sl) {
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float a0 = __uint_as_float(va[8 * sl + 2 * j]), a1 = __uint_as_float(va[8 * sl + 2 * j + 1]);
                  if (is_tanh) {
                    const uint32_t y2 = reinterpret_cast<const uint32_t*>(&yt[sl])[j];
                    const float y0 = __uint_as_float(y2 << 16), y1 = __uint_as_float(y2 & 0xFFFF0000u);
                    o[j] = pack_bf16x2(a0 * (1.f - y0 * y0), a1 * (1.f - y1 * y1));
                  } else {
                    const int e = 8 * sl + 2 * j;
                    o[j] = pack_bf16x2(((bits >> e) & 1u) ? a0 : 0.f, ((bits >> (e + 1)) & 1u) ? a1 : 0.f);
                  }
                }
                if (hand_on) store_slab_u(s_act, (cb >> 3) + sl, row, o[0], o[1], o[2], o[3]);
                if (!st_tile && t_ok && !(kdbg & 8))
                  *stash_unit(p, st.stash_out, tile, (cb >> 3) + sl, row) = make_uint4(o[0], o[1], o[2], o[3]);
              }
            }
            if (hand_on) {
              fence_proxy_async_smem();
              tc_fence_before_sync();
              publish(bg, true, st_tile ? &st_ready[bg] : nullptr);
            }
          }
          if (st_tile) ++su;
          tc_fence_before_sync();
        } else {
        TCK(t_other);
        // the step's facts are fetched from the parameter bank BEFORE the accumulator wait (the empty asm pins the loads
        // there): behind it every cycle is hand-off time -- tensor-core idle time
        const int st_epi = st.epi, st_n = st.N, st_stash = st.stash_out, st_outlo = st.out_lo, st_f16 = st.a_f16;
        asm volatile("" ::"r"(st_epi), "r"(st_n), "r"(st_stash), "r"(st_outlo), "r"(st_f16));
        mbar_wait(&acc_full[gstep & 1], (gstep >> 1) & 1);
        TCK(t_acc);
        if (tl) p.dbg_buf[64 + si * 8 + 2 + 3 * h] = clock64();
        tc_fence_after_sync();
        if (st_epi == TC_EPI_RELU || st_epi == TC_EPI_FEAT) {
          const bool feat = st_epi == TC_EPI_FEAT;
          const bool to_act = !feat || render;
          const int stN = st_n;
          const bool st_on = kStash && st_stash >= 0 && tile < num_tiles;
          const bool st_tile = kStash && st_stash >= 0 && to_act;     // this step's output leaves through warp 14
          // the last step of a program has no consumer in the activation tile: its output is only stashed (training)
          // and/or reduced to a 3-wide output by the prologue warps (TcStep::dot), and it must not arrive on the
          // column-group barriers (every arrival set is matched by exactly one wait of the MMA issuer)
          const bool consumer = si + 1 < prog.n_steps;
          const bool store = consumer || st_tile;
          if (to_act) {
            // Each warp owns two of the four 64-column groups (= K chunks of the next layer): h=0 -> groups 0 and 2,
            // h=1 -> groups 1 and 3.  One generic->async proxy fence (a MEMBAR.ALL.CTA) and one barrier arrival per
            // 64 columns: the fence, not the math, is the expensive part of this loop.
            // Split-precision tile: a step spends three times as long in the tensor core, so the extra fences are free
            // and what counts is how soon the FIRST K chunk of the next layer exists: all eight warps work on the same
            // group (32 columns each: h selects the half), groups in K order, four fences + arrivals per warp
            // (measured the other way round on the plain bf16 tile, profiles/r01_forward_kernel_experiments.md "v4b").
            constexpr int kIts = kSerialGroups ? 4 : 2;
#pragma unroll 1
            for (int it = 0; it < kIts; ++it) {
              const int bg = kSerialGroups ? it : 2 * it + h;
              const int c0 = kSerialGroups ? bg * 64 + 32 * h : bg * 64;
              if (c0 >= stN) {                                 // narrow layer: nothing to write, the barriers still count us
                if (st_tile) mbar_wait(&st_done[bg], (su & 1) ^ 1);
                publish(bg, consumer, st_tile ? &st_ready[bg] : nullptr);
                continue;
              }
              // nothing leaves through the activation tile (inference: the last step only feeds the prologue warps' dot)
              if (!store) continue;
              const bool second = !kSerialGroups && c0 + 32 < stN;
              uint32_t va[32], vb[32];
              TCK(t_other);
              if (tl && it == 0 && warp == 2) p.dbg_buf[64 + si * 8 + 5] = clock64();
              tmem_ld32(acc + c0, va);
              if (second) tmem_ld32(acc + c0 + 32, vb);
              // the bulk copy of this group's previous contents (last stashed step) must have read them before they are
              // overwritten: checked while the accumulator load is in flight
              if (st_tile) mbar_wait(&st_done[bg], (su & 1) ^ 1);
              tmem_ld_wait();
              TCK(t_ld);
              if (tl && it == 0 && warp == 2) p.dbg_buf[64 + si * 8 + 6] = clock64();
              if (kF8 && st_f16 && !feat) {
                // fp16 + fp8-remainder hand-off of a VF hidden layer: y = relu(acc) leaves as fp16(y) in the main columns,
                // e5m2(y - fp16(y)) and e5m2(2^-12 y) in the lo region (16 columns per 16-byte unit there)
                const uint32_t* v = va;            // (kSerialGroups: 32 columns per warp and iteration)
                uint32_t r8[8], h8[8];
#pragma unroll
                for (int sl = 0; sl < 4; ++sl) {
                  uint32_t hi[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float y0 = fmaxf(__uint_as_float(v[8 * sl + 2 * j]), 0.f), y1 = fmaxf(__uint_as_float(v[8 * sl + 2 * j + 1]), 0.f);
                    const __half2 hh = __floats2half2_rn(y0, y1);
                    hi[j] = *reinterpret_cast<const uint32_t*>(&hh);
                    const float2 hf2 = __half22float2(hh);
                    const uint32_t lo8 = __nv_cvt_float2_to_fp8x2(make_float2(y0 - hf2.x, y1 - hf2.y), __NV_SATFINITE, __NV_E5M2);
                    const uint32_t hi8 = __nv_cvt_float2_to_fp8x2(make_float2(y0 * (1.f / kF8ScaleHi), y1 * (1.f / kF8ScaleHi)),
                                                                  __NV_SATFINITE, __NV_E5M2);
                    const int e = 4 * sl + j;        // 16-bit pair e of the 32 columns -> half of 32-bit word e / 2
                    if (e & 1) { r8[e >> 1] |= lo8 << 16; h8[e >> 1] |= hi8 << 16; }
                    else { r8[e >> 1] = lo8; h8[e >> 1] = hi8; }
                  }
                  EPI_STORE(s_act, (c0 >> 3) + sl, row, hi[0], hi[1], hi[2], hi[3]);
                }
                if (st_outlo) {       // (not for the last hidden layer inside render(): its lo columns hold the aux inputs)
                  const int u8 = L::lo / 8 + (c0 >> 4);
                  EPI_STORE(s_act, u8, row, r8[0], r8[1], r8[2], r8[3]);
                  EPI_STORE(s_act, u8 + 1, row, r8[4], r8[5], r8[6], r8[7]);
                  EPI_STORE(s_act, u8 + 16, row, h8[0], h8[1], h8[2], h8[3]);
                  EPI_STORE(s_act, u8 + 17, row, h8[4], h8[5], h8[6], h8[7]);
                }
              } else if constexpr (kX3) {
                // split-precision hand-off: y = relu(acc) leaves as hi = bf16(y) in the main columns and, for the VF
                // layers, lo = bf16(y - hi) in the lo columns (same slab, kX3ColLo further on)
                const bool out_lo = st_outlo != 0;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                  if (hf == 1 && !second) break;
                  const uint32_t* v = hf == 0 ? va : vb;
#pragma unroll
                  for (int sl = 0; sl < 4; ++sl) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const float a0 = __uint_as_float(v[8 * sl + 2 * j]), a1 = __uint_as_float(v[8 * sl + 2 * j + 1]);
                      if (feat) {
                        hi[j] = tanh_bf16x2(pack_bf16x2(a0, a1));
                        lo[j] = 0u;
                      } else {
                        const float y0 = fmaxf(a0, 0.f), y1 = fmaxf(a1, 0.f);
                        hi[j] = pack_bf16x2(y0, y1);
                        lo[j] = pack_bf16x2(y0 - __uint_as_float(hi[j] << 16), y1 - __uint_as_float(hi[j] & 0xFFFF0000u));
                      }
                    }
                    const int slab = (c0 >> 3) + 4 * hf + sl;
                    EPI_STORE(s_act, slab, row, hi[0], hi[1], hi[2], hi[3]);
                    if (out_lo) EPI_STORE(s_act, slab + L::lo / 8, row, lo[0], lo[1], lo[2], lo[3]);
                  }
                }
              } else if (!(kdbg & 2)) {
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  pk[j] = feat ? tanh_bf16x2(pack_bf16x2(__uint_as_float(va[2 * j]), __uint_as_float(va[2 * j + 1])))
                               : pack_relu_bf16x2(__uint_as_float(va[2 * j]), __uint_as_float(va[2 * j + 1]));
                }
#pragma unroll
                for (int sl = 0; sl < 4; ++sl)
                  store_slab_u(s_act, (c0 >> 3) + sl, row, pk[4 * sl], pk[4 * sl + 1], pk[4 * sl + 2], pk[4 * sl + 3]);
                if (second) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    pk[j] = feat ? tanh_bf16x2(pack_bf16x2(__uint_as_float(vb[2 * j]), __uint_as_float(vb[2 * j + 1])))
                                 : pack_relu_bf16x2(__uint_as_float(vb[2 * j]), __uint_as_float(vb[2 * j + 1]));
                  }
#pragma unroll
                  for (int sl = 0; sl < 4; ++sl)
                    store_slab_u(s_act, (c0 >> 3) + 4 + sl, row, pk[4 * sl], pk[4 * sl + 1], pk[4 * sl + 2], pk[4 * sl + 3]);
                }
              }
              TCK(t_math);
              if (tl && it == 0 && warp == 2) p.dbg_buf[64 + si * 8 + 7] = clock64();
              fence_proxy_async_smem();
              tc_fence_before_sync();
              publish(bg, consumer, st_tile ? &st_ready[bg] : nullptr);
              TCK(t_sig);
              if constexpr (kStash && !kX3) {
                // training forward: the ReLU gates of this group as bits, AFTER the group has been handed on (the next
                // layer's MMAs wait for the arrival above, not for this).  Gate bit e = "pre-activation of column c0 + e is
                // not negative": one funnel shift per element collects the fp32 sign bits (an exact zero passes the gate;
                // its gradient contribution is zero or belongs to a padded channel).  The backward gates on the sign
                // only: 8 bytes per row and 64-column group instead of 128.
                if (st_on && !feat && !(kdbg & 16) && !(kdbg & 2)) {
                  uint32_t gate_lo = 0, gate_hi = 0;
#pragma unroll
                  for (int j = 31; j >= 0; --j) gate_lo = __funnelshift_l(va[j], gate_lo, 1);
                  gate_lo = ~gate_lo;
                  if (second) {
#pragma unroll
                    for (int j = 31; j >= 0; --j) gate_hi = __funnelshift_l(vb[j], gate_hi, 1);
                    gate_hi = ~gate_hi;
                  }
                  *gate_unit(p, st.stash_out, tile, bg, row) = make_uint2(gate_lo, gate_hi);
                }
              }
              if (tl && it == 0) p.dbg_buf[64 + si * 8 + 3 + 3 * h] = clock64();
              if (tl) p.dbg_buf[64 + si * 8 + 4 + 3 * h] = clock64();
            }
            if (st_tile) ++su;
          } else {
            // VF_FULL: features go to global memory as fp32 (module-call output), 32 columns at a time
#pragma unroll 1
            for (int gi = 0; gi < 4; ++gi) {
              const int c0 = (2 * gi + h) * 32;
              if (c0 >= stN) continue;
              uint32_t v[32];
              tmem_ld32(acc + c0, v);
              tmem_ld_wait();
              float f[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = kX3 ? tanhf(__uint_as_float(v[j])) : tanh_fast(__uint_as_float(v[j]));
              if (st_on) {       // module call kept for a backward: the features are the tanh gate of the dgrad chain
#pragma unroll
                for (int sl = 0; sl < 4; ++sl)
                  *stash_unit(p, st.stash_out, tile, (c0 >> 3) + sl, row) =
                      make_uint4(pack_bf16x2(f[8 * sl], f[8 * sl + 1]), pack_bf16x2(f[8 * sl + 2], f[8 * sl + 3]),
                                 pack_bf16x2(f[8 * sl + 4], f[8 * sl + 5]), pack_bf16x2(f[8 * sl + 6], f[8 * sl + 7]));
              }
              if (p.out_feat && valid) {
                float* o = p.out_feat + pi * p.feat_ld + c0;
                if ((reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
                  for (int j4 = 0; j4 < 8; ++j4)
                    reinterpret_cast<float4*>(o)[j4] = make_float4(f[4 * j4], f[4 * j4 + 1], f[4 * j4 + 2], f[4 * j4 + 3]);
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j) o[j] = f[j];
                }
              }
            }
          }
        }
        tc_fence_before_sync();
        }   // forward steps
      }
    }
    TCK(t_other);
    if (prof) {
      long long* o = p.dbg_buf + (warp == 2 ? 8 : 16);
      o[0] = t_acc; o[1] = t_ld; o[2] = t_math; o[3] = t_sig; o[4] = t_other;
    }
  }
  tc_fence_before_sync();
  cluster_sync_all();                       // neither CTA may exit (or free TMEM) while its peer can still touch it
  if (stamp) p.dbg_buf[244] = clock64();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int round16(int x) { return (x + 15) / 16 * 16; }

static int build_programs(int multires, int multires_view, int skip_layer, const vfnerf_mlp_desc& vf,
                          const vfnerf_mlp_desc* rn, TcPlan& plan, int x3) {
  const int E = 3 + 6 * multires, Epad = round16(E);
  const int L = vf.n_layers;
  VFN_REQUIRE(Epad <= 48 && multires <= kMaxRes && multires_view <= kMaxRes, "tensor-core path: embedding too wide");
  VFN_REQUIRE(L >= 3 && L + (rn ? rn->n_layers - 1 : 0) <= kTcTableSteps, "tensor-core path: too many layers");
  VFN_REQUIRE(skip_layer < 0 || (skip_layer >= 2 && skip_layer < L - 1), "tensor-core path: skip_layer=%d unsupported", skip_layer);
  for (int l = 0; l < L; ++l) {
    const int want_in = (l == 0) ? E : 256;
    const int want_out = (l == L - 1) ? 259 : ((l + 1 == skip_layer) ? 256 - E : 256);
    VFN_REQUIRE(vf.in_dim[l] == want_in && vf.out_dim[l] == want_out,
                "tensor-core path supports the shipped 256-wide VF net (layer %d is %d->%d); use precision fp32",
                l, vf.in_dim[l], vf.out_dim[l]);
  }
  TcProgram pr{};
  const int f8 = x3 == 2;
  auto layout = [&](TcProgram& q, int is_x3) {
    q.x3 = is_x3 ? 1 : 0; q.f8 = is_x3 == 2;
    q.col_aux = is_x3 ? kX3ColAux : kColAux; q.col_skip = is_x3 ? -1 : kColSkip; q.col_ones = is_x3 ? kX3ColOnes : kColOnes;
    q.col_emb0 = is_x3 ? kX3ColEmb0 : kColEmb0; q.col_lo = is_x3 ? kX3ColLo : 0;
  };
  layout(pr, x3);
  pr.emb_w = E; pr.emb_pad = Epad; pr.multires = multires; pr.multires_view = multires_view;
  pr.small_w = 3 + (3 + 6 * multires_view) + 3;
  pr.skip_step = skip_layer;
  pr.aux_step = -1;
  pr.emb0_last_step = (x3 && skip_layer > 0) ? skip_layer : 0;
  int ns = 0;
  long long woff = 0;
  // the 16 KiB ring slots of the split-precision tile hold 64 K columns of a 256-channel layer
  const int CK = x3 ? 64 : 128;
  // appends a step; the bias (ones) segment is added automatically as the last segment.  seglo (may be null): per segment,
  // column distance to the lo copy of the A operand (split-precision segment); segw (may be null): extra weight factor
  auto add = [&](int N, int n_valid, int nseg, const int* col0, const int* segk, int chunk_k, int epi, int fresh,
                 int net, int layer, int row0, int colmap, int src_split, float post, int no_bias = 0,
                 const int* seglo = nullptr, const float* segw = nullptr) {
    TcStep& s = pr.s[ns];
    s = TcStep{};
    s.N = N; s.n_valid = n_valid; s.n_seg = nseg + (no_bias ? 0 : 1); s.K = 0;
    for (int i = 0; i < kTcMaxSegs; ++i) { s.seg_lo[i] = 0; s.seg_wscale[i] = 1.f; s.seg_f8[i] = 0; s.seg_bar[i] = 0; }
    for (int i = 0; i < nseg; ++i) {
      s.seg_col0[i] = col0[i]; s.seg_k[i] = segk[i];
      s.seg_lo[i] = seglo ? seglo[i] : 0; s.seg_wscale[i] = segw ? segw[i] : 1.f;
      // fp16 + fp8 remainders: the 256-wide main segments of the VF steps (the small embedding segments keep three
      // 16-bit products, in fp16)
      s.seg_f8[i] = (f8 && net == 0 && s.seg_lo[i] == kX3ColLo) ? 1 : 0;
      s.K += segk[i] * (s.seg_lo[i] ? 2 : 1);
    }
    s.a_f16 = (f8 && net == 0) ? 1 : 0;
    if (!no_bias) { s.seg_col0[nseg] = pr.col_ones; s.seg_k[nseg] = 16; s.K += 16; }
    s.use_lo = seglo ? 1 : 0; s.out_lo = 0;
    s.no_bias = no_bias; s.stash_out = -1; s.mask_src = -1;
    s.chunk_k = chunk_k; s.epi = epi; s.fresh_mask = fresh; s.pre_wait_mask = 0; s.w_off = woff;
    s.net = net; s.layer = layer; s.row0 = row0; s.colmap = colmap; s.src_split = src_split; s.post_scale = post;
    woff += (long long)align_up((int64_t)N * s.K * 2, 128);
    ++ns;
  };
  const int main0[1] = {0};
  const int lo_main[1] = {kX3ColLo};
  for (int l = 0; l < L - 1; ++l) {
    const float post = (l + 1 == skip_layer) ? kInvSqrt2 : 1.f;
    const int N = round16(vf.out_dim[l]);
    if (x3) {
      // every VF hidden layer: three products per K chunk (mlp_tc.cuh), output written as (hi, lo)
      if (l == 0) {
        const int c0[1] = {kX3ColEmb0}, k[1] = {Epad}, lo[1] = {Epad};
        add(N, vf.out_dim[l], 1, c0, k, CK, TC_EPI_RELU, 1 << kBarEmb0, 0, l, 0, 0, 0, post, 0, lo);
      } else if (l == skip_layer) {
        // [previous layer (already / sqrt(2)) | embedding]: the embedding comes from the layer-0 operand columns
        // (their barrier was consumed by step 0: not "fresh"), its weights carry the 1 / sqrt(2)
        const int prev = vf.out_dim[l - 1];
        const int c[2] = {0, kX3ColEmb0}, k[2] = {round16(prev), Epad}, lo[2] = {kX3ColLo, Epad};
        const float w[2] = {1.f, kInvSqrt2};
        add(N, vf.out_dim[l], 2, c, k, CK, TC_EPI_RELU, 0xF, 0, l, 0, 3, prev, post, 0, lo, w);
      } else {
        const int k[1] = {256};
        add(N, vf.out_dim[l], 1, main0, k, CK, TC_EPI_RELU, 0xF, 0, l, 0, 0, 0, post, 0, lo_main);
      }
      pr.s[ns - 1].out_lo = 1;
    } else if (l == 0) {
      const int k[1] = {2 * Epad};
      const int c0[1] = {kColEmb0};
      add(N, vf.out_dim[l], 1, c0, k, CK, TC_EPI_RELU, 1 << kBarEmb0, 0, l, 0, 1, 0, post);
    } else if (l == skip_layer) {
      const int prev = vf.out_dim[l - 1];
      const int c[2] = {0, kColSkip}, k[2] = {round16(prev), 48};
      add(N, vf.out_dim[l], 2, c, k, CK, TC_EPI_RELU, 0xF | (1 << kBarSkip), 0, l, 0, 3, prev, post);
    } else {
      const int k[1] = {256};
      add(N, vf.out_dim[l], 1, main0, k, CK, TC_EPI_RELU, 0xF, 0, l, 0, 0, 0, post);
    }
  }
  const int k256[1] = {256};
  // the 3 vector rows of the VF output layer are evaluated by the epilogue of the last hidden layer (TcStep::dot)
  VFN_REQUIRE(vf.out_dim[L - 2] == 256, "tensor-core path: the last hidden VF layer must be 256 wide");
  pr.s[ns - 1].dot = 1;
  const int n_v = ns;
  // feature rows of the VF output layer: the image is split-precision as well; render() uses only its hi products (the
  // features are rounded to bf16 for the colour net anyway), the module call (VF_FULL) all three
  add(256, 256, 1, main0, k256, CK, TC_EPI_FEAT, 0xF, 0, L - 1, 3, 0, 0, 1.f, 0, x3 ? lo_main : nullptr);
  const int n_full = ns;
  if (rn) {
    const int Lr = rn->n_layers;
    VFN_REQUIRE(pr.small_w + 5 <= 48, "tensor-core path: colour-net small inputs %d too wide", pr.small_w);
    for (int l = 0; l < Lr; ++l) {
      const int want_in = (l == 0) ? pr.small_w + 256 : 256;
      const int want_out = (l == Lr - 1) ? 3 : 256;
      VFN_REQUIRE(rn->in_dim[l] == want_in && rn->out_dim[l] == want_out,
                  "tensor-core path supports the shipped 256-wide colour net (layer %d is %d->%d); use precision fp32",
                  l, rn->in_dim[l], rn->out_dim[l]);
    }
    const int c[2] = {0, pr.col_aux}, k[2] = {256, 48};
    pr.aux_step = ns;
    add(256, 256, 2, c, k, CK, TC_EPI_RELU, 0xF | (1 << kBarAux) | (1 << kBarAuxStatic), 1, 0, 0, 2, pr.small_w, 1.f);
    if (x3) pr.s[ns - 1].seg_bar[1] = (1 << kBarAux) | (1 << kBarAuxStatic);   // aux columns alias lo columns (mlp_tc.cuh)
    for (int l = 1; l < Lr - 1; ++l) add(256, 256, 1, main0, k256, CK, TC_EPI_RELU, 0xF, 1, l, 0, 0, 0, 1.f);
    VFN_REQUIRE(Lr >= 2, "tensor-core path: the colour net needs a hidden layer");
    pr.s[ns - 1].dot = 2;         // the 3 colour rows: epilogue of the last hidden colour layer
  }
  plan.dot_off = align_up(woff, 128);
  woff = plan.dot_off + kTcDotFloats * 4;
  plan.wpack_bytes = woff;
  pr.n_steps = ns; pr.render = 1;
  // a dot step's accumulator row is read by the prologue warps: the step that next writes the same TMEM buffer (two
  // steps later, possibly in the next tile) waits for their kBarDot arrival
  auto set_guards = [&](TcProgram& q) {
    q.dot_step[0] = q.dot_step[1] = -1;
    int nd = 0;
    for (int i = 0; i < q.n_steps; ++i) { q.s[i].dot_guard_next = 0; q.s[i].pre_wait_mask &= ~(1 << kBarDot); }
    for (int i = 0; i < q.n_steps; ++i) {
      if (!q.s[i].dot) continue;
      q.dot_step[nd++] = i;
      const int g = i + 2;
      if (g < q.n_steps) q.s[g].pre_wait_mask |= 1 << kBarDot; else q.s[g - q.n_steps].dot_guard_next = 1;
    }
  };
  for (int i = 0; i < ns; ++i)
    for (int g = 0; g < pr.s[i].n_seg; ++g)
      VFN_REQUIRE(!pr.s[i].seg_f8[g] || pr.s[i].seg_k[g] % 32 == 0, "tensor-core path (fp16f8): segment of %d columns is not a "
                  "multiple of 32", pr.s[i].seg_k[g]);
  for (int i = 0; i < ns; ++i) {
    int nc = 0;
    for (int g = 0; g < pr.s[i].n_seg; ++g)
      nc += (pr.s[i].seg_k[g] + pr.s[i].chunk_k - 1) / pr.s[i].chunk_k * (pr.s[i].seg_lo[g] ? 2 : 1);
    VFN_REQUIRE(nc <= kTcMaxChunks, "tensor-core path: step %d needs %d pipeline chunks (max %d)", i, nc, kTcMaxChunks);
  }
  plan.render = pr;
  if (x3 && rn) {
    plan.render.s[n_v].use_lo = 0;
    // ... so the last hidden layer writes no lo copy inside render(): its lo columns are the colour net's aux columns
    plan.render.s[n_v - 1].out_lo = 0;
  }
  plan.vf_full = pr; plan.vf_full.n_steps = n_full; plan.vf_full.render = 0; plan.vf_full.aux_step = -1;
  plan.v_only = pr; plan.v_only.n_steps = n_v; plan.v_only.render = 0; plan.v_only.aux_step = -1;
  set_guards(plan.render); set_guards(plan.vf_full); set_guards(plan.v_only);

  // ---------------- training: stash tensor numbering ----------------
  const int Lr = rn ? rn->n_layers : 1;     // VF-only plans have no colour tensors
  TcStash& S = plan.stash;
  S.n_y = L + Lr - 1;                       // Y_s[0..L-2], Y_FEAT, Y_c[0..Lr-2]
  S.idx_emb0 = S.n_y; S.idx_skip = S.n_y + 1; S.idx_aux = S.n_y + 2; S.idx_d0 = S.n_y + 3;
  S.idx_dcolu = 2 * S.n_y + 3; S.idx_dvu = 2 * S.n_y + 4;
  S.n_tensors = 2 * S.n_y + 5;
  VFN_REQUIRE(S.n_tensors <= kTcMaxStash, "tensor-core path: too many layers for the activation stash");
  for (int i = 0; i < S.n_tensors; ++i) S.slabs[i] = 32;
  S.slabs[S.idx_emb0] = 2 * Epad / 8; S.slabs[S.idx_skip] = 6; S.slabs[S.idx_aux] = 6;
  S.slabs[S.idx_dcolu] = 2; S.slabs[S.idx_dvu] = 2;
  auto yS = [&](int l) { return l; };                  // VF hidden layer l
  const int yFeat = L - 1;
  auto yC = [&](int l) { return L + l; };              // colour hidden layer l
  for (int l = 0; l < L - 1; ++l) plan.render.s[l].stash_out = plan.vf_full.s[l].stash_out = plan.v_only.s[l].stash_out = yS(l);
  plan.render.s[n_v].stash_out = plan.vf_full.s[n_v].stash_out = yFeat;
  for (int l = 0; l < Lr - 1; ++l) plan.render.s[n_full + l].stash_out = yC(l);
  const int D0 = S.idx_d0;
  TcProgram fwd = pr;                          // keep the forward description for post_scale look-ups

  // ---------------- training: the dgrad program of the VF net alone (module call with gradients) ----------------
  {
    pr = TcProgram{};
    layout(pr, 0);
    pr.emb_w = E; pr.emb_pad = Epad; pr.multires = multires; pr.multires_view = multires_view;
    pr.small_w = fwd.small_w; pr.bwd = 2; pr.aux_step = -1;
    ns = 0; woff = 0;
    // VF output layer: A = [d(feature pre-tanh) (main, copied from the stash by the prologue) | d(vector pre-tanh) unit]
    const int c[2] = {0, kColSkip}, k[2] = {256, 16};
    pr.skip_step = ns;
    add(256, 256, 2, c, k, 128, TC_EPI_BWD_RELU, 0xF | (1 << kBarSkip), 0, L - 1, 3, 12, 0, 1.f, 1);
    pr.s[ns - 1].mask_src = yS(L - 2); pr.s[ns - 1].stash_out = D0 + yS(L - 2);
    for (int l = L - 2; l >= 1; --l) {
      const int kk[1] = {round16(vf.out_dim[l])};
      const int Nn = round16(vf.out_dim[l - 1]);
      add(Nn, vf.out_dim[l - 1], 1, main0, kk, 128, TC_EPI_BWD_RELU, 0xF, 0, l, 0, 10, 0, fwd.s[l].post_scale, 1);
      pr.s[ns - 1].mask_src = yS(l - 1); pr.s[ns - 1].stash_out = D0 + yS(l - 1);
    }
    pr.n_steps = ns;
    pr.emb0_last_step = ns - 1;      // VF-only dgrad: reg_free[0] guards the whole main region, free after the last step
    plan.bwd_vf = pr;
    plan.wpack_bwd_vf_bytes = woff;
  }
  if (!rn) return 0;

  // ---------------- training: the dgrad program (weights transposed, no bias, gate with the stash) ----------------
  pr = TcProgram{};
  layout(pr, 0);
  pr.emb_w = E; pr.emb_pad = Epad; pr.multires = multires; pr.multires_view = multires_view;
  pr.small_w = fwd.small_w; pr.bwd = 1;
  ns = 0; woff = 0;
  {
    // colour output layer: A = d(colour pre-sigmoid) as a bf16 (hi, lo) unit in the aux columns
    const int c[1] = {kColAux}, k[1] = {16};
    add(256, 256, 1, c, k, 128, TC_EPI_BWD_RELU, 1 << kBarAuxStatic, 1, Lr - 1, 0, 11, 0, 1.f, 1);
    pr.s[ns - 1].mask_src = yC(Lr - 2); pr.s[ns - 1].stash_out = D0 + yC(Lr - 2);
  }
  for (int l = Lr - 2; l >= 1; --l) {
    add(256, 256, 1, main0, k256, 128, TC_EPI_BWD_RELU, 0xF, 1, l, 0, 10, 0, 1.f, 1);
    pr.s[ns - 1].mask_src = yC(l - 1); pr.s[ns - 1].stash_out = D0 + yC(l - 1);
  }
  // colour layer 0: only the feature columns of its input carry gradient (points / view dirs have none, the normal
  // is detached, rendering_network.py:76-77); gate with tanh'
  add(256, 256, 1, main0, k256, 128, TC_EPI_BWD_TANH, 0xF, 1, 0, 0, 10, fwd.small_w, 1.f, 1);
  pr.s[ns - 1].mask_src = yFeat; pr.s[ns - 1].stash_out = D0 + yFeat;
  {
    // VF output layer: A = [d(feature pre-tanh) (main) | d(vector pre-tanh) (hi, lo) unit in the skip columns]
    const int c[2] = {0, kColSkip}, k[2] = {256, 16};
    pr.skip_step = ns;
    add(256, 256, 2, c, k, 128, TC_EPI_BWD_RELU, 0xF | (1 << kBarSkip), 0, L - 1, 3, 12, 0, 1.f, 1);
    pr.s[ns - 1].mask_src = yS(L - 2); pr.s[ns - 1].stash_out = D0 + yS(L - 2);
  }
  for (int l = L - 2; l >= 1; --l) {
    const int kk[1] = {round16(vf.out_dim[l])};
    const int Nn = round16(vf.out_dim[l - 1]);
    add(Nn, vf.out_dim[l - 1], 1, main0, kk, 128, TC_EPI_BWD_RELU, 0xF, 0, l, 0, 10, 0, fwd.s[l].post_scale, 1);
    pr.s[ns - 1].mask_src = yS(l - 1); pr.s[ns - 1].stash_out = D0 + yS(l - 1);
  }
  pr.n_steps = ns; pr.aux_step = 0;
  plan.bwd = pr;
  plan.wpack_bwd_bytes = woff;
  return 0;
}

int tc_carve(char* base, int64_t& off, int multires, int multires_view, int skip_layer,
             const vfnerf_mlp_desc& vf, const vfnerf_mlp_desc* rn, TcPlan& plan, int64_t n_points, int keep, int x3) {
  VFN_REQUIRE(!(x3 && keep), "precisions bf16x3 / fp16f8 are forward-only: train with precision bf16 or fp32");
  if (int e = build_programs(multires, multires_view, skip_layer, vf, rn, plan, x3)) return e;
  off = align_up(off, 1024);
  plan.wpack = base ? reinterpret_cast<uint8_t*>(base + off) : nullptr;
  off += align_up(plan.wpack_bytes, 1024);
  if (keep) {
    if (!rn) plan.wpack_bwd_bytes = plan.wpack_bwd_vf_bytes;
    plan.wpack_bwd = base ? reinterpret_cast<uint8_t*>(base + off) : nullptr;
    off += align_up(plan.wpack_bwd_bytes, 1024);
    const int64_t tiles = (n_points + kTileM - 1) / kTileM, tiles2 = (tiles + 1) / 2 * 2;
    TcStash& S = plan.stash;
    int64_t so = 0;
    for (int i = 0; i < S.n_tensors; ++i) { S.off[i] = so; so += tiles2 * S.slabs[i] * (kTileM * 16); }
    for (int i = 0; i < S.n_tensors; ++i) { S.gate_off[i] = so; if (i < S.n_y) so += tiles2 * 4 * kTileM * 8; }
    S.bytes = so;
    plan.stash_buf = base ? reinterpret_cast<uint8_t*>(base + off) : nullptr;
    off += align_up(so, 1024);
    plan.gbuf_floats = (int64_t)(S.n_y + 2) * (256 * 320 + 256);
    plan.gbuf = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += align_up(plan.gbuf_floats * 4, 1024);
    plan.d3 = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += align_up(n_points * 6 * 4, 1024);
  }
  return 0;
}

// test support (host only): the chunk records the MMA issuer, the weight producer and the relay lane will read for one
// program of a plan -- records[step][chunk][4], n_chunks[step], step_facts[step][6] = N, K, chunk_k, n_seg, fresh_mask, use_lo
int tc_debug_chunk_table(const TcPlan& plan, int program, uint32_t* records, int* n_chunks, int* step_facts, int* n_steps) {
  const TcProgram& pr = program == 0 ? plan.render : program == 1 ? plan.vf_full : program == 2 ? plan.v_only
                        : program == 3 ? plan.bwd : plan.bwd_vf;
  *n_steps = pr.n_steps;
  for (int si = 0; si < pr.n_steps; ++si) {
    TcChunk rec[kTcMaxChunks];
    n_chunks[si] = tc_chunk_records(pr, si, rec);
    for (int c = 0; c < n_chunks[si]; ++c) {
      uint32_t* o = records + ((size_t)si * kTcMaxChunks + c) * 4;
      o[0] = rec[c].x; o[1] = rec[c].y; o[2] = rec[c].z; o[3] = rec[c].w;
    }
    const TcStep& st = pr.s[si];
    int* f = step_facts + si * 6;
    f[0] = st.N; f[1] = st.K; f[2] = st.chunk_k; f[3] = st.n_seg; f[4] = st.fresh_mask; f[5] = st.use_lo;
  }
  return 0;
}

int tc_prepare(const vfnerf_mlp_desc& vf, const float* vf_arena, const vfnerf_mlp_desc* rn,
               const float* rn_arena, float bn_eps, const TcPlan& plan, cudaStream_t s) {
  const TcProgram& pr = rn ? plan.render : plan.vf_full;
  vfnerf_mlp_desc none{};
  tc_pack_kernel<<<dim3(32, pr.n_steps), 256, 0, s>>>(pr, vf, vf_arena, rn ? *rn : none, rn_arena, bn_eps, plan.wpack);
  VFN_LAUNCH_CHECK();
  tc_pack_dot_kernel<<<2, 256, 0, s>>>(vf, vf_arena, rn ? *rn : none, rn_arena, rn ? 1 : 0, bn_eps,
                                       reinterpret_cast<float*>(plan.wpack + plan.dot_off));
  VFN_LAUNCH_CHECK();
  if (plan.wpack_bwd) {
    const TcProgram& pb = rn ? plan.bwd : plan.bwd_vf;
    tc_pack_kernel<<<dim3(32, pb.n_steps), 256, 0, s>>>(pb, vf, vf_arena, rn ? *rn : none, rn_arena, bn_eps, plan.wpack_bwd);
    VFN_LAUNCH_CHECK();
  }
  return 0;
}


int tc_forward(const TcPlan& plan, int mode, const float* points, const GridSpec* grid, int grid_res,
               int64_t grid_i0, int64_t n, const float* ray_dirs, int samples_per_ray, float* out_v,
               int64_t v_ld, float* out_feat, int64_t feat_ld, float* colors, cudaStream_t s, int64_t stash_tile0) {
  if (n <= 0) return 0;
  TcParams p{};
  const bool is_bwd = mode == TC_MODE_BWD || mode == TC_MODE_VF_BWD;
  const bool stashing = mode == TC_MODE_RENDER_STASH || mode == TC_MODE_V_ONLY_STASH || mode == TC_MODE_VF_FULL_STASH;
  p.prog = (mode == TC_MODE_RENDER || mode == TC_MODE_RENDER_STASH) ? plan.render
           : (mode == TC_MODE_VF_FULL || mode == TC_MODE_VF_FULL_STASH) ? plan.vf_full
           : mode == TC_MODE_BWD ? plan.bwd : mode == TC_MODE_VF_BWD ? plan.bwd_vf : plan.v_only;
  p.wpack = is_bwd ? plan.wpack_bwd : plan.wpack;
  p.dot_off = plan.dot_off;
  VFN_REQUIRE(p.prog.n_steps <= kTcTableSteps, "tc_forward: program of %d steps exceeds the chunk table", p.prog.n_steps);
  for (int si = 0; si < p.prog.n_steps; ++si) p.cnum[si] = tc_chunk_records(p.prog, si, p.ctab + si * kTcMaxChunks);
  if (stashing || is_bwd) {
    VFN_REQUIRE(plan.stash_buf, "tc_forward: this mode needs the training workspace (keep_for_backward)");
    p.stash = plan.stash_buf; p.sinfo = plan.stash;
    if (stash_tile0) {
      // this launch fills tiles [stash_tile0, ...) of every tensor: the kernel numbers its tiles from 0, so shift the bases
      VFN_REQUIRE(stashing, "tc_forward: stash_tile0 only applies to the forward *_STASH modes");
      for (int t = 0; t < p.sinfo.n_tensors; ++t) {
        p.sinfo.off[t] += stash_tile0 * (long long)p.sinfo.slabs[t] * (kTileM * 16);
        p.sinfo.gate_off[t] += stash_tile0 * 4ll * kTileM * 8;
      }
    }
  }
  if (is_bwd) { p.dcol_pre = points; p.dv_pre = ray_dirs; }
  p.points = points; p.use_grid = grid ? 1 : 0;
  if (grid) p.grid = *grid;
  p.grid_res = grid_res; p.grid_i0 = grid_i0; p.n_points = n;
  p.ray_dirs = ray_dirs; p.samples_per_ray = samples_per_ray > 0 ? samples_per_ray : 1;
  p.out_v = out_v; p.v_ld = v_ld; p.out_feat = out_feat; p.feat_ld = feat_ld; p.colors = colors;
#ifdef VFNERF_TC_PROFILE
  // experiment switches and in-kernel cycle counters: profile builds only (the product library reads no environment)
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("VFNERF_TC_DBG"); dbg = e ? atoi(e) : 0; }
  static long long* dbg_buf = nullptr;
  if ((dbg & 64) && !dbg_buf) VFN_CHECK_CUDA(cudaMalloc(&dbg_buf, 256 * sizeof(long long)));
  p.dbg_buf = (dbg & 64) ? dbg_buf : nullptr;
  p.dbg = dbg & 63;
  if (p.dbg_buf) VFN_CHECK_CUDA(cudaMemsetAsync(dbg_buf, 0, 256 * sizeof(long long), s));
#endif
  VFN_REQUIRE(out_v || is_bwd, "tc_forward: out_v is null");
  VFN_REQUIRE((mode != TC_MODE_RENDER && mode != TC_MODE_RENDER_STASH) || (colors && ray_dirs),
              "tc_forward: RENDER mode needs colors and ray_dirs");
  // per-device facts (a process may drive several GPUs): SM count and the opt-in to > 48 KiB of dynamic shared memory
  int dev = 0;
  VFN_CHECK_CUDA(cudaGetDevice(&dev));
  VFN_REQUIRE(dev >= 0 && dev < kMaxDevices, "tc_forward: device ordinal %d unsupported", dev);
  static int num_sms[kMaxDevices] = {0};
  if (num_sms[dev] == 0) {
    VFN_CHECK_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VFN_CHECK_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VFN_CHECK_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VFN_CHECK_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VFN_CHECK_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VFN_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms[dev], cudaDevAttrMultiProcessorCount, dev));
  }
  const int64_t tiles = (n + kTileM - 1) / kTileM;
  const int64_t pairs = (tiles + 1) / 2;
  const int grid_x = 2 * (int)std::min<int64_t>(pairs, (int64_t)(num_sms[dev] / 2));
  if (p.prog.x3) {
    VFN_REQUIRE(!is_bwd && !stashing, "tc_forward: precisions bf16x3 / fp16f8 are forward-only");
    if (p.prog.f8) mlp_tc_kernel<false, false, true, true><<<grid_x, kTcThreads, tc_smem_bytes<true>(), s>>>(p);
    else mlp_tc_kernel<false, false, true><<<grid_x, kTcThreads, tc_smem_bytes<true>(), s>>>(p);
  } else {
    const size_t smem = tc_smem_bytes<false>();
    if (is_bwd) mlp_tc_kernel<true, true, false><<<grid_x, kTcThreads, smem, s>>>(p);
    else if (stashing) mlp_tc_kernel<false, true, false><<<grid_x, kTcThreads, smem, s>>>(p);
    else mlp_tc_kernel<false, false, false><<<grid_x, kTcThreads, smem, s>>>(p);
  }
  VFN_LAUNCH_CHECK();
#ifdef VFNERF_TC_PROFILE
  if (p.dbg_buf) {
    long long h[256];
    VFN_CHECK_CUDA(cudaStreamSynchronize(s));
    VFN_CHECK_CUDA(cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost));
    const long long tc = (pairs + grid_x / 2 - 1) / (grid_x / 2);
    fprintf(stderr, "[tc dbg] tiles/CTA %lld steps %d | MMA thread cycles/tile: wait_grp %lld wait_full %lld issue %lld | "
            "epi h0: acc %lld ld %lld math %lld sig %lld other %lld | epi h1: acc %lld ld %lld math %lld sig %lld other %lld\n",
            tc, p.prog.n_steps, h[0] / tc, h[1] / tc, h[2] / tc, h[8] / tc, h[9] / tc, h[10] / tc, h[11] / tc, h[12] / tc,
            h[16] / tc, h[17] / tc, h[18] / tc, h[19] / tc, h[20] / tc);
    const long long base = h[64];
    fprintf(stderr, "[tc launch] CTA 0 cycles from entry: set-up done %lld, first MMA %lld, last step issued %lld, exit %lld\n",
            h[241] - h[240], h[242] - h[240], h[243] - h[240], h[244] - h[240]);
    for (int ci = 0; ci < 10 && h[168 + 4 * ci]; ++ci)
      fprintf(stderr, "[tc chunks] step 2 chunk %d: top->grp_done %6lld  full_done +%4lld  mmas issued +%4lld  commit +%4lld\n", ci,
              h[168 + 4 * ci] - base, h[168 + 4 * ci + 1] - h[168 + 4 * ci], h[168 + 4 * ci + 3] - h[168 + 4 * ci + 1],
              h[168 + 4 * ci + 2] - h[168 + 4 * ci + 3]);


    for (int si = 0; si < p.prog.n_steps && base; ++si) {
      const long long* e = h + 64 + si * 8;
      fprintf(stderr, "[tc timeline] step %2d: first_mma %6lld commit %6lld | acc_seen %6lld first_arrive %6lld last_arrive %6lld | "
              "first group of warp 2: region free %6lld, tcgen05.ld done %6lld, converted + stored %6lld\n", si, e[0] - base, e[1] - base,
              e[2] - base, e[3] ? e[3] - base : 0, e[4] ? e[4] - base : 0, e[5] ? e[5] - base : 0, e[6] ? e[6] - base : 0,
              e[7] ? e[7] - base : 0);
    }
  }
#endif
  return 0;
}

}  // namespace vfn
