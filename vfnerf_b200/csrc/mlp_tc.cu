// Fused tensor-core MLP chain (VFNERF_PREC_BF16): the VF MLP (positional encoding -> 9 Linear+BN+ReLU
// layers with skip -> tanh) and, in RENDER mode, the colour MLP (5 layers -> sigmoid) evaluated for a
// tile of 128 points entirely on chip.  SURVEY.md §8 rows a3 + a8.
//
// Per CTA (192 threads, 2 CTAs per SM so one CTA's epilogue overlaps the other's MMAs):
//   warp 0      weight producer: one lane streams pre-tiled bf16 weight chunks (<= 16 KiB) from L2 into a
//               shared-memory ring with cp.async.bulk, completion on mbarriers            (SASS UBLKCP)
//   warp 1      MMA issuer: one lane issues tcgen05.mma (M=128, N<=256, K=16, bf16 x bf16 -> fp32 in TMEM),
//               tcgen05.commit releases ring slots and publishes the accumulator          (SASS UTCHMMA)
//   warps 2..5  epilogue: tcgen05.ld the 128 x N fp32 accumulator (one row per thread), add the folded
//               BatchNorm shift (the scale is folded into the weights), ReLU / tanh / sigmoid, convert to
//               bf16 and write the NEXT layer's A operand straight into the activation tile in the K-slab
//               UMMA layout (tc_common.cuh) -- activations never leave the SM.
// Activation tile: 128 x 256 bf16 (64 KiB) [+ 48 aux columns holding the colour net's small inputs].
// Weights: 1.1 MB (VF) + 0.6 MB (colour) bf16, L2 resident, re-streamed per tile: 128 KiB per 256x256
// layer per tile = 64 B/clk/SM at full tensor rate (DESIGN.md discusses this limiter).
#include "mlp_tc.cuh"
#include "tc_common.cuh"

namespace vfn {
using namespace tc;

constexpr int kTileM = 128;
constexpr int kStageBytes = 16384;
constexpr int kTcThreads = 192;
constexpr int kAccCols = 256;
constexpr float kInvSqrt2 = 0.70710678118654752f;

struct TcParams {
  TcProgram prog;
  const uint8_t* wpack;
  const float* affine;
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
};

// ---------------------------------------------------------------------------------------------
// weight packing: fp32 Linear weights (+ folded BatchNorm scale) -> bf16 K-slab images, one per step
// ---------------------------------------------------------------------------------------------
__global__ void tc_pack_kernel(TcProgram prog, vfnerf_mlp_desc vf, const float* __restrict__ vf_arena,
                               vfnerf_mlp_desc rn, const float* __restrict__ rn_arena, float eps,
                               uint8_t* __restrict__ wpack, float* __restrict__ affine) {
  const TcStep st = prog.s[blockIdx.y];
  const vfnerf_mlp_desc& d = st.net == 0 ? vf : rn;
  const float* arena = st.net == 0 ? vf_arena : rn_arena;
  const int l = st.layer, in_dim = d.in_dim[l];
  const bool bn = d.gamma_off[l] >= 0;
  const int total = st.N * st.K;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int n = e / st.K, k = e - n * st.K;
    float w = 0.f;
    if (n < st.n_valid) {
      int src = -1;
      if (st.colmap == 0) src = k;
      else if (st.colmap == 1) src = (k < st.dup_w) ? k : k - st.dup_w;
      else src = (k < 256) ? st.dup_w + k : k - 256;
      if (st.colmap == 2 && k >= 256 && src >= st.dup_w) src = -1;
      if (src >= 0 && src < in_dim) {
        const int row = st.row0 + n;
        float sc = 1.f;
        if (bn) sc = arena[d.gamma_off[l] + row] / sqrtf(arena[d.var_off[l] + row] + eps);
        w = arena[d.w_off[l] + (int64_t)row * in_dim + src] * sc * st.post_scale;
      }
    }
    const int64_t off = st.w_off + (int64_t)(k / st.chunk_k) * st.N * st.chunk_k * 2 +
                        (int64_t)((k % st.chunk_k) / 8) * st.N * 16 + n * 16 + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(wpack + off) = __float2bfloat16(w);
  }
  if (blockIdx.x == 0) {
    for (int n = threadIdx.x; n < 256; n += blockDim.x) {
      float sh = 0.f;
      if (n < st.n_valid) {
        const int row = st.row0 + n;
        const float b = arena[d.b_off[l] + row];
        sh = b;
        if (bn) {
          const float sc = arena[d.gamma_off[l] + row] / sqrtf(arena[d.var_off[l] + row] + eps);
          sh = arena[d.beta_off[l] + row] + (b - arena[d.mean_off[l] + row]) * sc;
        }
        sh *= st.post_scale;
      }
      affine[st.aff_off + n] = sh;
    }
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

// 8 consecutive columns of one row -> one 16-byte store into slab `slab` of the activation tile
__device__ __forceinline__ void store_slab(uint8_t* s_act, int slab, int row, const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(s_act + slab * (kTileM * 16) + row * 16) = u;
}

// positional encoding of embedder.py:11-37 into e[0 .. 3+6*L)
__device__ __forceinline__ void embed3(const float* p, int L, float* e) {
  e[0] = p[0]; e[1] = p[1]; e[2] = p[2];
  float f = 1.f;
  for (int k = 0; k < L; ++k) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float sv, cv;
      sincosf(p[c] * f, &sv, &cv);
      e[3 + 6 * k + c] = sv;
      e[6 + 6 * k + c] = cv;
    }
    f *= 2.f;
  }
}

// ---------------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------------
template <int NSTAGE>
__global__ void __launch_bounds__(kTcThreads, 2) mlp_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcProgram& prog = p.prog;
  uint8_t* s_act = smem;
  uint8_t* s_stage = smem + prog.act_cols * (kTileM * 2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + NSTAGE * kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + NSTAGE;
  uint64_t* acc_full = bars + 2 * NSTAGE;
  uint64_t* act_ready = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(act_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(act_ready, 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kAccCols>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const long long num_tiles = (p.n_points + kTileM - 1) / kTileM;

  if (warp == 0) {
    // ===================== weight producer =====================
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int si = 0; si < prog.n_steps; ++si) {
          const TcStep& st = prog.s[si];
          const uint8_t* src = p.wpack + st.w_off;
          const int full_bytes = st.N * st.chunk_k * 2;
          for (int c = 0; c < st.n_chunks; ++c) {
            const int kc = min(st.chunk_k, st.K - c * st.chunk_k);
            const uint32_t bytes = (uint32_t)(st.N * kc * 2);
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], bytes);
            bulk_g2s(s_stage + stage * kStageBytes, src + (int64_t)c * full_bytes, bytes, &full[stage]);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0, phase = 0, par_act = 0;
      const uint32_t act_base = smem_u32(s_act), stage_base = smem_u32(s_stage);
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int si = 0; si < prog.n_steps; ++si) {
          const TcStep& st = prog.s[si];
          const uint32_t idesc = make_idesc_bf16(kTileM, st.N);
          const uint32_t b_lbo = st.N * 16;
          mbar_wait(act_ready, par_act);
          par_act ^= 1;
          tc_fence_after_sync();
          uint32_t accumulate = 0;
          for (int c = 0; c < st.n_chunks; ++c) {
            const int k0 = c * st.chunk_k;
            const int kc = min(st.chunk_k, st.K - k0);
            mbar_wait(&full[stage], phase);
            tc_fence_after_sync();
            for (int kk = 0; kk < kc; kk += 16) {
              const uint32_t a_addr = act_base + ((st.a_col0 + k0 + kk) >> 3) * (kTileM * 16);
              const uint32_t b_addr = stage_base + stage * kStageBytes + (kk >> 3) * b_lbo;
              umma_bf16(tmem, make_smem_desc(a_addr, kTileM * 16, 128), make_smem_desc(b_addr, b_lbo, 128), idesc,
                        accumulate);
              accumulate = 1;
            }
            umma_commit(&empty[stage]);          // ring slot reusable once these MMAs have read it
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
          umma_commit(acc_full);                 // accumulator complete -> epilogue
        }
      }
    }
  } else {
    // ===================== epilogue warps (one row per thread) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    const int E = prog.emb_w, Epad = prog.emb_pad;
    int it = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const long long pi = tile * kTileM + row;
      const bool valid = pi < p.n_points;
      float pt[3] = {0.f, 0.f, 0.f};
      if (valid) {
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
      // ---- prologue: positional encoding as a bf16 hi/lo pair -> A columns [0, 2*Epad)
      float emb[48];
      embed3(pt, prog.multires, emb);
      for (int i = E; i < 48; ++i) emb[i] = 0.f;
      {
        const int nsl = Epad >> 3;
        for (int sl = 0; sl < nsl; ++sl) {
          float hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float v = emb[sl * 8 + j];
            hi[j] = __bfloat162float(__float2bfloat16(v));
            lo[j] = v - hi[j];
          }
          store_slab(s_act, sl, row, hi);
          store_slab(s_act, nsl + sl, row, lo);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(act_ready);

      for (int si = 0; si < prog.n_steps; ++si) {
        const TcStep& st = prog.s[si];
        const float4* sh4 = reinterpret_cast<const float4*>(p.affine + st.aff_off);
        mbar_wait(acc_full, it & 1);
        ++it;
        tc_fence_after_sync();
        if (st.epi == TC_EPI_RELU || st.epi == TC_EPI_RELU_SKIPFILL) {
          const bool fill = st.epi == TC_EPI_RELU_SKIPFILL;
          for (int g = 0; g < 8; ++g) {
            const int c0 = g * 32;
            if (c0 >= st.N && !fill) break;
            float f[32];
            if (c0 < st.N) {
              uint32_t v[32];
              tmem_ld32(t_lane + c0, v);
              tmem_ld_wait();
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 s4 = __ldg(sh4 + (c0 >> 2) + j4);
                f[4 * j4 + 0] = fmaxf(__uint_as_float(v[4 * j4 + 0]) + s4.x, 0.f);
                f[4 * j4 + 1] = fmaxf(__uint_as_float(v[4 * j4 + 1]) + s4.y, 0.f);
                f[4 * j4 + 2] = fmaxf(__uint_as_float(v[4 * j4 + 2]) + s4.z, 0.f);
                f[4 * j4 + 3] = fmaxf(__uint_as_float(v[4 * j4 + 3]) + s4.w, 0.f);
              }
            }
            if (fill) {
              // skip connection: columns >= n_valid of the next layer's input are emb / sqrt(2)
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int col = c0 + j;
                if (col >= st.n_valid) f[j] = (col - st.n_valid < E) ? emb[col - st.n_valid] * kInvSqrt2 : 0.f;
              }
            }
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) store_slab(s_act, (c0 >> 3) + sl, row, f + 8 * sl);
          }
          fence_proxy_async_smem();
        } else if (st.epi == TC_EPI_V) {
          uint32_t v[16];
          tmem_ld16(t_lane, v);
          tmem_ld_wait();
          float nv[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) nv[j] = tanhf(__uint_as_float(v[j]) + __ldg(p.affine + st.aff_off + j));
          if (valid) {
#pragma unroll
            for (int j = 0; j < 3; ++j) p.out_v[pi * p.v_ld + j] = nv[j];
          }
          if (prog.act_cols > 256) {
            // colour-net small inputs [p(3), embed(view dir)(3+6*Lv), n(3), 0...] -> aux columns 256..303
            float a[48];
#pragma unroll
            for (int j = 0; j < 48; ++j) a[j] = 0.f;
            a[0] = pt[0]; a[1] = pt[1]; a[2] = pt[2];
            float d[3] = {0.f, 0.f, 0.f};
            if (valid) {
              const long long r = pi / p.samples_per_ray;
              d[0] = __ldg(p.ray_dirs + 3 * r); d[1] = __ldg(p.ray_dirs + 3 * r + 1); d[2] = __ldg(p.ray_dirs + 3 * r + 2);
            }
            embed3(d, prog.multires_view, a + 3);
            const int ev = 3 + 6 * prog.multires_view;
            a[3 + ev] = nv[0]; a[4 + ev] = nv[1]; a[5 + ev] = nv[2];
#pragma unroll
            for (int sl = 0; sl < 6; ++sl) store_slab(s_act, 32 + sl, row, a + 8 * sl);
            fence_proxy_async_smem();
          }
        } else if (st.epi == TC_EPI_FEAT) {
          const bool to_act = prog.act_cols > 256;
          for (int g = 0; g < 8; ++g) {
            const int c0 = g * 32;
            uint32_t v[32];
            tmem_ld32(t_lane + c0, v);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 s4 = __ldg(sh4 + (c0 >> 2) + j4);
              f[4 * j4 + 0] = tanh_fast(__uint_as_float(v[4 * j4 + 0]) + s4.x);
              f[4 * j4 + 1] = tanh_fast(__uint_as_float(v[4 * j4 + 1]) + s4.y);
              f[4 * j4 + 2] = tanh_fast(__uint_as_float(v[4 * j4 + 2]) + s4.z);
              f[4 * j4 + 3] = tanh_fast(__uint_as_float(v[4 * j4 + 3]) + s4.w);
            }
            if (to_act) {
#pragma unroll
              for (int sl = 0; sl < 4; ++sl) store_slab(s_act, (c0 >> 3) + sl, row, f + 8 * sl);
            }
            if (p.out_feat && valid) {
              float* o = p.out_feat + pi * p.feat_ld + c0;
#pragma unroll
              for (int j = 0; j < 32; ++j) o[j] = f[j];
            }
          }
          if (to_act) fence_proxy_async_smem();
        } else {  // TC_EPI_RGB
          uint32_t v[16];
          tmem_ld16(t_lane, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const float x = __uint_as_float(v[j]) + __ldg(p.affine + st.aff_off + j);
              p.colors[3 * pi + j] = 1.f / (1.f + expf(-x));
            }
          }
        }
        // accumulator drained (and the next A operand written): release the MMA issuer
        tc_fence_before_sync();
        if (si + 1 < prog.n_steps) mbar_arrive(act_ready);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kAccCols>(tmem);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int round16(int x) { return (x + 15) / 16 * 16; }

static int build_programs(int multires, int multires_view, int skip_layer, const vfnerf_mlp_desc& vf,
                          const vfnerf_mlp_desc* rn, TcPlan& plan) {
  const int E = 3 + 6 * multires, Epad = round16(E);
  const int L = vf.n_layers;
  VFN_REQUIRE(Epad <= 48, "tensor-core path: embedding width %d > 48", E);
  VFN_REQUIRE(L >= 3 && L + 1 + (rn ? rn->n_layers : 0) <= kTcMaxSteps, "tensor-core path: too many layers");
  for (int l = 0; l < L; ++l) {
    const int want_in = (l == 0) ? E : 256;
    const int want_out = (l == L - 1) ? 259 : ((l + 1 == skip_layer) ? 256 - E : 256);
    VFN_REQUIRE(vf.in_dim[l] == want_in && vf.out_dim[l] == want_out,
                "tensor-core path supports the shipped 256-wide VF net (layer %d is %d->%d); use precision fp32",
                l, vf.in_dim[l], vf.out_dim[l]);
  }
  TcProgram pr{};
  pr.emb_w = E; pr.emb_pad = Epad; pr.multires = multires; pr.multires_view = multires_view;
  pr.small_w = 3 + (3 + 6 * multires_view) + 3;
  int ns = 0;
  long long woff = 0;
  auto add = [&](int K, int N, int n_valid, int chunk_k, int epi, int net, int layer, int row0, int colmap,
                 int dup_w, float post) {
    TcStep& s = pr.s[ns];
    s.K = K; s.a_col0 = 0; s.N = N; s.n_valid = n_valid; s.chunk_k = chunk_k;
    s.n_chunks = (K + chunk_k - 1) / chunk_k; s.epi = epi; s.aff_off = ns * 256; s.w_off = woff;
    s.net = net; s.layer = layer; s.row0 = row0; s.colmap = colmap; s.dup_w = dup_w; s.post_scale = post;
    woff += (long long)align_up((int64_t)N * K * 2, 128);
    ++ns;
  };
  for (int l = 0; l < L - 1; ++l) {
    const bool pre_skip = (l + 1 == skip_layer);
    add(l == 0 ? 2 * Epad : 256, round16(vf.out_dim[l]), vf.out_dim[l], 32,
        pre_skip ? TC_EPI_RELU_SKIPFILL : TC_EPI_RELU, 0, l, 0, l == 0 ? 1 : 0, Epad, pre_skip ? kInvSqrt2 : 1.f);
  }
  add(256, 16, 3, 256, TC_EPI_V, 0, L - 1, 0, 0, 0, 1.f);
  const int n_v = ns;
  add(256, 256, 256, 32, TC_EPI_FEAT, 0, L - 1, 3, 0, 0, 1.f);
  const int n_full = ns;
  if (rn) {
    const int Lr = rn->n_layers;
    VFN_REQUIRE(pr.small_w <= 48, "tensor-core path: colour-net small inputs %d > 48", pr.small_w);
    for (int l = 0; l < Lr; ++l) {
      const int want_in = (l == 0) ? pr.small_w + 256 : 256;
      const int want_out = (l == Lr - 1) ? 3 : 256;
      VFN_REQUIRE(rn->in_dim[l] == want_in && rn->out_dim[l] == want_out,
                  "tensor-core path supports the shipped 256-wide colour net (layer %d is %d->%d); use precision fp32",
                  l, rn->in_dim[l], rn->out_dim[l]);
    }
    add(256 + 48, 256, 256, 32, TC_EPI_RELU, 1, 0, 0, 2, pr.small_w, 1.f);
    for (int l = 1; l < Lr - 1; ++l) add(256, 256, 256, 32, TC_EPI_RELU, 1, l, 0, 0, 0, 1.f);
    add(256, 16, 3, 256, TC_EPI_RGB, 1, Lr - 1, 0, 0, 0, 1.f);
  }
  plan.wpack_bytes = woff;
  pr.n_steps = ns; pr.act_cols = 304; pr.n_stages = 2;
  plan.render = pr;
  plan.vf_full = pr; plan.vf_full.n_steps = n_full; plan.vf_full.act_cols = 256; plan.vf_full.n_stages = 3;
  plan.v_only = pr; plan.v_only.n_steps = n_v; plan.v_only.act_cols = 256; plan.v_only.n_stages = 3;
  return 0;
}

int tc_carve(char* base, int64_t& off, int multires, int multires_view, int skip_layer,
             const vfnerf_mlp_desc& vf, const vfnerf_mlp_desc* rn, TcPlan& plan) {
  if (int e = build_programs(multires, multires_view, skip_layer, vf, rn, plan)) return e;
  off = align_up(off, 1024);
  plan.wpack = base ? reinterpret_cast<uint8_t*>(base + off) : nullptr;
  off += align_up(plan.wpack_bytes, 1024);
  plan.affine = base ? reinterpret_cast<float*>(base + off) : nullptr;
  off += kTcMaxSteps * 256 * sizeof(float);
  return 0;
}

int tc_prepare(const vfnerf_mlp_desc& vf, const float* vf_arena, const vfnerf_mlp_desc* rn,
               const float* rn_arena, float bn_eps, const TcPlan& plan, cudaStream_t s) {
  const TcProgram& pr = rn ? plan.render : plan.vf_full;
  vfnerf_mlp_desc none{};
  tc_pack_kernel<<<dim3(32, pr.n_steps), 256, 0, s>>>(pr, vf, vf_arena, rn ? *rn : none, rn_arena, bn_eps,
                                                       plan.wpack, plan.affine);
  VFN_LAUNCH_CHECK();
  return 0;
}

static int g_num_sms = 0;

int tc_forward(const TcPlan& plan, int mode, const float* points, const GridSpec* grid, int grid_res,
               int64_t grid_i0, int64_t n, const float* ray_dirs, int samples_per_ray, float* out_v,
               int64_t v_ld, float* out_feat, int64_t feat_ld, float* colors, cudaStream_t s) {
  if (n <= 0) return 0;
  TcParams p{};
  p.prog = mode == TC_MODE_RENDER ? plan.render : (mode == TC_MODE_VF_FULL ? plan.vf_full : plan.v_only);
  p.wpack = plan.wpack; p.affine = plan.affine;
  p.points = points; p.use_grid = grid ? 1 : 0;
  if (grid) p.grid = *grid;
  p.grid_res = grid_res; p.grid_i0 = grid_i0; p.n_points = n;
  p.ray_dirs = ray_dirs; p.samples_per_ray = samples_per_ray > 0 ? samples_per_ray : 1;
  p.out_v = out_v; p.v_ld = v_ld; p.out_feat = out_feat; p.feat_ld = feat_ld; p.colors = colors;
  VFN_REQUIRE(out_v, "tc_forward: out_v is null");
  VFN_REQUIRE(mode != TC_MODE_RENDER || (colors && ray_dirs), "tc_forward: RENDER mode needs colors and ray_dirs");
  if (g_num_sms == 0) {
    int dev = 0;
    VFN_CHECK_CUDA(cudaGetDevice(&dev));
    VFN_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int64_t tiles = (n + kTileM - 1) / kTileM;
  const int grid_x = (int)std::min<int64_t>(tiles, 2LL * g_num_sms);
  const size_t smem = (size_t)p.prog.act_cols * kTileM * 2 + (size_t)p.prog.n_stages * kStageBytes + 128;
  if (p.prog.n_stages == 3) {
    static bool attr3 = false;
    if (!attr3) {
      VFN_CHECK_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 116 * 1024));
      attr3 = true;
    }
    mlp_tc_kernel<3><<<grid_x, kTcThreads, smem, s>>>(p);
  } else {
    static bool attr2 = false;
    if (!attr2) {
      VFN_CHECK_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 116 * 1024));
      attr2 = true;
    }
    mlp_tc_kernel<2><<<grid_x, kTcThreads, smem, s>>>(p);
  }
  VFN_LAUNCH_CHECK();
  return 0;
}

}  // namespace vfn
