// Training backward on the tensor cores (VFNERF_PREC_BF16), SURVEY.md §8 row a12.
//
//   1. d(colour pre-sigmoid), d(vector pre-tanh)                      elementwise (fp32)
//   2. fused dgrad chain: mlp_tc_kernel<true> (mlp_tc.cu) walks the layers backwards with transposed weight
//      images, gates with the stashed forward activations and stashes dL/d(pre-activation) of every layer (bf16)
//   3. weight gradients: wgrad_tc_kernel below -- G[layer] = dY^T X over all points.  Both operands are read from the
//      activation stash exactly as stored (tile-major K-slab bf16) through MN-major UMMA descriptors, i.e. the
//      transposition needed by dY^T X is done by the tensor core's operand fetch, not by a memory pass.  HBM-bound:
//      one pass over the stash (1 KB per point per layer).
//   4. the 3-row gradients of the two output layers ride along as 16-channel (hi, lo) unit tasks of the same kernel;
//      every column sum of dY (bias / BatchNorm terms) is reduced from the shared-memory tiles it streams anyway
//   5. finalize: map G back to the reference's parameter layout and apply the BatchNorm-eval chain rule.
#include "mlp_tc.cuh"
#include "tc_common.cuh"

namespace vfn {
using namespace tc;

constexpr int kTileM = 128;
constexpr int kGLd = 320;                          // columns of one layer's G block: [0,256) main input, [256,304) side input
constexpr int kGSlot = 256 * kGLd + 256;           // + 256 column sums
constexpr int kWgDGran = 32768;                    // dY granule: 16 slabs = 128 channels of a 128-point tile = one M = 128 half
constexpr int kWgDSlots = 3;
constexpr int kWgXBytes = 65536;                   // X slot: one 128-point tile of a 256-channel tensor
constexpr int kWgXSlots = 2;
constexpr int kWgThreads = 192;                    // warp 0: loader lane, warp 1: MMA lane, warps 2..5: final epilogue

struct WgTask {
  long long d_off, x_off;      // byte offsets of the dY / X tensors inside the stash
  int d_slabs, x_slabs;        // slabs per tile of each tensor
  int x_slab0, n_xslabs;       // slab range of X used (N = 8 * n_xslabs <= 256)
  int g_off;                   // float offset of the destination block (row stride kGLd) + column offset
  int colsum_off;              // float offset of the 256 column sums of dY (bias / BatchNorm terms); -1: not this task
  int m_rows;                  // rows of G that are meaningful (128 per half; 16 for the (hi, lo) unit tasks)
  int cta0, nctas;             // CTAs [cta0, cta0 + nctas) work on this task, tiles split evenly
};
struct WgParams {
  const uint8_t* stash;
  float* gbuf;
  long long n_tiles;
  int n_tasks;
  WgTask t[24];
};

// Shared memory: a ring of three 32 KB dY granules and a ring of two 64 KB X tiles.  The kernel streams the stash once and
// is bound by HBM latency x bytes in flight: with the former ring of three 64 KB slots only ONE slot could be loading while
// a tile (two slots) was being consumed -- 64 KB in flight per SM, 4.6 TB/s.  Now everything tile t+1 needs first (X and
// the first dY half, 96 KB) is in flight while tile t is in the tensor core, and the second dY half follows as soon as
// the first half of tile t has been consumed.
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* dring = smem;
  uint8_t* xring = smem + kWgDSlots * kWgDGran;
  uint64_t* bars = reinterpret_cast<uint64_t*>(xring + kWgXSlots * kWgXBytes);
  uint64_t* dfull = bars;
  uint64_t* dempty = dfull + kWgDSlots;
  uint64_t* xfull = dempty + kWgDSlots;
  uint64_t* xempty = xfull + kWgXSlots;
  uint64_t* done = xempty + kWgXSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // which task / tile range
  int ti = 0;
  while (ti + 1 < p.n_tasks && (int)blockIdx.x >= p.t[ti + 1].cta0) ++ti;
  const WgTask& T = p.t[ti];
  const int local = blockIdx.x - T.cta0;
  const long long per = (p.n_tiles + T.nctas - 1) / T.nctas;
  const long long t0 = (long long)local * per, t1 = min(p.n_tiles, t0 + per);
  if (threadIdx.x == 0) {
    // a dY granule is free when the MMAs that read it have completed and, for tasks that also reduce dY over the
    // points, when the four reduction warps are done with it
    for (int i = 0; i < kWgDSlots; ++i) { mbar_init(&dfull[i], 1); mbar_init(&dempty[i], T.colsum_off >= 0 ? 5 : 1); }
    for (int i = 0; i < kWgXSlots; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const int N = 8 * T.n_xslabs;
  const int halves = (T.d_slabs + 15) / 16;
  if (warp == 0 && lane == 0) {
    // loader: per tile dY half 0, the used slabs of X, dY half 1
    int ds = 0, dph = 0, xs = 0, xph = 0;
    const uint32_t x_bytes = (uint32_t)T.n_xslabs * 2048u;
    for (long long t = t0; t < t1; ++t) {
      const uint8_t* dsrc = p.stash + T.d_off + t * (long long)T.d_slabs * 2048;
      for (int h = 0; h < halves; ++h) {
        const uint32_t bytes = (uint32_t)min(16, T.d_slabs - 16 * h) * 2048u;
        mbar_wait(&dempty[ds], dph ^ 1);
        mbar_arrive_expect_tx(&dfull[ds], bytes);
        bulk_g2s(dring + ds * kWgDGran, dsrc + (long long)h * kWgDGran, bytes, &dfull[ds]);
        if (++ds == kWgDSlots) { ds = 0; dph ^= 1; }
        if (h == 0) {
          mbar_wait(&xempty[xs], xph ^ 1);
          mbar_arrive_expect_tx(&xfull[xs], x_bytes);
          bulk_g2s(xring + xs * kWgXBytes, p.stash + T.x_off + (t * (long long)T.x_slabs + T.x_slab0) * 2048, x_bytes, &xfull[xs]);
          if (++xs == kWgXSlots) { xs = 0; xph ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // MMA issuer: G[half] (128 channels x N) += dY_tile[half]^T (channels x points) * X_tile (points x N)
    // both operands MN-major: 16-byte unit = 8 consecutive channels of one point, units of consecutive points are
    // 16 bytes apart (LBO field = 128 B per group of 8 points), channel groups are one slab (2048 B) apart (SBO field)
    int ds = 0, dph = 0, xs = 0, xph = 0;
    const uint32_t idesc = make_idesc_bf16(128, N) | (1u << 15) | (1u << 16);
    const uint32_t desc_hi = (2048u >> 4) | (1u << 14);
    const uint32_t lo_hi = (128u >> 4) << 16;
    const uint32_t dbase = smem_u32(dring), xbase = smem_u32(xring);
    uint32_t acc_started = 0;
    for (long long t = t0; t < t1; ++t) {
      const int sx = xs;
      mbar_wait(&xfull[xs], xph);
      if (++xs == kWgXSlots) { xs = 0; xph ^= 1; }
      const uint32_t b0 = (((xbase + sx * kWgXBytes) >> 4) & 0x3FFF) | lo_hi;
      for (int h = 0; h < halves; ++h) {
        const int sd = ds;
        mbar_wait(&dfull[ds], dph);
        if (++ds == kWgDSlots) { ds = 0; dph ^= 1; }
        tc_fence_after_sync();
        const uint32_t a0 = (((dbase + sd * kWgDGran) >> 4) & 0x3FFF) | lo_hi;
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k)      // 16 points per MMA: start address advances 16 points x 16 B
          umma_bf16_split(tmem + h * 256, a0 + k * 16u, desc_hi, b0 + k * 16u, desc_hi, idesc, acc_started | k);
        umma_commit(&dempty[sd]);
      }
      acc_started = 1;
      umma_commit(&xempty[sx]);
    }
    umma_commit(done);
  } else if (warp >= 2) {
    const int w = warp - 2;
    if (T.colsum_off >= 0) {
      // column sums of dY straight from the shared-memory granules the tensor core is reading.  Every warp visits EVERY
      // granule in ring order (a warp that skipped uses of a slot could mistake an older completed phase of its barrier
      // for the one it waits for) and reduces 4 of its 16 slabs: warp w owns slabs 4w..4w+3 of each half; lane l owns
      // points l, l+32, l+64, l+96 (each LDS.128 of a warp covers 512 contiguous bytes).  Partial sums stay in registers
      // across this CTA's tiles.  Rows past the end of the batch are exact zeros in the gradient stash.
      float acc[2][4][8];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[h][a][j] = 0.f;
      int ds = 0, dph = 0;
      for (long long t = t0; t < t1; ++t) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h < halves) {
            mbar_wait(&dfull[ds], dph);
            const uint8_t* dt = dring + ds * kWgDGran;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              const int sl = 4 * w + a;                     // slab inside the granule
              if (16 * h + sl < T.d_slabs) {
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                  const uint4 u = *reinterpret_cast<const uint4*>(dt + sl * 2048 + (lane + 32 * rr) * 16);
                  const uint32_t x[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    acc[h][a][2 * j] += __uint_as_float(x[j] << 16);
                    acc[h][a][2 * j + 1] += __uint_as_float(x[j] & 0xFFFF0000u);
                  }
                }
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&dempty[ds]);
            if (++ds == kWgDSlots) { ds = 0; dph ^= 1; }
          }
        }
      }
      if (t1 > t0) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float v = warp_sum(acc[h][a][j]);
              const int slab = 16 * h + 4 * w + a;
              if (lane == 0 && slab < T.d_slabs) atomicAdd(p.gbuf + T.colsum_off + slab * 8 + j, v);
            }
      }
    }
    // final epilogue: TMEM -> fp32 atomics into this layer's G block (a handful of CTAs share a block)
    mbar_wait(done, 0);
    tc_fence_after_sync();
    if (t1 > t0) {
      const int q = warp & 3, row = q * 32 + lane;
      for (int h = 0; h < halves; ++h) {
        if (q * 32 >= T.m_rows) continue;          // unit tasks: only the first 16 accumulator rows mean anything
        float* g = p.gbuf + T.g_off + (long long)(h * 128 + row) * kGLd;
        for (int c0 = 0; c0 < N; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(tmem + h * 256 + (((uint32_t)(q * 32)) << 16) + c0, v);
          tmem_ld_wait();
          if (row < T.m_rows) {
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(g + c0 + j, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ---------------------------------------------------------------------------------------------
// column sums of a [n,3] fp32 tensor -> out[0..2]
__global__ void sum3_kernel(const float* __restrict__ d3, long long n, float* __restrict__ out) {
  float a[3] = {0.f, 0.f, 0.f};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    for (int j = 0; j < 3; ++j) a[j] += d3[3 * i + j];
  for (int j = 0; j < 3; ++j) {
    a[j] = warp_sum(a[j]);
    if ((threadIdx.x & 31) == 0) atomicAdd(out + j, a[j]);
  }
}

// ---------------------------------------------------------------------------------------------
// finalize: G (gradient wrt the folded weights, in the packed column order) + column sums -> gradient arena
//   y = (x W^T + b - mean) * gamma * istd + beta (then * post):  W_eff = W * gamma*istd*post
//   dW = gamma*istd*post * G,  db = gamma*istd*post * s,  dbeta = post * s,
//   dgamma = istd*post * (rowdot(W, G) + (b - mean) * s)
// ---------------------------------------------------------------------------------------------
enum { FK_IDENT = 0, FK_EMB = 1, FK_SKIP = 2, FK_C0 = 3, FK_HILO = 4 };
struct FinLayer {
  int net, layer, kind, split, epad, ev;   // kind-specific: split = columns fed by the previous layer (FK_SKIP) / small_w (FK_C0)
  int g_slot;          // gbuf slot holding G rows for output channels [row0, ..)
  int row0;            // first output channel served by the G block (3 for the VF output layer: rows 0..2 come from thin3)
  int thin_slot;       // slot whose first 3 G rows hold the thin gradient of channels 0..2 (-1: none)
  float post;
};
struct FinParams {
  vfnerf_mlp_desc vf, rn;
  const float* vf_arena; const float* rn_arena;
  float* vf_grad; float* rn_grad;
  const float* gbuf;
  float eps;
  int accumulate;      // 1: add to the gradient arenas (supervision points after render()), 0: overwrite
  int n_layers;
  FinLayer L[VFNERF_MAX_LAYERS * 2];
};

__global__ void __launch_bounds__(128) tc_finalize_kernel(const __grid_constant__ FinParams p) {
  const FinLayer& F = p.L[blockIdx.y];
  const vfnerf_mlp_desc& d = F.net == 0 ? p.vf : p.rn;
  const float* arena = F.net == 0 ? p.vf_arena : p.rn_arena;
  float* grad = F.net == 0 ? p.vf_grad : p.rn_grad;
  const int l = F.layer, n = blockIdx.x;
  if (n >= d.out_dim[l]) return;
  const int K = d.in_dim[l];
  const bool bn = d.gamma_off[l] >= 0;
  // source rows
  const bool thin = F.thin_slot >= 0 && n < F.row0;
  const bool hilo = thin || F.kind == FK_HILO;      // rows n (hi part) and n + 3 (lo part) of a unit task
  const float* G = thin ? p.gbuf + (long long)F.thin_slot * kGSlot + (long long)n * kGLd
                        : p.gbuf + (long long)F.g_slot * kGSlot + (long long)(n - F.row0) * kGLd;
  const float s = thin ? p.gbuf[(long long)F.thin_slot * kGSlot + 256 * kGLd + n]
                       : p.gbuf[(long long)F.g_slot * kGSlot + 256 * kGLd + (n - F.row0)];
  float istd = 1.f, sc = 1.f;
  if (bn) {
    istd = 1.f / sqrtf(arena[d.var_off[l] + n] + p.eps);
    sc = arena[d.gamma_off[l] + n] * istd;
  }
  const float* Wn = arena + d.w_off[l] + (long long)n * K;
  float* dWn = grad + d.w_off[l] + (long long)n * K;
  float dot = 0.f;
  for (int j = threadIdx.x; j < K; j += blockDim.x) {
    float g;
    if (hilo) g = G[j] + G[3 * kGLd + j];
    else if (F.kind == FK_IDENT) g = G[j];
    else if (F.kind == FK_EMB) g = G[j] + G[F.epad + j];
    else if (F.kind == FK_SKIP) g = j < F.split ? G[j] : G[256 + (j - F.split)];
    else {  // FK_C0: reference input [p(3), embed(view)(ev), n(3), feat]; packed [feat | n(3), 0 x5, p(3), embed(view)]
      if (j < 3) g = G[256 + 8 + j];
      else if (j < 3 + F.ev) g = G[256 + 11 + (j - 3)];
      else if (j < F.split) g = G[256 + (j - 3 - F.ev)];
      else g = G[j - F.split];
    }
    dot += Wn[j] * g;
    dWn[j] = (p.accumulate ? dWn[j] : 0.f) + sc * F.post * g;
  }
  __shared__ float red[4];
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float tot = red[0] + red[1] + red[2] + red[3];
    auto put = [&](float* dst, float v) { *dst = (p.accumulate ? *dst : 0.f) + v; };
    put(grad + d.b_off[l] + n, sc * F.post * s);
    if (bn) {
      const float b = arena[d.b_off[l] + n];
      put(grad + d.gamma_off[l] + n, istd * F.post * (tot + (b - arena[d.mean_off[l] + n]) * s));
      put(grad + d.beta_off[l] + n, F.post * s);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// orchestration
// ---------------------------------------------------------------------------------------------

// steps 2-5 of the backward.  rn == nullptr: the VF net alone (dcol_pre unused; d(feature pre-tanh) is already in the
// gradient stash).  dcol_pre / dv_pre: [n,3] gradients wrt the pre-activations of the two 3-wide outputs.
static int backward_common(const TcPlan& plan, const vfnerf_mlp_desc& vf, const float* vf_arena, const vfnerf_mlp_desc* rnp,
                           const float* rn_arena, float bn_eps, int64_t n, const float* dcol_pre, const float* dv_pre,
                           float* vf_grad, float* rn_grad, int accumulate, cudaStream_t s) {
  const TcStash& S = plan.stash;
  const bool with_rn = rnp != nullptr;
  vfnerf_mlp_desc none{};
  const vfnerf_mlp_desc& rn = with_rn ? *rnp : none;
  const int L = vf.n_layers, Lr = with_rn ? rn.n_layers : 1;
  const int64_t tiles = (n + kTileM - 1) / kTileM;
  VFN_CHECK_CUDA(cudaMemsetAsync(plan.gbuf, 0, plan.gbuf_floats * sizeof(float), s));
  // 2. fused dgrad chain (also stashes dL/d(pre-activation) of every layer)
  if (int e = tc_forward(plan, with_rn ? TC_MODE_BWD : TC_MODE_VF_BWD, dcol_pre, nullptr, 0, 0, n, dv_pre, 1, nullptr, 0,
                         nullptr, 0, nullptr, s)) return e;

  // stash tensor numbering (mlp_tc.cu build_programs)
  auto yS = [&](int l) { return l; };
  const int yFeat = L - 1;
  auto yC = [&](int l) { return L + l; };
  const int D0 = S.idx_d0;
  // gbuf slots: VF layer l -> l, colour layer l -> L + l; thin results: VF v rows -> slot L + Lr (rows 0..2)
  auto slotVF = [&](int l) { return l; };
  auto slotRN = [&](int l) { return L + l; };
  const int slotThinV = L + Lr;
  VFN_REQUIRE((int64_t)(L + Lr + 1) * kGSlot <= plan.gbuf_floats, "tc_backward: gradient scratch too small");

  // 3. weight-gradient GEMMs
  WgParams wp{};
  wp.stash = plan.stash_buf; wp.gbuf = plan.gbuf; wp.n_tiles = tiles;
  int nt = 0;
  // the main task of a layer (col0 == 0) also produces the column sums of its dY tensor
  auto task = [&](int d_t, int x_t, int x_slab0, int n_xslabs, int slot, int col0, int d_slabs_used) {
    WgTask& T = wp.t[nt++];
    T.d_off = S.off[d_t]; T.x_off = S.off[x_t]; T.d_slabs = S.slabs[d_t]; T.x_slabs = S.slabs[x_t];
    T.x_slab0 = x_slab0; T.n_xslabs = n_xslabs; T.g_off = slot * kGSlot + col0;
    T.colsum_off = col0 == 0 ? slot * kGSlot + 256 * kGLd : -1;
    T.m_rows = 128;
    (void)d_slabs_used;
  };
  // 3-row gradients of the two output layers: G[16 x 256] = unit^T Y with unit = [hi(3), lo(3), 0...]; rows j and j+3
  // are summed by the finalize kernel.  The accumulator rows past 16 read stale shared memory and are never stored.
  auto unit_task = [&](int u_t, int x_t, int slot) {
    WgTask& T = wp.t[nt++];
    T.d_off = S.off[u_t]; T.x_off = S.off[x_t]; T.d_slabs = S.slabs[u_t]; T.x_slabs = S.slabs[x_t];
    T.x_slab0 = 0; T.n_xslabs = 32; T.g_off = slot * kGSlot; T.colsum_off = -1; T.m_rows = 16;
  };
  const int skip = plan.render.skip_step;
  task(D0 + yS(0), S.idx_emb0, 0, S.slabs[S.idx_emb0], slotVF(0), 0, 32);
  for (int l = 1; l <= L - 2; ++l) {
    task(D0 + yS(l), yS(l - 1), 0, 32, slotVF(l), 0, 32);
    if (l == skip) task(D0 + yS(l), S.idx_skip, 0, 6, slotVF(l), 256, 32);
  }
  task(D0 + yFeat, yS(L - 2), 0, 32, slotVF(L - 1), 0, 32);
  if (with_rn) {
    task(D0 + yC(0), yFeat, 0, 32, slotRN(0), 0, 32);
    task(D0 + yC(0), S.idx_aux, 0, 6, slotRN(0), 256, 32);
    for (int l = 1; l <= Lr - 2; ++l) task(D0 + yC(l), yC(l - 1), 0, 32, slotRN(l), 0, 32);
    unit_task(S.idx_dcolu, yC(Lr - 2), slotRN(Lr - 1));
  }
  unit_task(S.idx_dvu, yS(L - 2), slotThinV);
  wp.n_tasks = nt;
  // per-device facts: a process may drive several GPUs
  int dev = 0;
  VFN_CHECK_CUDA(cudaGetDevice(&dev));
  VFN_REQUIRE(dev >= 0 && dev < kMaxDevices, "tc_backward: device ordinal %d unsupported", dev);
  static int sms_of[kMaxDevices] = {0};
  const bool first_on_device = sms_of[dev] == 0;
  if (first_on_device) VFN_CHECK_CUDA(cudaDeviceGetAttribute(&sms_of[dev], cudaDevAttrMultiProcessorCount, dev));
  const int g_sms = sms_of[dev];
  {
    // CTAs per task proportional to the bytes it streams: floor of the proportional share first, then the CTAs that
    // rounding left over go, one at a time, to the task whose CTAs carry the most bytes (the kernel ends with its
    // slowest CTA, so the maximum per-CTA load is what counts)
    double total = 0;
    int cnt[24];
    for (int i = 0; i < nt; ++i) total += wp.t[i].d_slabs + wp.t[i].n_xslabs;
    int used = 0;
    for (int i = 0; i < nt; ++i) {
      int c = std::max(1, (int)((wp.t[i].d_slabs + wp.t[i].n_xslabs) / total * g_sms));
      cnt[i] = (int)std::min<int64_t>(c, std::max<int64_t>(1, tiles));
      used += cnt[i];
    }
    auto load = [&](int i) {   // bytes (in slabs) the busiest CTA of task i streams
      return (double)((tiles + cnt[i] - 1) / cnt[i]) * (wp.t[i].d_slabs + wp.t[i].n_xslabs);
    };
    while (used < g_sms) {
      int best = -1;
      for (int i = 0; i < nt; ++i)
        if (cnt[i] < tiles && (best < 0 || load(i) > load(best))) best = i;
      if (best < 0) break;
      ++cnt[best]; ++used;
    }
    used = 0;
    for (int i = 0; i < nt; ++i) { wp.t[i].cta0 = used; wp.t[i].nctas = cnt[i]; used += cnt[i]; }
    const size_t smem = (size_t)kWgDSlots * kWgDGran + (size_t)kWgXSlots * kWgXBytes + 256;
    if (first_on_device)
      VFN_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_tc_kernel<<<used, kWgThreads, smem, s>>>(wp);
    VFN_LAUNCH_CHECK();
  }
  // 4. bias gradients of the two output layers (the 3-row weight gradients and every column sum of dY come out of the
  //    wgrad kernel)
  {
    if (with_rn) {
      sum3_kernel<<<64, 256, 0, s>>>(dcol_pre, n, plan.gbuf + (int64_t)slotRN(Lr - 1) * kGSlot + 256 * kGLd);
      VFN_LAUNCH_CHECK();
    }
    sum3_kernel<<<64, 256, 0, s>>>(dv_pre, n, plan.gbuf + (int64_t)slotThinV * kGSlot + 256 * kGLd);
    VFN_LAUNCH_CHECK();
  }
  // 5. finalize into the gradient arenas
  {
    FinParams fp{};
    fp.vf = vf; fp.rn = rn; fp.vf_arena = vf_arena; fp.rn_arena = rn_arena; fp.vf_grad = vf_grad; fp.rn_grad = rn_grad;
    fp.gbuf = plan.gbuf; fp.eps = bn_eps; fp.accumulate = accumulate;
    int k = 0, maxrows = 0;
    const int Epad = plan.render.emb_pad, ev = 3 + 6 * plan.render.multires_view;
    for (int l = 0; l < L; ++l) {
      FinLayer& F = fp.L[k++];
      F.net = 0; F.layer = l; F.kind = l == 0 ? FK_EMB : (l == skip ? FK_SKIP : FK_IDENT);
      F.split = l == skip ? vf.out_dim[l - 1] : 0; F.epad = Epad; F.ev = ev;
      F.g_slot = slotVF(l); F.row0 = l == L - 1 ? 3 : 0; F.thin_slot = l == L - 1 ? slotThinV : -1;
      F.post = l < L - 1 ? plan.render.s[l].post_scale : 1.f;
      maxrows = std::max(maxrows, vf.out_dim[l]);
    }
    for (int l = 0; with_rn && l < Lr; ++l) {
      FinLayer& F = fp.L[k++];
      F.net = 1; F.layer = l; F.kind = l == 0 ? FK_C0 : (l == Lr - 1 ? FK_HILO : FK_IDENT); F.split = plan.render.small_w; F.epad = Epad; F.ev = ev;
      F.g_slot = slotRN(l); F.row0 = 0; F.thin_slot = -1; F.post = 1.f;
      maxrows = std::max(maxrows, rn.out_dim[l]);
    }
    fp.n_layers = k;
    tc_finalize_kernel<<<dim3(maxrows, k), 128, 0, s>>>(fp);
    VFN_LAUNCH_CHECK();
  }
  return 0;
}

int tc_backward(const TcPlan& plan, const vfnerf_mlp_desc& vf, const float* vf_arena, const vfnerf_mlp_desc& rn,
                const float* rn_arena, float bn_eps, int64_t n, const float* colors, const float* normals,
                const float* d_colors, const float* d_v, float* vf_grad, float* rn_grad, cudaStream_t s) {
  VFN_REQUIRE(plan.stash_buf && plan.wpack_bwd && plan.gbuf && plan.d3, "tc_backward: training workspace missing");
  float* dcol_pre = plan.d3;
  float* dv_pre = plan.d3 + 3 * n;
  // 1. gradients wrt the pre-activations of the two output layers
  if (int e = launch_act_bwd(colors, 3, d_colors, 3, n, 3, ACT_SIGMOID, dcol_pre, 3, s)) return e;
  if (int e = launch_act_bwd(normals, 3, d_v, 3, n, 3, ACT_TANH, dv_pre, 3, s)) return e;
  return backward_common(plan, vf, vf_arena, &rn, rn_arena, bn_eps, n, dcol_pre, dv_pre, vf_grad, rn_grad, 0, s);
}

// d(vector pre-tanh) -> dv_pre [n,3] fp32; d(feature pre-tanh) -> the gradient twin of the feature tensor in the stash
// (bf16 tile image, the first A operand of the VF-only dgrad chain).  One thread per (point, 8-channel slab).
__global__ void vf_dout_prepare_kernel(const float* __restrict__ out, long long out_ld, const float* __restrict__ d_out,
                                       long long d_ld, int n_cols, long long n, long long n_padded, float* __restrict__ dv_pre,
                                       uint8_t* __restrict__ dfeat_tiles) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_padded * 32) return;
  const long long pnt = i >> 5;
  const int sl = (int)(i & 31);
  float g[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = 0.f;
  if (pnt < n) {
    if (n_cols > 3) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = 3 + sl * 8 + j;
        if (c < n_cols) {
          const float y = out[pnt * out_ld + c];
          g[j] = d_out[pnt * d_ld + c] * (1.f - y * y);
        }
      }
    }
    if (sl == 0) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float y = out[pnt * out_ld + j];
        dv_pre[3 * pnt + j] = d_out[pnt * d_ld + j] * (1.f - y * y);
      }
    }
  }
  const long long tile = pnt / kTileM;
  const int row = (int)(pnt % kTileM);
  *reinterpret_cast<uint4*>(dfeat_tiles + tile * (32LL * 2048) + (long long)sl * 2048 + row * 16) =
      make_uint4(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]), pack_bf16x2(g[4], g[5]), pack_bf16x2(g[6], g[7]));
}

int tc_backward_vf(const TcPlan& plan, const vfnerf_mlp_desc& vf, const float* vf_arena, float bn_eps, int64_t n,
                   const float* out, int64_t out_ld, const float* d_out, int64_t d_ld, int n_out_cols, float* vf_grad,
                   int accumulate, cudaStream_t s) {
  VFN_REQUIRE(plan.stash_buf && plan.wpack_bwd && plan.gbuf && plan.d3, "tc_backward_vf: training workspace missing");
  const TcStash& S = plan.stash;
  const int t_dfeat = S.idx_d0 + vf.n_layers - 1;
  const int64_t n_padded = (n + kTileM - 1) / kTileM * kTileM;
  float* dv_pre = plan.d3;
  vf_dout_prepare_kernel<<<(unsigned)((n_padded * 32 + 255) / 256), 256, 0, s>>>(out, out_ld, d_out, d_ld, n_out_cols, n, n_padded,
                                                                                dv_pre, plan.stash_buf + S.off[t_dfeat]);
  VFN_LAUNCH_CHECK();
  return backward_common(plan, vf, vf_arena, nullptr, nullptr, bn_eps, n, nullptr, dv_pre, vf_grad, nullptr, accumulate, s);
}

// ---------------------------------------------------------------------------------------------
// test support: one stash tensor -> row-major fp32 [n, 8 * slabs]
// ---------------------------------------------------------------------------------------------
__global__ void stash_read_kernel(const uint8_t* __restrict__ t, int slabs, long long n, float* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // (point, slab)
  if (i >= n * slabs) return;
  const long long pnt = i / slabs;
  const int sl = (int)(i % slabs);
  const long long tile = pnt / kTileM;
  const int row = (int)(pnt % kTileM);
  const __nv_bfloat16* u = reinterpret_cast<const __nv_bfloat16*>(t + tile * (long long)slabs * 2048 + (long long)sl * 2048 + row * 16);
  for (int j = 0; j < 8; ++j) out[pnt * (slabs * 8) + sl * 8 + j] = __bfloat162float(u[j]);
}

int tc_debug_stash_read(const TcPlan& plan, int tensor, int64_t n, float* out, int* n_cols, cudaStream_t s) {
  VFN_REQUIRE(plan.stash_buf, "stash_read: no training workspace");
  const TcStash& S = plan.stash;
  if (tensor == 1000) {       // [n, 6]: d(colour pre-sigmoid), d(vector pre-tanh)
    if (n_cols) *n_cols = 6;
    if (!out) return 0;
    VFN_CHECK_CUDA(cudaMemcpy2DAsync(out, 24, plan.d3, 12, 12, n, cudaMemcpyDeviceToDevice, s));
    VFN_CHECK_CUDA(cudaMemcpy2DAsync(out + 3, 24, plan.d3 + 3 * n, 12, 12, n, cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  VFN_REQUIRE(tensor >= 0 && tensor < S.n_tensors, "stash_read: tensor %d out of range [0, %d)", tensor, S.n_tensors);
  if (n_cols) *n_cols = S.slabs[tensor] * 8;
  if (!out || n == 0) return 0;
  const long long work = n * S.slabs[tensor];
  stash_read_kernel<<<(unsigned)((work + 255) / 256), 256, 0, s>>>(plan.stash_buf + S.off[tensor], S.slabs[tensor], n, out);
  VFN_LAUNCH_CHECK();
  return 0;
}

}  // namespace vfn
