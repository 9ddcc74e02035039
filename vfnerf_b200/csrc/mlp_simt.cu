// CUDA-core fp32 building blocks of the two MLPs (VFNERF_PREC_FP32): a strided GEMM with a fused
// BatchNorm-affine / activation / skip-divide / ReLU-mask epilogue, the positional encoding, the
// BatchNorm fold, and the small reductions the backward needs.
//
// This is the generic-width, fp32-accurate path: it carries the 1e-3 parity contract, serves any
// layer widths (the goldens use a 64-wide net), and is the on-device reference the tcgen05 kernels
// (mlp_tc.cu) are validated against at full problem size.  It is not the throughput path.
#include "common.cuh"

namespace vfn {

// ---------------------------------------------------------------------------------------------
// strided SGEMM: C[m,n] (+)= sum_k A(m,k) * B(k,n)
// 128x128x16 tiles, 256 threads, 8x8 register blocking.
// ---------------------------------------------------------------------------------------------
constexpr int BM = 128, BN = 128, BK = 16, GT = 256;

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_TANH: return tanhf(v);
    case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

__global__ void __launch_bounds__(GT) gemm_kernel(GemmArgs g) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int64_t kchunk = (g.K + g.split_k - 1) / g.split_k;
  const int64_t kbeg = (int64_t)blockIdx.z * kchunk;
  const int64_t kend = min(g.K, kbeg + kchunk);
  const int ty = t >> 4, tx = t & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const bool a_kfast = (g.a_cs == 1);
  const bool b_kfast = (g.b_rs == 1);
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- A tile (BM x BK) -> As[k][m]
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      int kk, mm;
      if (a_kfast) { kk = t & 15; mm = (t >> 4) + 16 * p; }
      else { mm = t & 127; kk = (t >> 7) + 2 * p; }
      int64_t m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < g.M && k < kend) {
        v = __ldg(g.A + m * g.a_rs + k * g.a_cs);
        if (g.a_kscale) v *= __ldg(g.a_kscale + k);
      }
      As[kk][mm] = v;
    }
    // ---- B tile (BK x BN) -> Bs[k][n]
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      int kk, nn;
      if (b_kfast) { kk = t & 15; nn = (t >> 4) + 16 * p; }
      else { nn = t & 127; kk = (t >> 7) + 2 * p; }
      int64_t k = k0 + kk;
      int n = n0 + nn;
      float v = 0.f;
      if (n < g.N && k < kend) v = __ldg(g.B + k * g.b_rs + (int64_t)n * g.b_cs);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t m = m0 + ty * 8 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = n0 + tx * 8 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      float* c = g.C + m * g.c_rs + n;
      if (g.split_k > 1 || g.accumulate) { atomicAdd(c, v); continue; }
      if (g.scale) v *= __ldg(g.scale + n);
      if (g.shift) v += __ldg(g.shift + n);
      v = apply_act(v, g.act);
      if (g.post_div != 0.f) v = __fdiv_rn(v, g.post_div);
      if (g.mask && !(__ldg(g.mask + m * g.mask_rs + n) > 0.f)) v = 0.f;
      *c = v;
    }
  }
}

int launch_gemm(const GemmArgs& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0) return 0;
  VFN_REQUIRE(g.split_k >= 1, "gemm: split_k must be >= 1");
  VFN_REQUIRE(g.a_cs == 1 || g.a_rs == 1, "gemm: A needs a unit stride");
  VFN_REQUIRE(g.b_cs == 1 || g.b_rs == 1, "gemm: B needs a unit stride");
  dim3 grid((unsigned)ceil_div64(g.M, BM), (unsigned)((g.N + BN - 1) / BN), (unsigned)g.split_k);
  gemm_kernel<<<grid, GT, 0, s>>>(g);
  VFN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// BatchNorm1d (eval) folded into a per-channel affine: y = (x W^T) * scale + shift
// ---------------------------------------------------------------------------------------------
__global__ void fold_bn_kernel(vfnerf_mlp_desc d, const float* __restrict__ arena, float eps,
                               float* __restrict__ scale, float* __restrict__ shift) {
  int l = blockIdx.y;
  int off = 0;
  for (int i = 0; i < l; ++i) off += d.out_dim[i];
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < d.out_dim[l]; n += gridDim.x * blockDim.x) {
    float b = arena[d.b_off[l] + n];
    float sc = 1.f, sh = b;
    if (d.gamma_off[l] >= 0) {
      sc = arena[d.gamma_off[l] + n] / sqrtf(arena[d.var_off[l] + n] + eps);
      sh = arena[d.beta_off[l] + n] + (b - arena[d.mean_off[l] + n]) * sc;
    }
    scale[off + n] = sc;
    shift[off + n] = sh;
  }
}

int launch_fold_bn(const vfnerf_mlp_desc& d, const float* arena, float bn_eps, float* scale,
                   float* shift, cudaStream_t s) {
  fold_bn_kernel<<<dim3(2, d.n_layers), 256, 0, s>>>(d, arena, bn_eps, scale, shift);
  VFN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// positional encoding, embedder.py:11-37: [x, sin(2^k x), cos(2^k x)]_k / div
// ---------------------------------------------------------------------------------------------
__global__ void embed_kernel(const float* __restrict__ x, int64_t x_ld, int64_t n, int multires,
                             float div, float* __restrict__ out, int64_t out_ld) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p[3] = {x[i * x_ld], x[i * x_ld + 1], x[i * x_ld + 2]};
  float* o = out + i * out_ld;
  auto put = [&](int c, float v) { o[c] = (div != 0.f) ? __fdiv_rn(v, div) : v; };
#pragma unroll
  for (int c = 0; c < 3; ++c) put(c, p[c]);
  float f = 1.f;
  for (int k = 0; k < multires; ++k) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float a = __fmul_rn(p[c], f);
      put(3 + 6 * k + c, sinf(a));
      put(6 + 6 * k + c, cosf(a));
    }
    f *= 2.f;
  }
}

int launch_embed(const float* x, int64_t x_ld, int64_t n, int multires, float div, float* out,
                 int64_t out_ld, cudaStream_t s) {
  if (n <= 0) return 0;
  embed_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(x, x_ld, n, multires, div, out, out_ld);
  VFN_LAUNCH_CHECK();
  return 0;
}

// colour-net input head: columns [0,3) = point, [3, 3+3+6*mv) = embed(view dir); the remaining
// columns (normals, features) are written in place by the VF net's last layer.  Also materialises
// NerfOutput.ray_dirs (unit direction repeated per sample) when asked.
__global__ void color_input_head_kernel(const float* __restrict__ points, const float* __restrict__ ray_dirs,
                                        int64_t total, int N, int mv, float* __restrict__ cin, int64_t ld,
                                        float* __restrict__ rep) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t r = i / N;
  float d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    d[c] = __ldg(ray_dirs + 3 * r + c);
    if (rep) rep[3 * i + c] = d[c];
  }
  if (!cin) return;
  float* o = cin + i * ld;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o[c] = points[3 * i + c];
    o[3 + c] = d[c];
  }
  float f = 1.f;
  for (int k = 0; k < mv; ++k) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float a = __fmul_rn(d[c], f);
      o[6 + 6 * k + c] = sinf(a);
      o[9 + 6 * k + c] = cosf(a);
    }
    f *= 2.f;
  }
}

int launch_color_input_head(const float* points, const float* ray_dirs, int n_rays, int n_samples,
                            int multires_view, float* color_in, int64_t ld, float* ray_dirs_rep,
                            cudaStream_t s) {
  int64_t total = (int64_t)n_rays * n_samples;
  if (total <= 0) return 0;
  color_input_head_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(
      points, ray_dirs, total, n_samples, multires_view, color_in, ld, ray_dirs_rep);
  VFN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// backward helpers
// ---------------------------------------------------------------------------------------------
// out[i, c] = dy[i, c] * act'(y[i, c])   (dy may be NULL -> treated as zero)
__global__ void act_bwd_kernel(const float* __restrict__ y, int64_t y_ld, const float* __restrict__ dy,
                               int64_t dy_ld, int64_t n, int cols, int act, float* __restrict__ out,
                               int64_t out_ld) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * cols) return;
  int64_t i = e / cols;
  int c = (int)(e - i * cols);
  float g = dy ? dy[i * dy_ld + c] : 0.f;
  float v = y[i * y_ld + c];
  float d = 1.f;
  if (act == ACT_TANH) d = 1.f - v * v;
  else if (act == ACT_SIGMOID) d = v * (1.f - v);
  else if (act == ACT_RELU) d = v > 0.f ? 1.f : 0.f;
  out[i * out_ld + c] = g * d;
}

int launch_act_bwd(const float* y, int64_t y_ld, const float* dy, int64_t dy_ld, int64_t n, int cols,
                   int act, float* out, int64_t out_ld, cudaStream_t s) {
  if (n <= 0 || cols <= 0) return 0;
  act_bwd_kernel<<<(unsigned)ceil_div64(n * cols, 256), 256, 0, s>>>(y, y_ld, dy, dy_ld, n, cols, act, out, out_ld);
  VFN_LAUNCH_CHECK();
  return 0;
}

// out[c] += sum_i a[i, c]   (out pre-zeroed by the caller); 32 columns x 256-row slabs per block
__global__ void colsum_kernel(const float* __restrict__ a, int64_t a_rs, int64_t rows, int cols,
                              float* __restrict__ out) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.y * 2048;
  const int64_t r1 = min(rows, r0 + 2048);
  float acc = 0.f;
  if (c < cols)
    for (int64_t i = r0 + ry; i < r1; i += 8) acc += a[i * a_rs + c];
  red[ry][threadIdx.x & 31] = acc;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[k][threadIdx.x];
    atomicAdd(out + c, v);
  }
}

int launch_colsum(const float* a, int64_t a_rs, int64_t rows, int cols, const float* /*colscale*/,
                  float* out, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return 0;
  VFN_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, s));
  dim3 grid((cols + 31) / 32, (unsigned)ceil_div64(rows, 2048));
  colsum_kernel<<<grid, 256, 0, s>>>(a, a_rs, rows, cols, out);
  VFN_LAUNCH_CHECK();
  return 0;
}

// Turns G = dY^T X (unscaled weight gradient wrt the *folded* layer) and s = colsum(dY) into the
// gradients of Linear.weight/bias and BatchNorm1d.weight/bias of layer l (eval-mode BN):
//   y = (x W^T + b - mean) * gamma * istd + beta,   istd = 1/sqrt(var + eps)
//   dW = gamma*istd * G,  db = gamma*istd * s,  dbeta = s,
//   dgamma = istd * (rowdot(W, G) + (b - mean) * s)
__global__ void grad_finalize_kernel(vfnerf_mlp_desc d, int l, const float* __restrict__ arena, float eps,
                                     const float* __restrict__ s, const float* __restrict__ G,
                                     float* __restrict__ grad, int accumulate) {
  const int n = blockIdx.x;
  const int K = d.in_dim[l];
  const bool bn = d.gamma_off[l] >= 0;
  float istd = 1.f, sc = 1.f;
  if (bn) {
    istd = 1.f / sqrtf(arena[d.var_off[l] + n] + eps);
    sc = arena[d.gamma_off[l] + n] * istd;
  }
  float dotwg = 0.f;
  const float* Wn = arena + d.w_off[l] + (int64_t)n * K;
  const float* Gn = G + (int64_t)n * K;
  float* dWn = grad + d.w_off[l] + (int64_t)n * K;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float gv = Gn[k];
    dotwg += Wn[k] * gv;
    float v = sc * gv;
    dWn[k] = accumulate ? dWn[k] + v : v;
  }
  __shared__ float red[8];
  dotwg = warp_sum(dotwg);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dotwg;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
    float sn = s[n];
    float* db = grad + d.b_off[l] + n;
    float v = sc * sn;
    *db = accumulate ? *db + v : v;
    if (bn) {
      float b = arena[d.b_off[l] + n];
      float dg = istd * (tot + (b - arena[d.mean_off[l] + n]) * sn);
      float* pg = grad + d.gamma_off[l] + n;
      float* pb = grad + d.beta_off[l] + n;
      *pg = accumulate ? *pg + dg : dg;
      *pb = accumulate ? *pb + sn : sn;
    }
  }
}

int launch_grad_finalize(const vfnerf_mlp_desc& d, int layer, const float* arena, float bn_eps,
                         const float* colsum_dy, float* grad_arena, int accumulate,
                         const float* G_tmp, cudaStream_t s) {
  grad_finalize_kernel<<<d.out_dim[layer], 256, 0, s>>>(d, layer, arena, bn_eps, colsum_dy, G_tmp,
                                                          grad_arena, accumulate);
  VFN_LAUNCH_CHECK();
  return 0;
}

__global__ void copy_cols_kernel(const float* __restrict__ src, int64_t src_ld, float* __restrict__ dst,
                                 int64_t dst_ld, int64_t rows, int cols) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * cols) return;
  int64_t i = e / cols;
  int c = (int)(e - i * cols);
  dst[i * dst_ld + c] = src[i * src_ld + c];
}

int launch_copy_cols(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int64_t rows,
                     int cols, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return 0;
  copy_cols_kernel<<<(unsigned)ceil_div64(rows * cols, 256), 256, 0, s>>>(src, src_ld, dst, dst_ld, rows, cols);
  VFN_LAUNCH_CHECK();
  return 0;
}

// grid coordinates of evaluation/methods.py:194-208: index i -> (ix, iy, iz) with z fastest,
// position = ((index * voxel + origin) + translation) + centroid, separate fp32 ops like the reference.
__global__ void grid_points_kernel(int res, int64_t i0, int64_t n, GridSpec gs, float* __restrict__ pts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t g = i0 + i;
  int64_t idx[3] = {(g / res / res) % res, (g / res) % res, g % res};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = __fadd_rn(__fmul_rn((float)idx[c], gs.voxel), gs.origin[c]);
    v = __fadd_rn(v, gs.translation[c]);
    pts[3 * i + c] = __fadd_rn(v, gs.centroid[c]);
  }
}

int launch_grid_points(int res, int64_t i0, int64_t n, const GridSpec& gs, float* pts, cudaStream_t s) {
  if (n <= 0) return 0;
  grid_points_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(res, i0, n, gs, pts);
  VFN_LAUNCH_CHECK();
  return 0;
}

}  // namespace vfn
