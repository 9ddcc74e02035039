// Train-mode path of the two MLPs: what VectorFieldNerf.render() and the VF-only call do after
// model.train() -- vector_field_network.py:140-175 (autograd "Jacobian"), :177-208 / rendering_network.py:62-108 with
// BatchNorm1d normalising by BATCH statistics, vector_field_nerf.py:264-270,301-305,476-498 (directional derivatives).
//
// Batch statistics couple every sample of a layer's batch, so this is a layer-wise path: per hidden layer one GEMM
// (mlp_simt.cu, fp32), a two-level column reduction (per-slab pivoted moments, combined in double in a fixed order:
// reproducible and without the E[x^2]-E[x]^2 cancellation), the running-statistics update, and the normalise + ReLU
// pass.  The backward is the exact BatchNorm-training backward (column means of dy and dy*xhat between the GEMMs).  The
// reference's "Jacobian" is three autograd.grad calls of COLUMN SUMS over the batch through those statistics, i.e. three
// reverse sweeps with a one-hot seed on the output column -- reproduced as such (a per-sample forward-mode Jacobian would
// miss the cross-sample terms of the batch mean / variance and disagree with the reference).
#include "common.cuh"
#include "host_plan.cuh"

namespace vfn {

constexpr int kSlabRows = 1024;

// ---------------------------------------------------------------------------------------------
// column statistics: per 1024-row slab the pivoted moments (pivot, sum(x - p), sum((x - p)^2)) of every column
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) col_stats_partial_kernel(const float* __restrict__ z, int64_t ld, int64_t rows,
                                                                int cols, float* __restrict__ partial) {
  __shared__ float r1[8][33], r2[8][33];
  const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int64_t r0 = (int64_t)blockIdx.y * kSlabRows, re = min(rows, r0 + kSlabRows);
  float p = 0.f, s1 = 0.f, s2 = 0.f;
  if (c < cols) {
    p = z[r0 * ld + c];
    for (int64_t i = r0 + ry; i < re; i += 8) {
      const float d = z[i * ld + c] - p;
      s1 += d;
      s2 = fmaf(d, d, s2);
    }
  }
  r1[ry][lane] = s1;
  r2[ry][lane] = s2;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += r1[k][lane]; b += r2[k][lane]; }
    float* o = partial + ((int64_t)blockIdx.y * cols + c) * 3;
    o[0] = p; o[1] = a; o[2] = b;
  }
}

// Combines the slabs (Chan's formula, double, slab order) into the batch mean and BIASED variance of every column,
// stores mean / 1/sqrt(var+eps) / gamma/sqrt(var+eps) for the normalise pass and the backward, and folds the batch into
// the running statistics like nn.BatchNorm1d.forward in training mode: running = (1-m) running + m batch, with the
// UNBIASED variance.
__global__ void bn_stats_finalize_kernel(const float* __restrict__ partial, int slabs, int64_t rows, int cols,
                                         const float* __restrict__ gamma, float eps, float momentum,
                                         float* __restrict__ mean_o, float* __restrict__ istd_o, float* __restrict__ scale_o,
                                         float* __restrict__ run_mean, float* __restrict__ run_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double tot = 0.0;
  for (int s = 0; s < slabs; ++s) {
    const double ns = (double)min((int64_t)kSlabRows, rows - (int64_t)s * kSlabRows);
    const float* o = partial + ((int64_t)s * cols + c) * 3;
    tot += ns * (double)o[0] + (double)o[1];
  }
  const double mean = tot / (double)rows;
  double m2 = 0.0;
  for (int s = 0; s < slabs; ++s) {
    const double ns = (double)min((int64_t)kSlabRows, rows - (int64_t)s * kSlabRows);
    const float* o = partial + ((int64_t)s * cols + c) * 3;
    const double ms = (double)o[0] + (double)o[1] / ns;
    const double dm = ms - mean;
    m2 += ((double)o[2] - (double)o[1] * (double)o[1] / ns) + ns * dm * dm;
  }
  const double var = fmax(m2 / (double)rows, 0.0);
  const float istd = 1.f / sqrtf((float)var + eps);
  mean_o[c] = (float)mean;
  istd_o[c] = istd;
  scale_o[c] = gamma[c] * istd;
  if (run_mean) {
    const double unbiased = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
    run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * (float)mean;
    run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)unbiased;
  }
}

// out = relu((z - mean) * (gamma * istd) + beta) [/ post_div]
__global__ void bn_act_kernel(const float* __restrict__ z, int64_t rows, int cols, const float* __restrict__ mean,
                              const float* __restrict__ scale, const float* __restrict__ beta, float post_div,
                              float* __restrict__ out, int64_t out_ld) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * cols) return;
  const int64_t i = e / cols;
  const int c = (int)(e - i * cols);
  float v = fmaxf(fmaf(z[e] - mean[c], scale[c], beta[c]), 0.f);
  if (post_div != 0.f) v = __fdiv_rn(v, post_div);
  out[i * out_ld + c] = v;
}

// ---------------------------------------------------------------------------------------------
// BatchNorm-training backward:  dz = gamma*istd * (dy - mean(dy) - xhat * mean(dy * xhat)),  dgamma = sum dy*xhat,
// dbeta = sum dy;  xhat = (z - mean) * istd
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_bwd_partial_kernel(const float* __restrict__ dy, int64_t dy_ld,
                                                             const float* __restrict__ z, int64_t rows, int cols,
                                                             const float* __restrict__ mean, const float* __restrict__ istd,
                                                             float* __restrict__ partial) {
  __shared__ float r1[8][33], r2[8][33];
  const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int64_t r0 = (int64_t)blockIdx.y * kSlabRows, re = min(rows, r0 + kSlabRows);
  float s1 = 0.f, s2 = 0.f;
  if (c < cols) {
    const float m = mean[c], is = istd[c];
    for (int64_t i = r0 + ry; i < re; i += 8) {
      const float g = dy[i * dy_ld + c];
      s1 += g;
      s2 = fmaf(g, (z[i * (int64_t)cols + c] - m) * is, s2);
    }
  }
  r1[ry][lane] = s1;
  r2[ry][lane] = s2;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += r1[k][lane]; b += r2[k][lane]; }
    float* o = partial + ((int64_t)blockIdx.y * cols + c) * 2;
    o[0] = a; o[1] = b;
  }
}

__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int slabs, int64_t rows, int cols,
                                       float* __restrict__ c1, float* __restrict__ c2, float* __restrict__ d_gamma,
                                       float* __restrict__ d_beta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double a = 0.0, b = 0.0;
  for (int s = 0; s < slabs; ++s) {
    const float* o = partial + ((int64_t)s * cols + c) * 2;
    a += (double)o[0];
    b += (double)o[1];
  }
  c1[c] = (float)(a / (double)rows);
  c2[c] = (float)(b / (double)rows);
  if (d_gamma) { d_gamma[c] += (float)b; d_beta[c] += (float)a; }
}

__global__ void bn_bwd_apply_kernel(float* __restrict__ dy, int64_t dy_ld, const float* __restrict__ z, int64_t rows, int cols,
                                    const float* __restrict__ mean, const float* __restrict__ istd,
                                    const float* __restrict__ scale, const float* __restrict__ c1,
                                    const float* __restrict__ c2) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * cols) return;
  const int64_t i = e / cols;
  const int c = (int)(e - i * cols);
  const float xh = (z[e] - mean[c]) * istd[c];
  float* p = dy + i * dy_ld + c;
  *p = scale[c] * (*p - c1[c] - xh * c2[c]);
}

__global__ void axpy_kernel(float* __restrict__ out, const float* __restrict__ in, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += in[i];
}

// ---------------------------------------------------------------------------------------------
// the reference's "Jacobian": seed of output column k, and the way back through the positional encoding
// ---------------------------------------------------------------------------------------------
// gradient of sum_j tanh(u_j)[k] wrt the last hidden activation, masked by that layer's ReLU:
//   dy[i, c] = (1 - y[i,k]^2) * W_last[k, c] * [a[i,c] > 0]
__global__ void jac_seed_kernel(const float* __restrict__ y, int64_t y_ld, int k, const float* __restrict__ w_row,
                                const float* __restrict__ a, int64_t a_ld, int64_t rows, int cols, float* __restrict__ dy) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * cols) return;
  const int64_t i = e / cols;
  const int c = (int)(e - i * cols);
  const float t = y[i * y_ld + k];
  dy[e] = a[i * a_ld + c] > 0.f ? (1.f - t * t) * w_row[c] : 0.f;
}

// d/dx of [x, sin(2^k x), cos(2^k x)]_k contracted with the gradient of the embedding (the sum of the layer-0 path and,
// if the net has a skip layer, the skip path)
__global__ void embed_bwd_kernel(const float* __restrict__ x, int64_t n, int multires, const float* __restrict__ g0,
                                 const float* __restrict__ g1, int E, float* __restrict__ out, int64_t out_ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* a = g0 + i * E;
  const float* b = g1 ? g1 + i * E : nullptr;
  auto g = [&](int c) { return b ? a[c] + b[c] : a[c]; };
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float p = x[3 * i + c];
    float acc = g(c);
    float f = 1.f;
    for (int k = 0; k < multires; ++k) {
      const float ang = __fmul_rn(p, f);
      acc += f * (cosf(ang) * g(3 + 6 * k + c) - sinf(ang) * g(6 + 6 * k + c));
      f *= 2.f;
    }
    out[i * out_ld + c] = acc;
  }
}

// compute_directional_derivatives (vector_field_nerf.py:476-498) + the reference's bookkeeping around it (:268,305,337):
// rows (i,0), (i,1) = J_i n1_i, J_i n2_i with n1 = normalize(n.y, -n.x, 0), n2 = normalize(n x n1); the [2P,3] block is
// concatenated with ITSELF and the per-row norms are returned: out[0..2P) = out[2P..4P).
__global__ void dir_derivs_kernel(const float* __restrict__ nrm, int64_t nrm_ld, const float* __restrict__ jac, int64_t n,
                                  float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float nx = nrm[i * nrm_ld], ny = nrm[i * nrm_ld + 1], nz = nrm[i * nrm_ld + 2];
  float a[3] = {ny, -nx, 0.f};
  float b[3] = {ny * a[2] - nz * a[1], nz * a[0] - nx * a[2], nx * a[1] - ny * a[0]};
  const float la = fmaxf(sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), 1e-12f);
  const float lb = fmaxf(sqrtf(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]), 1e-12f);
  const float* J = jac + 9 * i;
  float sa = 0.f, sb = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float va = 0.f, vb = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      va += J[3 * r + c] * (a[c] / la);
      vb += J[3 * r + c] * (b[c] / lb);
    }
    sa += va * va;
    sb += vb * vb;
  }
  sa = sqrtf(sa); sb = sqrtf(sb);
  out[2 * i] = sa; out[2 * i + 1] = sb;
  out[2 * n + 2 * i] = sa; out[2 * n + 2 * i + 1] = sb;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct TrainBufs {
  MlpBufs m;                          // act[l]: post-activation outputs, skip concat in place; scale: gamma * istd
  float* z[VFNERF_MAX_LAYERS];        // pre-BatchNorm outputs of the hidden layers [n, out_dim[l]]
  float *mean, *istd;                 // concatenated over layers like m.scale (m.soff)
  float* partial;                     // slab partials of the column reductions
  float *c1, *c2;
};

static void carve_train(Carver& c, const vfnerf_mlp_desc& d, int64_t n, TrainBufs& b) {
  carve_mlp(c, d, n, b.m);
  const int so = sum_out(d);
  b.mean = c.f(so);
  b.istd = c.f(so);
  for (int l = 0; l + 1 < d.n_layers; ++l) b.z[l] = c.f(n * d.out_dim[l]);
  const int w = max_dim(d);
  b.partial = c.f(ceil_div64(std::max<int64_t>(n, 1), kSlabRows) * w * 3);
  b.c1 = c.f(w);
  b.c2 = c.f(w);
}

static int validate_train(const vfnerf_mlp_desc& d, int skip_layer, const char* who) {
  for (int l = 0; l + 1 < d.n_layers; ++l)
    VFN_REQUIRE(d.gamma_off[l] >= 0 && d.mean_off[l] >= 0, "%s: hidden layer %d has no BatchNorm (train mode is the batch_norm "
                "configuration of confs/vf_nerf.conf)", who, l);
  VFN_REQUIRE(skip_layer != d.n_layers - 1, "%s: a skip connection into the last layer is not supported in train mode", who);
  return 0;
}

// Forward of one MLP with batch statistics.  The last layer (no BatchNorm) writes act(x W^T + b)[:, :n_out_cols] to out.
static int mlp_forward_train(const vfnerf_mlp_desc& d, float* arena, const TrainBufs& b, int skip_layer, const float* x0,
                             int64_t x0_ld, int64_t n, float* out, int64_t out_ld, int n_out_cols, int last_act, float eps,
                             float momentum, cudaStream_t s) {
  const float* x = x0;
  int64_t x_ld = x0_ld;
  const int slabs = (int)ceil_div64(n, kSlabRows);
  for (int l = 0; l < d.n_layers; ++l) {
    const bool last = (l == d.n_layers - 1);
    const int No = d.out_dim[l];
    GemmArgs g{};
    g.A = x; g.a_rs = x_ld; g.a_cs = 1;
    g.B = arena + d.w_off[l]; g.b_rs = 1; g.b_cs = d.in_dim[l];
    g.M = n; g.K = d.in_dim[l];
    g.shift = arena + d.b_off[l]; g.split_k = 1;
    if (last) {
      g.C = out; g.c_rs = out_ld; g.N = n_out_cols; g.act = last_act;
      return launch_gemm(g, s);
    }
    g.C = b.z[l]; g.c_rs = No; g.N = No; g.act = ACT_NONE;
    if (int e = launch_gemm(g, s)) return e;
    col_stats_partial_kernel<<<dim3((No + 31) / 32, slabs), 256, 0, s>>>(b.z[l], No, n, No, b.partial);
    VFN_LAUNCH_CHECK();
    const int so = b.m.soff[l];
    bn_stats_finalize_kernel<<<(No + 127) / 128, 128, 0, s>>>(b.partial, slabs, n, No, arena + d.gamma_off[l], eps, momentum,
                                                             b.mean + so, b.istd + so, b.m.scale + so,
                                                             arena + d.mean_off[l], arena + d.var_off[l]);
    VFN_LAUNCH_CHECK();
    bn_act_kernel<<<(unsigned)ceil_div64(n * No, 256), 256, 0, s>>>(b.z[l], n, No, b.mean + so, b.m.scale + so,
                                                                    arena + d.beta_off[l],
                                                                    (l + 1 == skip_layer) ? kSqrt2 : 0.f, b.m.act[l],
                                                                    d.in_dim[l + 1]);
    VFN_LAUNCH_CHECK();
    x = b.m.act[l];
    x_ld = d.in_dim[l + 1];
  }
  return 0;
}

// Backward of one MLP through the batch statistics, from layer l_start down.
//   l_start == n_layers-1: dY [n, out_dim] is the gradient wrt the last layer's PRE-activation output (read only);
//   l_start <  n_layers-1: dY == w.dB [n, out_dim[l_start]] is the gradient wrt hidden layer l_start's BatchNorm output,
//                          ReLU mask already applied (the Jacobian seed); overwritten.
// grad_arena (optional): parameter gradients are ADDED to it.  d_in0 (optional): gradient wrt columns
// [in0_col0, in0_col0+in0_cols) of x0.  d_skip_tail (optional, [n, E]): gradient wrt the embedding through the skip
// connection (the tail columns of the skip layer's input).
static int mlp_backward_train(const vfnerf_mlp_desc& d, const float* arena, const TrainBufs& b, int skip_layer,
                              const float* x0, int64_t x0_ld, int64_t n, int l_start, const float* dY, int64_t dy_ld,
                              const BwdBufs& w, float* grad_arena, float* d_in0, int64_t d_in0_ld, int in0_col0,
                              int in0_cols, float* d_skip_tail, cudaStream_t s) {
  const float* dy = dY;
  int64_t ld = dy_ld;
  float* pp[2] = {w.dA, w.dB};
  int flip = 0;
  const int split = (int)std::min<int64_t>(128, std::max<int64_t>(1, n / 1024));
  const int slabs = (int)ceil_div64(n, kSlabRows);
  for (int l = l_start; l >= 0; --l) {
    const float* x = (l == 0) ? x0 : b.m.act[l - 1];
    const int64_t x_ld = (l == 0) ? x0_ld : d.in_dim[l];
    const int No = d.out_dim[l], Ki = d.in_dim[l];
    const bool last = (l == d.n_layers - 1);
    if (!last) {
      // dy -> dz in place (dy is one of our ping-pong buffers below the last layer)
      float* dyw = const_cast<float*>(dy);
      const int so = b.m.soff[l];
      bn_bwd_partial_kernel<<<dim3((No + 31) / 32, slabs), 256, 0, s>>>(dy, ld, b.z[l], n, No, b.mean + so, b.istd + so,
                                                                        b.partial);
      VFN_LAUNCH_CHECK();
      bn_bwd_finalize_kernel<<<(No + 127) / 128, 128, 0, s>>>(b.partial, slabs, n, No, b.c1, b.c2,
                                                             grad_arena ? grad_arena + d.gamma_off[l] : nullptr,
                                                             grad_arena ? grad_arena + d.beta_off[l] : nullptr);
      VFN_LAUNCH_CHECK();
      bn_bwd_apply_kernel<<<(unsigned)ceil_div64(n * No, 256), 256, 0, s>>>(dyw, ld, b.z[l], n, No, b.mean + so, b.istd + so,
                                                                            b.m.scale + so, b.c1, b.c2);
      VFN_LAUNCH_CHECK();
    }
    if (grad_arena) {
      if (last) {
        // Linear.bias of the output layer.  (A bias in front of a batch-statistics BatchNorm has an identically zero
        // gradient -- sum_i dz_i = 0 -- which autograd reproduces only up to rounding noise; it is left at zero here.)
        if (int e = launch_colsum(dy, ld, n, No, nullptr, w.colsum, s)) return e;
        axpy_kernel<<<(No + 255) / 256, 256, 0, s>>>(grad_arena + d.b_off[l], w.colsum, No);
        VFN_LAUNCH_CHECK();
      }
      GemmArgs g{};
      g.A = dy; g.a_rs = 1; g.a_cs = ld;            // A(m = out channel, k = point)
      g.B = x; g.b_rs = x_ld; g.b_cs = 1;           // B(k = point, n = in channel)
      g.C = grad_arena + d.w_off[l]; g.c_rs = Ki; g.M = No; g.N = Ki; g.K = n; g.split_k = split; g.accumulate = 1;
      if (int e = launch_gemm(g, s)) return e;
    }
    if (l > 0) {
      GemmArgs h{};
      h.A = dy; h.a_rs = ld; h.a_cs = 1;
      h.B = arena + d.w_off[l]; h.b_rs = Ki; h.b_cs = 1;
      h.C = pp[flip]; h.c_rs = d.out_dim[l - 1]; h.M = n; h.N = d.out_dim[l - 1]; h.K = No; h.split_k = 1;
      h.post_div = (l == skip_layer) ? kSqrt2 : 0.f;
      h.mask = b.m.act[l - 1]; h.mask_rs = d.in_dim[l];
      if (int e = launch_gemm(h, s)) return e;
      if (l == skip_layer && d_skip_tail) {
        GemmArgs t{};
        t.A = dy; t.a_rs = ld; t.a_cs = 1;
        t.B = arena + d.w_off[l] + d.out_dim[l - 1]; t.b_rs = Ki; t.b_cs = 1;
        t.C = d_skip_tail; t.c_rs = Ki - d.out_dim[l - 1]; t.M = n; t.N = Ki - d.out_dim[l - 1]; t.K = No; t.split_k = 1;
        t.post_div = kSqrt2;
        if (int e = launch_gemm(t, s)) return e;
      }
      dy = pp[flip]; ld = d.out_dim[l - 1];
      flip ^= 1;
    } else if (d_in0) {
      GemmArgs h{};
      h.A = dy; h.a_rs = ld; h.a_cs = 1;
      h.B = arena + d.w_off[0] + in0_col0; h.b_rs = Ki; h.b_cs = 1;
      h.C = d_in0; h.c_rs = d_in0_ld; h.M = n; h.N = in0_cols; h.K = No; h.split_k = 1;
      if (int e = launch_gemm(h, s)) return e;
    }
  }
  return 0;
}

// jac[i, 3k..3k+3) = d(sum_j y[j,k]) / d points[i], k = 0..2 (vector_field_network.py:146-171), after a train-mode forward
// whose buffers are still in `b`.
static int vf_jacobian_train(const vfnerf_mlp_desc& d, const float* arena, const TrainBufs& b, int skip_layer, int multires,
                             const float* points, const float* emb, int64_t n, const float* y, int64_t y_ld,
                             const BwdBufs& w, float* d_emb0, float* d_tail, float* jac, int64_t jac_ld, cudaStream_t s) {
  const int L = d.n_layers, E = 3 + 6 * multires;
  const int Kl = d.in_dim[L - 1];            // == out_dim[L-2] (no skip into the last layer)
  for (int k = 0; k < 3; ++k) {
    jac_seed_kernel<<<(unsigned)ceil_div64(n * Kl, 256), 256, 0, s>>>(y, y_ld, k, arena + d.w_off[L - 1] + (int64_t)k * Kl,
                                                                      b.m.act[L - 2], Kl, n, Kl, w.dB);
    VFN_LAUNCH_CHECK();
    if (int e = mlp_backward_train(d, arena, b, skip_layer, emb, E, n, L - 2, w.dB, Kl, w, nullptr, d_emb0, E, 0, E,
                                   skip_layer > 0 ? d_tail : nullptr, s)) return e;
    embed_bwd_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(points, n, multires, d_emb0, skip_layer > 0 ? d_tail : nullptr,
                                                                  E, jac + 3 * k, jac_ld);
    VFN_LAUNCH_CHECK();
  }
  return 0;
}

static int vf_embed(const vfnerf_mlp_desc& d, const MlpBufs& b, int multires, int skip_layer, const float* points, int64_t n,
                    float* emb, cudaStream_t s) {
  const int E = 3 + 6 * multires;
  if (int e = launch_embed(points, 3, n, multires, 0.f, emb, E, s)) return e;
  if (skip_layer > 0)
    if (int e = launch_embed(points, 3, n, multires, kSqrt2, b.act[skip_layer - 1] + d.out_dim[skip_layer - 1],
                             d.in_dim[skip_layer], s)) return e;
  return 0;
}

// ---- render() in train mode ----------------------------------------------------------------------------------------
struct TrainPlan {
  int R, Nc, Nf, N, E, Ev, F, cin_ld;
  int64_t P, Pc;
  float *directions, *ray_dirs, *cam_loc, *z_c, *pts_c, *y_c, *jac_c, *w_c, *weights, *cin, *emb, *d_emb0, *d_tail;
  TrainBufs vf, rn;
  BwdBufs bw;
  float *d_out, *d_colors;
  int64_t bytes;
};

static int make_train_plan(const vfnerf_render_cfg& cfg, const vfnerf_mlp_desc& vf, const vfnerf_mlp_desc& rn, void* ws,
                           TrainPlan& p) {
  p.R = cfg.n_rays; p.Nc = cfg.n_coarse; p.Nf = cfg.n_fine; p.N = p.Nc + p.Nf;
  p.P = (int64_t)p.R * p.N; p.Pc = (int64_t)p.R * p.Nc;
  p.E = 3 + 6 * cfg.multires; p.Ev = 3 + 6 * cfg.multires_view;
  p.F = vf.out_dim[vf.n_layers - 1] - 3;
  p.cin_ld = 3 + p.Ev + 3 + p.F;
  VFN_REQUIRE(cfg.precision == VFNERF_PREC_FP32, "train mode (batch-statistic BatchNorm) runs on the fp32 layer-wise path: "
              "precision must be VFNERF_PREC_FP32");
  VFN_REQUIRE(p.R >= 0 && p.Nc >= 2 && p.Nf >= 2, "render_train: n_rays=%d n_coarse=%d n_fine=%d invalid", p.R, p.Nc, p.Nf);
  VFN_REQUIRE(p.N <= VFNERF_MAX_SAMPLES, "render_train: %d samples per ray exceed %d", p.N, VFNERF_MAX_SAMPLES);
  VFN_REQUIRE(!(cfg.flags & (VFNERF_FLAG_WHITE_BG | VFNERF_FLAG_NERF_WEIGHTS)), "render_train: the white-background / "
              "nerf-weights fixes are eval-mode options");
  if (int e = validate_vf(vf, cfg.multires, cfg.skip_layer)) return e;
  if (int e = validate_train(vf, cfg.skip_layer, "VF net")) return e;
  if (int e = validate_train(rn, -1, "colour net")) return e;
  VFN_REQUIRE(rn.in_dim[0] == p.cin_ld, "colour net: in_dim[0]=%d, expected %d (mode 'idr')", rn.in_dim[0], p.cin_ld);
  VFN_REQUIRE(rn.out_dim[rn.n_layers - 1] == 3, "colour net: output dim must be 3");
  Carver c(ws);
  p.directions = c.f(3 * p.R); p.ray_dirs = c.f(3 * p.R); p.cam_loc = c.f(3 * p.R);
  p.z_c = c.f(p.Pc); p.pts_c = c.f(3 * p.Pc); p.y_c = c.f(3 * p.Pc); p.jac_c = c.f(9 * p.Pc); p.w_c = c.f(p.Pc);
  p.weights = c.f(p.P);
  p.cin = c.f(p.P * p.cin_ld);
  p.emb = c.f(p.P * p.E);
  p.d_emb0 = c.f(p.Pc * p.E); p.d_tail = c.f(p.Pc * p.E);
  carve_train(c, vf, p.P, p.vf);
  carve_train(c, rn, p.P, p.rn);
  carve_bwd(c, vf, rn, p.P, p.bw);
  p.d_out = c.f(p.P * (3 + p.F));
  p.d_colors = c.f(3 * p.P);
  p.bytes = c.off;
  return 0;
}

// ---- VF-only call in train mode -----------------------------------------------------------------------------------------
struct VfTrainPlan {
  float *emb, *d_emb0, *d_tail, *d_pre;
  TrainBufs b;
  BwdBufs bw;
  int64_t bytes;
};
static void make_vf_train_plan(const vfnerf_mlp_desc& vf, int64_t n, int multires, void* ws, VfTrainPlan& p) {
  Carver c(ws);
  const int E = 3 + 6 * multires;
  p.emb = c.f(n * E); p.d_emb0 = c.f(n * E); p.d_tail = c.f(n * E);
  carve_train(c, vf, n, p.b);
  vfnerf_mlp_desc none{};
  carve_bwd(c, vf, none, n, p.bw);
  p.d_pre = c.f(n * vf.out_dim[vf.n_layers - 1]);
  p.bytes = c.off;
}

}  // namespace vfn

using namespace vfn;

extern "C" {

int64_t vfnerf_render_train_workspace_bytes(const vfnerf_render_cfg* cfg, const vfnerf_mlp_desc* vf,
                                            const vfnerf_mlp_desc* rn) {
  if (!cfg || !vf || !rn) { set_error("null argument"); return -1; }
  TrainPlan p;
  if (make_train_plan(*cfg, *vf, *rn, nullptr, p)) return -1;
  return p.bytes + 256;
}

int vfnerf_render_train_fwd(const vfnerf_render_cfg* cfg, const vfnerf_mlp_desc* vf, float* vf_arena,
                            const vfnerf_mlp_desc* rn, float* rn_arena, const float* density_params, const float* uv,
                            const float* pose, const float* intrinsics, const float* t_vals, const float* U1,
                            const float* U2, const float* U3, const float* z_override, const vfnerf_render_out* out,
                            float* dir_derivs, float bn_momentum, void* workspace, int64_t workspace_bytes, void* stream) {
  VFN_REQUIRE(cfg && vf && rn && out && vf_arena && rn_arena, "render_train_fwd: null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DeviceGuard dev_guard(s);
  TrainPlan p;
  if (int e = make_train_plan(*cfg, *vf, *rn, workspace, p)) return e;
  if (p.R == 0) return 0;
  VFN_REQUIRE(workspace && workspace_bytes >= p.bytes, "render_train_fwd: workspace %lld B < required %lld B",
              (long long)workspace_bytes, (long long)p.bytes);
  VFN_REQUIRE(out->points && out->normals && out->rgb && out->depth && out->z_vals && out->colors,
              "render_train_fwd: required output pointer is null");
  if (p.R == 0) return 0;
  float* weights = out->weights ? out->weights : p.weights;
  float* z_c = out->z_coarse ? out->z_coarse : p.z_c;
  float* w_c = out->weights_coarse ? out->weights_coarse : p.w_c;
  if (int e = launch_ray_geometry(p.R, cfg->pose_is_quat, uv, pose, intrinsics, p.directions, p.ray_dirs, p.cam_loc, s)) return e;
  if (int e = launch_coarse_sample(p.R, p.Nc, cfg->near_, cfg->far_, cfg->perturb, t_vals, U1, p.directions, p.cam_loc, z_c,
                                   p.pts_c, s)) return e;
  // ---- coarse pass (vector_field_nerf.py:252-272): batch statistics of the R*Nc coarse points (the running statistics
  // take this batch too, although the pass is under no_grad), the three seeds of the "Jacobian", directional derivatives
  if (int e = vf_embed(*vf, p.vf.m, cfg->multires, cfg->skip_layer, p.pts_c, p.Pc, p.emb, s)) return e;
  if (int e = mlp_forward_train(*vf, vf_arena, p.vf, cfg->skip_layer, p.emb, p.E, p.Pc, p.y_c, 3, 3, ACT_TANH, cfg->bn_eps,
                                bn_momentum, s)) return e;
  if (dir_derivs) {
    if (int e = vf_jacobian_train(*vf, vf_arena, p.vf, cfg->skip_layer, cfg->multires, p.pts_c, p.emb, p.Pc, p.y_c, 3, p.bw,
                                  p.d_emb0, p.d_tail, p.jac_c, 9, s)) return e;
    dir_derivs_kernel<<<(unsigned)ceil_div64(p.Pc, 256), 256, 0, s>>>(p.y_c, 3, p.jac_c, p.Pc, dir_derivs);
    VFN_LAUNCH_CHECK();
  }
  if (int e = launch_density_weights(*cfg, p.R, p.Nc, density_params, p.y_c, 3, p.ray_dirs, z_c, nullptr, nullptr, w_c, s)) return e;
  if (int e = launch_fine_sample(p.R, p.Nc, p.Nf, cfg->fine_near_, cfg->fine_far_, cfg->fine_range, cfg->perturb, z_c, w_c,
                                 U2, U3, z_override, p.directions, p.cam_loc, out->z_vals, out->points, nullptr, nullptr, s)) return e;
  // ---- merged pass (:289-323); its own "Jacobian" is computed and dropped upstream (:305 concatenates the coarse block
  // with itself), so it is not computed here
  if (int e = launch_color_input_head(out->points, p.ray_dirs, p.R, p.N, cfg->multires_view, p.cin, p.cin_ld,
                                      out->ray_dirs_rep, s)) return e;
  float* vf_out = p.cin + 3 + p.Ev;
  if (int e = vf_embed(*vf, p.vf.m, cfg->multires, cfg->skip_layer, out->points, p.P, p.emb, s)) return e;
  if (int e = mlp_forward_train(*vf, vf_arena, p.vf, cfg->skip_layer, p.emb, p.E, p.P, vf_out, p.cin_ld, 3 + p.F, ACT_TANH,
                                cfg->bn_eps, bn_momentum, s)) return e;
  if (int e = launch_copy_cols(vf_out, p.cin_ld, out->normals, 3, p.P, 3, s)) return e;
  if (int e = launch_density_weights(*cfg, p.R, p.N, density_params, vf_out, p.cin_ld, p.ray_dirs, out->z_vals, nullptr,
                                     nullptr, weights, s)) return e;
  if (int e = mlp_forward_train(*rn, rn_arena, p.rn, -1, p.cin, p.cin_ld, p.P, out->colors, 3, 3, ACT_SIGMOID, cfg->bn_eps,
                                bn_momentum, s)) return e;
  return launch_composite(p.R, p.N, weights, out->colors, out->z_vals, out->rgb, out->depth, s, 0);
}

int vfnerf_render_train_bwd(const vfnerf_render_cfg* cfg, const vfnerf_mlp_desc* vf, const float* vf_arena,
                            const vfnerf_mlp_desc* rn, const float* rn_arena, const float* density_params,
                            const vfnerf_render_out* out, const float* d_rgb, const float* d_depth, const float* d_normals,
                            const float* d_colors, float* vf_grad_arena, float* rn_grad_arena, float* d_density,
                            void* workspace, int64_t workspace_bytes, void* stream) {
  VFN_REQUIRE(cfg && vf && rn && out && d_rgb && d_depth && vf_grad_arena && rn_grad_arena && d_density,
              "render_train_bwd: null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DeviceGuard dev_guard(s);
  TrainPlan p;
  if (int e = make_train_plan(*cfg, *vf, *rn, workspace, p)) return e;
  VFN_REQUIRE(workspace && workspace_bytes >= p.bytes, "render_train_bwd: workspace %lld B < required %lld B",
              (long long)workspace_bytes, (long long)p.bytes);
  VFN_CHECK_CUDA(cudaMemsetAsync(d_density, 0, 3 * sizeof(float), s));
  VFN_CHECK_CUDA(cudaMemsetAsync(vf_grad_arena, 0, sizeof(float) * vf->arena_floats, s));
  VFN_CHECK_CUDA(cudaMemsetAsync(rn_grad_arena, 0, sizeof(float) * rn->arena_floats, s));
  if (p.R == 0) return 0;
  const float* vf_out = p.cin + 3 + p.Ev;
  const int Dv = 3 + p.F;
  if (int e = launch_render_tail_bwd(*cfg, p.R, p.N, density_params, vf_out, p.cin_ld, p.ray_dirs, out->z_vals, out->colors,
                                     d_rgb, d_depth, d_normals, d_colors, p.d_colors, p.d_out, Dv, d_density, s)) return e;
  if (int e = launch_act_bwd(out->colors, 3, p.d_colors, 3, p.P, 3, ACT_SIGMOID, p.d_colors, 3, s)) return e;
  if (int e = mlp_backward_train(*rn, rn_arena, p.rn, -1, p.cin, p.cin_ld, p.P, rn->n_layers - 1, p.d_colors, 3, p.bw,
                                 rn_grad_arena, p.d_out + 3, Dv, 3 + p.Ev + 3, p.F, nullptr, s)) return e;
  if (int e = launch_act_bwd(vf_out, p.cin_ld, p.d_out, Dv, p.P, Dv, ACT_TANH, p.d_out, Dv, s)) return e;
  return mlp_backward_train(*vf, vf_arena, p.vf, cfg->skip_layer, p.emb, p.E, p.P, vf->n_layers - 1, p.d_out, Dv, p.bw,
                            vf_grad_arena, nullptr, 0, 0, 0, nullptr, s);
}

int64_t vfnerf_vf_train_workspace_bytes(const vfnerf_mlp_desc* vf, int64_t n_points, int multires) {
  if (!vf) { set_error("null argument"); return -1; }
  VfTrainPlan p;
  make_vf_train_plan(*vf, n_points, multires, nullptr, p);
  return p.bytes + 256;
}

int vfnerf_vf_train_fwd(const vfnerf_mlp_desc* vf, float* vf_arena, int multires, int skip_layer, float bn_eps,
                        float bn_momentum, const float* points, int64_t n_points, float* out, int64_t out_ld,
                        int n_out_cols, float* jacobian, int64_t jacobian_ld, void* workspace, int64_t workspace_bytes,
                        void* stream) {
  VFN_REQUIRE(vf && vf_arena && out && points, "vf_train_fwd: null argument");
  if (int e = validate_vf(*vf, multires, skip_layer)) return e;
  if (int e = validate_train(*vf, skip_layer, "VF net")) return e;
  VFN_REQUIRE(n_out_cols >= 3 && n_out_cols <= vf->out_dim[vf->n_layers - 1], "vf_train_fwd: n_out_cols=%d invalid", n_out_cols);
  if (n_points == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DeviceGuard dev_guard(s);
  VfTrainPlan p;
  make_vf_train_plan(*vf, n_points, multires, workspace, p);
  VFN_REQUIRE(workspace && workspace_bytes >= p.bytes, "vf_train_fwd: workspace %lld B < required %lld B",
              (long long)workspace_bytes, (long long)p.bytes);
  if (int e = vf_embed(*vf, p.b.m, multires, skip_layer, points, n_points, p.emb, s)) return e;
  if (int e = mlp_forward_train(*vf, vf_arena, p.b, skip_layer, p.emb, 3 + 6 * multires, n_points, out, out_ld, n_out_cols,
                                ACT_TANH, bn_eps, bn_momentum, s)) return e;
  if (jacobian)
    if (int e = vf_jacobian_train(*vf, vf_arena, p.b, skip_layer, multires, points, p.emb, n_points, out, out_ld, p.bw,
                                  p.d_emb0, p.d_tail, jacobian, jacobian_ld, s)) return e;
  return 0;
}

int vfnerf_vf_train_bwd(const vfnerf_mlp_desc* vf, const float* vf_arena, int multires, int skip_layer, int64_t n_points,
                        const float* out, int64_t out_ld, const float* d_out, int64_t d_ld, int n_out_cols,
                        float* vf_grad_arena, int accumulate, void* workspace, int64_t workspace_bytes, void* stream) {
  VFN_REQUIRE(vf && vf_arena && out && d_out && vf_grad_arena, "vf_train_bwd: null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DeviceGuard dev_guard(s);
  const int Do = vf->out_dim[vf->n_layers - 1];
  VFN_REQUIRE(n_out_cols >= 1 && n_out_cols <= Do, "vf_train_bwd: n_out_cols invalid");
  VfTrainPlan p;
  make_vf_train_plan(*vf, n_points, multires, workspace, p);
  VFN_REQUIRE(workspace && workspace_bytes >= p.bytes, "vf_train_bwd: workspace too small");
  if (!accumulate) VFN_CHECK_CUDA(cudaMemsetAsync(vf_grad_arena, 0, sizeof(float) * vf->arena_floats, s));
  if (n_points == 0) return 0;
  if (n_out_cols < Do) VFN_CHECK_CUDA(cudaMemsetAsync(p.d_pre, 0, sizeof(float) * n_points * Do, s));
  if (int e = launch_act_bwd(out, out_ld, d_out, d_ld, n_points, n_out_cols, ACT_TANH, p.d_pre, Do, s)) return e;
  return mlp_backward_train(*vf, vf_arena, p.b, skip_layer, p.emb, 3 + 6 * multires, n_points, vf->n_layers - 1, p.d_pre, Do,
                            p.bw, vf_grad_arena, nullptr, 0, 0, 0, nullptr, s);
}

}  // extern "C"
