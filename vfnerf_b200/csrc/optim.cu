// Clip + Adam on the flat parameter arenas (trainer glue of train/vector_field_nerf_train.py:252-258).
// The reference's clip_grad_norm_ + Adam.step touch 55 tensors with ~25 foreach launches (0.55 ms of a 2 ms step on a
// B200).  Every parameter of a network lives in ONE flat fp32 arena here (networks.ParamArena) and so does its gradient,
// so the whole update is one squared-norm reduction and one elementwise launch per arena.
#include "common.cuh"

namespace vfn {

// Squared 2-norm with a reproducible result: every block reduces its grid-stride share in a fixed order and stores ONE
// partial; the block that takes the last ticket adds the partials up in index order.  (Float atomics would make the clip
// coefficient -- and with it every parameter after the step -- depend on block scheduling; a resumed run could then not
// continue bit for bit.)  scratch: kSqnormBlocks partials + the ticket counter, zero before the first call and left zero.
constexpr int kSqnormBlocks = 592;          // 148 SMs x 4
__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ g, long long n, float pre_scale,
                                                     float* __restrict__ out, float* __restrict__ scratch) {
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = g[i] * pre_scale;
    s += v * v;
  }
  __shared__ float red[8];
  __shared__ bool last;
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + kSqnormBlocks);
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    scratch[blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float t = 0.f;                               // fixed order: thread k takes partials k, k+256, ...; then a fixed tree
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) t += __ldcg(scratch + b);
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    out[0] += tot;                             // calls are stream-ordered: no other writer
    *ticket = 0u;
  }
}

// torch.optim.Adam (amsgrad=False, maximize=False) with the gradient first scaled by pre_scale (1/world after a summing
// all-reduce; the norm in total_sq is of the scaled gradient) and by the global clip coefficient
// min(1, max_norm / (||g||_2 + 1e-6)) of torch.nn.utils.clip_grad_norm_.  step[0] holds t (already incremented).
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, const uint8_t* __restrict__ mask, long long n,
                                                   const float* __restrict__ lr, const float* __restrict__ step, float beta1,
                                                   float beta2, float eps, float wd, float max_norm,
                                                   const float* __restrict__ total_sq, float pre_scale) {
  const float t = step[0];
  const float bc1 = 1.f - powf(beta1, t), bc2 = 1.f - powf(beta2, t);
  const float step_size = lr[0] / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  float coef = 1.f;
  if (max_norm > 0.f) coef = fminf(1.f, max_norm / (sqrtf(total_sq[0]) + 1e-6f));
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (mask && !mask[i]) continue;                 // BatchNorm running statistics share the arena: not parameters
    float gi = g[i] * pre_scale * coef;
    g[i] = gi;                                      // like clip_grad_norm_, the stored gradient is the clipped one
    const float pi = p[i];
    gi += wd * pi;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] = pi - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}

}  // namespace vfn

using namespace vfn;

extern "C" int vfnerf_sqnorm_accumulate(const float* g, int64_t n, float pre_scale, float* out_sq, float* scratch, void* stream) {
  VFN_REQUIRE(g && out_sq && scratch, "sqnorm: null argument");
  static_assert(VFNERF_SQNORM_SCRATCH_FLOATS >= kSqnormBlocks + 1, "scratch too small");
  if (n <= 0) return 0;
  DeviceGuard dev_guard(reinterpret_cast<cudaStream_t>(stream));
  const int grid = (int)std::min<int64_t>((n + 255) / 256, kSqnormBlocks);
  sqnorm_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, n, pre_scale, out_sq, scratch);
  VFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vfnerf_adam_step(float* p, float* g, float* m, float* v, const uint8_t* mask, int64_t n, const float* lr,
                                const float* step, float beta1, float beta2, float eps, float weight_decay, float max_norm,
                                const float* total_sqnorm, float pre_scale, void* stream) {
  VFN_REQUIRE(p && g && m && v && lr && step, "adam_step: null argument");
  VFN_REQUIRE(max_norm <= 0.f || total_sqnorm, "adam_step: clipping needs the squared gradient norm");
  if (n <= 0) return 0;
  DeviceGuard dev_guard(reinterpret_cast<cudaStream_t>(stream));
  const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  adam_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, g, m, v, mask, n, lr, step, beta1, beta2, eps,
                                                                       weight_decay, max_norm, total_sqnorm, pre_scale);
  VFN_LAUNCH_CHECK();
  return 0;
}
