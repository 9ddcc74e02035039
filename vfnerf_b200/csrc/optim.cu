// Clip + Adam on the flat parameter arenas (trainer glue of train/vector_field_nerf_train.py:252-258).
// The reference's clip_grad_norm_ + Adam.step touch 55 tensors with ~25 foreach launches (0.55 ms of a 2 ms step on a
// B200).  Every parameter of a network lives in ONE flat fp32 arena here (networks.ParamArena) and so does its gradient,
// so the whole update is one squared-norm reduction and one elementwise launch per arena.
#include "common.cuh"

namespace vfn {

__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += g[i] * g[i];
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t);
  }
}

// torch.optim.Adam (amsgrad=False, maximize=False) with the gradient first scaled by the global clip coefficient
// min(1, max_norm / (||g||_2 + 1e-6)) of torch.nn.utils.clip_grad_norm_.  step[0] holds t (already incremented).
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, const uint8_t* __restrict__ mask, long long n,
                                                   const float* __restrict__ lr, const float* __restrict__ step, float beta1,
                                                   float beta2, float eps, float wd, float max_norm,
                                                   const float* __restrict__ total_sq) {
  const float t = step[0];
  const float bc1 = 1.f - powf(beta1, t), bc2 = 1.f - powf(beta2, t);
  const float step_size = lr[0] / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  float coef = 1.f;
  if (max_norm > 0.f) coef = fminf(1.f, max_norm / (sqrtf(total_sq[0]) + 1e-6f));
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (mask && !mask[i]) continue;                 // BatchNorm running statistics share the arena: not parameters
    float gi = g[i] * coef;
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

extern "C" int vfnerf_sqnorm_accumulate(const float* g, int64_t n, float* out_sq, void* stream) {
  VFN_REQUIRE(g && out_sq, "sqnorm: null argument");
  if (n <= 0) return 0;
  DeviceGuard dev_guard(reinterpret_cast<cudaStream_t>(stream));
  const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 4);
  sqnorm_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, n, out_sq);
  VFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vfnerf_adam_step(float* p, float* g, float* m, float* v, const uint8_t* mask, int64_t n, const float* lr,
                                const float* step, float beta1, float beta2, float eps, float weight_decay, float max_norm,
                                const float* total_sqnorm, void* stream) {
  VFN_REQUIRE(p && g && m && v && lr && step, "adam_step: null argument");
  VFN_REQUIRE(max_norm <= 0.f || total_sqnorm, "adam_step: clipping needs the squared gradient norm");
  if (n <= 0) return 0;
  DeviceGuard dev_guard(reinterpret_cast<cudaStream_t>(stream));
  const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  adam_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, g, m, v, mask, n, lr, step, beta1, beta2, eps,
                                                                       weight_decay, max_norm, total_sqnorm);
  VFN_LAUNCH_CHECK();
  return 0;
}
