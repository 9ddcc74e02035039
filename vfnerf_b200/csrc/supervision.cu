// Supervision-point sampling of the reference trainer on the device (second half).
//   train/vector_field_nerf_train.py:180-216   which points are supervised and with what target
//   models/helpers/functions.py:75-157         get_border_indices_and_gt / sample_border_points /
//                                              sample_center_points / get_center_indices_and_gt
//   models/samplers/sampler.py:160-193         SphereSampler.sample (numpy, float64)
// The reference draws three float64 uniform arrays with numpy on the host, builds the points with float64 numpy
// arithmetic, casts to fp32 and uploads; and selects ray samples near the border / centre with a chain of small torch
// kernels.  Here the draws are an INPUT (the host mirror makes them with numpy in the reference's order, or on the
// device), the sphere arithmetic runs in float64 like numpy's, and the selection is one pass over the points.
// HBM-bound, a few hundred KB per training step: the point is fewer launches and no numpy in the step, not bandwidth.
#include "common.cuh"

namespace vfn {

__device__ __forceinline__ void normalize3(float x, float y, float z, float* o) {
  // F.normalize(v, dim=1): v / max(||v||_2, 1e-12)
  const float n = fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
  o[0] = x / n; o[1] = y / n; o[2] = z / n;
}

// phi in [0, 2 pi), cos_theta in [-1, 1), u in [0, 1): exactly what np.random.uniform hands SphereSampler.sample
__global__ void sphere_points_kernel(int64_t n, const double* __restrict__ phi, const double* __restrict__ cos_theta,
                                     const double* __restrict__ u, double r_max, double r_min, float cx, float cy, float cz,
                                     int inward, float* __restrict__ points, float* __restrict__ gt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double theta = acos(cos_theta[i]);                           // sampler.py:174
    const double r = cbrt(u[i]) * (r_max - r_min) + r_min;             // :177
    const double st = sin(theta), ph = phi[i];
    const double x = r * st * cos(ph), y = r * st * sin(ph), z = r * cos(theta);   // :180-182
    // functions.py:109 / :125: torch.from_numpy(...).float() + centroid (an fp32 add)
    const float px = __fadd_rn((float)x, cx), py = __fadd_rn((float)y, cy), pz = __fadd_rn((float)z, cz);
    points[3 * i] = px; points[3 * i + 1] = py; points[3 * i + 2] = pz;
    // border points look at the centroid (functions.py:112), centre points away from it (:128)
    if (inward) normalize3(__fsub_rn(cx, px), __fsub_rn(cy, py), __fsub_rn(cz, pz), gt + 3 * i);
    else normalize3(__fsub_rn(px, cx), __fsub_rn(py, cy), __fsub_rn(pz, cz), gt + 3 * i);
  }
}

// mode 0: border samples, distance to the centroid > threshold, target = normalize(centroid - p)   (functions.py:86-97)
// mode 1: centre samples, distance < threshold,                 target = normalize(p - centroid)   (functions.py:143-154)
__global__ void select_supervised_kernel(int64_t n, const float* __restrict__ points, float cx, float cy, float cz,
                                         float threshold, int mode, uint8_t* __restrict__ flag, float* __restrict__ gt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float dx = __fsub_rn(points[3 * i], cx), dy = __fsub_rn(points[3 * i + 1], cy), dz = __fsub_rn(points[3 * i + 2], cz);
    const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
    flag[i] = mode == 0 ? (dist > threshold) : (dist < threshold);
    if (mode == 0) normalize3(-dx, -dy, -dz, gt + 3 * i); else normalize3(dx, dy, dz, gt + 3 * i);
  }
}

}  // namespace vfn

using namespace vfn;

extern "C" int vfnerf_sphere_points(int64_t n, const double* phi, const double* cos_theta, const double* u, double r_max,
                                    double r_min, const float* centroid3_host, int inward, float* points, float* gt,
                                    void* stream) {
  VFN_REQUIRE(n >= 0 && centroid3_host && (n == 0 || (phi && cos_theta && u && points && gt)), "sphere_points: null argument");
  if (n == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DeviceGuard dev_guard(s);
  const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  sphere_points_kernel<<<grid, 256, 0, s>>>(n, phi, cos_theta, u, r_max, r_min, centroid3_host[0], centroid3_host[1],
                                            centroid3_host[2], inward, points, gt);
  VFN_LAUNCH_CHECK();
  return 0;
}

extern "C" int vfnerf_select_supervised(int64_t n, const float* points, const float* centroid3_host, float threshold, int mode,
                                        uint8_t* flag, float* gt, void* stream) {
  VFN_REQUIRE(n >= 0 && centroid3_host && (n == 0 || (points && flag && gt)), "select_supervised: null argument");
  VFN_REQUIRE(mode == 0 || mode == 1, "select_supervised: mode %d (0 border, 1 centre)", mode);
  if (n == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DeviceGuard dev_guard(s);
  const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  select_supervised_kernel<<<grid, 256, 0, s>>>(n, points, centroid3_host[0], centroid3_host[1], centroid3_host[2], threshold,
                                                mode, flag, gt);
  VFN_LAUNCH_CHECK();
  return 0;
}
