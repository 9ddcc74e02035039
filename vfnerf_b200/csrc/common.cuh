// Shared helpers for the vfnerf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vfnerf_b200.h"

namespace vfn {

// thread-local error text behind vfnerf_last_error()
void set_error(const char* fmt, ...);

#define VFN_CHECK_CUDA(expr)                                                            \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      vfn::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                         \
    }                                                                                   \
  } while (0)

#define VFN_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      vfn::set_error(__VA_ARGS__);    \
      return 2;                       \
    }                                 \
  } while (0)

// every kernel launch goes through this macro: it also feeds vfnerf_launch_count()
extern long long g_launches;
#define VFN_LAUNCH_CHECK()               \
  do {                                   \
    ++vfn::g_launches;                   \
    VFN_CHECK_CUDA(cudaGetLastError());  \
  } while (0)

// Every C-ABI entry point runs on the device that owns its stream: kernels launch on the calling thread's CURRENT device,
// which in a multi-GPU process need not be the device of the caller's pointers and stream (the reference accepts any
// config.cuda_config.device).  NULL / legacy / per-thread streams belong to the current device by definition.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(cudaStream_t s) {
    if (s == nullptr || s == cudaStreamLegacy || s == cudaStreamPerThread) return;
    // a capturing stream must not be queried for its device (that invalidates the capture); a capture is driven from
    // the thread whose current device owns the stream anyway
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cap) != cudaSuccess) { cudaGetLastError(); return; }
    if (cap != cudaStreamCaptureStatusNone) return;
    int cur = 0, dev = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) return;
    if (cudaStreamGetDevice(s, &dev) != cudaSuccess) { cudaGetLastError(); return; }
    if (dev != cur && cudaSetDevice(dev) == cudaSuccess) prev = cur;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxDevices = 64;     // per-device caches (SM count, shared-memory opt-in) are indexed by device ordinal

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// inclusive prefix sum over the 32 lanes of a warp
__device__ __forceinline__ float warp_inclusive_scan(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// ---- kernels / launchers implemented across the translation units --------------------------

// geometry_sampler.cu
int launch_ray_geometry(int n_rays, int pose_is_quat, const float* uv, const float* pose,
                        const float* K, float* directions, float* ray_dirs, float* cam_loc,
                        cudaStream_t s);
int launch_coarse_sample(int n_rays, int n_coarse, double near_, double far_, int perturb,
                         const float* t_vals, const float* U1, const float* directions,
                         const float* cam_loc, float* z, float* points, cudaStream_t s);
int launch_fine_sample(int n_rays, int n_coarse, int n_fine, double near_, double far_,
                       double fine_range, int perturb, const float* z_coarse, const float* w_coarse,
                       const float* U2, const float* U3, const float* z_override,
                       const float* directions, const float* cam_loc, float* z, float* points,
                       uint8_t* src, float* points_fine, cudaStream_t s);
int launch_sample_pdf(int n_rays, int n_bins, int n_samples, const float* bins, const float* weights, const float* u,
                      int u_per_ray, float* samples, cudaStream_t s);
int launch_pdf_fine_sample(int n_rays, int n_coarse, int n_fine, const float* z_coarse, const float* w_coarse,
                           const float* u, int u_per_ray, const float* directions, const float* cam_loc, float* z,
                           float* points, cudaStream_t s);
int launch_rows_to_merged_order(int n_rays, int n_coarse, int n_fine, const uint8_t* src, const float* in, float* out,
                                int cols, cudaStream_t s);
// out[r, j, :] = src[r,j] < n_coarse ? coarse[r, src, :] : fine[r, src - n_coarse, :] for two [.,3] tensor pairs
int launch_merge_samples(int n_rays, int n_coarse, int n_fine, const uint8_t* src, const float* a_coarse,
                         const float* a_fine, float* a_out, const float* b_coarse, const float* b_fine,
                         float* b_out, cudaStream_t s);

int launch_volume_weights(int n_rays, int n_samples, int mode, int normalize, const float* sigma, const float* z,
                          float* weights, cudaStream_t s);
// mc_preprocess.cu
int launch_mc_count(const float* pred, int N, uint8_t* keep, int* cta_counts, float* div_raw, uint8_t* choice,
                    const uint8_t* surface, cudaStream_t s);
int launch_mc_emit(const float* pred, int N, const uint8_t* keep, const int64_t* cta_offsets, int* cells, float* comb,
                   float* udf, cudaStream_t s);

int launch_smooth_vf(const float* in, float* tmp, float* out, int N, int k, const float* w_host, cudaStream_t s);

// render_fused.cu: the per-ray stages fused around the MLP launches (same device code as the stage kernels)
int launch_ray_head(int n_rays, int pose_is_quat, const float* uv, const float* pose, const float* K, int n_coarse,
                    double near_, double far_, int perturb, const float* t_vals, const float* U1, float* directions,
                    float* ray_dirs, float* cam_loc, float* z, float* points, cudaStream_t s);
int launch_coarse_to_fine(const vfnerf_render_cfg& cfg, int n_rays, int n_coarse, int n_fine, const float* density_params,
                          const float* normals_c, int64_t normals_ld, const float* ray_dirs, const float* z_c,
                          const float* U2, const float* U3, const float* directions, const float* cam_loc, float* w_c,
                          float* z, float* points, uint8_t* src, float* points_fine, cudaStream_t s);
int launch_render_tail(const vfnerf_render_cfg& cfg, int n_rays, int n_samples, int n_coarse, const float* density_params,
                       const uint8_t* src, const float* normals, int64_t normals_ld, const float* colors,
                       const float* ray_dirs, const float* z, float* out_normals, float* out_colors, float* weights,
                       float* rgb, float* depth, float* rep_dirs, int white, cudaStream_t s);

// density_composite.cu
int launch_density_weights(const vfnerf_render_cfg& cfg, int n_rays, int n_samples,
                           const float* density_params, const float* normals, int64_t normals_ld,
                           const float* ray_dirs, const float* z, float* cosw, float* sigma,
                           float* weights, cudaStream_t s);
int launch_composite(int n_rays, int n_samples, const float* weights, const float* colors,
                     const float* z, float* rgb, float* depth, cudaStream_t s, int white = 0);
// fused backward of composite + volsdf weights + Laplace density + windowed cosine
int launch_render_tail_bwd(const vfnerf_render_cfg& cfg, int n_rays, int n_samples,
                           const float* density_params, const float* normals, int64_t normals_ld,
                           const float* ray_dirs, const float* z, const float* colors,
                           const float* d_rgb, const float* d_depth, const float* d_normals_up,
                           const float* d_colors_up, float* d_colors, float* d_normals,
                           int64_t d_normals_ld, float* d_density, cudaStream_t s,
                           const uint8_t* src = nullptr, int n_coarse = 0);

// mlp_simt.cu
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2, ACT_SIGMOID = 3 };
struct GemmArgs {
  // C[m,n] (+)= sum_k A(m,k) * B(k,n);  A(m,k) = A[m*a_rs + k*a_cs] * (a_kscale ? a_kscale[k] : 1)
  const float* A; int64_t a_rs, a_cs; const float* a_kscale;
  const float* B; int64_t b_rs, b_cs;
  float* C; int64_t c_rs;             // C is row-major with row stride c_rs, written at columns [0,N)
  int64_t M; int N; int64_t K;
  // epilogue: v = acc * scale[n] + shift[n]; v = act(v); v = v / post_div (if post_div != 0);
  //           if (mask) v = mask[m*mask_rs + n] > 0 ? v : 0
  const float* scale; const float* shift; int act; float post_div;
  const float* mask; int64_t mask_rs;
  int split_k;                        // >1: partial sums are atomically added into C (C pre-zeroed)
  int accumulate;                     // !=0: C += result (atomic adds, plain epilogue only) also when split_k == 1
};
int launch_gemm(const GemmArgs& g, cudaStream_t s);
int launch_fold_bn(const vfnerf_mlp_desc& d, const float* arena, float bn_eps, float* scale,
                   float* shift, cudaStream_t s);   // scale/shift: concatenated over layers
int launch_embed(const float* x, int64_t x_ld, int64_t n, int multires, float div, float* out,
                 int64_t out_ld, cudaStream_t s);
int launch_color_input_head(const float* points, const float* ray_dirs, int n_rays, int n_samples,
                            int multires_view, float* color_in, int64_t ld, float* ray_dirs_rep,
                            cudaStream_t s);
int launch_act_bwd(const float* y, int64_t y_ld, const float* dy, int64_t dy_ld, int64_t n, int cols,
                   int act, float* out, int64_t out_ld, cudaStream_t s);
int launch_colsum(const float* a, int64_t a_rs, int64_t rows, int cols, const float* colscale,
                  float* out, cudaStream_t s);
int launch_grad_finalize(const vfnerf_mlp_desc& d, int layer, const float* arena, float bn_eps,
                         const float* colsum_dy, float* grad_arena, int accumulate,
                         const float* G_tmp, cudaStream_t s);
int launch_copy_cols(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int64_t rows,
                     int cols, cudaStream_t s);
struct GridSpec { float origin[3], translation[3], centroid[3], voxel; };
int launch_grid_points(int res, int64_t i0, int64_t n, const GridSpec& gs, float* pts, cudaStream_t s);

}  // namespace vfn
