// The per-ray stages of render() fused around the two MLP launches: three kernels
// instead of seven, one warp per ray, nothing re-read from HBM between stages.
//
//   ray_head_kernel        a1 + a2   rays -> directions / unit directions / origin, coarse z values and points
//   coarse_to_fine_kernel  a4-a7     coarse VF vectors -> windowed cosine -> density -> weights (registers) -> argmax ->
//                                    fine candidates -> value-exact merge sort -> merged z / points (/ candidate map)
//   render_tail_kernel     a4-a6, a9 (gather of the per-candidate MLP results into sample order) -> weights -> composite
//
// The device code is the stand-alone stage kernels' (ray_ops.cuh), so results are bit-identical to running
// vfnerf_ray_geometry / coarse_sample / density_weights / fine_sample / composite one after the other -- the stage entry
// points stay, and the tests compare the two.  HBM-bound byte work: per ray and 64+64 samples the three kernels move
// 0.3 + 1.0 + 2.6 KB (head) + 2.1 + 2.1 KB (coarse-to-fine: vectors, z in; merged z, points, fine points, map out) +
// 5.4 KB (tail: vectors, colours, map, z in; vectors, colours, weights out) -- each tensor once.
#include "common.cuh"
#include "ray_ops.cuh"

namespace vfn {

__global__ void __launch_bounds__(kRayWarps * 32)
ray_head_kernel(int n_rays, int pose_is_quat, const float* __restrict__ uv, const float* __restrict__ pose,
                const float* __restrict__ K, int n_coarse, float nearf, float farf, int perturb,
                const float* __restrict__ t_vals, const float* __restrict__ U1, float* __restrict__ directions,
                float* __restrict__ ray_dirs, float* __restrict__ cam_loc, float* __restrict__ z,
                float* __restrict__ points) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * kRayWarps + wid;
  if (r >= n_rays) return;
  float d[3], rd[3], o[3];
  ray_geometry_one(r, pose_is_quat, uv, pose, K, d, rd, o);      // warp-uniform: every lane holds the ray
  if (lane < 3) {
    directions[3 * r + lane] = d[lane];
    ray_dirs[3 * r + lane] = rd[lane];
    cam_loc[3 * r + lane] = o[lane];
  }
  for (int i = lane; i < n_coarse; i += 32) {
    const int64_t idx = r * n_coarse + i;
    const float zi = coarse_z_one(i, n_coarse, nearf, farf, perturb, t_vals, U1, idx);
    z[idx] = zi;
#pragma unroll
    for (int c = 0; c < 3; ++c) points[3 * idx + c] = __fadd_rn(o[c], __fmul_rn(zi, d[c]));
  }
}

int launch_ray_head(int n_rays, int pose_is_quat, const float* uv, const float* pose, const float* K, int n_coarse,
                    double near_, double far_, int perturb, const float* t_vals, const float* U1, float* directions,
                    float* ray_dirs, float* cam_loc, float* z, float* points, cudaStream_t s) {
  if (n_rays <= 0) return 0;
  VFN_REQUIRE(!perturb || U1 != nullptr, "ray_head: perturb=1 needs U1");
  ray_head_kernel<<<(n_rays + kRayWarps - 1) / kRayWarps, kRayWarps * 32, 0, s>>>(
      n_rays, pose_is_quat, uv, pose, K, n_coarse, (float)near_, (float)far_, perturb, t_vals, U1, directions, ray_dirs,
      cam_loc, z, points);
  VFN_LAUNCH_CHECK();
  return 0;
}

template <int KP>
__global__ void __launch_bounds__(kRayWarps * 32)
coarse_to_fine_kernel(vfnerf_render_cfg cfg, int n_rays, FineCfg fc, const float* __restrict__ dparams,
                      const float* __restrict__ normals_c, int64_t ld, const float* __restrict__ ray_dirs,
                      const float* __restrict__ z_c, const float* __restrict__ U2, const float* __restrict__ U3,
                      const float* __restrict__ directions, const float* __restrict__ cam_loc, float* __restrict__ w_c,
                      float* __restrict__ z_out, float* __restrict__ points, uint8_t* __restrict__ src,
                      float* __restrict__ points_fine) {
  __shared__ __align__(16) float s_u[kRayWarps][kUS * VFNERF_MAX_SAMPLES];
  __shared__ float sbuf[kRayWarps][VFNERF_MAX_SAMPLES];
  __shared__ float tbuf[kRayWarps][VFNERF_MAX_SAMPLES];
  __shared__ uint8_t sidx[kRayWarps][VFNERF_MAX_SAMPLES];
  __shared__ uint8_t tidx[kRayWarps][VFNERF_MAX_SAMPLES];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * kRayWarps + wid;
  if (r >= n_rays) return;
  const int Nc = fc.n_coarse, N = fc.n_coarse + fc.n_fine;
  float* su = s_u[wid];
  float* s = sbuf[wid];
  float* t = tbuf[wid];
  const Laplace lap = load_laplace(cfg, dparams);
  const Window win = make_window(cfg.window, Nc);
  stage_unit_vectors(normals_c + r * Nc * ld, ld, Nc, lane, su, nullptr);
  for (int j = lane; j < Nc; j += 32) s[j] = z_c[r * Nc + j];
  float d[3];
  unit_dir(ray_dirs, r, d);
  __syncwarp();
  float what[KP];
  const float inv = ray_weights<KP>(cfg, lap, win, su, d, s, Nc, lane, nullptr, nullptr, what);
  // argmax of the coarse weights (the values the reference holds: unnormalised weight times the normalisation factor),
  // first index on ties
  float best = -INFINITY;
  int bi = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < KP; ++i) {
    const int j = lane + 32 * i;
    if (j < Nc) {
      const float w = what[i] * inv;
      if (w_c) w_c[r * Nc + j] = w;
      if (w > best) { best = w; bi = j; }
    }
  }
  warp_argmax_first(best, bi);
  const float z_star = s[bi];
  const float dv[3] = {directions[3 * r], directions[3 * r + 1], directions[3 * r + 2]};
  const float o[3] = {cam_loc[3 * r], cam_loc[3 * r + 1], cam_loc[3 * r + 2]};
  fine_candidates_sorted(fc, r, bi, z_star, U2, U3, o, dv, s, sidx[wid], t, tidx[wid], points_fine, lane);
  if (src) for (int j = lane; j < N; j += 32) src[r * N + j] = tidx[wid][j];
  __syncwarp();
  write_merged_samples(t, N, r, o, dv, z_out, points, lane);
}

int launch_coarse_to_fine(const vfnerf_render_cfg& cfg, int n_rays, int n_coarse, int n_fine, const float* density_params,
                          const float* normals_c, int64_t normals_ld, const float* ray_dirs, const float* z_c,
                          const float* U2, const float* U3, const float* directions, const float* cam_loc, float* w_c,
                          float* z, float* points, uint8_t* src, float* points_fine, cudaStream_t s) {
  if (n_rays <= 0) return 0;
  VFN_REQUIRE(n_coarse >= 2 && n_coarse + n_fine <= VFNERF_MAX_SAMPLES, "coarse_to_fine: n_coarse=%d n_fine=%d out of range",
              n_coarse, n_fine);
  VFN_REQUIRE(n_fine >= 2, "coarse_to_fine: n_fine must be >= 2 (the reference divides by n_fine-1)");
  VFN_REQUIRE(cfg.window >= 1 && cfg.window <= 63, "coarse_to_fine: window=%d unsupported", cfg.window);
  VFN_REQUIRE(U3 && (!cfg.perturb || U2), "coarse_to_fine: missing uniform draws");
  FineCfg fc;
  fc.n_coarse = n_coarse; fc.n_fine = n_fine; fc.perturb = cfg.perturb;
  fc.nearf = (float)cfg.fine_near_; fc.far_minus_near = (float)(cfg.fine_far_ - cfg.fine_near_);
  fc.rangef = (float)cfg.fine_range;
  fc.stepf = (float)(2.0 * cfg.fine_range / (double)(n_fine - 1));   // python double, then fp32
  const dim3 grid((n_rays + kRayWarps - 1) / kRayWarps), block(kRayWarps * 32);
#define VFN_CF(KP) coarse_to_fine_kernel<KP><<<grid, block, 0, s>>>( \
      cfg, n_rays, fc, density_params, normals_c, normals_ld, ray_dirs, z_c, U2, U3, directions, cam_loc, w_c, z, points, src, \
      points_fine)
  if (n_coarse <= 64) VFN_CF(2); else if (n_coarse <= 128) VFN_CF(4); else VFN_CF(kMaxPerLane);
#undef VFN_CF
  VFN_LAUNCH_CHECK();
  return 0;
}

// src != NULL: normals / colors are the per-candidate results in EVALUATION order (all coarse candidates of all rays,
// then all fine candidates; src[r, j] = candidate at merged position j) and are written out in sample order;
// src == NULL: they already are in sample order (normals with row stride ld) and are only read.
template <int KP>
__global__ void __launch_bounds__(kRayWarps * 32)
render_tail_kernel(vfnerf_render_cfg cfg, int n_rays, int N, int n_coarse, const float* __restrict__ dparams,
                   const uint8_t* __restrict__ src, const float* __restrict__ normals, int64_t ld,
                   const float* __restrict__ colors, const float* __restrict__ ray_dirs, const float* __restrict__ z,
                   float* __restrict__ out_normals, float* __restrict__ out_colors, float* __restrict__ weights,
                   float* __restrict__ rgb, float* __restrict__ depth, float* __restrict__ rep_dirs, int white) {
  __shared__ __align__(16) float s_u[kRayWarps][kUS * VFNERF_MAX_SAMPLES];
  __shared__ float s_z[kRayWarps][VFNERF_MAX_SAMPLES];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * kRayWarps + wid;
  if (r >= n_rays) return;
  float* su = s_u[wid];
  float* sz = s_z[wid];
  const Laplace lap = load_laplace(cfg, dparams);
  const Window win = make_window(cfg.window, N);
  const int n_fine = N - n_coarse;
  const float rd[3] = {ray_dirs[3 * r], ray_dirs[3 * r + 1], ray_dirs[3 * r + 2]};
  float col[KP][3];
#pragma unroll
  for (int i = 0; i < KP; ++i) {
    const int j = lane + 32 * i;
    col[i][0] = col[i][1] = col[i][2] = 0.f;
    if (j < N) {
      const int64_t idx = r * N + j;
      int64_t row = idx;
      if (src) {
        const int k = src[idx];
        row = k < n_coarse ? r * n_coarse + k : (int64_t)n_rays * n_coarse + r * n_fine + (k - n_coarse);
      }
      const float* p = normals + row * ld;
      const float x = p[0], y = p[1], zz = p[2];
      col[i][0] = colors[3 * row]; col[i][1] = colors[3 * row + 1]; col[i][2] = colors[3 * row + 2];
      if (src) {
        out_normals[3 * idx] = x; out_normals[3 * idx + 1] = y; out_normals[3 * idx + 2] = zz;
        out_colors[3 * idx] = col[i][0]; out_colors[3 * idx + 1] = col[i][1]; out_colors[3 * idx + 2] = col[i][2];
      }
      if (rep_dirs) { rep_dirs[3 * idx] = rd[0]; rep_dirs[3 * idx + 1] = rd[1]; rep_dirs[3 * idx + 2] = rd[2]; }
      const float n = fmaxf(sqrtf(x * x + y * y + zz * zz), 1e-8f);
      *reinterpret_cast<float4*>(su + kUS * j) = make_float4(x / n, y / n, zz / n, 0.f);
      sz[j] = z[idx];
    }
  }
  float d[3];
  unit_dir(ray_dirs, r, d);
  __syncwarp();
  float what[KP];
  const float inv = ray_weights<KP>(cfg, lap, win, su, d, sz, N, lane, nullptr, nullptr, what);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, ad = 0.f, aw = 0.f;
#pragma unroll
  for (int i = 0; i < KP; ++i) {
    const int j = lane + 32 * i;
    if (j < N) {
      const float wj = what[i] * inv;
      if (weights) weights[r * N + j] = wj;
      a0 += wj * col[i][0]; a1 += wj * col[i][1]; a2 += wj * col[i][2];
      ad += wj * sz[j];
      aw += wj;
    }
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); ad = warp_sum(ad);
  if (white) {             // rgb + (1 - acc_map), vector_field_nerf.py:325-329
    aw = warp_sum(aw);
    const float bg = 1.f - aw;
    a0 = a0 + bg; a1 = a1 + bg; a2 = a2 + bg;
  }
  if (lane == 0) {
    rgb[3 * r] = a0; rgb[3 * r + 1] = a1; rgb[3 * r + 2] = a2;
    depth[r] = ad;
  }
}

int launch_render_tail(const vfnerf_render_cfg& cfg, int n_rays, int n_samples, int n_coarse, const float* density_params,
                       const uint8_t* src, const float* normals, int64_t normals_ld, const float* colors,
                       const float* ray_dirs, const float* z, float* out_normals, float* out_colors, float* weights,
                       float* rgb, float* depth, float* rep_dirs, int white, cudaStream_t s) {
  if (n_rays <= 0) return 0;
  VFN_REQUIRE(n_samples >= 2 && n_samples <= VFNERF_MAX_SAMPLES, "render_tail: n_samples=%d out of [2,%d]", n_samples,
              VFNERF_MAX_SAMPLES);
  VFN_REQUIRE(cfg.window >= 1 && cfg.window <= 63, "render_tail: window=%d unsupported", cfg.window);
  VFN_REQUIRE(!src || (out_normals && out_colors), "render_tail: the gather needs its two outputs");
  const dim3 grid((n_rays + kRayWarps - 1) / kRayWarps), block(kRayWarps * 32);
#define VFN_RT(KP) render_tail_kernel<KP><<<grid, block, 0, s>>>( \
      cfg, n_rays, n_samples, n_coarse, density_params, src, normals, normals_ld, colors, ray_dirs, z, out_normals, out_colors, \
      weights, rgb, depth, rep_dirs, white)
  if (n_samples <= 64) VFN_RT(2); else if (n_samples <= 128) VFN_RT(4); else VFN_RT(kMaxPerLane);
#undef VFN_RT
  VFN_LAUNCH_CHECK();
  return 0;
}

}  // namespace vfn
