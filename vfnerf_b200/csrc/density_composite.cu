// Density, transmittance and compositing and their fused backward.
//
// One warp owns one ray: N <= 256 samples, lane l holds samples l, l+32, ... (coalesced row access).
// Unit VF vectors are staged in the warp's shared-memory slice so the 11-tap windowed cosine is a
// shared-memory stencil; transmittance is a chunked warp scan; all sums are warp shuffles.
// HBM-bound: forward reads 12N+4N bytes/ray and writes 4N..12N; nothing is re-read.
#include "common.cuh"
#include "ray_ops.cuh"

namespace vfn {

// ---------------------------------------------------------------------------------------------
// forward: c, sigma, weights
// ---------------------------------------------------------------------------------------------
// KP = samples per lane (N <= 32 * KP): instantiated for 2, 4 and 8 so a 128-sample ray does not pay for 256.
template <int KP>
__global__ void __launch_bounds__(kRayWarps * 32)
density_weights_kernel(vfnerf_render_cfg cfg, int n_rays, int N, const float* __restrict__ dparams,
                       const float* __restrict__ normals, int64_t ld, const float* __restrict__ ray_dirs,
                       const float* __restrict__ z, float* __restrict__ cosw, float* __restrict__ sigma_out,
                       float* __restrict__ weights) {
  __shared__ __align__(16) float s_u[kRayWarps][kUS * VFNERF_MAX_SAMPLES];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r = blockIdx.x * kRayWarps + wid;
  if (r >= n_rays) return;
  float* su = s_u[wid];
  const Laplace lap = load_laplace(cfg, dparams);
  const Window win = make_window(cfg.window, N);
  stage_unit_vectors(normals + (int64_t)r * N * ld, ld, N, lane, su, nullptr);
  float d[3];
  unit_dir(ray_dirs, r, d);
  __syncwarp();
  float what[KP];
  const float inv = ray_weights<KP>(cfg, lap, win, su, d, z + (int64_t)r * N, N, lane,
                                    cosw ? cosw + (int64_t)r * (N - 1) : nullptr,
                                    sigma_out ? sigma_out + (int64_t)r * N : nullptr, what);
#pragma unroll
  for (int i = 0; i < KP; ++i) {
    const int j = lane + 32 * i;
    if (j < N) weights[(int64_t)r * N + j] = what[i] * inv;
  }
}

int launch_density_weights(const vfnerf_render_cfg& cfg, int n_rays, int n_samples,
                           const float* density_params, const float* normals, int64_t normals_ld,
                           const float* ray_dirs, const float* z, float* cosw, float* sigma,
                           float* weights, cudaStream_t s) {
  if (n_rays <= 0) return 0;
  VFN_REQUIRE(n_samples >= 2 && n_samples <= VFNERF_MAX_SAMPLES, "density_weights: n_samples=%d out of [2,%d]",
              n_samples, VFNERF_MAX_SAMPLES);
  VFN_REQUIRE(cfg.window >= 1 && cfg.window <= 63, "density_weights: window=%d unsupported", cfg.window);
  const dim3 grid((n_rays + kRayWarps - 1) / kRayWarps), block(kRayWarps * 32);
#define VFN_DW(KP) density_weights_kernel<KP><<<grid, block, 0, s>>>( \
      cfg, n_rays, n_samples, density_params, normals, normals_ld, ray_dirs, z, cosw, sigma, weights)
  if (n_samples <= 64) VFN_DW(2); else if (n_samples <= 128) VFN_DW(4); else VFN_DW(kMaxPerLane);
#undef VFN_DW
  VFN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Stand-alone compositing weights from a given density (utils/rendering.py): one warp per ray, chunked warp scans.
//   mode 0  volsdf_volume_rendering (:122-148): w_j = (1 - exp(-E_j)) * exp(-sum_{i<j} E_i),  E_j = (z_{j+1} - z_j) * sigma_j,
//           last interval 1e10 -- what density_weights_kernel fuses behind the density
//   mode 1  nerf_volume_rendering (:98-119): w_j = a_j * prod_{i<=j} (1 - a_i + 1e-10), a_j = 1 - exp(-E_j) (the reference's
//           cumprod is INCLUSIVE of sample j).  render() upstream passes this function its arguments swapped
//           (SURVEY.md §8a), so it is offered as the corrected stand-alone op only.
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(kRayWarps * 32)
volume_weights_kernel(int n_rays, int N, int mode, int normalize, const float* __restrict__ sigma,
                      const float* __restrict__ z, float* __restrict__ weights) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r = blockIdx.x * kRayWarps + wid;
  if (r >= n_rays) return;
  const float* zr = z + (int64_t)r * N;
  const float* sr = sigma + (int64_t)r * N;
  float* wr = weights + (int64_t)r * N;
  float carry = mode ? 1.f : 0.f, wsum = 0.f;
  for (int j0 = 0; j0 < N; j0 += 32) {
    const int j = j0 + lane;
    float E = 0.f;
    if (j < N) E = ((j < N - 1) ? (zr[j + 1] - zr[j]) : 1e10f) * sr[j];
    const float a = 1.f - expf(-E);
    float w;
    if (mode == 0) {
      // exclusive prefix by shifting the inclusive scan (inc - E would cancel catastrophically on the 1e10 interval)
      const float inc = warp_inclusive_scan(E, lane);
      float exc = __shfl_up_sync(kFull, inc, 1);
      if (lane == 0) exc = 0.f;
      w = a * expf(-(carry + exc));
      carry += __shfl_sync(kFull, inc, 31);
    } else {
      const float f = (j < N) ? (1.f - a + 1e-10f) : 1.f;
      const float inc = warp_inclusive_prod(f, lane);
      w = a * (carry * inc);
      carry *= __shfl_sync(kFull, inc, 31);
    }
    if (j < N) { if (normalize) wsum += w; else wr[j] = w; }
    if (normalize && j < N) wr[j] = w;
  }
  if (normalize) {
    wsum = warp_sum(wsum);
    const float inv = 1.f / (wsum + 1e-5f);
    __syncwarp();
    for (int j = lane; j < N; j += 32) wr[j] *= inv;
  }
}

int launch_volume_weights(int n_rays, int n_samples, int mode, int normalize, const float* sigma, const float* z,
                          float* weights, cudaStream_t s) {
  if (n_rays <= 0 || n_samples <= 0) return 0;
  VFN_REQUIRE(sigma && z && weights, "volume_weights: null argument");
  VFN_REQUIRE(mode == 0 || mode == 1, "volume_weights: mode %d (0 volsdf, 1 nerf)", mode);
  volume_weights_kernel<<<(n_rays + kRayWarps - 1) / kRayWarps, kRayWarps * 32, 0, s>>>(n_rays, n_samples, mode, normalize,
                                                                                     sigma, z, weights);
  VFN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// a9: rgb = sum_j w_j c_j, depth = sum_j w_j z_j
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRayWarps * 32)
composite_kernel(int n_rays, int N, const float* __restrict__ w, const float* __restrict__ colors,
                 const float* __restrict__ z, float* __restrict__ rgb, float* __restrict__ depth, int white) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r = blockIdx.x * kRayWarps + wid;
  if (r >= n_rays) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, ad = 0.f, aw = 0.f;
  for (int j = lane; j < N; j += 32) {
    int64_t idx = (int64_t)r * N + j;
    float wj = w[idx];
    a0 += wj * colors[3 * idx]; a1 += wj * colors[3 * idx + 1]; a2 += wj * colors[3 * idx + 2];
    ad += wj * z[idx];
    aw += wj;
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); ad = warp_sum(ad);
  if (white) {             // rgb + (1 - acc_map), vector_field_nerf.py:325-329
    aw = warp_sum(aw);
    const float bg = 1.f - aw;
    a0 = a0 + bg; a1 = a1 + bg; a2 = a2 + bg;
  }
  if (lane == 0) {
    rgb[3 * (int64_t)r] = a0; rgb[3 * (int64_t)r + 1] = a1; rgb[3 * (int64_t)r + 2] = a2;
    depth[r] = ad;
  }
}

int launch_composite(int n_rays, int n_samples, const float* weights, const float* colors,
                     const float* z, float* rgb, float* depth, cudaStream_t s, int white) {
  if (n_rays <= 0) return 0;
  composite_kernel<<<(n_rays + kRayWarps - 1) / kRayWarps, kRayWarps * 32, 0, s>>>(
      n_rays, n_samples, weights, colors, z, rgb, depth, white);
  VFN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// fused backward of rows a9 -> a6 -> a5 -> a4.  Recomputes the forward quantities from the VF
// vectors (cheaper than stashing c / sigma / T) and produces
//   d_colors [R*N,3] = w_j * d_rgb (+ upstream),   d_normals [R,N,ld] = dL/dv_j (+ upstream),
//   d_density[3] += d(beta, scale, mean)   (atomicAdd of one value per ray; caller zeroes it).
// ---------------------------------------------------------------------------------------------
// Row that receives the gradient of merged sample idx = r*N + j.  Without a map: idx itself.  With the candidate map of
// the fine sampler (src[idx] = k: coarse candidate k < n_coarse, fine candidate k - n_coarse otherwise) the rows follow the
// order in which the forward evaluated the points: all coarse candidates of all rays, then all fine candidates.
__device__ __forceinline__ int64_t out_row(const uint8_t* __restrict__ src, int64_t idx, int r, int n_rays, int N, int n_coarse) {
  if (!src) return idx;
  const int k = src[idx];
  return k < n_coarse ? (int64_t)r * n_coarse + k
                      : (int64_t)n_rays * n_coarse + (int64_t)r * (N - n_coarse) + (k - n_coarse);
}

template <int KP>
__global__ void __launch_bounds__(kRayWarps * 32)
render_tail_bwd_kernel(vfnerf_render_cfg cfg, int n_rays, int N, const float* __restrict__ dparams,
                       const float* __restrict__ normals, int64_t ld, const float* __restrict__ ray_dirs,
                       const float* __restrict__ z, const float* __restrict__ colors,
                       const float* __restrict__ d_rgb, const float* __restrict__ d_depth,
                       const float* __restrict__ d_normals_up, const float* __restrict__ d_colors_up,
                       float* __restrict__ d_colors, float* __restrict__ d_normals, int64_t dn_ld,
                       float* __restrict__ d_density, const uint8_t* __restrict__ src, int n_coarse) {
  __shared__ __align__(16) float s_u[kRayWarps][kUS * VFNERF_MAX_SAMPLES];
  __shared__ float s_inv[kRayWarps][VFNERF_MAX_SAMPLES];
  __shared__ float s_dc[kRayWarps][VFNERF_MAX_SAMPLES];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r = blockIdx.x * kRayWarps + wid;
  if (r >= n_rays) return;
  float* su = s_u[wid];
  float* sinv = s_inv[wid];
  float* sdc = s_dc[wid];
  const Laplace lap = load_laplace(cfg, dparams);
  const Window win = make_window(cfg.window, N);
  stage_unit_vectors(normals + (int64_t)r * N * ld, ld, N, lane, su, sinv);
  float d[3] = {ray_dirs[3 * (int64_t)r], ray_dirs[3 * (int64_t)r + 1], ray_dirs[3 * (int64_t)r + 2]};
  {
    float n = fmaxf(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), 1e-8f);
    d[0] /= n; d[1] /= n; d[2] /= n;
  }
  __syncwarp();
  const float* zr = z + (int64_t)r * N;
  const float g0 = d_rgb[3 * (int64_t)r], g1 = d_rgb[3 * (int64_t)r + 1], g2 = d_rgb[3 * (int64_t)r + 2];
  const float gd = d_depth[r];

  // ---- forward recompute
  const bool nerf_w = (cfg.flags & VFNERF_FLAG_NERF_WEIGHTS) != 0;
  const bool white = (cfg.flags & VFNERF_FLAG_WHITE_BG) != 0;
  float cj[KP], E[KP], T[KP], what[KP], dw[KP], delta[KP];
  bool active[KP];
  float carry = 0.f, wsum = 0.f, pcarry = 1.f;
#pragma unroll
  for (int i = 0; i < KP; ++i) {
    const int j = lane + 32 * i;
    float e = 0.f, c = 1.f, dl = 0.f;
    bool act = false;
    if (j < N - 1) {
      c = window_cos(su, j, win);
      float cdir = dot3(su + kUS * j, d);
      float sg = lap.cdf(-c) - lap.L0;
      act = sg > 0.f && !(cdir < cfg.dir_to_normal_th && c < 0.f);
      sg = act ? sg : 0.f;
      dl = zr[j + 1] - zr[j];
      e = dl * sg;
    }
    cj[i] = c; E[i] = e; active[i] = act; delta[i] = dl;
    if (nerf_w) {
      // T holds the INCLUSIVE product P_j = prod_{i<=j} (1 - a_i + 1e-10), formed like the forward forms it
      const float a_ = 1.f - expf(-e);
      const float pinc = warp_inclusive_prod((j < N) ? (1.f - a_ + 1e-10f) : 1.f, lane);
      T[i] = pcarry * pinc;
      pcarry *= __shfl_sync(kFull, pinc, 31);
    } else {
      float inc = warp_inclusive_scan(e, lane);
      T[i] = expf(-(carry + inc - e));
      carry += __shfl_sync(kFull, inc, 31);
    }
    what[i] = (j < N) ? (1.f - expf(-e)) * T[i] : 0.f;
    wsum += what[i];
  }
  wsum = warp_sum(wsum);
  const float inv = cfg.normalize ? 1.f / (wsum + 1e-5f) : 1.f;

  // ---- composite backward: d_colors and dL/dw
  float dot_dw_w = 0.f;
#pragma unroll
  for (int i = 0; i < KP; ++i) {
    const int j = lane + 32 * i;
    dw[i] = 0.f;
    if (j < N) {
      const int64_t idx = (int64_t)r * N + j;
      const float w = what[i] * inv;
      const float c0 = colors[3 * idx], c1 = colors[3 * idx + 1], c2 = colors[3 * idx + 2];
      dw[i] = g0 * c0 + g1 * c1 + g2 * c2 + gd * zr[j];
      if (white) dw[i] -= g0 + g1 + g2;              // rgb_c += 1 - sum_j w_j
      float o0 = w * g0, o1 = w * g1, o2 = w * g2;
      if (d_colors_up) { o0 += d_colors_up[3 * idx]; o1 += d_colors_up[3 * idx + 1]; o2 += d_colors_up[3 * idx + 2]; }
      const int64_t od = out_row(src, idx, r, n_rays, N, n_coarse);
      d_colors[3 * od] = o0; d_colors[3 * od + 1] = o1; d_colors[3 * od + 2] = o2;
      dot_dw_w += dw[i] * w;
    }
  }
  dot_dw_w = warp_sum(dot_dw_w);
  // w = what / (S + eps)  =>  d what_j = (dw_j - sum_i dw_i w_i) / (S + eps)
  float q[KP], qtot = 0.f;
#pragma unroll
  for (int i = 0; i < KP; ++i) {
    const int j = lane + 32 * i;
    float dwh = 0.f;
    if (j < N) dwh = cfg.normalize ? (dw[i] - dot_dw_w) * inv : dw[i];
    dw[i] = dwh;                 // now d/d what_j
    q[i] = dwh * what[i];
    qtot += q[i];
  }
  qtot = warp_sum(qtot);
  // dE_j = d what_j * T_j * exp(-E_j) - sum_{i>j} d what_i * what_i
  float ds_acc = 0.f, dm_acc = 0.f, db_acc = 0.f, qcarry = 0.f;
#pragma unroll
  for (int i = 0; i < KP; ++i) {
    const int j = lane + 32 * i;
    float inc = warp_inclusive_scan(q[i], lane);
    float suffix = qtot - (qcarry + inc);            // sum over i > j
    qcarry += __shfl_sync(kFull, inc, 31);
    float dc = 0.f;
    if (j < N - 1 && active[i]) {
      const float ex = expf(-E[i]);
      // volsdf: w_i = a_i exp(-sum_{k<i} E_k): E_j enters a_j and every later transmittance.
      // nerf:   w_i = a_i prod_{k<=i} (exp(-E_k) + 1e-10): E_j enters a_j and every product from i = j on.
      // (the factor of the product is the ROUNDED 1 - a + 1e-10 the forward multiplied with, not exp(-E) + 1e-10:
      //  they differ once exp(-E) drops below the fp32 spacing at 1, which a density scale of 100 reaches easily)
      float dE = nerf_w ? dw[i] * T[i] * ex - (suffix + q[i]) * ex / (1.f - (1.f - ex) + 1e-10f)
                        : dw[i] * T[i] * ex - suffix;
      float dsig = delta[i] * dE;
      float x = -cj[i];
      float lp = lap.dcdf_dx(x);
      dc = -dsig * lp;
      ds_acc += dsig * (lap.cdf(x) - lap.L0) / lap.scale;
      dm_acc += dsig * (lap.Lp0 - lp);
      db_acc += dsig * (lap.dcdf_dbeta(x) - lap.Lb0);
    }
    if (j < N) sdc[j] = dc;
  }
  __syncwarp();

  // ---- windowed-cosine backward (gather form): for sample k collect every pair (k,m) in which
  // k is the centre or the partner.  d cos(a,b)/da = (b^ - cos * a^)/|a|.
  const int reach = win.nb + 1;
  for (int k = lane; k < N; k += 32) {
    float gx = 0.f, gy = 0.f, gz = 0.f;
    const float* uk = su + kUS * k;
    const bool k_band = (k >= win.lo && k < win.hi);
    for (int m = max(0, k - reach); m <= min(N - 1, k + reach); ++m) {
      if (m == k) continue;
      float om = 0.f;
      if (k < N - 1) {                                  // k is a centre
        if (k_band) { if ((m - k >= 1 && m - k <= win.nb + 1) || (k - m >= 1 && k - m <= win.nb)) om += sdc[k] * win.coef; }
        else if (m == k + 1) om += sdc[k];
      }
      if (m < N - 1) {                                  // m is a centre, k its partner
        const bool m_band = (m >= win.lo && m < win.hi);
        if (m_band) { if ((k - m >= 1 && k - m <= win.nb + 1) || (m - k >= 1 && m - k <= win.nb)) om += sdc[m] * win.coef; }
        else if (k == m + 1) om += sdc[m];
      }
      if (om != 0.f) {
        const float* um = su + kUS * m;
        float cs = dot3(uk, um);
        gx += om * (um[0] - cs * uk[0]); gy += om * (um[1] - cs * uk[1]); gz += om * (um[2] - cs * uk[2]);
      }
    }
    const float iv = sinv[k];
    gx *= iv; gy *= iv; gz *= iv;
    const int64_t idx = (int64_t)r * N + k;
    if (d_normals_up) { gx += d_normals_up[3 * idx]; gy += d_normals_up[3 * idx + 1]; gz += d_normals_up[3 * idx + 2]; }
    float* o = d_normals + out_row(src, idx, r, n_rays, N, n_coarse) * dn_ld;
    o[0] = gx; o[1] = gy; o[2] = gz;
  }

  // ---- density parameter grads through the clamps (density_functions.py:169-204)
  ds_acc = warp_sum(ds_acc); dm_acc = warp_sum(dm_acc); db_acc = warp_sum(db_acc);
  if (lane == 0) {
    const float b = dparams[0], sc = dparams[1], mu = dparams[2];
    if (b >= cfg.beta_lo && b <= cfg.beta_hi && db_acc != 0.f) atomicAdd(d_density + 0, db_acc);
    float as = fabsf(sc);
    float route = (as > cfg.scale_min) ? 1.f : ((as == cfg.scale_min) ? 0.5f : 0.f);
    float sgn = (sc > 0.f) ? 1.f : ((sc < 0.f) ? -1.f : 0.f);
    if (route != 0.f && ds_acc != 0.f) atomicAdd(d_density + 1, ds_acc * route * sgn);
    if (mu >= cfg.mean_lo && mu <= cfg.mean_hi && dm_acc != 0.f) atomicAdd(d_density + 2, dm_acc);
  }
}

int launch_render_tail_bwd(const vfnerf_render_cfg& cfg, int n_rays, int n_samples,
                           const float* density_params, const float* normals, int64_t normals_ld,
                           const float* ray_dirs, const float* z, const float* colors,
                           const float* d_rgb, const float* d_depth, const float* d_normals_up,
                           const float* d_colors_up, float* d_colors, float* d_normals,
                           int64_t d_normals_ld, float* d_density, cudaStream_t s, const uint8_t* src, int n_coarse) {
  if (n_rays <= 0) return 0;
  VFN_REQUIRE(n_samples >= 2 && n_samples <= VFNERF_MAX_SAMPLES, "render_tail_bwd: n_samples=%d out of range", n_samples);
  const dim3 grid((n_rays + kRayWarps - 1) / kRayWarps), block(kRayWarps * 32);
#define VFN_TB(KP) render_tail_bwd_kernel<KP><<<grid, block, 0, s>>>( \
      cfg, n_rays, n_samples, density_params, normals, normals_ld, ray_dirs, z, colors, d_rgb, d_depth, \
      d_normals_up, d_colors_up, d_colors, d_normals, d_normals_ld, d_density, src, n_coarse)
  if (n_samples <= 64) VFN_TB(2); else if (n_samples <= 128) VFN_TB(4); else VFN_TB(kMaxPerLane);
#undef VFN_TB
  VFN_LAUNCH_CHECK();
  return 0;
}

}  // namespace vfn
