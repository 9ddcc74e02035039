// Ray generation and the two samplers.
//
// Contract: sample positions are BIT-EXACT with the reference given the same uniform draws.  The
// reference evaluates every expression as separate fp32 tensor ops on the CPU, so every product and
// sum here is an explicit round-to-nearest intrinsic (__fmul_rn/__fadd_rn/...): the compiler may not
// contract them into FMAs.  HBM-bound byte work: one thread per output element, fully coalesced.
#include "common.cuh"
#include "ray_ops.cuh"

namespace vfn {

// ---------------------------------------------------------------------------------------------
// a1: utils/rendering.py:12-60 + utils/pinhole_model.py:9-63
// ---------------------------------------------------------------------------------------------
__global__ void ray_geometry_kernel(int n_rays, int pose_is_quat, const float* __restrict__ uv,
                                    const float* __restrict__ pose, const float* __restrict__ K,
                                    float* __restrict__ directions, float* __restrict__ ray_dirs,
                                    float* __restrict__ cam_loc) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  float d[3], rd[3], o[3];
  ray_geometry_one(r, pose_is_quat, uv, pose, K, d, rd, o);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    directions[3 * (int64_t)r + i] = d[i];
    ray_dirs[3 * (int64_t)r + i] = rd[i];
    cam_loc[3 * (int64_t)r + i] = o[i];
  }
}

int launch_ray_geometry(int n_rays, int pose_is_quat, const float* uv, const float* pose,
                        const float* K, float* directions, float* ray_dirs, float* cam_loc,
                        cudaStream_t s) {
  if (n_rays <= 0) return 0;
  ray_geometry_kernel<<<(n_rays + 127) / 128, 128, 0, s>>>(n_rays, pose_is_quat, uv, pose, K,
                                                             directions, ray_dirs, cam_loc);
  VFN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// a2: UniformSampler.get_z_vals (ray_sampler.py:113-142) + RaySampler.sample (:49-80)
// ---------------------------------------------------------------------------------------------
__global__ void coarse_sample_kernel(int64_t total, int n_coarse, float nearf, float farf, int perturb,
                                     const float* __restrict__ t_vals, const float* __restrict__ U1,
                                     const float* __restrict__ directions,
                                     const float* __restrict__ cam_loc, float* __restrict__ z,
                                     float* __restrict__ points) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int64_t r = idx / n_coarse;
  int i = (int)(idx - r * n_coarse);
  const float zi = coarse_z_one(i, n_coarse, nearf, farf, perturb, t_vals, U1, idx);
  z[idx] = zi;
  if (points) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      points[3 * idx + c] = __fadd_rn(__ldg(cam_loc + 3 * r + c), __fmul_rn(zi, __ldg(directions + 3 * r + c)));
  }
}

int launch_coarse_sample(int n_rays, int n_coarse, double near_, double far_, int perturb,
                         const float* t_vals, const float* U1, const float* directions,
                         const float* cam_loc, float* z, float* points, cudaStream_t s) {
  int64_t total = (int64_t)n_rays * n_coarse;
  if (total <= 0) return 0;
  VFN_REQUIRE(!perturb || U1 != nullptr, "coarse_sample: perturb=1 needs U1");
  coarse_sample_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(
      total, n_coarse, (float)near_, (float)far_, perturb, t_vals, U1, directions, cam_loc, z, points);
  VFN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// a7: RangeFineSampler.get_z_vals (ray_sampler.py:264-302) + sample.  One warp per ray:
// warp argmax (first index on ties), candidate generation, warp sort of <= 256 values in the warp's
// shared-memory slice (value-exact, so identical to torch.sort's values), point generation.
// ---------------------------------------------------------------------------------------------
constexpr int kFineWarps = 4;

__global__ void __launch_bounds__(kFineWarps * 32)
fine_sample_kernel(int n_rays, int n_coarse, int n_fine, float nearf, float far_minus_near,
                   float rangef, float stepf, int perturb, const float* __restrict__ z_coarse,
                   const float* __restrict__ w_coarse, const float* __restrict__ U2,
                   const float* __restrict__ U3, const float* __restrict__ z_override,
                   const float* __restrict__ directions, const float* __restrict__ cam_loc,
                   float* __restrict__ z_out, float* __restrict__ points,
                   uint8_t* __restrict__ src, float* __restrict__ points_fine) {
  __shared__ float sbuf[kFineWarps][VFNERF_MAX_SAMPLES];
  __shared__ float tbuf[kFineWarps][VFNERF_MAX_SAMPLES];
  __shared__ uint8_t sidx[kFineWarps][VFNERF_MAX_SAMPLES];
  __shared__ uint8_t tidx[kFineWarps][VFNERF_MAX_SAMPLES];   // candidate index (coarse 0.., fine n_coarse..) per output
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r = blockIdx.x * kFineWarps + wid;
  if (r >= n_rays) return;
  float* s = sbuf[wid];
  float* t = tbuf[wid];
  const int N = n_coarse + n_fine;
  const float dv[3] = {directions[3 * (int64_t)r], directions[3 * (int64_t)r + 1], directions[3 * (int64_t)r + 2]};
  const float o[3] = {cam_loc[3 * (int64_t)r], cam_loc[3 * (int64_t)r + 1], cam_loc[3 * (int64_t)r + 2]};
  if (z_override) {
    for (int j = lane; j < N; j += 32) t[j] = z_override[(int64_t)r * N + j];
  } else {
    // argmax of the coarse weights, first index on ties (torch.argmax)
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = lane; j < n_coarse; j += 32) {
      float w = w_coarse[(int64_t)r * n_coarse + j];
      if (w > best) { best = w; bi = j; }
    }
    warp_argmax_first(best, bi);
    const float z_star = z_coarse[(int64_t)r * n_coarse + bi];
    for (int j = lane; j < n_coarse; j += 32) s[j] = z_coarse[(int64_t)r * n_coarse + j];
    const FineCfg fc{n_coarse, n_fine, perturb, nearf, far_minus_near, rangef, stepf};
    fine_candidates_sorted(fc, r, bi, z_star, U2, U3, o, dv, s, sidx[wid], t, tidx[wid], points_fine, lane);
    if (src) for (int j = lane; j < N; j += 32) src[(int64_t)r * N + j] = tidx[wid][j];
  }
  __syncwarp();
  write_merged_samples(t, N, r, o, dv, z_out, points, lane);
}

int launch_fine_sample(int n_rays, int n_coarse, int n_fine, double near_, double far_,
                       double fine_range, int perturb, const float* z_coarse, const float* w_coarse,
                       const float* U2, const float* U3, const float* z_override,
                       const float* directions, const float* cam_loc, float* z, float* points,
                       uint8_t* src, float* points_fine, cudaStream_t s) {
  if (n_rays <= 0) return 0;
  VFN_REQUIRE(!(z_override && (src || points_fine)), "fine_sample: z_override carries no candidate indices");
  VFN_REQUIRE(n_coarse + n_fine <= VFNERF_MAX_SAMPLES, "fine_sample: n_coarse+n_fine=%d exceeds %d",
              n_coarse + n_fine, VFNERF_MAX_SAMPLES);
  VFN_REQUIRE(n_fine >= 2, "fine_sample: n_fine must be >= 2 (the reference divides by n_fine-1)");
  VFN_REQUIRE(z_override || (U3 && (!perturb || U2)), "fine_sample: missing uniform draws");
  float stepf = (float)(2.0 * fine_range / (double)(n_fine - 1));   // python double, then fp32
  fine_sample_kernel<<<(n_rays + kFineWarps - 1) / kFineWarps, kFineWarps * 32, 0, s>>>(
      n_rays, n_coarse, n_fine, (float)near_, (float)(far_ - near_), (float)fine_range, stepf, perturb,
      z_coarse, w_coarse, U2, U3, z_override, directions, cam_loc, z, points, src, points_fine);
  VFN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Inverse-CDF importance sampler: FineSampler.sample_pdf / get_z_vals (ray_sampler.py:163-237).  Not called by the
// reference's render(); offered as the alternative fine sampler.  One warp per ray:
//   w + 1e-5 -> warp-shuffle sum -> pdf -> chunked warp prefix scan = cdf (cdf[0] = 0) in shared memory,
//   per sample a binary search for the first cdf entry > u (torch.searchsorted right=True), the reference's
//   clamped gather / guarded division / lerp, then (merged mode) the warp bitonic sort of cat(z_coarse, samples).
// The reference forms its sum and cumulative sum with aten's CPU orders (vectorised cascade sum, double-accumulated
// cumsum); a warp tree / fp32 scan differs from those in the last bit of the cdf, so parity of this kernel is
// tolerance-based: |cdf_ref(z) - u| <= 1e-6 and |z - z_ref| <= 1e-5 wherever the pdf is not degenerate (tests).
// ---------------------------------------------------------------------------------------------
constexpr int kPdfWarps = 4;

template <bool kMerged>
__global__ void __launch_bounds__(kPdfWarps * 32)
pdf_sample_kernel(int n_rays, int n_bins, int n_new, const float* __restrict__ bins_or_z,
                  const float* __restrict__ weights, const float* __restrict__ u, int u_per_ray,
                  const float* __restrict__ directions, const float* __restrict__ cam_loc,
                  float* __restrict__ out, float* __restrict__ points) {
  __shared__ float s_cdf[kPdfWarps][VFNERF_MAX_SAMPLES];
  __shared__ float s_bin[kPdfWarps][VFNERF_MAX_SAMPLES];
  __shared__ float s_z[kPdfWarps][VFNERF_MAX_SAMPLES];
  __shared__ uint8_t s_i[kPdfWarps][VFNERF_MAX_SAMPLES], s_ti[kPdfWarps][VFNERF_MAX_SAMPLES];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r = blockIdx.x * kPdfWarps + wid;
  if (r >= n_rays) return;
  float* cdf = s_cdf[wid];
  float* bin = s_bin[wid];
  float* z = s_z[wid];
  // merged mode: bins_or_z = z_coarse [R, Nc], n_bins = Nc - 1 midpoints, weights = w_coarse [R, Nc] (inner Nc - 2 used)
  // plain mode:  bins_or_z = bins [R, n_bins], weights [R, n_bins - 1]
  const int Nc = kMerged ? n_bins + 1 : 0;
  const int n_w = n_bins - 1;
  const float* wrow = kMerged ? weights + (int64_t)r * Nc + 1 : weights + (int64_t)r * n_w;
  if (kMerged) {
    const float* zr = bins_or_z + (int64_t)r * Nc;
    for (int j = lane; j < Nc; j += 32) z[j] = zr[j];
    __syncwarp();
    for (int j = lane; j < n_bins; j += 32) bin[j] = __fmul_rn(0.5f, __fadd_rn(z[j + 1], z[j]));
  } else {
    for (int j = lane; j < n_bins; j += 32) bin[j] = bins_or_z[(int64_t)r * n_bins + j];
  }
  float part = 0.f;
  for (int j = lane; j < n_w; j += 32) part = __fadd_rn(part, __fadd_rn(wrow[j], 1e-5f));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part = __fadd_rn(part, __shfl_xor_sync(kFull, part, o));
  const float total = part;
  // cdf[0] = 0, cdf[j + 1] = pdf[0] + ... + pdf[j]: inclusive warp scan per 32-entry chunk, carry between chunks
  float carry = 0.f;
  if (lane == 0) cdf[0] = 0.f;
  for (int j0 = 0; j0 < n_w; j0 += 32) {
    const int j = j0 + lane;
    float v = (j < n_w) ? __fdiv_rn(__fadd_rn(wrow[j], 1e-5f), total) : 0.f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float t = __shfl_up_sync(kFull, v, o);
      if (lane >= o) v = __fadd_rn(v, t);
    }
    v = __fadd_rn(v, carry);
    if (j < n_w) cdf[j + 1] = v;
    carry = __shfl_sync(kFull, v, 31);
  }
  __syncwarp();
  for (int k = lane; k < n_new; k += 32) {
    const float uk = u_per_ray ? u[(int64_t)r * n_new + k] : u[k];
    int lo = 0, hi = n_bins;               // first index in [0, n_bins] with cdf[idx] > uk
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (cdf[mid] > uk) hi = mid; else lo = mid + 1;
    }
    const int below = max(lo - 1, 0), above = min(lo, n_bins - 1);
    float denom = __fsub_rn(cdf[above], cdf[below]);
    if (denom < 1e-5f) denom = 1.f;
    const float t = __fdiv_rn(__fsub_rn(uk, cdf[below]), denom);
    const float smp = __fadd_rn(bin[below], __fmul_rn(t, __fsub_rn(bin[above], bin[below])));
    if (kMerged) z[Nc + k] = smp; else out[(int64_t)r * n_new + k] = smp;
  }
  if (!kMerged) return;
  const int N = Nc + n_new;
  __syncwarp();
  float* t = s_cdf[wid];                 // the cdf is dead: its slice receives the merged, sorted values
  warp_sort_two_runs(z, s_i[wid], t, s_ti[wid], Nc, n_new, lane);
  for (int j = lane; j < N; j += 32) out[(int64_t)r * N + j] = t[j];
  if (points) {
    const float dx = directions[3 * (int64_t)r], dy = directions[3 * (int64_t)r + 1], dz = directions[3 * (int64_t)r + 2];
    const float ox = cam_loc[3 * (int64_t)r], oy = cam_loc[3 * (int64_t)r + 1], oz = cam_loc[3 * (int64_t)r + 2];
    float* pr = points + (int64_t)r * N * 3;
    for (int e = lane; e < 3 * N; e += 32) {
      int j = e / 3, c = e - 3 * j;
      float d = (c == 0) ? dx : (c == 1 ? dy : dz);
      float o = (c == 0) ? ox : (c == 1 ? oy : oz);
      pr[e] = __fadd_rn(o, __fmul_rn(t[j], d));
    }
  }
}

int launch_sample_pdf(int n_rays, int n_bins, int n_samples, const float* bins, const float* weights, const float* u,
                      int u_per_ray, float* samples, cudaStream_t s) {
  if (n_rays <= 0 || n_samples <= 0) return 0;
  VFN_REQUIRE(n_bins >= 2 && n_bins <= VFNERF_MAX_SAMPLES, "sample_pdf: n_bins=%d outside [2, %d]", n_bins, VFNERF_MAX_SAMPLES);
  VFN_REQUIRE(bins && weights && u && samples, "sample_pdf: null argument");
  pdf_sample_kernel<false><<<(n_rays + kPdfWarps - 1) / kPdfWarps, kPdfWarps * 32, 0, s>>>(
      n_rays, n_bins, n_samples, bins, weights, u, u_per_ray, nullptr, nullptr, samples, nullptr);
  VFN_LAUNCH_CHECK();
  return 0;
}

int launch_pdf_fine_sample(int n_rays, int n_coarse, int n_fine, const float* z_coarse, const float* w_coarse,
                           const float* u, int u_per_ray, const float* directions, const float* cam_loc, float* z,
                           float* points, cudaStream_t s) {
  if (n_rays <= 0) return 0;
  VFN_REQUIRE(n_coarse >= 3, "pdf_fine_sample: n_coarse=%d; the reference needs at least one inner weight", n_coarse);
  VFN_REQUIRE(n_fine >= 1 && n_coarse + n_fine <= VFNERF_MAX_SAMPLES, "pdf_fine_sample: n_coarse+n_fine=%d exceeds %d",
              n_coarse + n_fine, VFNERF_MAX_SAMPLES);
  VFN_REQUIRE(z_coarse && w_coarse && u && z, "pdf_fine_sample: null argument");
  VFN_REQUIRE(!points || (directions && cam_loc), "pdf_fine_sample: points need directions and cam_loc");
  pdf_sample_kernel<true><<<(n_rays + kPdfWarps - 1) / kPdfWarps, kPdfWarps * 32, 0, s>>>(
      n_rays, n_coarse - 1, n_fine, z_coarse, w_coarse, u, u_per_ray, directions, cam_loc, z, points);
  VFN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Merge of per-sample results computed separately for the coarse and the fine candidates into the
// merged (sorted) sample order of vector_field_nerf.py:284-312.  The reference re-evaluates both MLPs
// on all merged points; the coarse half of those points is bit-identical to the coarse sweep's points,
// so their results are moved instead of recomputed.  One thread per merged sample, two [.,3] tensors per pass.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
merge_samples_kernel(int64_t n_points, int n_coarse, int n_fine, const uint8_t* __restrict__ src,
                     const float* __restrict__ a_coarse, const float* __restrict__ a_fine, float* __restrict__ a_out,
                     const float* __restrict__ b_coarse, const float* __restrict__ b_fine, float* __restrict__ b_out) {
  const int64_t pt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // merged sample index r * N + j
  if (pt >= n_points) return;
  const int N = n_coarse + n_fine;
  const int64_t r = pt / N;
  const int k = src[pt];
  const bool c = k < n_coarse;
  const int64_t from = 3 * (c ? r * n_coarse + k : r * n_fine + (k - n_coarse));
  const float* pa = (c ? a_coarse : a_fine) + from;
  const float a0 = __ldg(pa), a1 = __ldg(pa + 1), a2 = __ldg(pa + 2);
  a_out[3 * pt] = a0; a_out[3 * pt + 1] = a1; a_out[3 * pt + 2] = a2;
  if (b_out) {
    const float* pb = (c ? b_coarse : b_fine) + from;
    const float b0 = __ldg(pb), b1 = __ldg(pb + 1), b2 = __ldg(pb + 2);
    b_out[3 * pt] = b0; b_out[3 * pt + 1] = b1; b_out[3 * pt + 2] = b2;
  }
}

int launch_merge_samples(int n_rays, int n_coarse, int n_fine, const uint8_t* src, const float* a_coarse,
                         const float* a_fine, float* a_out, const float* b_coarse, const float* b_fine,
                         float* b_out, cudaStream_t s) {
  int64_t total = (int64_t)n_rays * (n_coarse + n_fine);
  if (total <= 0) return 0;
  merge_samples_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(total, n_coarse, n_fine, src, a_coarse, a_fine,
                                                                        a_out, b_coarse, b_fine, b_out);
  VFN_LAUNCH_CHECK();
  return 0;
}

// test support (vfnerf_debug_stash_read): rows in evaluation order -> rows in merged sample order
__global__ void rows_to_merged_order_kernel(int64_t total, int n_rays, int n_coarse, int n_fine, int cols,
                                            const uint8_t* __restrict__ src, const float* __restrict__ in,
                                            float* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int64_t pt = e / cols;
  const int c = (int)(e - pt * cols);
  const int N = n_coarse + n_fine;
  const int64_t r = pt / N;
  const int k = src[pt];
  const int64_t from = k < n_coarse ? r * n_coarse + k : (int64_t)n_rays * n_coarse + r * n_fine + (k - n_coarse);
  out[e] = in[from * cols + c];
}

int launch_rows_to_merged_order(int n_rays, int n_coarse, int n_fine, const uint8_t* src, const float* in, float* out,
                                int cols, cudaStream_t s) {
  const int64_t total = (int64_t)n_rays * (n_coarse + n_fine) * cols;
  if (total <= 0) return 0;
  rows_to_merged_order_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(total, n_rays, n_coarse, n_fine, cols, src, in, out);
  VFN_LAUNCH_CHECK();
  return 0;
}

}  // namespace vfn
