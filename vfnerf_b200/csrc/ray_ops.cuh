// Per-ray device building blocks shared by the stand-alone stage kernels (density_composite.cu, geometry_sampler.cu) and
// the fused per-ray kernels of the render path (render_fused.cu): Laplace density, windowed cosine, the weights of one
// ray held one-sample-per-lane in registers, and the warp sort of the fine sampler.  One warp owns one ray everywhere.
#pragma once
#include "common.cuh"

namespace vfn {

constexpr int kRayWarps = 4;
constexpr int kMaxPerLane = VFNERF_MAX_SAMPLES / 32;  // 8
constexpr int kUS = 4;   // floats per staged unit vector (x, y, z, pad): one LDS.128 per stencil partner

struct Laplace {
  float beta, scale, mean;   // effective (clamped) parameters
  float L0, Lp0, Lb0;        // cdf, d/dx and d/dbeta at the cutoff x0 = -0.5
  __device__ __forceinline__ float cdf(float x) const {
    float d = x - mean;
    float sg = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
    return scale * (0.5f + 0.5f * sg * (1.f - expf(-fabsf(d) / beta)));
  }
  __device__ __forceinline__ float dcdf_dx(float x) const {   // = -dcdf/dmean
    return scale * 0.5f * expf(-fabsf(x - mean) / beta) / beta;
  }
  __device__ __forceinline__ float dcdf_dbeta(float x) const {
    float d = x - mean, a = fabsf(d);
    float sg = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
    return scale * 0.5f * sg * (-expf(-a / beta) * a / (beta * beta));
  }
};

// get_beta / get_scale / get_mean, density_functions.py:169-204; the cutoff is ALWAYS -0.5 because
// Density.forward drops its cutoff argument (density_functions.py:20-34).
__device__ __forceinline__ Laplace load_laplace(const vfnerf_render_cfg& cfg, const float* __restrict__ dp) {
  Laplace l;
  l.beta = fminf(fmaxf(dp[0], cfg.beta_lo), cfg.beta_hi);
  l.scale = fmaxf(fabsf(dp[1]), cfg.scale_min);
  l.mean = fminf(fmaxf(dp[2], cfg.mean_lo), cfg.mean_hi);
  l.L0 = l.cdf(-0.5f);
  l.Lp0 = l.dcdf_dx(-0.5f);
  l.Lb0 = l.dcdf_dbeta(-0.5f);
  return l;
}

struct Window {
  int start, nb, lo, hi;     // band of centre indices j in [lo, hi) that get the full window
  float coef;                // (1/W) / sum_i |1/W|, as the reference forms it in fp32
};
__device__ __forceinline__ Window make_window(int W, int N) {
  Window w;
  w.start = (W + 1) / 2 + 1;         // int((W + 1) / 2 + 1), functions.py:52
  w.nb = w.start - 2;                // partners on each side besides j+1, functions.py:65
  const int L = N - 1;
  w.lo = w.start;
  w.hi = L - w.start;
  if (w.hi <= w.lo) { w.lo = 0; w.hi = 0; }
  float wu = 1.0f / (float)W, nrm = 0.f;
  for (int i = 0; i < W; ++i) nrm = __fadd_rn(nrm, wu);
  w.coef = wu / nrm;
  return w;
}

// Loads the N vectors of ray r, stores unit vectors (x / max(|x|, 1e-8), torch 2.x cosine_similarity)
// and 1/max(|x|,1e-8) into shared memory.
__device__ __forceinline__ void stage_unit_vectors(const float* __restrict__ nrm_row, int64_t ld, int N,
                                                   int lane, float* su, float* sinv) {
  for (int j = lane; j < N; j += 32) {
    const float* p = nrm_row + (int64_t)j * ld;
    float x = p[0], y = p[1], z = p[2];
    float n = fmaxf(sqrtf(x * x + y * y + z * z), 1e-8f);
    *reinterpret_cast<float4*>(su + kUS * j) = make_float4(x / n, y / n, z / n, 0.f);
    if (sinv) sinv[j] = 1.f / n;
  }
}

__device__ __forceinline__ float dot3(const float* a, const float* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
__device__ __forceinline__ float dot3(const float4& a, const float4& b) {
  return a.x * b.x + a.y * b.y + a.z * b.z;
}

// windowed cosine c_j (functions.py:41-72 with the uniform weights of vector_field_nerf.py:453)
__device__ __forceinline__ float window_cos(const float* su, int j, const Window& w) {
  const float4* u4 = reinterpret_cast<const float4*>(su);
  const float4 uj = u4[j];
  float base = dot3(uj, u4[j + 1]);
  if (j < w.lo || j >= w.hi) return base;
  float c = base * w.coef;
  if (w.nb == 5) {            // the shipped 11-tap window, unrolled (same order of operations as the loop below)
#pragma unroll
    for (int i = 1; i <= 5; ++i) {
      c = c + dot3(uj, u4[j + 1 + i]) * w.coef;
      c = c + dot3(uj, u4[j - i]) * w.coef;
    }
    return c;
  }
  for (int i = 1; i <= w.nb; ++i) {
    c = c + dot3(uj, u4[j + 1 + i]) * w.coef;
    c = c + dot3(uj, u4[j - i]) * w.coef;
  }
  return c;
}

__device__ __forceinline__ float warp_inclusive_prod(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}


// Rows a4-a6 for one ray whose unit vectors are staged in `su`: windowed cosine -> Laplace density (direction mask) ->
// free energy -> transmittance scan -> weights.  Lane l holds samples l, l+32, ...: what[i] is the UNNORMALISED weight of
// sample l + 32 i; the return value is the normalisation factor (1 when cfg.normalize is off), so the weight the
// reference would hold is what[i] * inv.  zr: the ray's z values (global or shared).  cosw / sigma_out: optional rows.
template <int KP>
__device__ __forceinline__ float ray_weights(const vfnerf_render_cfg& cfg, const Laplace& lap, const Window& win,
                                             const float* su, const float* d, const float* zr, int N, int lane,
                                             float* cosw_row, float* sigma_row, float (&what)[KP]) {
  const bool nerf_w = (cfg.flags & VFNERF_FLAG_NERF_WEIGHTS) != 0;
  float carry = 0.f, wsum = 0.f, pcarry = 1.f;
#pragma unroll
  for (int i = 0; i < KP; ++i) {
    const int j = lane + 32 * i;
    float E = 0.f, sg = 0.f;
    if (j < N - 1) {
      float c = window_cos(su, j, win);
      float cdir = dot3(su + kUS * j, d);
      sg = fmaxf(lap.cdf(-c) - lap.L0, 0.f);
      if (cdir < cfg.dir_to_normal_th && c < 0.f) sg = 0.f;
      E = (zr[j + 1] - zr[j]) * sg;
      if (cosw_row) cosw_row[j] = c;
    }
    if (j < N && sigma_row) sigma_row[j] = sg;
    // exclusive prefix of the free energy over this 32-sample chunk, plus the carry of earlier chunks
    float a = 1.f - expf(-E);
    if (nerf_w) {
      // nerf_volume_rendering (utils/rendering.py:98-119): inclusive cumprod of (1 - a + 1e-10)
      const float pinc = warp_inclusive_prod((j < N) ? (1.f - a + 1e-10f) : 1.f, lane);
      what[i] = (j < N) ? a * (pcarry * pinc) : 0.f;
      pcarry *= __shfl_sync(kFull, pinc, 31);
    } else {
      float inc = warp_inclusive_scan(E, lane);
      float T = expf(-(carry + inc - E));
      what[i] = (j < N) ? a * T : 0.f;
      carry += __shfl_sync(kFull, inc, 31);
    }
    wsum += what[i];
  }
  wsum = warp_sum(wsum);
  return cfg.normalize ? 1.f / (wsum + 1e-5f) : 1.f;
}

// unit view direction as F.cosine_similarity normalises it (|d| clamped at 1e-8)
__device__ __forceinline__ void unit_dir(const float* __restrict__ ray_dirs, int64_t r, float* d) {
  d[0] = ray_dirs[3 * r]; d[1] = ray_dirs[3 * r + 1]; d[2] = ray_dirs[3 * r + 2];
  const float n = fmaxf(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), 1e-8f);
  d[0] /= n; d[1] /= n; d[2] /= n;
}

// ---------------------------------------------------------------------------------------------
// Warp-level sort of cat(run A = s[0, na), run B = s[na, na + nb)) into t[0, na + nb), with the source index of every
// output in ti.  Values are only moved, so the result equals torch.sort's values bit for bit.
//  * fast path (both runs already ascending -- coarse z values are, and so is the fine ramp): merge by rank.  An element
//    of A lands at (its index + number of B elements smaller than it), an element of B at (its index + number of A
//    elements not larger than it): two binary searches per lane and sample instead of a full sort;
//  * otherwise (the uniform "z_add" candidates, random inverse-CDF draws, NaNs): bitonic sort in place, then copy.
// s / si need room for the next power of two >= na + nb (<= VFNERF_MAX_SAMPLES).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_sort_two_runs(float* s, uint8_t* si, float* t, uint8_t* ti, int na, int nb, int lane) {
  const int N = na + nb;
  bool ok = true;
  for (int j = lane; j < N; j += 32)
    if (j != 0 && j != na) ok = ok && (s[j - 1] <= s[j]);
  if (__all_sync(kFull, ok)) {
    // branch-free binary searches with a fixed trip count (log2 of the next power of two >= run length)
    int pa = 1, pb = 1;
    while (pa < na) pa <<= 1;
    while (pb < nb) pb <<= 1;
    for (int e = lane; e < N; e += 32) {
      const float v = s[e];
      int pos = 0;
      if (e < na) {                       // number of B elements smaller than v (lower bound)
        for (int st = pb; st > 0; st >>= 1) {
          const int q = pos + st;
          if (q <= nb && s[na + q - 1] < v) pos = q;
        }
        pos += e;
      } else {                            // number of A elements not larger than v (upper bound)
        for (int st = pa; st > 0; st >>= 1) {
          const int q = pos + st;
          if (q <= na && s[q - 1] <= v) pos = q;
        }
        pos += e - na;
      }
      t[pos] = v;
      ti[pos] = (uint8_t)e;
    }
    __syncwarp();
    return;
  }
  int P2 = 32;
  while (P2 < N) P2 <<= 1;
  for (int j = N + lane; j < P2; j += 32) s[j] = INFINITY;
  for (int j = lane; j < P2; j += 32) si[j] = (uint8_t)j;
  __syncwarp();
  for (int k = 2; k <= P2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int q = lane; q < (P2 >> 1); q += 32) {
        int lo = ((q & ~(j - 1)) << 1) | (q & (j - 1));   // index with bit j cleared
        int hi = lo | j;
        bool asc = (lo & k) == 0;
        float a = s[lo], b = s[hi];
        if ((a > b) == asc) {
          s[lo] = b; s[hi] = a;
          uint8_t ia = si[lo]; si[lo] = si[hi]; si[hi] = ia;
        }
      }
      __syncwarp();
    }
  }
  for (int j = lane; j < N; j += 32) { t[j] = s[j]; ti[j] = si[j]; }
  __syncwarp();
}


// a1 for one ray (utils/rendering.py:12-60 + utils/pinhole_model.py:9-63): world-space direction (unnormalised, what the
// samplers use), unit direction (F.normalize) and camera location.  Every product and sum is an explicit
// round-to-nearest intrinsic in the reference's order (torch CPU bmm: ((p0*x + p1*y) + p2*z) + p3*1, unfused).
__device__ __forceinline__ void ray_geometry_one(int64_t r, int pose_is_quat, const float* __restrict__ uv,
                                                 const float* __restrict__ pose, const float* __restrict__ K, float* d,
                                                 float* rd, float* o) {
  float p[3][4];
  if (pose_is_quat) {
    const float* q7 = pose + (int64_t)r * 7;
    float qr = q7[0], qi = q7[1], qj = q7[2], qk = q7[3];
    float nq = fmaxf(sqrtf(qr * qr + qi * qi + qj * qj + qk * qk), 1e-12f);
    qr /= nq; qi /= nq; qj /= nq; qk /= nq;
    p[0][0] = 1.f - 2.f * (qj * qj + qk * qk); p[0][1] = 2.f * (qj * qi - qk * qr); p[0][2] = 2.f * (qi * qk + qr * qj);
    p[1][0] = 2.f * (qj * qi + qk * qr); p[1][1] = 1.f - 2.f * (qi * qi + qk * qk); p[1][2] = 2.f * (qj * qk - qi * qr);
    p[2][0] = 2.f * (qk * qi - qj * qr); p[2][1] = 2.f * (qj * qk + qi * qr); p[2][2] = 1.f - 2.f * (qi * qi + qj * qj);
    p[0][3] = q7[4]; p[1][3] = q7[5]; p[2][3] = q7[6];
  } else {
    const float4* pm = reinterpret_cast<const float4*>(pose + (int64_t)r * 16);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float4 row = __ldg(pm + i);
      p[i][0] = row.x; p[i][1] = row.y; p[i][2] = row.z; p[i][3] = row.w;
    }
  }
  const float* Kr = K + (int64_t)r * 16;
  float fx = Kr[0], skew = Kr[1], cx = Kr[2], fy = Kr[5], cy = Kr[6];
  float k011 = __ldg(K + 5);                                   // intrinsics[0,1,1] of the FIRST ray
  float zc = (k011 > 0.f) ? 1.f : ((k011 < 0.f) ? -1.f : 0.f);  // ones * sign(.)
  float za = fabsf(zc);
  float u = uv[2 * (int64_t)r], v = uv[2 * (int64_t)r + 1];
  // x = (u - cx + cy*skew/fy - skew*v/fy) / fx * |z| ; y = (v - cy) / fy * |z|
  float t = __fsub_rn(u, cx);
  t = __fadd_rn(t, __fdiv_rn(__fmul_rn(cy, skew), fy));
  t = __fsub_rn(t, __fdiv_rn(__fmul_rn(skew, v), fy));
  float x = __fmul_rn(__fdiv_rn(t, fx), za);
  float y = __fmul_rn(__fdiv_rn(__fsub_rn(v, cy), fy), za);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    // bmm row: ((p0*x + p1*y) + p2*z) + p3*1, unfused (matches torch CPU bmm bit for bit)
    float acc = __fmul_rn(p[i][0], x);
    acc = __fadd_rn(acc, __fmul_rn(p[i][1], y));
    acc = __fadd_rn(acc, __fmul_rn(p[i][2], zc));
    acc = __fadd_rn(acc, __fmul_rn(p[i][3], 1.f));
    d[i] = __fsub_rn(acc, p[i][3]);
  }
  float nrm = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2]))), 1e-12f);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    rd[i] = __fdiv_rn(d[i], nrm);
    o[i] = p[i][3];
  }
}

// a2 for one sample (UniformSampler.get_z_vals, ray_sampler.py:113-142): z_i = near (1 - t_i) + far t_i, stratified by u
__device__ __forceinline__ float coarse_z_one(int i, int n_coarse, float nearf, float farf, int perturb,
                                              const float* __restrict__ t_vals, const float* __restrict__ U1, int64_t idx) {
  auto lin = [&](int k) {
    float tk = __ldg(t_vals + k);
    return __fadd_rn(__fmul_rn(nearf, __fsub_rn(1.f, tk)), __fmul_rn(farf, tk));
  };
  float zi = lin(i);
  if (perturb) {
    float lower = (i == 0) ? zi : __fmul_rn(0.5f, __fadd_rn(zi, lin(i - 1)));
    float upper = (i == n_coarse - 1) ? zi : __fmul_rn(0.5f, __fadd_rn(lin(i + 1), zi));
    zi = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), U1[idx]));
  }
  return zi;
}

// torch.argmax over a ray's weights held per lane: largest value, FIRST index on ties; rows without a finite maximum
// (all -inf / NaN) select index 0 like an all-equal row
__device__ __forceinline__ void warp_argmax_first(float& best, int& bi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ob = __shfl_xor_sync(kFull, best, o);
    int oi = __shfl_xor_sync(kFull, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (bi == 0x7fffffff) bi = 0;
}

// RangeFineSampler.get_z_vals (ray_sampler.py:264-302) for one ray, given the index `bi` of its largest coarse weight
// (first index on ties) : candidates (ramp around z*, stratified by U2, or the uniform z_add fallback from U3 when
// bi == 0), optional points of the fine candidates alone, then the value-exact warp sort of cat(coarse, candidates).
// s: the ray's coarse z values in s[0, n_coarse) on entry; t / ti: sorted values and candidate indices on return.
struct FineCfg { int n_coarse, n_fine, perturb; float nearf, far_minus_near, rangef, stepf; };
__device__ __forceinline__ void fine_candidates_sorted(const FineCfg& f, int64_t r, int bi, float z_star,
                                                       const float* __restrict__ U2, const float* __restrict__ U3,
                                                       const float* o, const float* dvec, float* s, uint8_t* si, float* t,
                                                       uint8_t* ti, float* __restrict__ points_fine, int lane) {
  const int n_coarse = f.n_coarse, n_fine = f.n_fine;
  const float base = __fsub_rn(z_star, f.rangef);
  auto ramp = [&](int i) { return __fadd_rn(base, __fmul_rn(f.stepf, (float)i)); };
  for (int i = lane; i < n_fine; i += 32) {
    float v;
    if (bi > 0) {
      v = ramp(i);
      if (f.perturb) {
        float lower = (i == 0) ? v : __fmul_rn(0.5f, __fadd_rn(v, ramp(i - 1)));
        float upper = (i == n_fine - 1) ? v : __fmul_rn(0.5f, __fadd_rn(ramp(i + 1), v));
        v = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), U2[r * n_fine + i]));
      }
    } else {
      v = __fadd_rn(__fmul_rn(U3[r * n_fine + i], f.far_minus_near), f.nearf);
    }
    s[n_coarse + i] = v;
  }
  __syncwarp();
  if (points_fine) {
    // the fine candidates alone, in candidate order: the only points of this ray the MLPs have not seen yet
    float* pf = points_fine + r * n_fine * 3;
    for (int j = lane; j < n_fine; j += 32) {
      const float zj = s[n_coarse + j];
      pf[3 * j] = __fadd_rn(o[0], __fmul_rn(zj, dvec[0]));
      pf[3 * j + 1] = __fadd_rn(o[1], __fmul_rn(zj, dvec[1]));
      pf[3 * j + 2] = __fadd_rn(o[2], __fmul_rn(zj, dvec[2]));
    }
    __syncwarp();
  }
  warp_sort_two_runs(s, si, t, ti, n_coarse, n_fine, lane);
}

// merged z row + points of one ray from the sorted values in t
__device__ __forceinline__ void write_merged_samples(const float* t, int N, int64_t r, const float* o, const float* dvec,
                                                     float* __restrict__ z_out, float* __restrict__ points, int lane) {
  for (int j = lane; j < N; j += 32) z_out[r * N + j] = t[j];
  if (points) {
    // one sample per lane and iteration, three 12-byte-strided stores (a warp still writes one contiguous 384-byte span)
    float* pr = points + r * N * 3;
    for (int j = lane; j < N; j += 32) {
      const float zj = t[j];
      pr[3 * j] = __fadd_rn(o[0], __fmul_rn(zj, dvec[0]));
      pr[3 * j + 1] = __fadd_rn(o[1], __fmul_rn(zj, dvec[1]));
      pr[3 * j + 2] = __fadd_rn(o[2], __fmul_rn(zj, dvec[2]));
    }
  }
}

}  // namespace vfn
