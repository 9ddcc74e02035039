// Marching-cubes preprocessing on the device: everything the reference computes on the CPU
// between the dense VF grid query and its contrastive marching cubes (evaluation/methods.py:209-278 with the default
// flags), fused per cell:
//   extract_divergence  mc_utils.py:34-85    unit vectors of the 8 corners projected on the corner directions, d|d| sum,
//                                            surface cell <=> divergence <= -0.5
//   unify_direction     mc_utils.py:107-166  most opposite pair of corner vectors, every corner sides with the nearer one
//   make_comb_format    mc_utils.py:169-223  28 corner pairs: "on different sides" flag + the two vector norms
//   compaction          methods.py:186-192, 260-278   cells in 2x2x2-block order, keep those with any differing pair
// Every quantity of a cell depends only on the 8 vectors at its corners, so the reference's three full-grid conv3d /
// gather passes (30 GB of fp32 intermediates at resolution 512) collapse into two launches over the [N^3,3] grid:
//   mc_count_kernel: one thread per cell in block order -> keep flag + per-CTA counts  (reads 12 B/point through L1/L2)
//   mc_emit_kernel:  recomputes the kept cells and writes them at their scanned offsets (348 B per kept cell, ~N^2 cells)
// The scan of the per-CTA counts between the two is a tiny host-side cumsum (one int per 256 cells).
// Arithmetic follows the reference's fp32 expression order with explicit round-to-nearest ops (no FMA contraction).
#include "common.cuh"

namespace vfn {

constexpr int kMcThreads = 256;

// corner s of the reference's `inc` table (methods.py:176-186) as the bit pattern m = 4a + 2b + c of its offset (a, b, c)
__device__ __constant__ int c_inc_m[8] = {0, 2, 6, 4, 1, 3, 7, 5};

struct McCell {
  float ux[8], uy[8], uz[8], nrm[8];   // unit vectors and norms of the 8 corners, indexed by m = 4a + 2b + c
};

__device__ __forceinline__ void mc_load(const float* __restrict__ pred, int N, int i, int j, int k, McCell& c) {
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int64_t p = ((int64_t)(i + ((m >> 2) & 1)) * N + (j + ((m >> 1) & 1))) * N + (k + (m & 1));
    const float x = __ldg(pred + 3 * p), y = __ldg(pred + 3 * p + 1), z = __ldg(pred + 3 * p + 2);
    const float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    const float dn = fmaxf(n, 1e-12f);                     // F.normalize eps
    c.ux[m] = __fdiv_rn(x, dn); c.uy[m] = __fdiv_rn(y, dn); c.uz[m] = __fdiv_rn(z, dn);
    c.nrm[m] = n;
  }
}

// mc_utils.py:40-76: output channel m of the conv3d = <unit vector at corner (a,b,c), normalize((2a-1, 2b-1, 2c-1))>
__device__ __forceinline__ float mc_divergence(const McCell& c) {
  const float fs = __fdiv_rn(1.f, sqrtf(3.f));
  const float face_area = 0.4330127018922193f, shape_volume = 0.47140452079103173f;   // sqrt(3)/4, sqrt(2)/3
  float acc = 0.f;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const float fx = (m & 4) ? fs : -fs, fy = (m & 2) ? fs : -fs, fz = (m & 1) ? fs : -fs;
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(c.ux[m], fx), __fmul_rn(c.uy[m], fy)), __fmul_rn(c.uz[m], fz));
    const float t = __fmul_rn(__fmul_rn(d, fabsf(d)), face_area);
    acc = (m == 0) ? t : __fadd_rn(acc, t);
  }
  return __fdiv_rn(acc, shape_volume);
}

// mc_utils.py:125-160: bit s of the result = side choice of corner s (`inc` order)
__device__ __forceinline__ uint32_t mc_choice(const McCell& c) {
  float best = -INFINITY;
  int bp = 0, bq = 0;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int mp = c_inc_m[p];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int mq = c_inc_m[q];
      const float dot = __fadd_rn(__fadd_rn(__fmul_rn(c.ux[mp], c.ux[mq]), __fmul_rn(c.uy[mp], c.uy[mq])),
                                  __fmul_rn(c.uz[mp], c.uz[mq]));
      const float dist = __fsub_rn(1.f, dot);
      if (dist > best) { best = dist; bp = p; bq = q; }      // first index on ties (torch.argmax)
    }
  }
  const int m1 = c_inc_m[bp], m2 = c_inc_m[bq];
  uint32_t bits = 0;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const int ms = c_inc_m[s];
    float ax = __fsub_rn(c.ux[m1], c.ux[ms]), ay = __fsub_rn(c.uy[m1], c.uy[ms]), az = __fsub_rn(c.uz[m1], c.uz[ms]);
    float bx = __fsub_rn(c.ux[m2], c.ux[ms]), by = __fsub_rn(c.uy[m2], c.uy[ms]), bz = __fsub_rn(c.uz[m2], c.uz[ms]);
    const float d1 = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az)));
    const float d2 = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(bx, bx), __fmul_rn(by, by)), __fmul_rn(bz, bz)));
    if (d2 < d1) bits |= 1u << s;                            // torch.argmin: the first vector wins ties
  }
  return bits;
}

// cell q of the block order (methods.py:186-192): block q / 8 in C order over (N/2)^3, cell q % 8 of the block in `inc` order
__device__ __forceinline__ void mc_cell_of(int64_t q, int N, int& i, int& j, int& k) {
  const int h = N >> 1;
  const int s = (int)(q & 7);
  const int64_t b = q >> 3;
  const int bz = (int)(b % h), by = (int)((b / h) % h), bx = (int)(b / ((int64_t)h * h));
  const int m = c_inc_m[s];
  i = 2 * bx + ((m >> 2) & 1); j = 2 * by + ((m >> 1) & 1); k = 2 * bz + (m & 1);
}

__global__ void __launch_bounds__(kMcThreads)
mc_count_kernel(const float* __restrict__ pred, int N, int64_t n_q, uint8_t* __restrict__ keep,
                int* __restrict__ cta_counts, float* __restrict__ div_raw, uint8_t* __restrict__ choice_out,
                const uint8_t* __restrict__ surface) {
  const int64_t q = (int64_t)blockIdx.x * kMcThreads + threadIdx.x;
  int kept = 0;
  if (q < n_q) {
    int i, j, k;
    mc_cell_of(q, N, i, j, k);
    uint32_t bits = 0;
    if (i < N - 1 && j < N - 1 && k < N - 1) {               // boundary cells never carry a surface (mc_utils.py:79-80)
      McCell c;
      mc_load(pred, N, i, j, k, c);
      const int64_t cell = ((int64_t)i * N + j) * N + k;
      bool surf;
      if (surface && !div_raw) surf = surface[cell] != 0;          // surface cells decided on another field (smooth_after)
      else {
        const float dv = mc_divergence(c);
        if (div_raw) div_raw[cell] = dv;
        surf = surface ? surface[cell] != 0 : dv <= -0.5f;
      }
      if (surf) {
        bits = mc_choice(c);
        kept = (bits != 0u && bits != 0xFFu) ? 1 : 0;        // some pair of corners on different sides
      }
    }
    if (choice_out) choice_out[((int64_t)i * N + j) * N + k] = (uint8_t)bits;
    keep[q] = (uint8_t)kept;
  }
  const int total = __syncthreads_count(kept);
  if (threadIdx.x == 0) cta_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kMcThreads)
mc_emit_kernel(const float* __restrict__ pred, int N, int64_t n_q, const uint8_t* __restrict__ keep,
               const int64_t* __restrict__ cta_offsets, int* __restrict__ cells, float* __restrict__ comb,
               float* __restrict__ udf) {
  __shared__ int s_warp[kMcThreads / 32];
  const int64_t q = (int64_t)blockIdx.x * kMcThreads + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kept = (q < n_q) ? keep[q] : 0;
  const uint32_t ball = __ballot_sync(kFull, kept);
  if (lane == 0) s_warp[warp] = __popc(ball);
  __syncthreads();
  if (!kept) return;
  int before = __popc(ball & ((1u << lane) - 1u));
  for (int w = 0; w < warp; ++w) before += s_warp[w];
  // cta_offsets is the INCLUSIVE scan of the per-CTA counts (what a plain cumsum gives)
  const int64_t row = (blockIdx.x ? cta_offsets[blockIdx.x - 1] : 0) + before;
  int i, j, k;
  mc_cell_of(q, N, i, j, k);
  McCell c;
  mc_load(pred, N, i, j, k, c);
  const uint32_t bits = mc_choice(c);
  cells[3 * row] = i; cells[3 * row + 1] = j; cells[3 * row + 2] = k;
  float* cr = comb + row * 28;
  float* ur = udf + row * 56;
  int t = 0;
#pragma unroll
  for (int a = 0; a < 7; ++a) {
#pragma unroll
    for (int b = a + 1; b < 8; ++b, ++t) {                   // pair order of mc_utils.py:205-211
      cr[t] = (((bits >> a) ^ (bits >> b)) & 1u) ? 1.f : 0.f;
      ur[2 * t] = c.nrm[c_inc_m[a]];
      ur[2 * t + 1] = c.nrm[c_inc_m[b]];
    }
  }
}

int launch_mc_count(const float* pred, int N, uint8_t* keep, int* cta_counts, float* div_raw, uint8_t* choice,
                    const uint8_t* surface, cudaStream_t s) {
  VFN_REQUIRE(pred && keep && cta_counts, "mc_count: null argument");
  VFN_REQUIRE(N >= 2 && N <= 2048, "mc_count: resolution %d out of range", N);
  const int64_t n_q = 8 * (int64_t)(N / 2) * (N / 2) * (N / 2);
  if (n_q == 0) return 0;
  mc_count_kernel<<<(unsigned)ceil_div64(n_q, kMcThreads), kMcThreads, 0, s>>>(pred, N, n_q, keep, cta_counts, div_raw, choice, surface);
  VFN_LAUNCH_CHECK();
  return 0;
}

int launch_mc_emit(const float* pred, int N, const uint8_t* keep, const int64_t* cta_offsets, int* cells, float* comb,
                   float* udf, cudaStream_t s) {
  VFN_REQUIRE(pred && keep && cta_offsets && cells && comb && udf, "mc_emit: null argument");
  const int64_t n_q = 8 * (int64_t)(N / 2) * (N / 2) * (N / 2);
  if (n_q == 0) return 0;
  mc_emit_kernel<<<(unsigned)ceil_div64(n_q, kMcThreads), kMcThreads, 0, s>>>(pred, N, n_q, keep, cta_offsets, cells, comb, udf);
  VFN_LAUNCH_CHECK();
  return 0;
}

}  // namespace vfn

// ---------------------------------------------------------------------------------------------
// smooth_vf (evaluation/utils/guassian_smoothing.py:81-97): depthwise 3-D gaussian of the [N,N,N,3] vector grid with
// replicate padding.  The reference's k^3-tap kernel is the outer product of one 1-D factor per axis (normalised as a
// whole, which equals normalising each factor), so it runs as three 1-D passes of k taps: 24 B/point of traffic per
// pass instead of 27 (k = 3) or 729 (k = 9) gathered taps per point.  One thread per grid point, taps clamped to the grid.
// ---------------------------------------------------------------------------------------------
namespace vfn {

constexpr int kMaxTaps = 31;
struct SmoothW { float w[kMaxTaps]; };

__global__ void __launch_bounds__(256)
smooth_axis_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int64_t stride, int k, SmoothW W) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)N * N * N;
  if (p >= total) return;
  const int c = (int)((p / stride) % N);           // coordinate along the smoothed axis
  const int h = k >> 1;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int t = 0; t < k; ++t) {
    const int cc = min(max(c + t - h, 0), N - 1);
    const float* q = in + 3 * (p + (int64_t)(cc - c) * stride);
    const float w = W.w[t];
    a0 = fmaf(w, __ldg(q), a0); a1 = fmaf(w, __ldg(q + 1), a1); a2 = fmaf(w, __ldg(q + 2), a2);
  }
  out[3 * p] = a0; out[3 * p + 1] = a1; out[3 * p + 2] = a2;
}

int launch_smooth_vf(const float* in, float* tmp, float* out, int N, int k, const float* w_host, cudaStream_t s) {
  VFN_REQUIRE(in && tmp && out && w_host, "smooth_vf: null argument");
  VFN_REQUIRE(k >= 1 && (k & 1) && k <= kMaxTaps, "smooth_vf: kernel size %d must be odd and <= %d", k, kMaxTaps);
  VFN_REQUIRE(N >= 1 && N <= 2048, "smooth_vf: resolution %d out of range", N);
  VFN_REQUIRE(in != tmp && tmp != out, "smooth_vf: tmp must not alias in / out");
  SmoothW W{};
  for (int t = 0; t < k; ++t) W.w[t] = w_host[t];
  const int64_t total = (int64_t)N * N * N;
  const unsigned grid = (unsigned)ceil_div64(total, 256);
  // x index is slowest (stride N^2), z fastest (stride 1): in -> out (z), out -> tmp (y), tmp -> out (x)
  smooth_axis_kernel<<<grid, 256, 0, s>>>(in, out, N, 1, k, W);
  VFN_LAUNCH_CHECK();
  smooth_axis_kernel<<<grid, 256, 0, s>>>(out, tmp, N, N, k, W);
  VFN_LAUNCH_CHECK();
  smooth_axis_kernel<<<grid, 256, 0, s>>>(tmp, out, N, (int64_t)N * N, k, W);
  VFN_LAUNCH_CHECK();
  return 0;
}

}  // namespace vfn
