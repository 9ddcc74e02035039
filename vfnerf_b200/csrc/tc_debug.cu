// Debug / regression kernel for the UMMA conventions of tc_common.cuh: one CTA computes
// D[128 x N] = bf16(A[128 x K]) * bf16(B[N x K])^T with tcgen05.mma (fp32 accumulate in TMEM).
// tests/test_gpu_tc.py compares it with a bf16-rounded fp32 matmul; `variant` 1 swaps LBO/SBO so a
// descriptor-convention mistake shows up as "variant 1 right, variant 0 wrong" instead of plain garbage.
// TEST-ONLY translation unit: part of libvfnerf_b200_debug.so (vfnerf_b200/_lib.py: build_debug), never of the product
// library.  Declarations: include/vfnerf_b200_debug.h.
#include "common.cuh"
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include "tc_common.cuh"
#include "../../include/vfnerf_b200_debug.h"

namespace vfn {
using namespace tc;

__global__ void __launch_bounds__(128) umma_debug_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                         float* __restrict__ D, int N, int K, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * K * 2;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 128 * K; e += 128) {
    int r = e / K, k = e % K;
    *reinterpret_cast<__nv_bfloat16*>(sA + slab_offset(128, r, k)) = __float2bfloat16(A[e]);
  }
  for (int e = tid; e < N * K; e += 128) {
    int r = e / K, k = e % K;
    *reinterpret_cast<__nv_bfloat16*>(sB + slab_offset(N, r, k)) = __float2bfloat16(B[e]);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t a_addr = smem_u32(sA) + (k0 / 8) * slab_bytes(128);
      uint32_t b_addr = smem_u32(sB) + (k0 / 8) * slab_bytes(N);
      uint64_t da = variant ? make_smem_desc(a_addr, 128, slab_bytes(128)) : make_smem_desc(a_addr, slab_bytes(128), 128);
      uint64_t db = variant ? make_smem_desc(b_addr, 128, slab_bytes(N)) : make_smem_desc(b_addr, slab_bytes(N), 128);
      umma_bf16(tmem, da, db, idesc, k0 > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(int64_t)row * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

}  // namespace vfn

extern "C" int vfnerf_debug_umma_gemm(const float* A, const float* B, float* D, int N, int K, int variant, void* stream) {
  using namespace vfn;
  VFN_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 256, "debug_umma_gemm: bad N/K");
  size_t smem = (size_t)(128 + N) * K * 2;
  VFN_CHECK_CUDA(cudaFuncSetAttribute(umma_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_debug_kernel<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(A, B, D, N, K, variant);
  VFN_LAUNCH_CHECK();
  return 0;
}

// ---- micro-benchmark: cycles per tcgen05.mma (M=128, N, K=16) issued back to back by one thread of every CTA.
// mode 0: MMAs only; mode 1: a tcgen05.commit after every second MMA (the ring-release pattern of mlp_tc.cu).
namespace vfn {
__global__ void __launch_bounds__(128) umma_bench_kernel(int N, int n_mma, int mode, int a_slabs, long long* out,
                                                         const uint8_t* gsrc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ volatile int done_flag;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < (64 + 128) * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1 << 20); mbar_init(&bar3, 1); done_flag = 0; fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 32 && (mode & 4)) {
    // background weight-ring traffic: 16 KB bulk copies global -> shared, back to back, until the MMAs are done
    int ph = 0;
    while (!done_flag) {
      mbar_arrive_expect_tx(&bar3, 16384);
      bulk_g2s(smem + 192 * 1024, gsrc + (size_t)(blockIdx.x % 64) * 16384, 16384, &bar3);
      mbar_wait(&bar3, ph);
      ph ^= 1;
    }
  }
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
    long long t0 = clock64();
    if (mode >= 8) {
      // lean issue loop: descriptor halves precomputed, two IADDs per MMA, 4 MMAs per iteration
      const uint64_t da0 = make_smem_desc(a0, slab_bytes(128), 128), db0 = make_smem_desc(b0, slab_bytes(N), 128);
      const uint32_t a_hi = (uint32_t)(da0 >> 32), b_hi = (uint32_t)(db0 >> 32);
      const uint32_t a_lo0 = (uint32_t)da0, b_lo0 = (uint32_t)db0;
      const uint32_t a_step = (2 * slab_bytes(128)) >> 4, b_step = (2 * slab_bytes(N)) >> 4;
      const bool stream = (mode & 2) != 0;    // walk over all 16 K-steps of the 64 KB A tile / 128 KB B tile instead of re-reading 4
      for (int i = 0; i < n_mma; i += 4) {
        const uint32_t base = stream ? (uint32_t)((i >> 2) & 3) * 4u : 0u;
        if (mode & 32) mbar_wait(&bar3, 1);          // already-complete phase: returns immediately
        if (mode & 16) tc_fence_after_sync();
        if (mode & 64) {
          // the runtime-trip-count inner loop of mlp_tc.cu
          uint32_t a_lo = a_lo0 + base * a_step, b_lo = b_lo0 + base * b_step, accumulate = i > 0;
          const int kc = (n_mma > 7) ? 64 : 48;
          for (int kk = 0; kk < kc; kk += 16) {
            umma_bf16_split(tmem, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
            accumulate = 1; a_lo += a_step; b_lo += b_step;
          }
        } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          umma_bf16_split(tmem, a_lo0 + (base + j) * a_step, a_hi, b_lo0 + (base + j) * b_step, b_hi, idesc, (i | j) > 0);
        }
        if (mode & 1) umma_commit(&bar2);
      }
    } else
    for (int i = 0; i < n_mma; ++i) {
      uint64_t da, db;
      const int lay = mode >> 1;          // 0: K-slab (no swizzle), 1: SWIZZLE_128B, 2: SWIZZLE_64B, 3: SWIZZLE_32B
      if (lay == 0) {
        da = make_smem_desc(a0 + ((2 * i) % a_slabs) * slab_bytes(128), slab_bytes(128), 128);
        db = make_smem_desc(b0 + ((2 * i) % 32) * slab_bytes(N), slab_bytes(N), 128);
      } else {
        const uint32_t rowb = lay == 1 ? 128 : (lay == 2 ? 64 : 32), lt = lay == 1 ? 2 : (lay == 2 ? 4 : 6);
        const uint32_t kstep = (i % (rowb / 32)) * 32;     // K advance inside the swizzle row
        const uint32_t blk = (i / (rowb / 32)) % 2;        // alternate between two K blocks
        da = make_smem_desc_sw(a0 + blk * 128 * rowb + kstep, 8 * rowb, lt);
        db = make_smem_desc_sw(b0 + blk * N * rowb + kstep, 8 * rowb, lt);
      }
      umma_bf16(tmem + (i & 256 ? 256 : 0), da, db, idesc, i > 0);
      if ((mode & 1) && (i & 1)) umma_commit(&bar2);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    done_flag = 1;
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}
}  // namespace vfn

extern "C" int vfnerf_debug_umma_bench(int N, int n_mma, int mode, int n_ctas, long long* cycles_dev, void* stream) {
  using namespace vfn;
  size_t smem = (64 + 128 + 16) * 1024;
  static uint8_t* gsrc = nullptr;
  if (!gsrc) { VFN_CHECK_CUDA(cudaMalloc(&gsrc, 64 * 16384)); VFN_CHECK_CUDA(cudaMemset(gsrc, 0, 64 * 16384)); }
  VFN_CHECK_CUDA(cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_bench_kernel<<<n_ctas, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(N, n_mma, mode, 32, cycles_dev, gsrc);
  VFN_LAUNCH_CHECK();
  return 0;
}

// ---- 2-CTA (cta_group::2) convention check: a cluster of two CTAs computes D[256 x N] = A[256 x K] * B[N x K]^T.
// CTA r holds rows [128r, 128r+128) of A and rows [N/2 * r, N/2 * (r+1)) of B at identical shared-memory offsets;
// the leader issues tcgen05.mma.cta_group::2 with M = 256; each CTA drains its own 128 accumulator rows.
namespace vfn {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
umma2_debug_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t rank = cluster_ctarank();
  const int Nh = N / 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * K * 2;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 128 * K; e += 128) {
    int r = e / K, k = e % K;
    *reinterpret_cast<__nv_bfloat16*>(sA + slab_offset(128, r, k)) = __float2bfloat16(A[(int64_t)(rank * 128 + r) * K + k]);
  }
  for (int e = tid; e < Nh * K; e += 128) {
    int r = e / K, k = e % K;
    *reinterpret_cast<__nv_bfloat16*>(sB + slab_offset(Nh, r, k)) = __float2bfloat16(B[(int64_t)(rank * Nh + r) * K + k]);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = make_idesc_bf16(256, N);
    for (int k0 = 0; k0 < K; k0 += 16) {
      const uint64_t da = make_smem_desc(smem_u32(sA) + (k0 / 8) * slab_bytes(128), slab_bytes(128), 128);
      const uint64_t db = make_smem_desc(smem_u32(sB) + (k0 / 8) * slab_bytes(Nh), slab_bytes(Nh), 128);
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
          "}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(k0 > 0)) : "memory");
    }
    // completion to the same barrier offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(int64_t)(rank * 128 + row) * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
}
}  // namespace vfn

extern "C" int vfnerf_debug_umma2_gemm(const float* A, const float* B, float* D, int N, int K, void* stream) {
  using namespace vfn;
  VFN_REQUIRE(N % 32 == 0 && N >= 32 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 256, "debug_umma2_gemm: bad N/K");
  size_t smem = (size_t)(128 + N / 2) * K * 2;
  VFN_CHECK_CUDA(cudaFuncSetAttribute(umma2_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma2_debug_kernel<<<2, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(A, B, D, N, K);
  VFN_LAUNCH_CHECK();
  return 0;
}

// ---- 2-CTA convention check for the other operand kinds the split-precision chain uses: kind::f16 with fp16 operands
// (fmt 2) and kind::f8f6f4 with 8-bit operands (a_fmt / b_fmt: 0 = e4m3, 1 = e5m2; K = 32 per instruction).  Same data
// placement as umma2_debug_kernel; the 8-bit K-slab holds 16 columns per 16-byte unit:
//   byte(r, k) = (k / 16) * rows * 16 + r * 16 + (k % 16)
namespace vfn {
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
umma2_alt_debug_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K,
                       int a_fmt, int b_fmt) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t rank = cluster_ctarank();
  const int Nh = N / 2;
  const bool f16 = a_fmt == 2;
  const int esz = f16 ? 2 : 1, per = 16 / esz;           // elements per 16-byte unit
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * K * esz;
  const int tid = threadIdx.x, warp = tid >> 5;
  auto put = [&](uint8_t* base, int rows, int r, int k, float v, int fmt) {
    uint8_t* q = base + (size_t)(k / per) * rows * 16 + r * 16 + (k % per) * esz;
    if (f16) *reinterpret_cast<__half*>(q) = __float2half_rn(v);
    else *q = (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, fmt ? __NV_E5M2 : __NV_E4M3);
  };
  for (int e = tid; e < 128 * K; e += 128) put(sA, 128, e / K, e % K, A[(int64_t)(rank * 128 + e / K) * K + e % K], a_fmt);
  for (int e = tid; e < Nh * K; e += 128) put(sB, Nh, e / K, e % K, B[(int64_t)(rank * Nh + e / K) * K + e % K], b_fmt);
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = f16 ? make_idesc_f16(256, N) : make_idesc_f8(256, N, a_fmt, b_fmt);
    const int kstep = f16 ? 16 : 32;                       // two 16-byte units per instruction either way
    for (int k0 = 0; k0 < K; k0 += kstep) {
      const uint64_t da = make_smem_desc(smem_u32(sA) + (k0 / per) * slab_bytes(128), slab_bytes(128), 128);
      const uint64_t db = make_smem_desc(smem_u32(sB) + (k0 / per) * slab_bytes(Nh), slab_bytes(Nh), 128);
      if (f16)
        umma2_bf16_split_w1(tmem, (uint32_t)da, (uint32_t)(da >> 32), (uint32_t)db, (uint32_t)(db >> 32), idesc, (uint32_t)(k0 > 0));
      else
        umma2_f8_split_w1(tmem, (uint32_t)da, (uint32_t)(da >> 32), (uint32_t)db, (uint32_t)(db >> 32), idesc, (uint32_t)(k0 > 0));
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(int64_t)(rank * 128 + row) * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
}
}  // namespace vfn

extern "C" int vfnerf_debug_umma2_alt_gemm(const float* A, const float* B, float* D, int N, int K, int a_fmt, int b_fmt,
                                           void* stream) {
  using namespace vfn;
  VFN_REQUIRE(N % 32 == 0 && N >= 32 && N <= 256 && K % 32 == 0 && K >= 32 && K <= 256, "debug_umma2_alt_gemm: bad N/K");
  VFN_REQUIRE((a_fmt == 2 && b_fmt == 2) || (a_fmt >= 0 && a_fmt <= 1 && b_fmt >= 0 && b_fmt <= 1), "debug_umma2_alt_gemm: formats");
  size_t smem = (size_t)(128 + N / 2) * K * 2;
  VFN_CHECK_CUDA(cudaFuncSetAttribute(umma2_alt_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma2_alt_debug_kernel<<<2, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(A, B, D, N, K, a_fmt, b_fmt);
  VFN_LAUNCH_CHECK();
  return 0;
}

// ---- layout probe: tcgen05.mma.cta_group::2 with M = 128 (64 rows per CTA).  Dumps all 128 TMEM lanes x 512 columns of
// both CTAs so the test can see where the 64 x N accumulator of each CTA lands, and whether it can be placed at a lane
// offset (two 64-row sub-tiles side by side in the lanes of the same columns).
namespace vfn {
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
umma2_m128_probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ dump, int N, int K,
                        int lane_off, int col_off) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t rank = cluster_ctarank();
  const int Nh = N / 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 64 * K * 2;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 64 * K; e += 128) {
    int r = e / K, k = e % K;
    *reinterpret_cast<__nv_bfloat16*>(sA + slab_offset(64, r, k)) = __float2bfloat16(A[(int64_t)(rank * 64 + r) * K + k]);
  }
  for (int e = tid; e < Nh * K; e += 128) {
    int r = e / K, k = e % K;
    *reinterpret_cast<__nv_bfloat16*>(sB + slab_offset(Nh, r, k)) = __float2bfloat16(B[(int64_t)(rank * Nh + r) * K + k]);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  // fill TMEM with a sentinel
  for (int c0 = 0; c0 < 512; c0 += 16) {
    const uint32_t z = __float_as_uint(-777.f);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
                 ::"r"(tmem + ((uint32_t)(warp * 32) << 16) + c0), "r"(z) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    for (int k0 = 0; k0 < K; k0 += 16) {
      const uint64_t da = make_smem_desc(smem_u32(sA) + (k0 / 8) * slab_bytes(64), slab_bytes(64), 128);
      const uint64_t db = make_smem_desc(smem_u32(sB) + (k0 / 8) * slab_bytes(Nh), slab_bytes(Nh), 128);
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
          "}" ::"r"(tmem + ((uint32_t)lane_off << 16) + col_off), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(k0 > 0)) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  const int lane_row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < 512; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) dump[((int64_t)rank * 128 + lane_row) * 512 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}
}  // namespace vfn

extern "C" int vfnerf_debug_umma2_m128_probe(const float* A, const float* B, float* dump, int N, int K, int lane_off,
                                             int col_off, void* stream) {
  using namespace vfn;
  VFN_REQUIRE(N % 32 == 0 && N >= 32 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 256, "debug_umma2_m128_probe: bad N/K");
  size_t smem = (size_t)(64 + N / 2) * K * 2;
  VFN_CHECK_CUDA(cudaFuncSetAttribute(umma2_m128_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma2_m128_probe_kernel<<<2, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(A, B, dump, N, K, lane_off, col_off);
  VFN_LAUNCH_CHECK();
  return 0;
}

// ---- 2-CTA micro-benchmark: cycles per tcgen05.mma.cta_group::2 (M=256, N=256, K=16), with an optional
// multicast tcgen05.commit after every 4th MMA (mode 1) or a non-multicast commit to the leader only (mode 2).
namespace vfn {
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
umma2_bench_kernel(int n_mma, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar, bar2;
  __shared__ uint32_t tmem_base_s;
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 128 * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1 << 20); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (rank == 0 && tid == 0) {
    const bool f8_all = (mode & 16) != 0;                                                    // mode bit 4: kind::f8f6f4 (e4m3 x e5m2, K = 32)
    const uint32_t idesc_8 = make_idesc_f8((mode & 4) ? 128 : 256, (mode & 8) ? 128 : 256, 0, 1);
    const uint32_t idesc_16 = (mode & 128) ? make_idesc_f16((mode & 4) ? 128 : 256, (mode & 8) ? 128 : 256)      // bit 7: fp16 operands
                                           : make_idesc_bf16((mode & 4) ? 128 : 256, (mode & 8) ? 128 : 256);   // mode bit 2: M = 128 (64 rows per CTA); bit 3: N = 128
    const uint64_t da0 = make_smem_desc(smem_u32(smem), slab_bytes(128), 128);
    const uint64_t db0 = make_smem_desc(smem_u32(smem + 64 * 1024), slab_bytes(128), 128);
    const uint32_t a_hi = (uint32_t)(da0 >> 32), b_hi = (uint32_t)(db0 >> 32), a_lo0 = (uint32_t)da0, b_lo0 = (uint32_t)db0;
    const uint32_t step = (2 * slab_bytes(128)) >> 4;
    long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 4) {
      const uint32_t base = (uint32_t)((i >> 2) & 3) * 4u;
      // mode bit 5: alternate the kind every 4 instructions (16-bit, 8-bit, 16-bit, ...); bit 6: every 16
      const bool f8 = (mode & 32) ? (((i >> 2) & 1) != 0) : ((mode & 64) ? (((i >> 4) & 1) != 0) : f8_all);
      const uint32_t idesc = f8 ? idesc_8 : idesc_16;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (f8)
          umma2_f8_split_w1(tmem, a_lo0 + (base + j) * step, a_hi, b_lo0 + (base + j) * step, b_hi, idesc, (uint32_t)((i | j) > 0));
        else
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
            ::"r"(tmem), "r"(a_lo0 + (base + j) * step), "r"(a_hi), "r"(b_lo0 + (base + j) * step), "r"(b_hi), "r"(idesc),
              "r"((uint32_t)((i | j) > 0)) : "memory");
      }
      if ((mode & 3) == 1)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar2)), "h"((uint16_t)3) : "memory");
      if ((mode & 3) == 2)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  if (rank == 1 && tid == 0) mbar_wait(&bar, 0);
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}
}  // namespace vfn

extern "C" int vfnerf_debug_umma2_bench(int n_mma, int mode, int n_ctas, long long* cycles_dev, void* stream) {
  using namespace vfn;
  size_t smem = 128 * 1024;
  VFN_CHECK_CUDA(cudaFuncSetAttribute(umma2_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma2_bench_kernel<<<n_ctas & ~1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(n_mma, mode, cycles_dev);
  VFN_LAUNCH_CHECK();
  return 0;
}

// ---- MN-major operand convention check (the weight-gradient GEMM): D[M=128, N] = At[K, 128]^T * Bt[K, N] where At / Bt
// are row-major with the REDUCTION index as the row (points x channels, exactly how activations are stashed).
// Shared-memory image: slab s holds channels 8s..8s+7 of every row k as one 16-byte unit, rows contiguous
// (the same K-slab image the forward uses, read here with the major bits set): unit(m/8, k) at (m/8)*K*16 + k*16.
// For an MN-major operand the canonical no-swizzle layout is ((8 m-elems),(8 k)) core matrices of 128 contiguous
// bytes; SBO = stride between m-groups (K*16 bytes here), LBO = stride between k-groups of 8 (128 bytes).
namespace vfn {
__global__ void __launch_bounds__(128) umma_mn_debug_kernel(const float* __restrict__ At, const float* __restrict__ Bt,
                                                            float* __restrict__ D, int N, int K, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * K * 2;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < K * 128; e += 128) {
    int k = e / 128, m = e % 128;
    *reinterpret_cast<__nv_bfloat16*>(sA + (m / 8) * K * 16 + k * 16 + (m % 8) * 2) = __float2bfloat16(At[e]);
  }
  for (int e = tid; e < K * N; e += 128) {
    int k = e / N, n = e % N;
    *reinterpret_cast<__nv_bfloat16*>(sB + (n / 8) * K * 16 + k * 16 + (n % 8) * 2) = __float2bfloat16(Bt[e]);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N) | (1u << 15) | (1u << 16);   // a_major = b_major = MN
    const uint32_t sbo = K * 16, lbo = 128;
    for (int k0 = 0; k0 < K; k0 += 16) {
      const uint32_t a_addr = smem_u32(sA) + k0 * 16, b_addr = smem_u32(sB) + k0 * 16;
      const uint64_t da = variant ? make_smem_desc(a_addr, sbo, lbo) : make_smem_desc(a_addr, lbo, sbo);
      const uint64_t db = variant ? make_smem_desc(b_addr, sbo, lbo) : make_smem_desc(b_addr, lbo, sbo);
      umma_bf16(tmem, da, db, idesc, k0 > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(int64_t)row * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}
}  // namespace vfn

extern "C" int vfnerf_debug_umma_mn_gemm(const float* At, const float* Bt, float* D, int N, int K, int variant, void* stream) {
  using namespace vfn;
  VFN_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 256, "debug_umma_mn_gemm: bad N/K");
  size_t smem = (size_t)(128 + N) * K * 2;
  VFN_CHECK_CUDA(cudaFuncSetAttribute(umma_mn_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_mn_debug_kernel<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(At, Bt, D, N, K, variant);
  VFN_LAUNCH_CHECK();
  return 0;
}
