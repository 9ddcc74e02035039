// Debug / regression kernel for the UMMA conventions of tc_common.cuh: one CTA computes
// D[128 x N] = bf16(A[128 x K]) * bf16(B[N x K])^T with tcgen05.mma (fp32 accumulate in TMEM).
// tests/test_gpu_tc.py compares it with a bf16-rounded fp32 matmul; `variant` 1 swaps LBO/SBO so a
// descriptor-convention mistake shows up as "variant 1 right, variant 0 wrong" instead of plain garbage.
#include "common.cuh"
#include "tc_common.cuh"

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
__global__ void __launch_bounds__(128) umma_bench_kernel(int N, int n_mma, int mode, int a_slabs, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < (64 + 128) * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1 << 20); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
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
      for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          umma_bf16_split(tmem, a_lo0 + j * a_step, a_hi, b_lo0 + j * b_step, b_hi, idesc, (i | j) > 0);
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
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}
}  // namespace vfn

extern "C" int vfnerf_debug_umma_bench(int N, int n_mma, int mode, int n_ctas, long long* cycles_dev, void* stream) {
  using namespace vfn;
  size_t smem = (64 + 128) * 1024;
  VFN_CHECK_CUDA(cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_bench_kernel<<<n_ctas, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(N, n_mma, mode, 32, cycles_dev);
  VFN_LAUNCH_CHECK();
  return 0;
}
