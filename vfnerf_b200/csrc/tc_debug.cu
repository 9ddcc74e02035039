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
