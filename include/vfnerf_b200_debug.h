/* vfnerf_b200_debug -- TEST-ONLY entry points, built by the test-suite into libvfnerf_b200_debug.so
 * (vfnerf_b200/_lib.py: build_debug()).  Nothing here ships in libvfnerf_b200.so. */
#ifndef VFNERF_B200_DEBUG_H
#define VFNERF_B200_DEBUG_H

#include "vfnerf_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* error text of the last failing call made INTO THIS library (it has its own copy of the error state) */
const char* vfnerf_debug_last_error(void);

/* One CTA: D[128,N] = bf16(A[128,K]) * bf16(B[N,K])^T through tcgen05.mma + TMEM; pins the UMMA
 * descriptor conventions of csrc/tc_common.cuh (variant 1 = LBO/SBO swapped, expected to be wrong). */
int vfnerf_debug_umma_gemm(const float* A, const float* B, float* D, int N, int K, int variant, void* stream);
/* MN-major operands (both with the reduction index as the row, as activations are stashed):
 * D[128,N] = bf16(At[K,128])^T * bf16(Bt[K,N]).  variant 1 swaps LBO/SBO (expected wrong). */
int vfnerf_debug_umma_mn_gemm(const float* At, const float* Bt, float* D, int N, int K, int variant, void* stream);
/* 2-CTA variant (cluster of two, tcgen05.mma.cta_group::2, M = 256): D[256,N] = bf16(A[256,K]) * bf16(B[N,K])^T */
int vfnerf_debug_umma2_gemm(const float* A, const float* B, float* D, int N, int K, void* stream);
/* Layout probe for cta_group::2 with M = 128 (64 rows per CTA): A [128,K], B [N,K]; the accumulator is placed at TMEM
 * (lane_off, col_off); dump receives all 128 lanes x 512 columns of both CTAs ([2,128,512] floats, sentinel -777). */
int vfnerf_debug_umma2_m128_probe(const float* A, const float* B, float* dump, int N, int K, int lane_off, int col_off,
                                  void* stream);
/* Micro-benchmark: every CTA issues n_mma back-to-back tcgen05.mma (M=128, N, K=16) from one thread; CTA 0 writes
 * the elapsed SM cycles to cycles_dev[0].  mode 1 adds a tcgen05.commit after every second MMA. */
int vfnerf_debug_umma_bench(int N, int n_mma, int mode, int n_ctas, long long* cycles_dev, void* stream);
/* 2-CTA convention check for kind::f16 with fp16 operands (a_fmt = b_fmt = 2) and kind::f8f6f4 with 8-bit operands
 * (0 = e4m3, 1 = e5m2; K multiple of 32): D[256,N] = round(A)[256,K] * round(B)[N,K]^T */
int vfnerf_debug_umma2_alt_gemm(const float* A, const float* B, float* D, int N, int K, int a_fmt, int b_fmt, void* stream);
/* Same for tcgen05.mma.cta_group::2 (M=256, N=256): mode 0 no commits, 1 multicast commit per 4 MMAs, 2 leader-only commit;
 * +4: M = 128, +8: N = 128, +16: kind::f8f6f4 (K = 32 per instruction) */
int vfnerf_debug_umma2_bench(int n_mma, int mode, int n_ctas, long long* cycles_dev, void* stream);
/* Test support for the bf16 training path: after vfnerf_render_fwd(keep_for_backward = 1) [and vfnerf_render_bwd] on
 * `workspace`, convert activation-stash tensor `tensor` to row-major fp32 out[n_rays * n_samples, *n_cols].
 * Numbering (L = VF layers, Lr = colour layers, n_y = L + Lr - 1): 0..L-2 VF hidden activations, L-1 features,
 * L..n_y-1 colour hidden activations, n_y layer-0 input (embedding hi | lo), n_y+1 skip input, n_y+2 small colour
 * inputs, n_y+3+i = dL/d(pre-activation) of tensor i (valid after the backward); 1000 = [n,6] gradients wrt the
 * pre-activations of the colour (cols 0..2) and vector (cols 3..5) outputs.  out may be NULL to query *n_cols. */
int vfnerf_debug_stash_read(const vfnerf_render_cfg* cfg, const vfnerf_mlp_desc* vf, const vfnerf_mlp_desc* rn,
                            void* workspace, int tensor, float* out, int* n_cols, void* stream);

/* Host only (no launch, no GPU): the per-chunk records the fused tensor-core kernel's MMA issuer / weight producer / relay
 * lane read from the kernel parameters, for one program of the render() plan (0 render, 1 VF + features, 2 VF vector only,
 * 3 dgrad chain of render() [keep_for_backward], 4 dgrad chain of the VF net).  records [n_steps][20][4] = {A offset,
 * K columns | barrier mask << 16, bytes | offset/16 << 16, flags}; n_chunks [n_steps]; step_facts [n_steps][6] = {N, K,
 * chunk_k, n_seg, fresh_mask, use_lo}. */
int vfnerf_debug_chunk_table(const vfnerf_render_cfg* cfg, const vfnerf_mlp_desc* vf, const vfnerf_mlp_desc* rn,
                             int keep_for_backward, int program, uint32_t* records, int* n_chunks, int* step_facts,
                             int* n_steps);

#ifdef __cplusplus
}
#endif
#endif /* VFNERF_B200_DEBUG_H */
