"""Probe (round 2): tcgen05.mma.cta_group::2 with fp16 operands (kind::f16) and 8-bit operands (kind::f8f6f4) in the
no-swizzle K-slab layout -- correctness against torch, cycles per instruction, and how 8-bit products accumulate into an
fp32 accumulator that already holds O(1) values (the question behind an fp16 + fp8-remainder split-precision scheme)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vfnerf_b200 import _lib
_lib.build_debug()
L = _lib.debug_lib()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
st = torch.cuda.current_stream().cuda_stream


def gemm(A, B, N, K, af, bf):
    D = torch.empty(256, N, device=dev)
    _lib.check_debug(L.vfnerf_debug_umma2_alt_gemm(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, af, bf, st), "alt_gemm")
    torch.cuda.synchronize()
    return D


def rnd(x, fmt):
    if fmt == 2:
        return x.half().float()
    return x.to(torch.float8_e5m2 if fmt else torch.float8_e4m3fn).float()


for (af, bf, name) in ((2, 2, "fp16 x fp16"), (0, 0, "e4m3 x e4m3"), (0, 1, "e4m3 x e5m2"), (1, 0, "e5m2 x e4m3")):
    for N, K in ((256, 256), (64, 32), (128, 96 if af == 2 else 64)):
        A = torch.randn(256, K, device=dev, generator=g)
        B = torch.randn(N, K, device=dev, generator=g) * 0.5
        D = gemm(A, B, N, K, af, bf)
        ref = rnd(A, af).double() @ rnd(B, bf).double().t()
        err = (D.double() - ref).abs().max().item()
        print(f"{name:14s} N={N:3d} K={K:3d}: max abs err vs fp64 product of the rounded operands {err:.3e} (|ref| max {ref.abs().max().item():.2f})")

# products of tiny 8-bit operands: exactness of the accumulation (each product exactly representable)
K, N = 256, 256
A = torch.full((256, K), 2.0 ** -6, device=dev); B = torch.full((N, K), 2.0 ** -8, device=dev)
A[:, 0] = 1.0; B[:, 0] = 1.0                     # one O(1) product followed by 255 products of 2^-14
D = gemm(A, B, N, K, 0, 1)
print("e4m3 x e5m2, 1 + 255 * 2^-14 =", repr(D[0, 0].item()), "exact:", 1 + 255 * 2.0 ** -14)

cyc = torch.zeros(1, dtype=torch.int64, device=dev)
for mode, name in ((0, "kind::f16   M=256 N=256 K=16"), (16, "kind::f8f6f4 M=256 N=256 K=32"),
                   (32, "alternating 4 x f16 / 4 x f8f6f4"), (64, "alternating 16 x f16 / 16 x f8f6f4"),
                   (128, "kind::f16 with FP16 operands"), (128 + 32, "alternating 4 x fp16 / 4 x f8f6f4")):
    for n_ctas in (2, 148):
        _lib.check_debug(L.vfnerf_debug_umma2_bench(4096, mode, n_ctas, cyc.data_ptr(), st), "bench")
        torch.cuda.synchronize()
        print(f"{name}, {n_ctas} CTAs: {cyc.item() / 4096:.1f} cycles per MMA")
