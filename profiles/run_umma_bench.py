"""cycles per tcgen05.mma (K-slab no-swizzle operands) -- see csrc/tc_debug.cu"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vfnerf_b200 import _lib
_lib.build()
L = _lib.lib()
out = torch.zeros(4, dtype=torch.int64, device="cuda")
for n_ctas in (1, 148):
    for N in (256, 128, 64, 16):
        for mode in (0, 8, 9):
            n = 2048
            _lib.check(L.vfnerf_debug_umma_bench(N, n, mode, n_ctas, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "bench")
            torch.cuda.synchronize()
            print(f"ctas={n_ctas:3d} N={N:3d} mode={mode} (layout {mode >> 1}, commit {mode & 1}): {out[0].item() / n:7.1f} cycles / MMA (ideal {128 * N / 256:.0f})")
