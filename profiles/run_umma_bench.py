"""cycles per tcgen05.mma (K-slab no-swizzle operands) -- see csrc/tc_debug.cu"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vfnerf_b200 import _lib
_lib.build_debug()
L = _lib.debug_lib()
out = torch.zeros(4, dtype=torch.int64, device="cuda")
for n_ctas in (148,):
    for N in (256,):
        for mode in (9, 9 + 16, 9 + 32, 9 + 64, 9 + 16 + 32 + 64):
            n = 2048
            _lib.check_debug(L.vfnerf_debug_umma_bench(N, n, mode, n_ctas, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "bench")
            torch.cuda.synchronize()
            print(f"ctas={n_ctas:3d} N={N:3d} mode={mode} (commit per 4; +16 tcgen05 fence, +32 mbarrier wait, +64 runtime kk loop): {out[0].item() / n:7.1f} cycles / MMA (ideal {128 * N / 256:.0f})")

for mode in (0, 1, 2):
    n = 2048
    _lib.check_debug(L.vfnerf_debug_umma2_bench(n, mode, 148, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "bench2")
    torch.cuda.synchronize()
    print(f"2-CTA pairs=74 N=256 M=256 mode={mode} (0 none, 1 multicast commit / 4 MMAs, 2 leader-only commit): "
          f"{out[0].item() / n:7.1f} cycles / MMA (ideal 128)")

for mode, what in ((0, "M=256 N=256"), (4, "M=128 (64 rows per CTA) N=256"), (8, "M=256 N=128"), (12, "M=128 N=128")):
    n = 2048
    _lib.check_debug(L.vfnerf_debug_umma2_bench(n, mode, 148, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "bench2")
    torch.cuda.synchronize()
    print(f"2-CTA {what}: {out[0].item() / n:7.1f} cycles / MMA")
