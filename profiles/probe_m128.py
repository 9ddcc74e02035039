"""Where does tcgen05.mma.cta_group::2 with M = 128 put its accumulator?  (layout probe, see tc_debug.cu)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vfnerf_b200 import _lib
L = _lib.lib()
N, K = 64, 64
g = torch.Generator().manual_seed(0)
A = torch.randn(128, K, generator=g).cuda()
B = torch.randn(N, K, generator=g).cuda()
ref = (A.bfloat16().float() @ B.bfloat16().float().T)          # [128, N]
for lane_off, col_off in ((0, 0), (0, 128)):
    dump = torch.zeros(2, 128, 512, device="cuda")
    _lib.check_debug(L.vfnerf_debug_umma2_m128_probe(A.data_ptr(), B.data_ptr(), dump.data_ptr(), N, K, lane_off, col_off,
                                               torch.cuda.current_stream().cuda_stream), "probe")
    torch.cuda.synchronize()
    d = dump.cpu()
    refc = ref.cpu()
    print(f"--- lane_off {lane_off} col_off {col_off}: where is D[row, col]?")
    for rank in range(2):
        for (r, c) in ((0, 0), (0, 1), (0, 31), (0, 32), (0, 63), (1, 0), (31, 5), (32, 5), (63, 5), (63, 40), (64, 0), (64, 33), (100, 7), (127, 63)):
            hit = (d[rank] - refc[r, c]).abs() < 1e-3 * max(1.0, abs(refc[r, c].item()))
            locs = hit.nonzero().tolist()
            print(f" CTA {rank} D[{r},{c}]={refc[r, c].item():+.3f} -> (lane, col) {locs[:3]}")
