"""Driver used under ncu: a few launches of the fused RENDER program (VF + colour MLPs, bf16 tcgen05) on one bench
chunk of merged sample points (65536 rays x 128 samples) -- the same call bench.py's roofline leg times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import vfn_testutil as U
from vfnerf_b200 import ops, _lib
if len(sys.argv) > 3:                 # A/B timing of a differently built library (e.g. NVCC_EXTRA=-DVFN_X3_SERIAL_GROUPS=0)
    _lib.LIB_PATH = os.path.abspath(sys.argv[3])

rays = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"          # bf16 | bf16x3
N = 128
case, z = U.load_golden("full_det")
model = U.make_model(case, U.case_state(case, z), "cuda", precision=prec)
pts = (torch.rand(rays * N, 3, device="cuda") - 0.5) * 6
dirs = torch.nn.functional.normalize(torch.randn(rays, 3, device="cuda"), dim=1)
ws = None
with torch.no_grad():
    for i in range(6):
        _, _, ws = ops.mlp_points(model.vector_field_network, model.rendering_network, pts, dirs, N, workspace=ws, repack=(i == 0))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.mlp_points(model.vector_field_network, model.rendering_network, pts, dirs, N, workspace=ws, repack=False)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"{prec}: {rays * N} points: {ms:.3f} ms/launch, {1592832 * rays * N / ms / 1e9:.1f} TFLOP/s algorithmic")
