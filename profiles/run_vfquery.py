"""Tiny driver used under ncu: a few VF-MLP launches (bf16 tcgen05 chain) on random points."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import vfn_testutil as U
from vfnerf_b200 import ops

P = int(sys.argv[1]) if len(sys.argv) > 1 else 128 * 296 * 8
ncols = int(sys.argv[2]) if len(sys.argv) > 2 else 3
case, z = U.load_golden("full_det")
model = U.make_model(case, U.case_state(case, z), "cuda", precision="bf16")
pts = (torch.rand(P, 3, device="cuda") - 0.5) * 6
with torch.no_grad():
    for _ in range(4):
        ops.vf_query(model.vector_field_network, pts, n_cols=ncols)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.vf_query(model.vector_field_network, pts, n_cols=ncols)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"P={P} ncols={ncols}: {ms:.3f} ms/launch, {P / ms / 1e3:.1f} Mpts/s, {1050112 * P / ms / 1e9:.1f} TFLOP/s")
