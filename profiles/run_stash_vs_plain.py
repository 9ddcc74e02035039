"""Under `ncu --metrics gpu__time_duration.sum`: the RENDER program on the same 8192-ray batch without a stash
(torch.no_grad) and with the activation stash (grad enabled), reuse schedule (two launches of 524 288 points each)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import vfn_testutil as U

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = "cuda"
case, z = U.load_golden("full_det")
case = dict(case, perturb=True, dir_to_normal_th=-2.0)
tm = U.make_model(case, U.case_state(case, z), dev, precision="bf16")
uv, pose, K = (t.to(dev) for t in U.S.synthetic_rays(R, seed=0, start=40000, stride=25013))
g2 = torch.Generator(device=dev).manual_seed(7)
draws = tuple(torch.rand(R, n, device=dev, generator=g2) for n in (64, 64, 64))
for _ in range(3):
    with torch.no_grad():
        tm.render(pose, uv, K, 0, draws=draws)
    out = tm.render(pose, uv, K, 0, draws=draws)
    del out
torch.cuda.synchronize()
print("ok")
