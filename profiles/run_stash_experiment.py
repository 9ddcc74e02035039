"""Experiment driver: device time of the stash-writing forward (RENDER_STASH) and of the backward (tail + dgrad chain +
wgrad + finalize) of one training batch, timed separately with CUDA events.  Run with VFNERF_TC_DBG=8 (no activation /
gradient stash stores; results garbage) or 24 (also no gate bits) to see what the stash stores cost the fused kernels.
Usage: python profiles/run_stash_experiment.py [rays] [literal: 0|1]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import vfn_testutil as U

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
literal = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
dev = "cuda"
case, z = U.load_golden("full_det")
case = dict(case, perturb=True, dir_to_normal_th=-2.0)
tm = U.make_model(case, U.case_state(case, z), dev, precision="bf16")
tm.recompute_coarse = literal
from vfnerf_b200 import optim
optim.use_arena_optimizer(tm, max_norm=0.5)
uv, pose, K = (t.to(dev) for t in U.S.synthetic_rays(R, seed=0, start=40000, stride=25013))
g2 = torch.Generator(device=dev).manual_seed(7)
Nc, Nf = case["n_coarse"], case["n_fine"]
draws = tuple(torch.rand(R, n, device=dev, generator=g2) for n in (Nc, Nf, Nf))
c_rgb, c_dep = torch.randn(R, 3, device=dev), torch.randn(R, 1, device=dev)
tf = tb = 0.0
n = 8
for i in range(n + 2):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    e[0].record()
    out = tm.render(pose, uv, K, 0, draws=draws)
    e[1].record()
    loss = (out.coarse_rgb_values * c_rgb).sum() + (out.coarse_depth_map * c_dep).sum()
    tm.optimizer.zero_grad()
    torch.cuda.synchronize()
    e[2].record()
    loss.backward()
    e[3].record()
    torch.cuda.synchronize()
    if i >= 2:
        tf += e[0].elapsed_time(e[1]); tb += e[2].elapsed_time(e[3])
print(f"VFNERF_TC_DBG={os.environ.get('VFNERF_TC_DBG', '0')} rays={R} literal={literal}: forward {tf / n:.3f} ms, backward {tb / n:.3f} ms")
