import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
import vfn_testutil as U
from vfnerf_b200 import graphed
DEV = "cuda"
case, z = U.load_golden("full_det")
st = U.case_state(case, z)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 96
uv, pose, K = (t.to(DEV) for t in U.S.synthetic_rays(R, seed=0, start=case["start"], stride=case["stride"]))
draws = tuple(d.to(DEV) for d in U.S.synthetic_draws(R, case["n_coarse"], case["n_fine"]))
g = torch.Generator().manual_seed(3)
rgb_gt, dep_gt = torch.rand(R, 3, generator=g).to(DEV), (torch.rand(R, 1, generator=g) * case["far"]).to(DEV)
def _loss(out, rgb_gt, depth_gt):
    nrm = torch.norm(out.coarse_normals.reshape(-1, 3), dim=1)
    return 2.0 * (out.coarse_rgb_values - rgb_gt).abs().mean() + 0.5 * (out.coarse_depth_map - depth_gt).abs().clamp(max=0.5).mean() + 0.1 * torch.mean((nrm - 1) ** 2)
m = U.make_model(case, st, DEV, precision="bf16")
def eager_grads():
    out = m.render(pose, uv, K, 0, draws=draws)
    loss = _loss(out, rgb_gt, dep_gt)
    m.optimizer.zero_grad()
    loss.backward()
    torch.cuda.synchronize()
    return [p.grad.detach().clone() for p in m.vector_field_network.parameters()] + [p.grad.detach().clone() for p in m.rendering_network.parameters()]
a = eager_grads(); b = eager_grads()
print("eager run-to-run worst rel:", max(((x - y).norm() / (x.norm() + 1e-20)).item() for x, y in zip(a, b)))
graphed.make_capturable(m)
step = graphed.GraphedTrainStep(m, _loss, R, dict(rgb_gt=rgb_gt, depth_gt=dep_gt), clip_norm=None, optimizer_step=False, given_draws=True)
for it in range(3):
    step(pose, uv, K, draws=draws, rgb_gt=rgb_gt, depth_gt=dep_gt)
    torch.cuda.synchronize()
    c = [p.grad.detach().clone() for p in m.vector_field_network.parameters()] + [p.grad.detach().clone() for p in m.rendering_network.parameters()]
    rels = [((x - y).norm() / (x.norm() + 1e-20)).item() for x, y in zip(a, c)]
    print("replay", it, "worst rel vs eager:", max(rels), [round(r, 3) for r in rels[:12]])
