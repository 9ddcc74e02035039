"""Driver for profiling one training step (render -> loss -> backward -> clip -> Adam) on 1024 rays.
Usage: python profiles/run_train_step.py [precision] [n_steps] [rays] [supervision points] [native|torch]
  native (default): fused VFLoss (losses.py) + ArenaAdam (optim.py); torch: inline aten loss + torch Adam / clip_grad_norm_
Prints event-timed ms/step and a host/device split (host time = wall time of the python step with no sync)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import vfn_testutil as U

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
Rt = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
SUP = int(sys.argv[4]) if len(sys.argv) > 4 else 0        # supervision points through the VF-only entry (trainer: 2 x 13107)
NATIVE = (sys.argv[5] if len(sys.argv) > 5 else "native") == "native"
dev = "cuda"
case, z = U.load_golden("full_det")
case = dict(case, perturb=True, dir_to_normal_th=-2.0)
st = U.case_state(case, z)
tm = U.make_model(case, st, dev, precision=prec)
if NATIVE:
    import types
    from vfnerf_b200 import optim
    from vfnerf_b200.losses import VFLoss
    optim.use_arena_optimizer(tm, max_norm=0.5)
    loss_mod = VFLoss(types.SimpleNamespace(norm_smaller_than_one_start=11000, depth_loss_clamp=0.5, directional_derivatives_start=100),
                      types.SimpleNamespace(rgb=2.0, depth=0.5, unit_norm=0.1, supervision=1.0, norm_smaller_than_one=0.1,
                                            directional_derivatives=0.0), sync=False)
uv, pose, K = (t.to(dev) for t in U.S.synthetic_rays(Rt, seed=0, start=40000, stride=25013))
g2 = torch.Generator(device=dev).manual_seed(7)
Nc, Nf = case["n_coarse"], case["n_fine"]
draws = (torch.rand(Rt, Nc, device=dev, generator=g2), torch.rand(Rt, Nf, device=dev, generator=g2),
         torch.rand(Rt, Nf, device=dev, generator=g2))
sup_pts = (torch.rand(max(SUP, 1), 3, device=dev, generator=g2) - 0.5) * 6
sup_tgt = torch.nn.functional.normalize(torch.randn(max(SUP, 1), 3, device=dev, generator=g2), dim=1)
rgb_gt = torch.rand(Rt, 3, device=dev, generator=g2)
dep_gt = torch.rand(Rt, 1, device=dev, generator=g2) * case["far"]


def train_step(parts=None):
    t = time.perf_counter()
    out = tm.render(pose, uv, K, 0, draws=draws)
    if parts is not None: torch.cuda.synchronize(); parts["render"] += time.perf_counter() - t; t = time.perf_counter()
    if NATIVE:
        sup = tm.vector_field_network(sup_pts)[:, :3] if SUP else None
        loss = loss_mod({"rgb": out.coarse_rgb_values, "depth": out.coarse_depth_map, "normals": out.coarse_normals.reshape(-1, 3),
                         "supervised_normals": sup, "directional_derivatives": None},
                        {"rgb": rgb_gt, "depth": dep_gt, "supervised_normals": sup_tgt}, 0)[0]
    else:
        nrm = torch.norm(out.coarse_normals.reshape(-1, 3), dim=1)
        loss = 2.0 * (out.coarse_rgb_values - rgb_gt).abs().mean() + \
            0.5 * (out.coarse_depth_map - dep_gt).abs().clamp(max=0.5).mean() + 0.1 * torch.mean((nrm - 1) ** 2)
        if SUP:
            loss = loss + ((tm.vector_field_network(sup_pts)[:, :3] - sup_tgt) ** 2).mean()
    tm.optimizer.zero_grad()
    if parts is not None: torch.cuda.synchronize(); parts["loss"] += time.perf_counter() - t; t = time.perf_counter()
    loss.backward()
    if parts is not None: torch.cuda.synchronize(); parts["backward"] += time.perf_counter() - t; t = time.perf_counter()
    if not NATIVE:
        torch.nn.utils.clip_grad_norm_(tm.parameters(), 0.5)
    if parts is not None: torch.cuda.synchronize(); parts["clip"] += time.perf_counter() - t; t = time.perf_counter()
    tm.optimizer.step()
    if parts is not None: torch.cuda.synchronize(); parts["adam"] += time.perf_counter() - t


for _ in range(3):
    train_step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(n_steps):
    train_step()
e1.record()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"{prec} R={Rt} sup={SUP} {'native' if NATIVE else 'torch'}: {e0.elapsed_time(e1) / n_steps:.3f} ms/step (events); host issue time {t_host / n_steps * 1e3:.3f} ms/step")
parts = dict(render=0.0, loss=0.0, backward=0.0, clip=0.0, adam=0.0)
for _ in range(n_steps):
    train_step(parts)
print("synchronised per-phase ms:", {k: round(v / n_steps * 1e3, 3) for k, v in parts.items()})
