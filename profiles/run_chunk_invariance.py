"""Chunk invariance at bench scale: one 65 536-ray chunk of the bench image rendered in one call, again in one call, and in
64 calls of 1024 rays (different tile -> cluster assignment, different timing of every hand-off inside the fused kernel) must
agree BIT FOR BIT in every output field -- a race in the activation tile (e.g. on the aliased aux / remainder columns of the
split-precision tile) would show up here as run-to-run or chunk-to-chunk differences.
Usage: python profiles/run_chunk_invariance.py [precision ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import vfn_testutil as U

dev = "cuda"
R = 65536
case, z = U.load_golden("full_det")
st = U.case_state(case, z)
uv, pose, K = (t.to(dev) for t in U.S.synthetic_rays(R, seed=0, start=0, stride=1))
g = torch.Generator(device=dev).manual_seed(3)
draws = tuple(torch.rand(R, n, device=dev, generator=g) for n in (case["n_coarse"], case["n_fine"], case["n_fine"]))
fields = ("coarse_rgb_values", "coarse_depth_map", "coarse_normals", "coarse_colors", "z_vals", "weights")
ok = True
for prec in (sys.argv[1:] or ["fp16f8", "bf16x3", "bf16"]):
    m = U.make_model(case, st, dev, precision=prec)
    with torch.no_grad():
        a = m.render(pose, uv, K, 0, draws=draws)
        a = {f: getattr(a, f).clone() for f in fields}
        b = m.render(pose, uv, K, 0, draws=draws)
        parts = [m.render(pose[i:i + 1024], uv[i:i + 1024], K[i:i + 1024], 0, draws=tuple(d[i:i + 1024] for d in draws))
                 for i in range(0, R, 1024)]
        parts = [{f: getattr(o, f).clone() for f in fields} for o in parts]
    torch.cuda.synchronize()
    for f in fields:
        same_run = torch.equal(a[f], getattr(b, f))
        cat = torch.cat([p[f] for p in parts])
        same_chunk = torch.equal(a[f].reshape(cat.shape), cat)
        ok &= same_run and same_chunk
        print(f"{prec:7s} {f:18s} run-to-run identical: {same_run}  65536-ray call == 64 x 1024-ray calls: {same_chunk}")
    assert (a["weights"] > 0).float().mean().item() > 0.01
print("chunk invariance:", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
