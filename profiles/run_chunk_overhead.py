"""Host-side cost of one render() call at the reference's evaluation chunk size (1024 rays): cProfile of 200 calls with
host inputs (the e2e path of bench.py), to see where the Python time goes."""
import cProfile, pstats, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import vfn_testutil as U

dev = "cuda"
case, z = U.load_golden("full_det")
model = U.make_model(case, U.case_state(case, z), dev, precision="bf16")
model.return_ray_dirs = False
R = 1024
uv, pose, K = U.S.synthetic_rays(R, seed=0, start=0, stride=797)
uv, pose, K = uv.pin_memory(), pose.pin_memory(), K.pin_memory()
rgb_h = torch.empty(R, 3).pin_memory()


def call():
    with torch.no_grad():
        out = model.render(pose.to(dev, non_blocking=True), uv.to(dev, non_blocking=True), K.to(dev, non_blocking=True), 0)
        rgb_h.copy_(out.coarse_rgb_values, non_blocking=True)


for _ in range(20):
    call()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(200):
    call()
t_issue = time.perf_counter() - t
torch.cuda.synchronize()
t_all = time.perf_counter() - t
print(f"200 calls: host issue {t_issue / 200 * 1e6:.0f} us/call, with final sync {t_all / 200 * 1e6:.0f} us/call")
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    call()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
