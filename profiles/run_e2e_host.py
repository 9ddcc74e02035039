"""Host-side cost of one big-chunk render() call through the e2e path (CPU-generator draws in pinned memory + upload)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vfnerf_b200 import synthetic as S, samplers
import bench
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
dev = torch.device("cuda")
st = S.synthetic_state(0, vf_gain=2.0)
m = S.make_model(bench.CASE, st, dev, precision=prec)
m.return_ray_dirs = False
R = 65536
pose1, K1 = S.synthetic_camera(seed=0)
uv = S.pixel_grid(680, 1200)[:R].contiguous().to(dev)
pose = pose1.repeat(R, 1, 1).contiguous().to(dev)
K = K1.repeat(R, 1, 1).contiguous().to(dev)
h = torch.empty(R, 64).pin_memory()
for _ in range(3):
    t = time.perf_counter(); samplers.cpu_generator_rand_(h); print(f"cpu_generator_rand_ 4.2M: {(time.perf_counter() - t) * 1e3:.2f} ms")
with torch.no_grad():
    for i in range(8):
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = m.render(pose, uv, K, 0)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"{prec} call {i}: host {1e3 * (t1 - t):.2f} ms, total {1e3 * (t2 - t):.2f} ms")
