"""Driver used under ncu: the HBM-bound kernels of the path (coarse / fine sampler, density + transmittance scan,
compositor) on 262 144 rays x (64 + 64) samples through their stage entry points -- the same calls bench.py's
`hbm_kernels` leg times."""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import vfn_testutil as U
from vfnerf_b200 import _lib

L = _lib.lib()
dev = "cuda"
case, z = U.load_golden("full_det")
model = U.make_model(case, U.case_state(case, z), dev, precision="bf16")
R, Nc, Nf = 262144, 64, 64
N = Nc + Nf
sp = torch.cuda.current_stream().cuda_stream
f = lambda *sh: torch.empty(*sh, device=dev)
dirs = torch.nn.functional.normalize(torch.randn(R, 3, device=dev), dim=1)
cam = torch.randn(R, 3, device=dev)
tv = torch.linspace(0., 1., Nc).to(dev)
U3 = torch.rand(R, Nf, device=dev)
z_c, pts_c, w_c = f(R, Nc), f(R, Nc, 3), torch.rand(R, Nc, device=dev)
z_m, pts_m = f(R, N), f(R, N, 3)
nrm = torch.tanh(torch.randn(R, N, 3, device=dev))
col = torch.rand(R, N, 3, device=dev)
wts, rgb, dep = f(R, N), f(R, 3), f(R, 1)
cfg = model._render_cfg(R, False)
dpar = model.density.flat()
near, far, fr = float(model.ray_sampler.near), float(model.ray_sampler.far), float(model.fine_sampler.range)
for _ in range(3):
    assert L.vfnerf_coarse_sample(R, Nc, near, far, 0, tv.data_ptr(), None, dirs.data_ptr(), cam.data_ptr(), z_c.data_ptr(),
                                  pts_c.data_ptr(), sp) == 0
    assert L.vfnerf_fine_sample(R, Nc, Nf, near, far, fr, 0, z_c.data_ptr(), w_c.data_ptr(), None, U3.data_ptr(),
                                dirs.data_ptr(), cam.data_ptr(), z_m.data_ptr(), pts_m.data_ptr(), sp) == 0
    assert L.vfnerf_density_weights(C.byref(cfg), N, dpar.data_ptr(), nrm.data_ptr(), 3, dirs.data_ptr(), z_m.data_ptr(), None,
                                    None, wts.data_ptr(), sp) == 0
    assert L.vfnerf_composite(R, N, wts.data_ptr(), col.data_ptr(), z_m.data_ptr(), rgb.data_ptr(), dep.data_ptr(), sp) == 0
torch.cuda.synchronize()
print("ok")
