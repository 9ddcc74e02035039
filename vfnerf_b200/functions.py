"""Host-side mirror of the supervision helpers of models/helpers/functions.py:75-157 -- same names, arguments and
return values as the reference trainer uses them (train/vector_field_nerf_train.py:180-216) -- running on the device
(csrc/supervision.cu).

The sphere samplers consume numpy's global generator in the reference's order (phi, cos_theta, u: three
np.random.uniform calls of `num_samples`, models/samplers/sampler.py:169-176), so np.random.seed reproduces the
reference's points; ``draws=`` injects them, ``on_device=True`` makes them with the device generator instead (no numpy, no
upload -- same distribution, a different stream)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib


def _c3(centroid) -> "C.Array":
    c = [float(x) for x in (centroid.detach().cpu().reshape(-1).tolist() if torch.is_tensor(centroid) else centroid)]
    if len(c) != 3:
        raise ValueError("centroid must have 3 components")
    return (C.c_float * 3)(*c)


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _sphere(r_max: float, r_min: float, num_samples: int, centroid, device, inward: bool, draws, on_device: bool):
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("vfnerf_b200.functions needs a CUDA device: there is no CPU fallback")
    n = int(num_samples)
    if draws is not None:
        phi, cos_t, u = (torch.as_tensor(d, dtype=torch.float64).to(dev).contiguous() for d in draws)
    elif on_device:
        phi = torch.rand(n, dtype=torch.float64, device=dev) * (2.0 * np.pi)
        cos_t = torch.rand(n, dtype=torch.float64, device=dev) * 2.0 - 1.0
        u = torch.rand(n, dtype=torch.float64, device=dev)
    else:
        # the reference's stream: numpy's global generator, three calls in this order (sampler.py:169-176)
        host = np.empty((3, n), dtype=np.float64)
        host[0] = np.random.uniform(0.0, 2.0 * np.pi, n)
        host[1] = np.random.uniform(-1.0, 1.0, n)
        host[2] = np.random.uniform(0.0, 1.0, n)
        d = torch.from_numpy(host).to(dev)
        phi, cos_t, u = d[0], d[1], d[2]
    if phi.numel() != n or cos_t.numel() != n or u.numel() != n:
        raise ValueError("draws must hold num_samples values each")
    points = torch.empty(n, 3, dtype=torch.float32, device=dev)
    gt = torch.empty(n, 3, dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().vfnerf_sphere_points(n, phi.data_ptr(), cos_t.data_ptr(), u.data_ptr(), float(r_max), float(r_min),
                                               _c3(centroid), 1 if inward else 0, points.data_ptr(), gt.data_ptr(),
                                               _stream(dev)), "vfnerf_sphere_points")
    return points, gt


def sample_border_points(r_min: float, r_max: float, num_samples: int, centroid: torch.Tensor,
                         device: torch.device = "cuda", *, draws=None, on_device: bool = False
                         ) -> Tuple[torch.Tensor, torch.Tensor]:
    """functions.py:99-114: points in the shell r_min..r_max around the centroid, targets pointing at the centroid."""
    return _sphere(r_max, r_min, num_samples, centroid, device, True, draws, on_device)


def sample_center_points(centroid: torch.Tensor, radius: float, num_samples: int, device: torch.device = "cuda", *,
                         draws=None, on_device: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """functions.py:116-130: points in the ball of `radius` around the centroid, targets pointing away from it."""
    return _sphere(radius, 0.0, num_samples, centroid, device, False, draws, on_device)


def _select(points: torch.Tensor, normals: torch.Tensor, centroid, threshold: float, mode: int):
    if not points.is_cuda:
        raise RuntimeError("vfnerf_b200.functions needs CUDA tensors: there is no CPU fallback")
    dev = points.device
    pts = points.detach().reshape(-1, 3).float().contiguous()
    n = pts.shape[0]
    flag = torch.empty(n, dtype=torch.uint8, device=dev)
    gt = torch.empty(n, 3, dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().vfnerf_select_supervised(n, pts.data_ptr(), _c3(centroid), float(np.float32(threshold)), mode,
                                                   flag.data_ptr(), gt.data_ptr(), _stream(dev)), "vfnerf_select_supervised")
    idx = flag.nonzero().squeeze(1)               # the one synchronisation the reference's boolean indexing has too
    return normals.reshape(-1, 3).index_select(0, idx), gt.index_select(0, idx)


def get_border_indices_and_gt(points: torch.Tensor, normals: torch.Tensor, far: float, radius: float,
                              centroid: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """functions.py:75-97: ray samples further than far/2 - radius from the centroid (their predicted vectors keep their
    autograd history) and normalize(centroid - p) as targets."""
    return _select(points, normals, centroid, far / 2 - radius, 0)


def get_center_indices_and_gt(points: torch.Tensor, normals: torch.Tensor, centroid: torch.Tensor,
                              radius: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """functions.py:132-154: ray samples closer than `radius` to the centroid and normalize(p - centroid) as targets."""
    return _select(points, normals, centroid, radius, 1)
