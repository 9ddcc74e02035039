"""Device-side mirrors of the stand-alone weight functions of utils/rendering.py (same names and arguments).

``volsdf_volume_rendering`` is what render() fuses behind the density (csrc/density_composite.cu); ``nerf_volume_rendering``
is the alternative the reference's render() calls with swapped arguments -- offered here
with the argument order of its own signature.  CUDA tensors in, CUDA tensor out, no gradient (the render() path has its
own fused backward); host tensors are rejected."""
from __future__ import annotations

import torch

from . import _lib, ops


def _weights(mode: int, sigma: torch.Tensor, z_vals: torch.Tensor, normalize: bool) -> torch.Tensor:
    z_vals = ops._require_cuda("z_vals", z_vals.detach())
    sigma = ops._require_cuda("sigma", sigma.detach()).reshape(z_vals.shape)
    R, N = z_vals.shape
    out = torch.empty(R, N, dtype=torch.float32, device=z_vals.device)
    _lib.check(_lib.lib().vfnerf_volume_weights(R, N, mode, int(bool(normalize)), sigma.data_ptr(), z_vals.data_ptr(),
                                                out.data_ptr(), ops._stream_ptr(z_vals.device)), "vfnerf_volume_weights")
    return out


def nerf_volume_rendering(sigma: torch.Tensor, z_vals: torch.Tensor, normalize: bool = False) -> torch.Tensor:
    """utils/rendering.py:98-119."""
    return _weights(1, sigma, z_vals, normalize)


def volsdf_volume_rendering(z_vals: torch.Tensor, density: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    """utils/rendering.py:122-148."""
    return _weights(0, density, z_vals, normalize)
