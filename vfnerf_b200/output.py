"""Return type of render(): field-for-field the reference's NerfOutput (models/nerf/output.py:7-22),
so callers that read ``outputs.coarse_rgb_values`` etc. work unchanged.  ``weights`` is an additive
extra (the reference does not return the compositing weights)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch


@dataclass
class NerfOutput:
    points_coarse: torch.Tensor
    coarse_normals: torch.Tensor
    coarse_rgb_values: torch.Tensor
    coarse_depth_map: torch.Tensor
    mask: Optional[torch.Tensor] = None
    z_vals: Optional[torch.Tensor] = None
    points_fine: Optional[torch.Tensor] = None
    fine_normals: Optional[torch.Tensor] = None
    fine_rgb_values: Optional[torch.Tensor] = None
    fine_depth_map: Optional[torch.Tensor] = None
    fine_mask: Optional[torch.Tensor] = None
    directional_derivtives: Optional[torch.Tensor] = None   # (sic) the reference's spelling
    ray_dirs: Optional[torch.Tensor] = None
    coarse_colors: Optional[torch.Tensor] = None
    weights: Optional[torch.Tensor] = None

    def fine_active(self) -> bool:
        return self.fine_normals is not None

    def get_normals(self, N_rays: int, N_coarse: int, N_fine: int):
        cn = self.coarse_normals.reshape(N_rays, N_coarse, 3)
        a, b = cn[:, :-1, :].reshape(-1, 3), cn[:, 1:, :].reshape(-1, 3)
        if self.fine_active():
            fn = self.fine_normals.reshape(N_rays, N_fine, 3)
            return a, b, fn[:, :-1, :].reshape(-1, 3), fn[:, 1:, :].reshape(-1, 3)
        return a, b, None, None

    def to_dict(self) -> Dict[str, Optional[torch.Tensor]]:
        keys = ("points_coarse", "coarse_normals", "coarse_rgb_values", "coarse_depth_map", "mask",
                "points_fine", "fine_normals", "fine_rgb_values", "fine_depth_map", "fine_mask")
        return {k: getattr(self, k) for k in keys}
