"""Host-side mirrors of the reference's ray samplers (models/samplers/ray_sampler.py).

They hold the knobs callers mutate between render() calls -- ``near``, ``far``, ``N_samples``,
``max_samples`` (train/vector_field_nerf_train.py:43-45,129-131,146-147; evaluation/evaluate.py:37-40)
-- and draw the uniform numbers exactly where the reference does: on the global CPU generator, in the
order U1 -> U2 -> U3 (U3 even when deterministic, ray_sampler.py:297), so that seeding
``torch.manual_seed`` reproduces the reference's sample positions bit for bit.  The arithmetic itself
runs in the sampler kernels (csrc/geometry_sampler.cu)."""
from __future__ import annotations

from typing import Optional

import torch


class RaySampler:
    def __init__(self, near: float, far: float, N_samples: int) -> None:
        self.near, self.far, self._N_samples = near, far, N_samples

    @property
    def N_samples(self) -> int:
        return self._N_samples

    @N_samples.setter
    def N_samples(self, n: int) -> None:
        self._N_samples = n

    def active_sampler(self) -> bool:
        return self.N_samples > 0


class UniformSampler(RaySampler):
    """ray_sampler.py:95-142."""

    def __init__(self, N_samples: int, near: float, far: float, deterministic: bool = False) -> None:
        super().__init__(near, far, N_samples)
        self.deterministic = deterministic

    def t_vals(self) -> torch.Tensor:
        # made on the host like the reference (:129): CPU linspace is not reproducible by a scalar formula
        return torch.linspace(0., 1., steps=self.N_samples)

    def draw(self, n_rays: int) -> Optional[torch.Tensor]:
        return None if self.deterministic else torch.rand([n_rays, self.N_samples])


class RangeFineSampler(RaySampler):
    """ray_sampler.py:240-302."""

    def __init__(self, N_samples: int, near: float, far: float, deterministic: bool = False,
                 range: float = 0.5, max_samples: int = 100, pytest: bool = False) -> None:
        super().__init__(near, far, N_samples)
        self.deterministic, self.pytest, self.range, self.max_samples = deterministic, pytest, range, max_samples

    def n_fine(self) -> int:
        return min(self.max_samples, self.N_samples)

    def draw(self, n_rays: int):
        nf = self.n_fine()
        U2 = None if self.deterministic else torch.rand([n_rays, nf])
        U3 = torch.rand((n_rays, nf))
        return U2, U3
