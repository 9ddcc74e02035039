"""Configuration objects for the render() hot path.

Field names follow the reference's dataclasses (config_parser/vf_nerf_config.py:10-124) so that
an object built by the reference's own parser can be handed to ``VectorFieldNerf`` unchanged:
the facade only reads attributes (duck typing) and never isinstance-checks these classes.
Only the knobs that reach the hot path are represented; loss / dataset / runner configs stay
with the caller ("Config / flags").
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence

import torch


@dataclass
class DensityConfig:
    beta_bounds: Sequence[float] = (1e-4, 1e9)
    mean_bounds: Sequence[float] = (0.6, 1.0)
    scale_min: float = 1.0
    params_init: Dict[str, float] = field(default_factory=lambda: dict(beta=0.5, scale=100.0, mean=0.7))
    cutoff: float = -2.0   # accepted and ignored, exactly like the reference (density_functions.py:20-34)

    def todict(self) -> Dict[str, Any]:
        return dict(beta_bounds=self.beta_bounds, mean_bounds=self.mean_bounds,
                    scale_min=self.scale_min, params_init=self.params_init)


@dataclass
class VFNetConfig:
    input_dims: int = 3
    output_dims: int = 3
    dimensions: List[int] = field(default_factory=lambda: [256] * 8)
    feature_vector_dims: int = 256
    embedder_multires: int = 6
    weight_norm: bool = False
    batch_norm: bool = True
    skip_connection_in: Optional[List[int]] = field(default_factory=lambda: [4])
    bias_init: float = 0.0
    dropout: bool = False
    dropout_probability: float = 0.2
    xavier_init: bool = False
    init: str = ""


@dataclass
class RenderingNetConfig:
    output_dims: int = 3
    dimensions: List[int] = field(default_factory=lambda: [256] * 4)
    feature_vector_dims: int = 256
    weight_norm: bool = False
    batch_norm: bool = True
    mode: str = "idr"
    embedder_multires: int = 4
    detach_normals: bool = True


@dataclass
class RaySamplerConfig:
    n_samples: int = 64
    n_importance: int = 64
    rays_per_batch: int = 1024
    perturb: bool = False
    near: float = 0.0
    far: float = 6.0
    fine_range: float = 0.3
    increase_every: int = 50
    max_samples: int = 100

    def fine_sampling(self) -> bool:
        return self.n_importance > 0


@dataclass
class CudaConfig:
    device: torch.device = torch.device("cuda")
    num_gpus: int = 1


@dataclass
class SchedulerConfig:
    lr: float = 5e-4
    lr_decay_factor: float = 0.1
    lr_decay_steps: int = 50000
    clip_norm: float = 0.5
    weight_decay: float = 0.0


@dataclass
class VFNerfConfig:
    vf_net_config: VFNetConfig = field(default_factory=VFNetConfig)
    rendering_net_config: RenderingNetConfig = field(default_factory=RenderingNetConfig)
    ray_sampler_config: RaySamplerConfig = field(default_factory=RaySamplerConfig)
    cuda_config: CudaConfig = field(default_factory=CudaConfig)
    scheduler_config: SchedulerConfig = field(default_factory=SchedulerConfig)
    density_config: DensityConfig = field(default_factory=DensityConfig)
    cos_sim_weights: Any = field(default_factory=lambda: [0.09] * 11)
    cos_sim_weights_anneal: str = "hard"
    anneal_start: int = 700
    anneal_end: int = 1400
    rendering: str = "volsdf"
    normalize_rendering: bool = True
    dir_to_normal_th: float = -2.0
    numerical_jacobian: bool = False
    border_supervision: bool = True
    center_supervision: bool = True

    def __post_init__(self) -> None:
        # same admissible values and error type as vf_nerf_config.py:120-124
        if self.cos_sim_weights_anneal not in ("none", "hard", "soft"):
            raise ValueError(f"Invalid cos_sim_weights_anneal: {self.cos_sim_weights_anneal}")
        if self.rendering not in ("nerf", "volsdf"):
            raise ValueError(f"Invalid rendering: {self.rendering}")
        self.cos_sim_weights = torch.as_tensor(self.cos_sim_weights, dtype=torch.float32)
