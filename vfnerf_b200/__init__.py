"""vfnerf_b200 -- B200-native volume-rendering hot path of VF-NeRF (render(), its backward, and the
VF-only grid query) behind the reference's Python API.  See DESIGN.md / INTEGRATION.md.

Importing the package never needs a GPU; calling the hot path does, and raises if the CUDA
library is missing (there is no CPU fallback)."""
from .config import (CudaConfig, DensityConfig, RaySamplerConfig, RenderingNetConfig, SchedulerConfig,  # noqa: F401
                     VFNerfConfig, VFNetConfig)
from .nerf import VectorFieldNerf                                                                     # noqa: F401
from .networks import LaplaceDensity, RenderingNetwork, VectorFieldNetwork                             # noqa: F401
from .output import NerfOutput                                                                        # noqa: F401
from .samplers import RangeFineSampler, UniformSampler                                                 # noqa: F401

__version__ = "0.1.0"
