"""vfnerf_b200 -- B200-native volume-rendering hot path of VF-NeRF (render(), its backward, and
the VF-only grid query) behind the reference's Python API.  See DESIGN.md."""
__version__ = "0.1.0"
