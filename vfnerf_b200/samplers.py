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


# ---- the reference's CPU-generator stream, faster -------------------------------------------------------------
_FAST_RNG = {"ok": None}


def _mt_fields(raw_np):
    """Views of the engine fields inside torch.get_rng_state() (CPUGeneratorImplStateLegacy: seed u64, left i32,
    seeded i32, next u64, state u64[624], ...)."""
    import numpy as np
    return raw_np[8:12].view(np.int32), raw_np[16:24].view(np.uint64), raw_np[24:24 + 624 * 8].view(np.uint64)


def _fast_rand_raw(host: torch.Tensor) -> None:
    import numpy as np
    from . import _lib
    raw = torch.get_rng_state()
    a = raw.numpy()
    left_v, next_v, state_v = _mt_fields(a)
    state = state_v.astype(np.uint32)
    left = np.array([left_v[0]], dtype=np.int32)
    nxt = np.array([next_v[0]], dtype=np.uint32)
    _lib.check(_lib.lib().vfnerf_mt19937_uniform(state.ctypes.data, left.ctypes.data, nxt.ctypes.data, host.numel(),
                                                 host.data_ptr()), "vfnerf_mt19937_uniform")
    state_v[:] = state
    left_v[0] = left[0]
    next_v[0] = nxt[0]
    torch.set_rng_state(raw)


def _fast_rng_self_check() -> bool:
    """The fast path is used only if, from a copy of the current generator state, it reproduces torch.rand bit for bit
    (values and final state) across a block boundary; the caller's generator state is left untouched."""
    saved = torch.get_rng_state()
    try:
        if saved.numel() != 5056 or saved.dtype != torch.uint8:
            return False
        n = 3 * 624 + 17
        want = torch.rand(n)
        after_want = torch.get_rng_state()
        torch.set_rng_state(saved)
        got = torch.empty(n)
        _fast_rand_raw(got)
        return bool(torch.equal(got, want) and torch.equal(torch.get_rng_state(), after_want))
    except Exception:
        return False
    finally:
        torch.set_rng_state(saved)


def cpu_generator_rand_(host: torch.Tensor) -> torch.Tensor:
    """``torch.rand(host.shape, out=host)`` on the global CPU generator -- same numbers, same generator state afterwards --
    through the vectorised MT19937 of csrc/host_rng.cu when ``host`` is a large contiguous float32 CPU tensor."""
    if host.numel() >= (1 << 16) and host.dtype == torch.float32 and host.is_contiguous() and not host.is_cuda:
        if _FAST_RNG["ok"] is None:
            _FAST_RNG["ok"] = _fast_rng_self_check()
        if _FAST_RNG["ok"]:
            _fast_rand_raw(host)
            return host
    return torch.rand(host.shape, out=host)


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


class FineSampler(RaySampler):
    """The inverse-CDF importance sampler, ray_sampler.py:145-237 (same constructor and methods).  The reference's
    render() never calls it; it is the alternative fine sampler.  The uniform numbers are
    made where the reference makes them -- ``linspace`` on the host when deterministic (:180), the global CPU
    generator otherwise (:183) -- and everything else (pdf, warp prefix-scan cdf, binary search, lerp, merge sort)
    is one kernel launch, ``pdf_sample_kernel`` in csrc/geometry_sampler.cu."""

    def __init__(self, N_samples: int, deterministic: bool = False, pytest: bool = False) -> None:
        super().__init__(near=-1.0, far=-1.0, N_samples=N_samples)
        self.deterministic, self.pytest = deterministic, pytest

    def _u(self, n_rays: int) -> torch.Tensor:
        if self.pytest:
            # :186-194 -- numpy's generator, seeded per call (float64 there; the kernel computes in fp32)
            import numpy as np
            np.random.seed(0)
            if self.deterministic:
                return torch.tensor(np.linspace(0., 1., self.N_samples)).float()
            return torch.tensor(np.random.rand(n_rays, self.N_samples)).float()
        if self.deterministic:
            return torch.linspace(0., 1., steps=self.N_samples)
        return torch.rand([n_rays, self.N_samples])

    def sample_pdf(self, bins: torch.Tensor, weights: torch.Tensor, u: Optional[torch.Tensor] = None) -> torch.Tensor:
        """bins [R,B], weights [R,B-1] on the GPU -> samples [R,N_samples] (detached, like :214)."""
        from . import _lib, ops
        bins, weights = ops._require_cuda("bins", bins.detach()), ops._require_cuda("weights", weights.detach())
        R, B = bins.shape
        if weights.shape != (R, B - 1):
            raise ValueError(f"weights must be [R, B-1] = {(R, B - 1)}, got {tuple(weights.shape)}")
        u = (self._u(R) if u is None else u).to(bins.device).float().contiguous()
        out = torch.empty(R, u.shape[-1], dtype=torch.float32, device=bins.device)
        _lib.check(_lib.lib().vfnerf_sample_pdf(R, B, u.shape[-1], bins.data_ptr(), weights.data_ptr(), u.data_ptr(),
                                                int(u.dim() == 2), out.data_ptr(), ops._stream_ptr(bins.device)),
                   "vfnerf_sample_pdf")
        return out

    def get_z_vals(self, ray_dirs=None, cam_loc=None, device=None, coarse_z_vals: torch.Tensor = None,
                   coarse_weights: torch.Tensor = None, u: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Merged, sorted z values [R, Nc + N_samples] (:217-237).  ``ray_dirs``/``cam_loc``/``device`` are accepted
        and unused, like in the reference."""
        return self._merged(coarse_z_vals, coarse_weights, u, None, None)[0]

    def sample(self, directions: torch.Tensor, cam_loc: torch.Tensor, coarse_z_vals: torch.Tensor,
               coarse_weights: torch.Tensor, u: Optional[torch.Tensor] = None):
        """get_z_vals + RaySampler.sample (:49-80) in the same launch -> (z [R,N], points [R,N,3])."""
        return self._merged(coarse_z_vals, coarse_weights, u, directions, cam_loc)

    def _merged(self, z_c, w_c, u, directions, cam_loc):
        from . import _lib, ops
        z_c, w_c = ops._require_cuda("coarse_z_vals", z_c.detach()), ops._require_cuda("coarse_weights", w_c.detach())
        R, Nc = z_c.shape
        if w_c.shape != (R, Nc):
            raise ValueError(f"coarse_weights must be {(R, Nc)}, got {tuple(w_c.shape)}")
        dev = z_c.device
        u = (self._u(R) if u is None else u).to(dev).float().contiguous()
        Nf = u.shape[-1]
        z = torch.empty(R, Nc + Nf, dtype=torch.float32, device=dev)
        pts = None
        if directions is not None:
            directions, cam_loc = ops._require_cuda("directions", directions), ops._require_cuda("cam_loc", cam_loc)
            pts = torch.empty(R, Nc + Nf, 3, dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().vfnerf_pdf_fine_sample(R, Nc, Nf, z_c.data_ptr(), w_c.data_ptr(), u.data_ptr(),
                                                     int(u.dim() == 2), _lib.ptr(directions), _lib.ptr(cam_loc),
                                                     z.data_ptr(), _lib.ptr(pts), ops._stream_ptr(dev)),
                   "vfnerf_pdf_fine_sample")
        return z, pts
