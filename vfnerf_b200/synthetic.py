"""Synthetic cameras, rays, uniform draws and model weights.

There is no dataset or checkpoint in the build environment (the reference's .pth files are
git-LFS pointers), so tests, goldens and the benchmark all use these generators.  Everything is
derived from ``numpy.random.default_rng(seed)`` (PCG64, stream-stable across numpy releases) so
that the GPU box, this container and the committed goldens see identical values.

Model recipe: plain random init renders exactly zero (neighbouring VF vectors are parallel, so
the windowed-cosine density never fires).  SURVEY's first suggestion (all VF
weights x8) does give densities, but the resulting ReLU net is chaotic: fp32 and fp64 evaluations
of it already disagree by 1e-2 on the normals, so no tolerance test means anything.  The recipe
used instead keeps the net well conditioned (fp32 vs fp64: ~5e-5):

  * every VF Linear weight x ``vf_gain`` (2.0, roughly variance preserving through ReLU), VF
    biases ~ U(-1,1), BatchNorm running stats / affine terms randomised (exercises BN folding);
  * the three vector outputs are *centred*: the last layer's first three rows are rescaled and
    their biases shifted so that the pre-tanh vector has zero mean and std ``out_std`` over a
    fixed set of probe points.  The field then changes direction in space (cos < 0.5 on ~20 % of
    samples), which is what the windowed-cosine Laplace density responds to.

The centring constants are computed in float64 numpy and rounded to 1e-3 so that they are
bit-stable across machines / BLAS builds.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch


def mlp_layer_dims(in_dim: int, hidden: Sequence[int], out_dim: int, skip_in: Sequence[int] = (),
                   skip_width: int = 0) -> List[Tuple[int, int]]:
    """(fan_in, fan_out) per Linear, following the reference's construction rule
    (vector_field_network.py:46-54): the layer *before* a skip layer is narrowed by the embedding
    width so that the concatenation has the nominal width again."""
    dims = [in_dim] + list(hidden) + [out_dim]
    out = []
    for i in range(len(dims) - 1):
        fan_out = dims[i + 1] - skip_width if (i + 1) in skip_in else dims[i + 1]
        out.append((dims[i], fan_out))
    return out


def _mlp_state(rng: np.random.Generator, layer_dims, gain: float, bias_unit: bool,
               randomize_bn: bool) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = {}
    n = len(layer_dims)
    for i, (fi, fo) in enumerate(layer_dims):
        bound = 1.0 / math.sqrt(fi)
        W = rng.uniform(-bound, bound, size=(fo, fi)).astype(np.float32) * np.float32(gain)
        b = rng.uniform(-1.0, 1.0, size=(fo,)).astype(np.float32) if bias_unit else \
            rng.uniform(-bound, bound, size=(fo,)).astype(np.float32)
        last = i == n - 1
        pre = f"layers.{i}" if last else f"layers.{i}.0"
        sd[f"{pre}.weight"] = torch.from_numpy(W)
        sd[f"{pre}.bias"] = torch.from_numpy(b)
        if not last:
            if randomize_bn:
                g = rng.uniform(0.5, 1.5, size=(fo,)).astype(np.float32)
                be = (0.1 * rng.standard_normal(fo)).astype(np.float32)
                rm = (0.1 * rng.standard_normal(fo)).astype(np.float32)
                rv = rng.uniform(0.5, 1.5, size=(fo,)).astype(np.float32)
            else:
                g, be = np.ones(fo, np.float32), np.zeros(fo, np.float32)
                rm, rv = np.zeros(fo, np.float32), np.ones(fo, np.float32)
            sd[f"layers.{i}.1.weight"] = torch.from_numpy(g)
            sd[f"layers.{i}.1.bias"] = torch.from_numpy(be)
            sd[f"layers.{i}.1.running_mean"] = torch.from_numpy(rm)
            sd[f"layers.{i}.1.running_var"] = torch.from_numpy(rv)
            sd[f"layers.{i}.1.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return sd


def _embed64(x: np.ndarray, multires: int) -> np.ndarray:
    parts = [x]
    for k in range(multires):
        parts += [np.sin(x * 2.0 ** k), np.cos(x * 2.0 ** k)]
    return np.concatenate(parts, axis=-1)


def _center_vector_outputs(sd: Dict[str, torch.Tensor], multires: int, skip_in: Sequence[int],
                           out_std: float, extent: float = 4.0, n_probe: int = 8192) -> None:
    """Rescale / shift the three vector rows of the last VF layer (in place) so the pre-tanh vector
    has zero mean and std ``out_std`` over probe points drawn uniformly in [-extent, extent]^3."""
    rng = np.random.default_rng(1)
    pts = rng.uniform(-extent, extent, size=(n_probe, 3))
    emb = _embed64(pts, multires)
    n_layers = max(int(k.split(".")[1]) for k in sd) + 1
    x = emb
    for i in range(n_layers - 1):
        if i in skip_in:
            x = np.concatenate([x, emb], axis=1) / np.sqrt(2.0)
        W = sd[f"layers.{i}.0.weight"].double().numpy()
        b = sd[f"layers.{i}.0.bias"].double().numpy()
        g = sd[f"layers.{i}.1.weight"].double().numpy()
        be = sd[f"layers.{i}.1.bias"].double().numpy()
        rm = sd[f"layers.{i}.1.running_mean"].double().numpy()
        rv = sd[f"layers.{i}.1.running_var"].double().numpy()
        x = np.maximum(((x @ W.T + b) - rm) / np.sqrt(rv + 1e-5) * g + be, 0.0)
    last = n_layers - 1
    if last in skip_in:
        x = np.concatenate([x, emb], axis=1) / np.sqrt(2.0)
    W8 = sd[f"layers.{last}.weight"][:3].double().numpy()
    pre = x @ W8.T
    scale = np.round(out_std / pre.std(axis=0), 3)
    shift = np.round(-pre.mean(axis=0) * scale, 3)
    sd[f"layers.{last}.weight"][:3] = torch.from_numpy((W8 * scale[:, None]).astype(np.float32))
    sd[f"layers.{last}.bias"][:3] = torch.from_numpy(shift.astype(np.float32))


def synthetic_state(seed: int = 0, vf_hidden: Sequence[int] = (256,) * 8, feat: int = 256,
                    rn_hidden: Sequence[int] = (256,) * 4, multires: int = 6, multires_view: int = 4,
                    skip_in: Sequence[int] = (4,), vf_gain: float = 2.0, randomize_bn: bool = True,
                    center_output: bool = True, out_std: float = 1.5
                    ) -> Dict[str, Dict[str, torch.Tensor]]:
    """State dicts with the reference's checkpoint keys (vector_field_nerf.py:195-210):
    'vf_net', 'rendering_net', 'density'."""
    rng = np.random.default_rng(seed)
    emb = 3 + 6 * multires
    vf_dims = mlp_layer_dims(emb, vf_hidden, 3 + feat, skip_in, emb)
    rn_in = 3 + (3 + 6 * multires_view) + 3 + feat
    rn_dims = mlp_layer_dims(rn_in, rn_hidden, 3)
    vf = _mlp_state(rng, vf_dims, vf_gain, True, randomize_bn)
    if center_output:
        _center_vector_outputs(vf, multires, skip_in, out_std)
    return {
        "vf_net": vf,
        "rendering_net": _mlp_state(rng, rn_dims, 1.0, False, randomize_bn),
        "density": {"beta": torch.tensor(0.5), "scale": torch.tensor(100.0), "mean": torch.tensor(0.7)},
    }


def synthetic_camera(seed: int = 0, height: int = 680, width: int = 1200, focal: float = 600.0
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """One pinhole camera: 4x4 camera-to-world pose (unit-quaternion rotation, N(0,1) translation)
    and a 4x4 intrinsics matrix with fx=fy=focal, principal point at the image centre."""
    rng = np.random.default_rng(seed + 7919)
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    r, i, j, k = q
    Rm = np.array([[1 - 2 * (j * j + k * k), 2 * (i * j - k * r), 2 * (i * k + j * r)],
                   [2 * (i * j + k * r), 1 - 2 * (i * i + k * k), 2 * (j * k - i * r)],
                   [2 * (i * k - j * r), 2 * (j * k + i * r), 1 - 2 * (i * i + j * j)]])
    pose = np.eye(4)
    pose[:3, :3] = Rm
    pose[:3, 3] = rng.standard_normal(3)
    K = np.eye(4)
    K[0, 0] = K[1, 1] = focal
    K[0, 2] = (width - 1) / 2.0
    K[1, 2] = (height - 1) / 2.0
    return torch.from_numpy(pose.astype(np.float32)), torch.from_numpy(K.astype(np.float32))


def pixel_grid(height: int, width: int) -> torch.Tensor:
    """All pixels in row-major order as float (u=x, v=y), the layout the reference's datasets
    emit with all_pixels=True (replica_dataset.py:153-155)."""
    v, u = torch.meshgrid(torch.arange(height, dtype=torch.float32),
                          torch.arange(width, dtype=torch.float32), indexing="ij")
    return torch.stack([u.reshape(-1), v.reshape(-1)], dim=-1)


def synthetic_rays(n_rays: int, seed: int = 0, height: int = 680, width: int = 1200,
                   focal: float = 600.0, start: int = 0, stride: int = 1
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """uv[R,2], pose[R,4,4], K[R,4,4] for pixels start, start+stride, ... of one synthetic image
    (per-ray repeated pose / intrinsics, as the datasets emit them, replica_dataset.py:160-161)."""
    pose, K = synthetic_camera(seed, height, width, focal)
    idx = (start + stride * torch.arange(n_rays)) % (height * width)
    uv = torch.stack([(idx % width).float(), (idx // width).float()], dim=-1)
    return uv, pose.repeat(n_rays, 1, 1).contiguous(), K.repeat(n_rays, 1, 1).contiguous()


def synthetic_draws(n_rays: int, n_coarse: int, n_fine: int, seed: int = 1234
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """U1 [R,Nc], U2 [R,Nf], U3 [R,Nf] uniform [0,1) fp32 (the three torch.rand calls of
    ray_sampler.py:138,292,297), generated with a numpy stream so they are reproducible."""
    rng = np.random.default_rng(seed)
    U1 = torch.from_numpy(rng.random((n_rays, n_coarse), dtype=np.float32))
    U2 = torch.from_numpy(rng.random((n_rays, n_fine), dtype=np.float32))
    U3 = torch.from_numpy(rng.random((n_rays, n_fine), dtype=np.float32))
    return U1, U2, U3


def synthetic_vector_grid(N: int, seed: int = 0, scale: float = 1.0) -> torch.Tensor:
    """[N^3, 3] vector field on the marching-cubes grid of evaluation/methods.py:194-208 (x index slowest): vectors
    pointing at the nearer of a sphere (radius 0.6) and a plane (z = -0.8), tanh-squashed like the VF net's output,
    plus seeded noise -- it flips direction across both surfaces, so the divergence test fires on two sheets of cells."""
    g = torch.Generator().manual_seed(seed)
    idx = torch.arange(N ** 3)
    p = torch.stack([((idx // N) // N) % N, (idx // N) % N, idx % N], dim=1).float() * (2.0 * scale / (N - 1)) - scale
    d = p.norm(dim=1, keepdim=True)
    to_sphere = (0.6 - d) * p / d.clamp(min=1e-6)
    to_plane = torch.zeros_like(p)
    to_plane[:, 2] = -0.8 - p[:, 2]
    v = torch.where(to_sphere.norm(dim=1, keepdim=True) < to_plane.norm(dim=1, keepdim=True), to_sphere, to_plane)
    return torch.tanh(3.0 * v) + 0.02 * torch.randn(N ** 3, 3, generator=g)


def make_config(case: dict, device):
    """VFNerfConfig of a synthetic / golden case (dict with vf_hidden, feat, rn_hidden, n_coarse, n_fine, perturb, near,
    far, fine_range, max_samples, window, dir_to_normal_th) -- the shipped confs/vf_nerf.conf shapes by default."""
    from .config import (CudaConfig, DensityConfig, RaySamplerConfig, RenderingNetConfig, SchedulerConfig, VFNerfConfig,
                         VFNetConfig)
    return VFNerfConfig(
        vf_net_config=VFNetConfig(dimensions=list(case["vf_hidden"]), feature_vector_dims=case["feat"]),
        rendering_net_config=RenderingNetConfig(dimensions=list(case["rn_hidden"]), feature_vector_dims=case["feat"]),
        ray_sampler_config=RaySamplerConfig(n_samples=case["n_coarse"], n_importance=case["n_fine"],
                                            perturb=case["perturb"], near=case["near"], far=case["far"],
                                            fine_range=case["fine_range"], max_samples=case["max_samples"]),
        cuda_config=CudaConfig(device=torch.device(device), num_gpus=1),
        scheduler_config=SchedulerConfig(), density_config=DensityConfig(),
        cos_sim_weights=[0.09] * case["window"], dir_to_normal_th=case["dir_to_normal_th"])


def make_model(case: dict, state: dict, device, precision: str = "fp32"):
    """A vfnerf_b200.VectorFieldNerf for `case` with the weights of `state` ({"vf_net", "rendering_net", "density"}), in
    eval mode."""
    from .nerf import VectorFieldNerf
    model = VectorFieldNerf(make_config(case, device), precision=precision)
    model.vector_field_network.load_state_dict(state["vf_net"])
    model.rendering_network.load_state_dict(state["rendering_net"])
    model.density.load_state_dict(state["density"])
    model.eval()
    return model
