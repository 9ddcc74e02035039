"""CPU oracle for the VF-NeRF render() hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  Nothing under ``vfnerf_b200/`` imports it, and the product
path raises when the CUDA library is missing instead of falling back to this code.

It is a clean-room restatement, in plain fp32 torch-CPU tensor arithmetic, of what the
reference (albertgassol1/vf-nerf, mounted at /root/reference while this was written)
computes on the path ``VectorFieldNerf.render()``:

    models/nerf/vector_field_nerf.py:216-338   render()
    models/nerf/vector_field_nerf.py:442-474   get_density()
    utils/rendering.py:12-60                   ray directions / camera location
    utils/pinhole_model.py:9-63                quat_to_rot / pixel2camera
    models/samplers/ray_sampler.py:49-80       RaySampler.sample
    models/samplers/ray_sampler.py:113-142     UniformSampler.get_z_vals
    models/samplers/ray_sampler.py:264-302     RangeFineSampler.get_z_vals
    models/samplers/ray_sampler.py:163-237     FineSampler.sample_pdf / get_z_vals (inverse-CDF sampler; not called
                                               by render() upstream)
    models/helpers/embedder.py:6-52            positional encoding
    models/vector_field/vector_field_network.py:177-208   VF MLP (eval mode)
    models/vector_field/rendering_network.py:62-108       colour MLP
    models/helpers/functions.py:41-72          window_cosine_similarity
    models/helpers/density_functions.py:20-34,129-204     LaplaceDensity (cutoff dropped)
    utils/rendering.py:122-148                 volsdf_volume_rendering

Parity pin: the reference ships no tests or golden vectors, so the oracle
is pinned against the *live reference* imported from /root/reference by
``tests/golden/make_golden.py`` (bit-exact on z_vals / points / directions, <= 1e-4 on
everything else) and the resulting vectors are committed under ``tests/golden/``.

Uniform draws are explicit arguments (U1: coarse stratification, U2: fine stratification,
U3: the "z_add" fallback samples) because the reference draws them from the global CPU
generator (ray_sampler.py:138,292,297).  Arithmetic order in the sampler functions is
kept identical to the reference's expression trees so that sample positions are
bit-exact; everything is written with out-of-place ops so torch autograd can provide the
gradient oracle for the backward kernels.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # torch.nn.BatchNorm1d default, vector_field_network.py:60 / rendering_network.py:52


# --------------------------------------------------------------------------------------
# rays
# --------------------------------------------------------------------------------------
def quat_pose_to_matrix(pose7: torch.Tensor) -> torch.Tensor:
    """[R,7] (qr,qi,qj,qk,tx,ty,tz) -> [R,4,4].  pinhole_model.py:9-33 + rendering.py:27-33."""
    q = F.normalize(pose7[:, :4], dim=1)
    qr, qi, qj, qk = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    p = torch.eye(4, dtype=pose7.dtype).repeat(pose7.shape[0], 1, 1)
    p[:, 0, 0] = 1 - 2 * (qj ** 2 + qk ** 2)
    p[:, 0, 1] = 2 * (qj * qi - qk * qr)
    p[:, 0, 2] = 2 * (qi * qk + qr * qj)
    p[:, 1, 0] = 2 * (qj * qi + qk * qr)
    p[:, 1, 1] = 1 - 2 * (qi ** 2 + qk ** 2)
    p[:, 1, 2] = 2 * (qj * qk - qi * qr)
    p[:, 2, 0] = 2 * (qk * qi - qj * qr)
    p[:, 2, 1] = 2 * (qj * qk + qi * qr)
    p[:, 2, 2] = 1 - 2 * (qi ** 2 + qj ** 2)
    p[:, :3, 3] = pose7[:, 4:]
    return p


def ray_geometry(uv: torch.Tensor, pose: torch.Tensor, intrinsics: torch.Tensor):
    """uv[R,2], pose[R,4,4] (or [R,7]), K[R,4,4] -> directions, ray_dirs, cam_loc (all [R,3]).

    rendering.py:12-60.  The batched 4x4 @ 4x1 product is written as the strict
    left-to-right, non-fused chain ((p0*x + p1*y) + p2*z) + p3*1, which is what torch's CPU
    bmm produces bit for bit (re-checked by make_golden.py).
    """
    if pose.shape[1] == 7:
        pose = quat_pose_to_matrix(pose)
    cam_loc = pose[:, :3, 3]
    fx, fy = intrinsics[:, 0, 0], intrinsics[:, 1, 1]
    cx, cy = intrinsics[:, 0, 2], intrinsics[:, 1, 2]
    skew = intrinsics[:, 0, 1]
    u, v = uv[:, 0], uv[:, 1]
    # depth 1 with the sign of K[0,1,1] of the FIRST ray (rendering.py:42)
    z = torch.ones_like(u) * torch.sign(intrinsics[0, 1, 1])
    x = (u - cx + cy * skew / fy - skew * v / fy) / fx * z.abs()   # pinhole_model.py:58-59
    y = (v - cy) / fy * z.abs()                                   # pinhole_model.py:60
    one = torch.ones_like(z)
    world = []
    for r in range(3):
        acc = pose[:, r, 0] * x
        acc = acc + pose[:, r, 1] * y
        acc = acc + pose[:, r, 2] * z
        acc = acc + pose[:, r, 3] * one
        world.append(acc)
    world = torch.stack(world, dim=-1)
    directions = world - cam_loc
    norm = torch.sqrt((directions * directions).sum(dim=1, keepdim=True))
    ray_dirs = directions / norm.clamp_min(1e-12)                 # F.normalize, rendering.py:58
    return directions, ray_dirs, cam_loc


# --------------------------------------------------------------------------------------
# samplers (bit-exact contract)
# --------------------------------------------------------------------------------------
def _stratify(z: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """ray_sampler.py:134-140 / :285-293."""
    mids = .5 * (z[..., 1:] + z[..., :-1])
    upper = torch.cat([mids, z[..., -1:]], -1)
    lower = torch.cat([z[..., :1], mids], -1)
    return lower + (upper - lower) * u


def coarse_z_vals(n_rays: int, near: float, far: float, t_vals: torch.Tensor,
                  perturb: bool, U1: Optional[torch.Tensor]) -> torch.Tensor:
    """ray_sampler.py:113-142.  t_vals = torch.linspace(0,1,Nc) made on the host by the caller."""
    near_c = near * torch.ones(n_rays, 1)
    far_c = far * torch.ones(n_rays, 1)
    z = near_c * (1. - t_vals) + far_c * t_vals
    if perturb:
        z = _stratify(z, U1)
    return z


def sample_points(cam_loc: torch.Tensor, z: torch.Tensor, directions: torch.Tensor) -> torch.Tensor:
    """ray_sampler.py:76-78: separate multiply then add (no FMA)."""
    return cam_loc.unsqueeze(1) + z.unsqueeze(2) * directions.unsqueeze(1)


def fine_z_vals(z_coarse: torch.Tensor, w_coarse: torch.Tensor, near: float, far: float,
                fine_range: float, n_fine: int, perturb: bool,
                U2: Optional[torch.Tensor], U3: torch.Tensor) -> torch.Tensor:
    """RangeFineSampler.get_z_vals, ray_sampler.py:264-302.  n_fine = min(max_samples, N_samples)."""
    R = z_coarse.shape[0]
    m = torch.argmax(w_coarse, dim=-1)                      # first index on ties
    z_star = z_coarse[torch.arange(R), m]
    ramp = 2 * fine_range / (n_fine - 1) * torch.arange(n_fine)
    z_f = z_star[:, None] - fine_range + ramp
    if perturb:
        z_f = _stratify(z_f, U2)
    z_add = U3 * (far - near) + near
    out = torch.sort(torch.cat([z_coarse, z_add], dim=-1), dim=-1)[0]
    alt = torch.sort(torch.cat([z_coarse, z_f], dim=-1), dim=-1)[0]
    return torch.where((m > 0)[:, None], alt, out)


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """FineSampler.sample_pdf, ray_sampler.py:163-215: inverse-CDF sampling of the piecewise-constant pdf that
    ``weights[R,B-1]`` define over ``bins[R,B]``.  ``u`` is the reference's uniform tensor made explicit: its
    ``linspace(0, 1, N_samples)`` when deterministic (:180-181), its ``torch.rand`` draw otherwise (:183);
    [n] or [R,n]."""
    weights = weights + 1e-5                                                # :174
    pdf = weights / torch.sum(weights, -1, keepdim=True)                    # :175
    cdf = torch.cumsum(pdf, -1)                                             # :176
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)              # :177
    u = u.expand(list(cdf.shape[:-1]) + [u.shape[-1]]).contiguous()         # :181, :197
    inds = torch.searchsorted(cdf, u, right=True)                           # :198
    below = torch.clamp(inds - 1, min=0)                                    # :199
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)                        # :200
    cdf_lo, cdf_hi = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)     # :205-207 (gather per ray)
    bin_lo, bin_hi = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_hi - cdf_lo                                                 # :209
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)        # :210
    t = (u - cdf_lo) / denom                                                # :211
    return bin_lo + t * (bin_hi - bin_lo)                                   # :212


def pdf_fine_z_vals(z_coarse: torch.Tensor, w_coarse: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """FineSampler.get_z_vals, ray_sampler.py:217-237: bins are the midpoints of the coarse samples, the weights drop
    their first and last entry, and the new samples are merged (sorted) with the coarse ones."""
    mid = .5 * (z_coarse[..., 1:] + z_coarse[..., :-1])                     # :231
    z_new = sample_pdf(mid, w_coarse[..., 1:-1], u)                         # :233
    return torch.sort(torch.cat([z_coarse, z_new], dim=-1), dim=-1)[0]      # :236


# --------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------
def embed(x: torch.Tensor, multires: int) -> torch.Tensor:
    """embedder.py:11-37: [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]."""
    if multires <= 0:
        return x
    parts = [x]
    for k in range(multires):
        f = float(2 ** k)
        parts.append(torch.sin(x * f))
        parts.append(torch.cos(x * f))
    return torch.cat(parts, dim=-1)


def _linear_bn(sd: Dict[str, torch.Tensor], i: int, x: torch.Tensor, has_bn: bool) -> torch.Tensor:
    if has_bn:
        W, b = sd[f"layers.{i}.0.weight"], sd[f"layers.{i}.0.bias"]
        g, be = sd[f"layers.{i}.1.weight"], sd[f"layers.{i}.1.bias"]
        rm, rv = sd[f"layers.{i}.1.running_mean"], sd[f"layers.{i}.1.running_var"]
        y = F.linear(x, W, b)
        return (y - rm) / torch.sqrt(rv + BN_EPS) * g + be       # BatchNorm1d in eval mode
    return F.linear(x, sd[f"layers.{i}.weight"], sd[f"layers.{i}.bias"])


def _num_layers(sd: Dict[str, torch.Tensor]) -> int:
    idx = {int(k.split(".")[1]) for k in sd if k.startswith("layers.")}
    return max(idx) + 1


def vf_network(sd: Dict[str, torch.Tensor], points: torch.Tensor, multires: int = 6,
               skip_in=(4,)) -> torch.Tensor:
    """VectorFieldNetwork._forward (eval), vector_field_network.py:177-208 -> [P, 3+feat]."""
    n_layers = _num_layers(sd)
    emb = embed(points, multires)
    x = emb
    for i in range(n_layers):
        if i in skip_in:
            x = torch.cat([x, emb], 1) / torch.sqrt(torch.tensor([2.0]))
        last = i == n_layers - 1
        x = _linear_bn(sd, i, x, has_bn=not last)
        x = torch.tanh(x) if last else torch.relu(x)
    return x


def color_network(sd: Dict[str, torch.Tensor], points, normals, view_dirs, feat,
                  multires_view: int = 4) -> torch.Tensor:
    """RenderingNetwork.forward, mode 'idr', rendering_network.py:62-108 -> [P,3]."""
    n_layers = _num_layers(sd)
    x = torch.cat([points, embed(view_dirs, multires_view), normals.detach(), feat], dim=-1)
    for i in range(n_layers):
        last = i == n_layers - 1
        x = _linear_bn(sd, i, x, has_bn=not last)
        if not last:
            x = torch.relu(x)
    return torch.sigmoid(x)


# --------------------------------------------------------------------------------------
# density + compositing
# --------------------------------------------------------------------------------------
def _cos(x: torch.Tensor, y: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """F.cosine_similarity of torch 2.x: each norm clamped separately."""
    nx = torch.sqrt((x * x).sum(-1, keepdim=True)).clamp_min(eps)
    ny = torch.sqrt((y * y).sum(-1, keepdim=True)).clamp_min(eps)
    return ((x / nx) * (y / ny)).sum(-1)


def window_cosine(normals: torch.Tensor, window: int) -> torch.Tensor:
    """functions.py:41-72 with uniform weights 1/W (vector_field_nerf.py:453). [R,N,3] -> [R,N-1]."""
    R, N, _ = normals.shape
    L = N - 1
    start = int((window + 1) / 2 + 1)
    nb = start - 2                                   # partners on each side besides j+1
    w = torch.ones(window) / window
    normalizer = torch.tensor(0.0)
    for i in range(window):
        normalizer = normalizer + w[i].abs()
    middle = int((window - 1) / 2)
    x, y = normals[:, :-1, :], normals[:, 1:, :]
    base = _cos(x, y)
    lo, hi = start, L - start                        # windowed band j in [lo, hi)
    if hi <= lo:
        return base
    band = base[:, lo:hi] * w[middle] / normalizer
    xc = x[:, lo:hi, :]
    for i in range(1, nb + 1):
        fwd = _cos(xc, y[:, lo + i:hi + i, :])           # n[j] . n[j+1+i]
        bwd = _cos(xc, y[:, lo - i - 1:hi - i - 1, :])   # n[j] . n[j-i]
        band = band + fwd * w[middle + i].abs() / normalizer + bwd * w[middle - i].abs() / normalizer
    return torch.cat([base[:, :lo], band, base[:, hi:]], dim=1)


def laplace_cdf(x, beta, scale, mean):
    """density_functions.py:153-167."""
    return scale * (0.5 + 0.5 * torch.sign(x - mean) * (1 - torch.exp(-torch.abs(x - mean) / beta)))


def effective_density_params(beta, scale, mean, beta_bounds, scale_min, mean_bounds):
    """get_beta/get_scale/get_mean, density_functions.py:169-204."""
    b = torch.clamp(beta, beta_bounds[0], beta_bounds[1])
    s = torch.max(scale.abs(), torch.tensor(float(scale_min)))
    m = torch.clamp(mean, mean_bounds[0], mean_bounds[1])
    return b, s, m


def laplace_density(x, beta, scale, mean, cutoff: float = -0.5):
    """LaplaceDensity.density_func with the cutoff the reference *actually* uses: Density.forward
    drops its cutoff argument (density_functions.py:20-34), so it is always -0.5 (:134)."""
    return torch.relu(laplace_cdf(x, beta, scale, mean) - laplace_cdf(torch.tensor([cutoff]), beta, scale, mean))


def get_density(normals, ray_dirs, beta, scale, mean, window: int, dir_to_normal_th: float):
    """vector_field_nerf.py:442-474.  normals [R,N,3], ray_dirs [R,3] (unit) -> sigma [R,N]."""
    R, N, _ = normals.shape
    c = window_cosine(normals, window)
    c_dir = _cos(normals[:, :-1, :], ray_dirs[:, None, :].expand(R, N - 1, 3))
    sigma = laplace_density(-c.reshape(-1, 1), beta, scale, mean).reshape(R, N - 1)
    kill = torch.logical_and(c_dir < dir_to_normal_th, c < 0)
    sigma = torch.where(kill, torch.zeros_like(sigma), sigma)
    return torch.cat([sigma, torch.zeros(R, 1)], dim=-1), c


def volsdf_weights(z: torch.Tensor, sigma: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    """rendering.py:122-148."""
    R = z.shape[0]
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((R, 1), 1e10)], dim=-1)
    energy = dists * sigma
    shifted = torch.cat([torch.zeros(R, 1), energy[:, :-1]], dim=-1)
    trans = torch.exp(-torch.cumsum(shifted, dim=-1))
    alpha = 1.0 - torch.exp(-energy)
    w = alpha * trans
    if normalize:
        w = w / (w.sum(dim=-1, keepdim=True) + 1e-5)
    return w


# --------------------------------------------------------------------------------------
# render()
# --------------------------------------------------------------------------------------
def nerf_weights(sigma: torch.Tensor, z: torch.Tensor, normalize: bool = False) -> torch.Tensor:
    """nerf_volume_rendering, utils/rendering.py:98-119 (note the INCLUSIVE cumprod and the +1e-10)."""
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((z.shape[0], 1), 1e10)], dim=-1)
    free_energy = dists * sigma
    alpha = 1.0 - torch.exp(-free_energy)
    w = alpha * torch.cumprod(1. - alpha + 1e-10, dim=-1, dtype=torch.float32)
    if normalize:
        w = w / (w.sum(dim=-1, keepdim=True) + 1e-5)
    return w


def render(vf_sd, rn_sd, density_params, cfg: dict, uv, pose, intrinsics, t_vals,
           U1=None, U2=None, U3=None, z_vals_override: Optional[torch.Tensor] = None) -> dict:
    """VectorFieldNerf.render() in eval mode with rendering='volsdf', vector_field_nerf.py:216-338.

    cfg keys: n_coarse, n_fine (= min(N_samples, max_samples)), near, far, fine_range, perturb,
    window, dir_to_normal_th, normalize, beta_bounds, scale_min, mean_bounds, multires,
    multires_view, skip_in.  Optional: rendering ("volsdf" | "nerf": nerf_volume_rendering with its
    arguments in the function's own order -- upstream's call swaps them), white (True:
    rgb += 1 - sum of weights after the final composite, vector_field_nerf.py:325-329 -- upstream the same
    statement in the coarse block raises first), fine_near / fine_far (the fine sampler's own range).
    density_params: dict(beta, scale, mean) of 0-d tensors.
    z_vals_override: if given, the second pass uses these merged z values instead of the ones
    derived from the coarse pass (the parity protocol of SURVEY.md §8c: the argmax that places
    the fine samples is discontinuous in the VF output).
    """
    R = uv.shape[0]
    beta, scale, mean = effective_density_params(
        density_params["beta"], density_params["scale"], density_params["mean"],
        cfg["beta_bounds"], cfg["scale_min"], cfg["mean_bounds"])
    directions, ray_dirs, cam_loc = ray_geometry(uv, pose, intrinsics)
    out = {"directions": directions, "ray_dirs": ray_dirs, "cam_loc": cam_loc}

    # coarse pass (no grad in the reference, :252)
    with torch.no_grad():
        z_c = coarse_z_vals(R, cfg["near"], cfg["far"], t_vals, cfg["perturb"], U1)
        pts_c = sample_points(cam_loc, z_c, directions)
        vf_c = vf_network(vf_sd, pts_c.reshape(-1, 3), cfg["multires"], cfg["skip_in"])
        n_c = vf_c[:, :3].reshape(R, -1, 3)
        sigma_c, _ = get_density(n_c, ray_dirs, beta, scale, mean, cfg["window"], cfg["dir_to_normal_th"])
        nerf_mode = cfg.get("rendering", "volsdf") == "nerf"
        weights_fn = (lambda zz, ss: nerf_weights(ss, zz, cfg["normalize"])) if nerf_mode else \
                     (lambda zz, ss: volsdf_weights(zz, ss, cfg["normalize"]))
        w_c = weights_fn(z_c, sigma_c)
        if z_vals_override is None:
            z = fine_z_vals(z_c, w_c, cfg.get("fine_near", cfg["near"]), cfg.get("fine_far", cfg["far"]),
                            cfg["fine_range"], cfg["n_fine"], cfg["perturb"], U2, U3)
        else:
            z = z_vals_override
        pts = sample_points(cam_loc, z, directions)
    out.update(z_coarse=z_c, points_coarse_pass=pts_c, normals_coarse_pass=n_c,
               sigma_coarse=sigma_c, weights_coarse=w_c, z_vals=z, points=pts)

    # merged pass (with grad)
    N = z.shape[1]
    vf = vf_network(vf_sd, pts.reshape(-1, 3), cfg["multires"], cfg["skip_in"])
    normals = vf[:, :3].reshape(R, N, 3)
    feat = vf[:, 3:]
    sigma, cosw = get_density(normals, ray_dirs, beta, scale, mean, cfg["window"], cfg["dir_to_normal_th"])
    w = weights_fn(z, sigma)
    rep_dirs = ray_dirs.unsqueeze(1).repeat(1, N, 1).reshape(-1, 3)
    colors = color_network(rn_sd, pts.reshape(-1, 3), vf[:, :3], rep_dirs, feat, cfg["multires_view"])
    rgb = torch.sum(w.unsqueeze(-1) * colors.reshape(R, N, 3), dim=1)
    depth = torch.sum(w.unsqueeze(-1) * z.unsqueeze(-1), dim=1)
    if cfg.get("white", False):
        rgb = rgb + (1. - torch.sum(w, -1)[..., None])                      # :325-329
    out.update(normals=normals, feat=feat, cosw=cosw, sigma=sigma, weights=w, colors=colors,
               rgb=rgb, depth=depth, rep_ray_dirs=rep_dirs)
    return out


def vf_loss(rgb, depth, normals_flat, rgb_gt, depth_gt, weights: dict, depth_clamp: float,
            supervised=None, supervised_gt=None, epoch: int = 0, norm_lt1_start: int = 11000):
    """VFLoss.forward, models/losses/vf_loss.py:34-87 (directional-derivative term omitted: its
    weight is 0.0 in the shipped config and the tensor carries no grad)."""
    loss = weights["rgb"] * (rgb - rgb_gt).abs().mean()
    loss = loss + weights["depth"] * (depth - depth_gt).abs().clamp(max=depth_clamp).mean()
    nrm = torch.norm(normals_flat, dim=1)
    loss = loss + weights["unit_norm"] * torch.mean((nrm - 1) ** 2)
    if supervised is not None and supervised.nelement() > 0:
        loss = loss + weights["supervision"] * torch.mean((supervised - supervised_gt) ** 2)
    if epoch >= norm_lt1_start:
        loss = loss + weights["norm_smaller_than_one"] * torch.mean(torch.relu(nrm - 1) ** 2)
    return loss


# --------------------------------------------------------------------------------------
# train mode: BatchNorm batch statistics, autograd Jacobian, directional derivatives
# --------------------------------------------------------------------------------------
BN_MOMENTUM = 0.1  # torch.nn.BatchNorm1d default


def _linear_bn_train(sd: Dict[str, torch.Tensor], i: int, x: torch.Tensor, stats: Optional[dict]) -> torch.Tensor:
    """Linear + BatchNorm1d in TRAINING mode: normalise with the batch's mean and biased variance; ``stats`` (if given)
    receives the values the module would fold into its running statistics (mean, UNBIASED variance)."""
    W, b = sd[f"layers.{i}.0.weight"], sd[f"layers.{i}.0.bias"]
    g, be = sd[f"layers.{i}.1.weight"], sd[f"layers.{i}.1.bias"]
    y = F.linear(x, W, b)
    mean = y.mean(dim=0)
    var = ((y - mean) ** 2).mean(dim=0)
    if stats is not None:
        n = y.shape[0]
        stats[i] = (mean.detach(), (var * n / max(n - 1, 1)).detach())
    return (y - mean) / torch.sqrt(var + BN_EPS) * g + be


def vf_network_train(sd: Dict[str, torch.Tensor], points: torch.Tensor, multires: int = 6, skip_in=(4,),
                     stats: Optional[dict] = None) -> torch.Tensor:
    """VectorFieldNetwork._forward with the module in train() mode, vector_field_network.py:177-208."""
    n_layers = _num_layers(sd)
    emb = embed(points, multires)
    x = emb
    for i in range(n_layers):
        if i in skip_in:
            x = torch.cat([x, emb], 1) / torch.sqrt(torch.tensor([2.0]))
        last = i == n_layers - 1
        x = _linear_bn(sd, i, x, has_bn=False) if last else _linear_bn_train(sd, i, x, stats)
        x = torch.tanh(x) if last else torch.relu(x)
    return x


def vf_network_train_with_jacobian(sd, points, multires: int = 6, skip_in=(4,), stats: Optional[dict] = None):
    """VectorFieldNetwork.forward in train() mode, vector_field_network.py:140-175: [y, d(sum_j y_j0)/dx, d(sum_j y_j1)/dx,
    d(sum_j y_j2)/dx].  The three gradients are taken of COLUMN SUMS over the batch, through the batch statistics, so
    they are not per-sample Jacobians: every sample's row also carries the other samples' dependence on it via the
    batch mean and variance."""
    with torch.enable_grad():
        pts = points.detach().clone().requires_grad_(True)
        y = vf_network_train(sd, pts, multires, skip_in, stats)
        rows = [torch.autograd.grad(y[:, k].sum(), pts, create_graph=False, retain_graph=True)[0] for k in range(3)]
    return y, torch.cat(rows, dim=-1)


def directional_derivatives(normals: torch.Tensor, jac9: torch.Tensor) -> torch.Tensor:
    """compute_directional_derivatives, vector_field_nerf.py:476-498: [P,3], [P,9] -> [P,2,3]."""
    J = jac9.reshape(-1, 3, 3)
    n1 = torch.stack([normals[:, 1], -normals[:, 0], torch.zeros_like(normals[:, 0])], dim=1)
    n2 = torch.cross(normals, n1, dim=-1)
    d1 = torch.bmm(J, F.normalize(n1, dim=-1).unsqueeze(-1)).squeeze(-1)
    d2 = torch.bmm(J, F.normalize(n2, dim=-1).unsqueeze(-1)).squeeze(-1)
    return torch.stack([d1, d2], dim=1)


def color_network_train(sd, points, normals, view_dirs, feat, multires_view: int = 4, stats: Optional[dict] = None):
    """RenderingNetwork.forward (mode 'idr') with the module in train() mode."""
    n_layers = _num_layers(sd)
    x = torch.cat([points, embed(view_dirs, multires_view), normals.detach(), feat], dim=-1)
    for i in range(n_layers):
        last = i == n_layers - 1
        x = _linear_bn(sd, i, x, has_bn=False) if last else torch.relu(_linear_bn_train(sd, i, x, stats))
    return torch.sigmoid(x)


def fold_running(sd: Dict[str, torch.Tensor], stats: dict) -> Dict[str, torch.Tensor]:
    """What one training-mode forward does to the BatchNorm buffers: running = (1 - m) running + m batch."""
    out = dict(sd)
    for i, (mean, var_unbiased) in stats.items():
        out[f"layers.{i}.1.running_mean"] = (1 - BN_MOMENTUM) * sd[f"layers.{i}.1.running_mean"] + BN_MOMENTUM * mean
        out[f"layers.{i}.1.running_var"] = (1 - BN_MOMENTUM) * sd[f"layers.{i}.1.running_var"] + BN_MOMENTUM * var_unbiased
        key = f"layers.{i}.1.num_batches_tracked"
        if key in sd:
            out[key] = sd[key] + 1
    return out


def render_train(vf_sd, rn_sd, density_params, cfg: dict, uv, pose, intrinsics, t_vals, U1=None, U2=None, U3=None,
                 z_vals_override: Optional[torch.Tensor] = None) -> dict:
    """VectorFieldNerf.render() after model.train() (numerical_jacobian False), vector_field_nerf.py:216-338: both passes
    of the VF net and the colour net normalise with batch statistics; the coarse pass also yields the directional
    derivatives, which the reference then duplicates (:305: cat([dir, dir]) -- the merged pass's own derivatives are
    computed and dropped) and returns as per-row norms [4 R Nc] without gradient.  Returns the eval-mode keys plus
    ``directional_derivatives``, ``vf_sd_after`` / ``rn_sd_after`` (state dicts with the updated running statistics)."""
    R = uv.shape[0]
    beta, scale, mean = effective_density_params(
        density_params["beta"], density_params["scale"], density_params["mean"],
        cfg["beta_bounds"], cfg["scale_min"], cfg["mean_bounds"])
    directions, ray_dirs, cam_loc = ray_geometry(uv, pose, intrinsics)
    out = {"directions": directions, "ray_dirs": ray_dirs, "cam_loc": cam_loc}
    weights_fn = lambda zz, ss: volsdf_weights(zz, ss, cfg["normalize"])       # noqa: E731
    with torch.no_grad():
        z_c = coarse_z_vals(R, cfg["near"], cfg["far"], t_vals, cfg["perturb"], U1)
        pts_c = sample_points(cam_loc, z_c, directions)
        st1: dict = {}
        y_c, jac_c = vf_network_train_with_jacobian(vf_sd, pts_c.reshape(-1, 3), cfg["multires"], cfg["skip_in"], st1)
        y_c, jac_c = y_c.detach(), jac_c.detach()
        vf_mid = fold_running(vf_sd, st1)
        n_c = y_c[:, :3].reshape(R, -1, 3)
        dd = directional_derivatives(y_c[:, :3], jac_c).reshape(-1, 3)
        sigma_c, _ = get_density(n_c, ray_dirs, beta, scale, mean, cfg["window"], cfg["dir_to_normal_th"])
        w_c = weights_fn(z_c, sigma_c)
        if z_vals_override is None:
            z = fine_z_vals(z_c, w_c, cfg.get("fine_near", cfg["near"]), cfg.get("fine_far", cfg["far"]),
                            cfg["fine_range"], cfg["n_fine"], cfg["perturb"], U2, U3)
        else:
            z = z_vals_override
        pts = sample_points(cam_loc, z, directions)
    out.update(z_coarse=z_c, normals_coarse_pass=n_c, jacobian_coarse=jac_c, weights_coarse=w_c, z_vals=z, points=pts)
    N = z.shape[1]
    st2: dict = {}
    vf = vf_network_train(vf_sd, pts.reshape(-1, 3), cfg["multires"], cfg["skip_in"], st2)
    normals = vf[:, :3].reshape(R, N, 3)
    feat = vf[:, 3:]
    sigma, cosw = get_density(normals, ray_dirs, beta, scale, mean, cfg["window"], cfg["dir_to_normal_th"])
    w = weights_fn(z, sigma)
    rep_dirs = ray_dirs.unsqueeze(1).repeat(1, N, 1).reshape(-1, 3)
    st3: dict = {}
    colors = color_network_train(rn_sd, pts.reshape(-1, 3), vf[:, :3], rep_dirs, feat, cfg["multires_view"], st3)
    rgb = torch.sum(w.unsqueeze(-1) * colors.reshape(R, N, 3), dim=1)
    depth = torch.sum(w.unsqueeze(-1) * z.unsqueeze(-1), dim=1)
    out.update(normals=normals, feat=feat, cosw=cosw, sigma=sigma, weights=w, colors=colors, rgb=rgb, depth=depth,
               rep_ray_dirs=rep_dirs, directional_derivatives=torch.cat([dd, dd], dim=0).norm(dim=-1),
               vf_sd_after=fold_running(vf_mid, st2), rn_sd_after=fold_running(rn_sd, st3))
    return out
