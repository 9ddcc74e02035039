"""Device-side replacement of the reference's marching-cubes preprocessing (evaluation/utils/mc_utils.py and the glue of
evaluation/methods.py:209-278): SURVEY.md §8f rank 3.

``get_set_predictions`` (the chunked VF query of mc_utils.py:88-104) lives in grid_query.py.  This module turns the dense
grid of VF vectors into what ``marching_cubes_vt.contrastive_marching_cubes`` consumes, without the grid ever leaving
the GPU: two kernel launches (csrc/mc_preprocess.cu) replace ``extract_divergence`` -> ``unify_direction`` ->
``make_comb_format`` -> block-ordered masking, whose fp32 intermediates are 250 B per grid point on the CPU.  Only the
~N^2 surface cells (348 B each) cross PCIe.  No CPU path: host tensors are rejected.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib, ops

MC_CTA = 256      # VFNERF_MC_CTA


def _count(pred: torch.Tensor, N: int, want_dense: bool, surface: Optional[torch.Tensor] = None):
    pred = ops._require_cuda("prediction", pred.detach()).reshape(-1, 3)
    if pred.shape[0] != N ** 3:
        raise ValueError(f"prediction must hold resolution^3 = {N ** 3} vectors, got {pred.shape[0]}")
    dev = pred.device
    n_q = 8 * (N // 2) ** 3
    n_cta = max((n_q + MC_CTA - 1) // MC_CTA, 1)
    keep = torch.empty(max(n_q, 1), dtype=torch.uint8, device=dev)
    counts = torch.zeros(n_cta, dtype=torch.int32, device=dev)
    div_raw = torch.zeros(N ** 3, dtype=torch.float32, device=dev) if want_dense else None
    choice = torch.zeros(N ** 3, dtype=torch.uint8, device=dev) if want_dense else None
    if surface is not None:
        surface = (surface.to(dev).reshape(-1) != 0).to(torch.uint8).contiguous()
        if surface.numel() != N ** 3:
            raise ValueError("surface must hold resolution^3 flags")
    _lib.check(_lib.lib().vfnerf_mc_count(pred.data_ptr(), N, keep.data_ptr(), counts.data_ptr(), _lib.ptr(div_raw),
                                          _lib.ptr(choice), None if surface is None else surface.data_ptr(),
                                          ops._stream_ptr(dev)), "vfnerf_mc_count")
    return pred, keep, counts, div_raw, choice


def mc_preprocess(prediction: torch.Tensor, resolution: int, surface: Optional[torch.Tensor] = None
                  ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """prediction [res^3,3] (CUDA; the output of the grid query, x index slowest) ->
    (selected_indices [M,3] int32, comb_values [M,28] float32, udf [M,28,2] float32), all on the device, rows in the
    reference's order.  ``comb_values.reshape(-1)``, ``selected_indices`` and ``udf.reshape(-1, 2)`` are the arguments
    of ``contrastive_marching_cubes`` (methods.py:272-283).  ``surface`` ([res^3] or [res,res,res], optional): surface-cell
    flags decided on another field (methods.py:213-218 tests the divergence before its k = 9 smoothing)."""
    N = int(resolution)
    pred, keep, counts, _, _ = _count(prediction, N, False, surface)
    dev = pred.device
    offsets = torch.cumsum(counts, dim=0, dtype=torch.int64)         # one entry per 256 cells
    M = int(offsets[-1].item())                                      # the one sync: sizes the outputs
    cells = torch.empty(M, 3, dtype=torch.int32, device=dev)
    comb = torch.empty(M, 28, dtype=torch.float32, device=dev)
    udf = torch.empty(M, 28, 2, dtype=torch.float32, device=dev)
    if M:
        _lib.check(_lib.lib().vfnerf_mc_emit(pred.data_ptr(), N, keep.data_ptr(), offsets.data_ptr(), cells.data_ptr(),
                                             comb.data_ptr(), udf.data_ptr(), ops._stream_ptr(dev)), "vfnerf_mc_emit")
    return cells, comb, udf


def extract_divergence(vt_values: torch.Tensor, N: int, return_raw: bool = False):
    """mc_utils.py:34-85 on the device: [N,N,N] float grid, 1 where the divergence of the cell is <= -0.5.
    (Cells of the last layer of an odd N are outside the block order the reference meshes, and are 0 here.)"""
    _, _, _, div_raw, _ = _count(vt_values, N, True)
    raw = div_raw.reshape(N, N, N)
    flag = (raw <= -0.5).float()
    flag[-1, :, :] = 0
    flag[:, -1, :] = 0
    flag[:, :, -1] = 0
    return (flag, raw) if return_raw else flag


def unify_direction(divergence_grid: Optional[torch.Tensor], vt_grid: torch.Tensor, N: int = 64) -> torch.Tensor:
    """mc_utils.py:107-166 on the device: [N^3, 8] int64 side choice per cell corner.  ``vt_grid`` is [3,N,N,N] like the
    reference's argument; the surface mask is recomputed from it (``divergence_grid`` is accepted for signature
    compatibility and, when given, applied as an additional mask)."""
    pred = vt_grid.permute(1, 2, 3, 0).reshape(-1, 3).contiguous()
    _, _, _, _, choice = _count(pred, N, True)
    bits = choice.to(torch.int64)
    out = torch.stack([(bits >> s) & 1 for s in range(8)], dim=1)
    if divergence_grid is not None:
        out = out * (divergence_grid.reshape(-1, 1) == 1).to(out.dtype).to(out.device)
    return out


def gaussian_taps(k: int, sigma: float):
    """The 1-D factor of GaussianSmoothing's kernel (guassian_smoothing.py:40-52: exp(-((i - mean) / (2 sigma))^2), the
    constant in front cancels), normalised to sum 1 -- the reference normalises the k^3 product, which is the same."""
    import math
    mean = (k - 1) / 2
    g = [math.exp(-(((i - mean) / (2 * sigma)) ** 2)) for i in range(k)]
    t = sum(g)
    return [x / t for x in g]


def smooth_vf(vf: torch.Tensor, k: int = 3, sigma: float = 1.0) -> torch.Tensor:
    """guassian_smoothing.py:81-97 on the device: vf [N,N,N,3] (CUDA) -> smoothed [N,N,N,3]."""
    import ctypes as C
    vf = ops._require_cuda("vf", vf.detach())
    if vf.dim() != 4 or vf.shape[3] != 3 or not (vf.shape[0] == vf.shape[1] == vf.shape[2]):
        raise ValueError(f"vf must be [N,N,N,3], got {tuple(vf.shape)}")
    N = vf.shape[0]
    taps = (C.c_float * k)(*gaussian_taps(int(k), float(sigma)))
    tmp, out = torch.empty_like(vf), torch.empty_like(vf)
    _lib.check(_lib.lib().vfnerf_smooth_vf(vf.data_ptr(), tmp.data_ptr(), out.data_ptr(), N, int(k), taps,
                                           ops._stream_ptr(vf.device)), "vfnerf_smooth_vf")
    return out


def grid_to_mc_inputs(decoder, resolution: int, scale: float = 1.0, translation=None, centroid=None,
                      smooth_after: bool = False, smooth_all: bool = False):
    """Grid query + preprocessing without leaving the GPU: the part of generate_mesh (methods.py:194-278) in front of
    contrastive_marching_cubes, including its optional smoothing steps (:211-218).  Returns numpy (comb_values [M*28],
    selected_indices [M,3], udf [M*28,2])."""
    from .grid_query import grid_query
    N = int(resolution)
    with torch.no_grad():
        pred = grid_query(decoder, resolution, scale, translation, centroid)
        if smooth_all:
            pred = smooth_vf(pred.reshape(N, N, N, 3), k=3, sigma=1).reshape(N ** 3, 3)
        if smooth_after or smooth_all:
            # the reference tests the divergence on the un-smoothed (resp. lightly smoothed) field and takes sides and
            # norms from the k = 9 smoothed one; the fused kernel does both from one field, so this variant goes stage by stage
            div = extract_divergence(pred, N)
            pred = smooth_vf(pred.reshape(N, N, N, 3), k=9, sigma=2).reshape(N ** 3, 3)
            cells, comb, udf = mc_preprocess(pred, resolution, surface=div)
        else:
            cells, comb, udf = mc_preprocess(pred, resolution)
    return comb.reshape(-1).cpu().numpy(), cells.cpu().numpy(), udf.reshape(-1, 2).cpu().numpy()
