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


def _count(pred: torch.Tensor, N: int, want_dense: bool):
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
    _lib.check(_lib.lib().vfnerf_mc_count(pred.data_ptr(), N, keep.data_ptr(), counts.data_ptr(), _lib.ptr(div_raw),
                                          _lib.ptr(choice), ops._stream_ptr(dev)), "vfnerf_mc_count")
    return pred, keep, counts, div_raw, choice


def mc_preprocess(prediction: torch.Tensor, resolution: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """prediction [res^3,3] (CUDA; the output of the grid query, x index slowest) ->
    (selected_indices [M,3] int32, comb_values [M,28] float32, udf [M,28,2] float32), all on the device, rows in the
    reference's order.  ``comb_values.reshape(-1)``, ``selected_indices`` and ``udf.reshape(-1, 2)`` are the arguments
    of ``contrastive_marching_cubes`` (methods.py:272-283)."""
    N = int(resolution)
    pred, keep, counts, _, _ = _count(prediction, N, False)
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


def grid_to_mc_inputs(decoder, resolution: int, scale: float = 1.0, translation=None, centroid=None):
    """Grid query + preprocessing without leaving the GPU: the part of generate_mesh (methods.py:194-278) in front of
    contrastive_marching_cubes.  Returns numpy (comb_values [M*28], selected_indices [M,3], udf [M*28,2])."""
    from .grid_query import grid_query
    with torch.no_grad():
        pred = grid_query(decoder, resolution, scale, translation, centroid)
        cells, comb, udf = mc_preprocess(pred, resolution)
    return comb.reshape(-1).cpu().numpy(), cells.cpu().numpy(), udf.reshape(-1, 2).cpu().numpy()
