"""VFLoss on the device (reference: models/losses/vf_loss.py:12-87).

Same constructor arguments, same ``forward(pred, gt, epoch) -> (loss, dict)`` contract and the same arithmetic as the
reference, but the six terms are ONE reduction launch and their gradients ONE elementwise launch
(``vfnerf_vf_loss_fwd`` / ``vfnerf_vf_loss_bwd``, csrc/loss.cu) instead of ~35 aten kernels, and the per-term values are
fetched with a single device->host copy (the reference calls ``.item()`` six times).  ``sync=False`` skips even that:
the dict then holds 0-d device tensors, which is what a CUDA-graph-captured step needs (graphed.py)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib

_NAMES = ("rgb_loss", "depth_loss", "unit_norm_loss", "supervision_loss", "norm_smaller_than_one_loss",
          "directional_derivatives_loss")


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous().float()


class _VFLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, depth, normals, sup, rgb_gt, depth_gt, sup_gt, dd, weights, clamp, lt1):
        L = _lib.lib()
        dev = rgb.device
        if dev.type != "cuda":
            raise RuntimeError("vfnerf_b200.losses.VFLoss needs CUDA tensors: there is no CPU fallback")
        rgb, normals, rgb_gt = _f32(rgb), _f32(normals), _f32(rgb_gt)
        depth = None if depth is None else _f32(depth)
        depth_gt = None if depth_gt is None or depth_gt.nelement() == 0 else _f32(depth_gt)
        n_sup = 0 if sup is None else sup.shape[0]
        sup = None if n_sup == 0 else _f32(sup)
        sup_gt = None if n_sup == 0 else _f32(sup_gt)
        dd = None if dd is None else _f32(dd.detach())
        terms = torch.empty(16, dtype=torch.float32, device=dev)
        w = (C.c_float * 6)(*weights)
        R, P = rgb.shape[0], normals.shape[0]
        _lib.check(L.vfnerf_vf_loss_fwd(R, P, n_sup, 0 if dd is None else dd.numel(), rgb.data_ptr(), rgb_gt.data_ptr(),
                                        _lib.ptr(depth), _lib.ptr(depth_gt), normals.data_ptr(), _lib.ptr(sup),
                                        _lib.ptr(sup_gt), _lib.ptr(dd), w, float(clamp), int(lt1), terms.data_ptr(),
                                        _stream(dev)), "vfnerf_vf_loss_fwd")
        ctx.save_for_backward(rgb, rgb_gt, normals, *[t for t in (depth, depth_gt, sup, sup_gt) if t is not None])
        ctx.have = (depth is not None, depth_gt is not None, sup is not None)
        ctx.cfg = (tuple(weights), float(clamp), int(lt1), R, P, n_sup)
        ctx.mark_non_differentiable(terms)
        return terms[6], terms

    @staticmethod
    def backward(ctx, g_loss, _g_terms):
        L = _lib.lib()
        saved = list(ctx.saved_tensors)
        rgb, rgb_gt, normals = saved[:3]
        rest = saved[3:]
        has_depth, has_dgt, has_sup = ctx.have
        depth = rest.pop(0) if has_depth else None
        depth_gt = rest.pop(0) if has_dgt else None
        sup = rest.pop(0) if has_sup else None
        sup_gt = rest.pop(0) if has_sup else None
        weights, clamp, lt1, R, P, n_sup = ctx.cfg
        dev = rgb.device
        need = ctx.needs_input_grad
        d_rgb = torch.empty_like(rgb) if need[0] else None
        d_depth = torch.empty_like(depth) if (need[1] and depth is not None) else None
        d_normals = torch.empty_like(normals) if need[2] else None
        d_sup = torch.empty_like(sup) if (need[3] and sup is not None) else None
        g = g_loss.contiguous().float()
        w = (C.c_float * 6)(*weights)
        _lib.check(L.vfnerf_vf_loss_bwd(R, P, n_sup, rgb.data_ptr(), rgb_gt.data_ptr(), _lib.ptr(depth), _lib.ptr(depth_gt),
                                        normals.data_ptr(), _lib.ptr(sup), _lib.ptr(sup_gt), w, clamp, lt1, g.data_ptr(),
                                        _lib.ptr(d_rgb), _lib.ptr(d_depth), _lib.ptr(d_normals), _lib.ptr(d_sup),
                                        _stream(dev)), "vfnerf_vf_loss_bwd")
        return d_rgb, d_depth, d_normals, d_sup, None, None, None, None, None, None, None


class VFLoss(torch.nn.Module):
    """Drop-in for the reference's ``VFLoss(config, weights)``: ``config`` needs ``norm_smaller_than_one_start``,
    ``depth_loss_clamp`` and ``directional_derivatives_start``; ``weights`` needs ``rgb, depth, unit_norm, supervision,
    norm_smaller_than_one, directional_derivatives`` (config_parser/vf_nerf_config.py:133-150)."""

    def __init__(self, config, weights, sync: bool = True) -> None:
        super().__init__()
        self.config, self.weights, self.sync = config, weights, sync

    def forward(self, pred: Dict[str, Optional[torch.Tensor]], gt: Dict[str, torch.Tensor], epoch: int
                ) -> Tuple[torch.Tensor, Dict[str, object]]:
        wt = self.weights
        w = (float(wt.rgb), float(wt.depth), float(wt.unit_norm), float(wt.supervision),
             float(wt.norm_smaller_than_one), float(wt.directional_derivatives))
        dd = pred.get("directional_derivatives")
        if dd is not None and epoch < self.config.directional_derivatives_start:
            dd = None
        sup = pred.get("supervised_normals")
        if sup is not None and sup.nelement() == 0:
            sup = None
        loss, terms = _VFLossFn.apply(pred["rgb"], pred["depth"], pred["normals"], sup, gt["rgb"], gt.get("depth"),
                                      gt.get("supervised_normals") if sup is not None else None, dd, w,
                                      float(self.config.depth_loss_clamp),
                                      epoch >= self.config.norm_smaller_than_one_start)
        if self.sync:
            vals = terms[:6].tolist()                       # one device->host copy for all six terms
            return loss, dict(zip(_NAMES, vals))
        return loss, {k: terms[i] for i, k in enumerate(_NAMES)}
