"""Dense VF queries for marching cubes (BASELINE config 5).

``get_set_predictions`` keeps the reference's signature (evaluation/utils/mc_utils.py:88-104): CPU
samples in, chunked over the GPU, CPU ``[P,3]`` out.  ``grid_query`` is the B200-first form of the same
work: grid coordinates are generated in-kernel from the linear index with the reference's fp32 op order
(evaluation/methods.py:194-208), only the three vector outputs are evaluated, and the result stays on
the device (optionally for a z-slab ``[i0, i0+n)`` of the grid -- the multi-GPU partition of §8e)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib, ops


_STAGING: dict = {}


def _staging(device: torch.device, rows: int):
    """Two pinned (input, output) staging pairs per (device, chunk size), allocated once: pinning memory costs
    milliseconds, far more than moving a chunk."""
    key = (str(device), rows)
    if key not in _STAGING:
        _STAGING.clear()                 # keep one size alive
        _STAGING[key] = [(torch.empty(rows, 3).pin_memory(), torch.empty(rows, 3).pin_memory(), torch.cuda.Event())
                         for _ in range(2)]
    return _STAGING[key]


def get_set_predictions(decoder, samples: torch.Tensor, max_batch: int, device: torch.device) -> torch.Tensor:
    """evaluation/utils/mc_utils.py:88-104: CPU samples [P, >=3] in, CPU vectors [P,3] out, chunked over the GPU.
    Chunks are double-buffered through pinned staging memory so the host copies, both PCIe directions and the tensor-core
    query of consecutive chunks overlap."""
    samples.requires_grad = False
    n = samples.shape[0]
    out = torch.zeros(n, 3, dtype=samples.dtype)
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("get_set_predictions needs a CUDA device: vfnerf_b200 has no CPU fallback")
    rows = min(max_batch, max(n, 1))
    slots = _staging(device, rows)
    pending = [None, None]                 # (lo, hi) of the chunk whose result sits in the slot's output buffer
    stream = torch.cuda.current_stream(device)

    def drain(k):
        if pending[k] is not None:
            lo, hi = pending[k]
            slots[k][2].synchronize()
            out[lo:hi].copy_(slots[k][1][:hi - lo])
            pending[k] = None
    with torch.no_grad():
        for i, head in enumerate(range(0, n, max_batch)):
            k = i & 1
            lo, hi = head, min(head + max_batch, n)
            drain(k)
            s_in, s_out, ev = slots[k]
            s_in[:hi - lo].copy_(samples[lo:hi, :3])
            sub = s_in[:hi - lo].to(device, non_blocking=True)
            s_out[:hi - lo].copy_(ops.vf_query(decoder, sub, n_cols=3), non_blocking=True)
            ev.record(stream)
            pending[k] = (lo, hi)
    drain(0)
    drain(1)
    return out


def grid_query(decoder, resolution: int, scale: float = 1.0, translation: Optional[torch.Tensor] = None,
               centroid: Optional[torch.Tensor] = None, i0: int = 0, n_points: Optional[int] = None,
               chunk: int = 1 << 20, return_points: bool = False):
    """VF vectors on the ``resolution^3`` grid of marching_cubes_mesh (methods.py:190-210).  Returns a CUDA
    tensor ``[n_points, 3]`` for grid indices ``[i0, i0 + n_points)`` (z fastest)."""
    L = _lib.lib()
    ar = decoder.arena()
    dev = ar.flat.device
    if dev.type != "cuda":
        raise RuntimeError("grid_query needs the network on a CUDA device: vfnerf_b200 has no CPU fallback")
    total = resolution ** 3
    n_points = total - i0 if n_points is None else n_points
    tr = [0.0, 0.0, 0.0] if translation is None else [float(x) for x in translation]
    ce = [0.0, 0.0, 0.0] if centroid is None else [float(x) for x in centroid]
    f3 = C.c_float * 3
    origin, tr_c, ce_c = f3(-scale, -scale, -scale), f3(*tr), f3(*ce)
    voxel = scale * 2.0 / (resolution - 1)
    prec = _lib.PRECISIONS[decoder.precision]
    out = torch.empty(n_points, 3, dtype=torch.float32, device=dev)
    chunk = min(chunk, max(n_points, 1))
    nbytes = L.vfnerf_vf_workspace_bytes(C.byref(ar.desc), chunk, decoder.multires, 0, prec)
    if nbytes < 0:
        _lib.check(1, "vfnerf_vf_workspace_bytes")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    pts = torch.empty(n_points, 3, dtype=torch.float32, device=dev) if return_points else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    from .ops import _on_device
    guard = _on_device(dev)
    guard.__enter__()
    head = 0
    while head < n_points:
        n = min(chunk, n_points - head)
        _lib.check(L.vfnerf_vf_grid_query(C.byref(ar.desc), ar.flat.data_ptr(), decoder.multires, decoder.skip_layer,
                                          1e-5, prec, resolution, i0 + head, n, origin, tr_c, ce_c, voxel,
                                          out[head:].data_ptr(), ws.data_ptr(), ws.numel(), stream), "vfnerf_vf_grid_query")
        if return_points:
            # fp32 path: the generated coordinates sit right after the embedding in the workspace
            E = 3 + 6 * decoder.multires
            off = ((n * E * 4 + 255) // 256) * 256
            pts[head:head + n] = ws[off:off + n * 12].view(torch.float32).view(n, 3)
        head += n
    guard.__exit__()
    return (out, pts) if return_points else out
