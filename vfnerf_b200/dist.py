"""Multi-GPU plumbing: one process per GPU, rays sharded by contiguous slices, weights replicated.

The data path has no collective: each rank renders its slice with its slice of the global uniform draws, so the
concatenation of the per-rank results is bitwise the single-GPU result.  The only exchanges are the final gather of
rgb + depth (16 B/ray) for image rendering and the all-reduce of the gradient arenas for training.  Works with any
torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of n units for `rank`: sizes differ by at most one, earlier ranks get the extras."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rays(tensors: Sequence[Optional[torch.Tensor]], rank: Optional[int] = None,
               world_size: Optional[int] = None) -> List[Optional[torch.Tensor]]:
    """Slice every per-ray tensor (uv, pose, intrinsics, U1, U2, U3, targets ...) along dim 0 for this rank."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    out = []
    for t in tensors:
        if t is None:
            out.append(None)
        else:
            lo, hi = shard_bounds(t.shape[0], rank, world_size)
            out.append(t[lo:hi])
    return out


def gather_render(rgb: torch.Tensor, depth: torch.Tensor, n_total: int, dst: int = 0):
    """Collect the per-rank [r_i,3] rgb and [r_i,1] depth on `dst` in ray order.  Returns (rgb, depth) on dst, (None, None)
    elsewhere.  Ranks may hold different slice sizes (shard_bounds), so slices are padded to the largest one."""
    rank, w = world()
    if w == 1:
        return rgb, depth
    both = torch.cat([rgb, depth], dim=1).contiguous()
    sizes = [shard_bounds(n_total, r, w) for r in range(w)]
    m = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(m, 4, dtype=both.dtype, device=both.device)
    pad[:both.shape[0]] = both
    bufs = [torch.empty_like(pad) for _ in range(w)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None, None
    full = torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)
    return full[:, :3].contiguous(), full[:, 3:].contiguous()


def allreduce_gradients(model, average: bool = True) -> None:
    """Sum (or average) the gradients of every parameter of a VectorFieldNerf across ranks.

    Gradients produced by render()'s backward are views of one flat arena per network, so they are packed into a
    single flat buffer and reduced with ONE collective; the result is scattered back in place.  With average=True
    the loss of the global batch is the mean over ranks of per-rank means (equal slice sizes assumed, like the
    weak-scaled training configuration of BASELINE.json)."""
    rank, w = world()
    if w == 1:
        return
    flats = [model.vector_field_network.arena().grad_flat, model.rendering_network.arena().grad_flat,
             getattr(model.density, "grad_flat", None)]
    if all(f is not None for f in flats):
        # flat-gradient mode (optim.ArenaAdam): the gradient tensor IS the communication buffer -- no packing, no scatter.
        # ArenaAdam keeps all three segments in one tensor: ONE 3.2 MB collective, averaged by NCCL itself
        # (ncclAvg) so no divide kernel follows; backends without AVG (gloo) sum and divide.
        whole = getattr(model.optimizer, "grad_all", None)
        lo, hi = (whole.data_ptr(), whole.data_ptr() + 4 * whole.numel()) if whole is not None else (0, 0)
        bufs = [whole] if whole is not None and all(lo <= f.data_ptr() < hi for f in flats) else flats
        avg_op = average and dist.get_backend() == "nccl"
        for f in bufs:
            dist.all_reduce(f, op=dist.ReduceOp.AVG if avg_op else dist.ReduceOp.SUM)
            if average and not avg_op:
                f /= w
        return
    seen, params = set(), []
    for p in model.parameters():            # the reference's list holds the VF parameters twice
        if id(p) not in seen and p.grad is not None:
            seen.add(id(p))
            params.append(p)
    if not params:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= w
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


def broadcast_parameters(model, src: int = 0) -> None:
    """Make every rank start from rank `src`'s weights (one broadcast per arena)."""
    rank, w = world()
    if w == 1:
        return
    for t in (model.vector_field_network.arena().flat, model.rendering_network.arena().flat, model.density.flat()):
        dist.broadcast(t, src=src)
