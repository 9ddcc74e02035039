"""CUDA-graph replay of the reference trainer's inner sequence for a fixed ray count
(train/vector_field_nerf_train.py:177-260: render -> loss -> backward -> clip_grad_norm_ -> optimizer.step).

At the reference's batch of 1024 rays the B200 finishes the device work of a step in ~2 ms while Python needs ~3 ms
to issue it (autograd bookkeeping, ~60 foreach launches of clip + Adam): the step is launch-bound.  Capturing the
whole sequence once and replaying it removes the host from the loop.  Everything on the path is capture-safe: the
C ABI is stream-ordered, allocates nothing and never synchronises; workspaces come from torch's graph memory pool.

Differences from the eager path, all explicit:
  * the sampler draws U1/U2/U3 come from the DEVICE generator inside the graph (the eager path draws them on the host
    like the reference, ray_sampler.py:138,292,297) -- statistically identical, not bit-identical to a CPU stream;
  * the optimizer must be capturable (device-side step counter, tensor learning rate): make_capturable() rebuilds
    model.optimizer / model.scheduler that way, with the reference's hyper-parameters and its duplicated VF entries.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from .output import NerfOutput


def make_capturable(model) -> None:
    """Rebuild model.optimizer (Adam) and model.scheduler (ExponentialLR) so that optimizer.step() can be captured:
    capturable=True and the learning rate held in a device tensor (schedulers fill_() it in place)."""
    sc = model.config.scheduler_config
    dev = next(iter(model.parameters())).device
    old = model.optimizer.param_groups[0]
    lr = old["lr"]
    lr_t = lr.clone() if isinstance(lr, torch.Tensor) else torch.tensor(float(lr), device=dev)
    model.optimizer = torch.optim.Adam(model.parameters(), lr=lr_t, weight_decay=old.get("weight_decay", 0.0),
                                       capturable=True)
    gamma = getattr(model.scheduler, "gamma", sc.lr_decay_factor ** (1. / sc.lr_decay_steps))
    model.scheduler = torch.optim.lr_scheduler.ExponentialLR(model.optimizer, gamma)


class GraphedTrainStep:
    """One captured training step.

    :param model: vfnerf_b200.VectorFieldNerf on a CUDA device, in eval() mode like the reference trainer.
    :param loss_fn: loss_fn(outputs: NerfOutput, **targets) -> scalar tensor; torch ops only, no host syncs.
    :param n_rays: rays per step (fixed by the capture).
    :param targets: example target tensors (name -> tensor); their shapes/dtypes fix the static buffers.
    :param clip_norm: max gradient norm (None: no clipping).  :param optimizer_step: include optimizer.step().
    :param quat_pose: poses are [R,7] instead of [R,4,4].
    :param allreduce: all-reduce the gradients across ranks inside the captured step (default: when torch.distributed is
        initialised with more than one rank).
    :param given_draws: False: U1/U2/U3 are drawn by the device generator inside the graph; True: the caller passes
        ``draws=(U1, U2, U3)`` to every call (e.g. host draws in the reference's order, or a rank's slice of global draws).
    """

    def __init__(self, model, loss_fn: Callable[..., torch.Tensor], n_rays: int, targets: Dict[str, torch.Tensor],
                 clip_norm: Optional[float] = 0.5, optimizer_step: bool = True, quat_pose: bool = False,
                 warmup: int = 3, seed: int = 0, given_draws: bool = False, allreduce: Optional[bool] = None) -> None:
        self.model, self.loss_fn, self.n_rays = model, loss_fn, n_rays
        self.clip_norm, self.optimizer_step = clip_norm, optimizer_step
        dev = next(iter(model.parameters())).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs the model on a CUDA device")
        from .optim import ArenaAdam
        from . import dist as vdist
        self._arena_opt = isinstance(model.optimizer, ArenaAdam)
        # multi-GPU (one process per GPU): the gradient all-reduce is part of the captured step
        self.allreduce = (vdist.world()[1] > 1) if allreduce is None else bool(allreduce)
        if optimizer_step and not self._arena_opt and not model.optimizer.param_groups[0].get("capturable", False):
            raise RuntimeError("optimizer is not capturable: call vfnerf_b200.graphed.make_capturable(model) or "
                               "vfnerf_b200.optim.use_arena_optimizer(model) first")
        f32 = dict(dtype=torch.float32, device=dev)
        self.pose = torch.zeros((n_rays, 7) if quat_pose else (n_rays, 4, 4), **f32)
        self.pixels = torch.zeros(n_rays, 2, **f32)
        self.intrinsics = torch.zeros(n_rays, 4, 4, **f32)
        self.targets = {k: torch.zeros_like(v, device=dev) for k, v in targets.items()}
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(seed)
        self.given_draws = given_draws
        nc, nf = model.ray_sampler.N_samples, model.fine_sampler.n_fine()
        self.draws = (torch.zeros(n_rays, nc, **f32), torch.zeros(n_rays, nf, **f32), torch.zeros(n_rays, nf, **f32)) \
            if given_draws else None
        self.loss = None
        self.outputs: Optional[NerfOutput] = None
        self.graph = torch.cuda.CUDAGraph()
        # clip over each tensor once.  The reference hands clip_grad_norm_ its parameter list with every VF tensor
        # listed twice (vector_field_nerf.py:132-137); torch's foreach kernels then read-modify-write the same memory
        # from two chunks concurrently and the result depends on timing (coef or coef^2 per chunk).  A replayed graph
        # should be reproducible, so duplicates are dropped here; the optimizer keeps the reference's list.
        seen, self._params = set(), []
        for p in model.parameters():
            if id(p) not in seen:
                seen.add(id(p))
                self._params.append(p)
        self._capture(warmup)

    def _draws(self):
        if self.given_draws:
            det = self.model.ray_sampler.deterministic
            return (None if det else self.draws[0], None if det else self.draws[1], self.draws[2])
        m, R, dev = self.model, self.n_rays, self.pixels.device
        nc, nf = m.ray_sampler.N_samples, m.fine_sampler.n_fine()
        u1 = None if m.ray_sampler.deterministic else torch.rand(R, nc, device=dev, generator=self.gen)
        u2 = None if m.ray_sampler.deterministic else torch.rand(R, nf, device=dev, generator=self.gen)
        u3 = torch.rand(R, nf, device=dev, generator=self.gen)       # drawn even when deterministic (ray_sampler.py:297)
        return u1, u2, u3

    def _step(self):
        m = self.model
        out = m.render(self.pose, self.pixels, self.intrinsics, 0, draws=self._draws())
        loss = self.loss_fn(out, **self.targets)
        loss.backward()
        if self.allreduce:
            from . import dist as vdist
            vdist.allreduce_gradients(m)
        if self._arena_opt:
            # flat arenas: clipping is part of the fused update (csrc/optim.cu)
            if self.optimizer_step:
                m.optimizer.step(max_norm=self.clip_norm)
            return out, loss
        if self.clip_norm is not None:
            torch.nn.utils.clip_grad_norm_(self._params, self.clip_norm)
        if self.optimizer_step:
            m.optimizer.step()
        return out, loss

    def _capture(self, warmup: int) -> None:
        m = self.model
        # warm-up on a side stream (torch's capture recipe): lazy one-time work -- kernel attributes, optimizer state,
        # the arenas -- happens here, outside the capture.  Parameters and optimizer state are restored afterwards.
        saved = [p.detach().clone() for p in self._params]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(max(1, warmup)):
                m.optimizer.zero_grad(set_to_none=not self._arena_opt)
                self._step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.no_grad():
            for p, q in zip(self._params, saved):
                p.copy_(q)
            for st in m.optimizer.state.values():
                for v in st.values():
                    if isinstance(v, torch.Tensor):
                        v.zero_()
            if self._arena_opt:
                for t in m.optimizer._m + m.optimizer._v + [m.optimizer._step]:
                    t.zero_()
        self.graph.register_generator_state(self.gen)
        m.optimizer.zero_grad(set_to_none=not self._arena_opt)
        # a captured collective: NCCL's watchdog thread polls events while we capture, which the default (global) capture
        # mode treats as an error
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local" if self.allreduce else "global"):
            if self._arena_opt:
                m.optimizer.zero_grad()            # persistent flat gradients: zeroed inside the graph, every replay
            self.outputs, self.loss = self._step()

    def __call__(self, pose: torch.Tensor, pixels: torch.Tensor, intrinsics: torch.Tensor, draws=None,
                 **targets: torch.Tensor) -> torch.Tensor:
        """Copy the step's inputs into the static buffers and replay.  Returns the (static) loss tensor; the
        NerfOutput of the step is in ``self.outputs`` (static tensors, overwritten by the next replay)."""
        self.pose.copy_(pose, non_blocking=True)
        self.pixels.copy_(pixels, non_blocking=True)
        self.intrinsics.copy_(intrinsics, non_blocking=True)
        for k, v in targets.items():
            self.targets[k].copy_(v, non_blocking=True)
        if self.given_draws:
            if draws is None:
                raise ValueError("this step was captured with given_draws=True: pass draws=(U1, U2, U3)")
            for dst, src in zip(self.draws, draws):
                if src is not None:
                    dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss
