"""ArenaAdam: clip_grad_norm_ + Adam over the flat parameter arenas (csrc/optim.cu).

The reference trainer ends every step with ``clip_grad_norm_(model.parameters(), clip_norm)`` and ``Adam.step()``
(train/vector_field_nerf_train.py:252-258): on a B200 those ~25 foreach launches over 55 tensors take 0.55 ms of a
2 ms step.  Here every network is one flat tensor, its gradient is one flat tensor (``ParamArena.enable_flat_grad``),
and a step is one squared-norm reduction + one elementwise launch per network.  Everything stays on the device
(learning rate and step counter are device scalars), so the step can be captured in a CUDA graph (graphed.py).

Semantics: exactly torch.optim.Adam (amsgrad off) after exactly torch's clip_grad_norm_ -- with every parameter counted
ONCE.  The reference lists the VF parameters twice in its optimizer (vector_field_nerf.py:132-137), which makes torch
clip and update them twice with racy foreach kernels; that quirk is deliberately not reproduced here (the default
``model.optimizer`` still does)."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


class ArenaAdam(torch.optim.Optimizer):
    def __init__(self, model, lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 max_norm: Optional[float] = None) -> None:
        self.model = model
        vf, rn = model.vector_field_network.arena(), model.rendering_network.arena()
        vf.sync(); rn.sync()
        dev = vf.flat.device
        if dev.type != "cuda":
            raise RuntimeError("ArenaAdam needs the model on a CUDA device")
        self._arenas = (vf, rn)
        self._density = model.density
        flats = [vf.flat, rn.flat, model.density.flat()]
        super().__init__([{"params": flats}], dict(lr=torch.tensor(float(lr), device=dev), betas=betas, eps=eps,
                                                   weight_decay=weight_decay))
        self.max_norm = max_norm
        # ONE gradient tensor for everything that trains: [VF arena | colour arena | beta, scale, mean], segments starting
        # on 256-byte boundaries.  clip reads it with one launch, and multi-GPU training all-reduces it with one collective
        # (dist.allreduce_gradients); grad_scale is folded into the update (1/world after a SUMMING all-reduce).
        seg = [(n + 63) // 64 * 64 for n in (vf.flat.numel(), rn.flat.numel())]
        self.grad_all = torch.zeros(seg[0] + seg[1] + 64, dtype=torch.float32, device=dev)
        self.grad_scale = 1.0
        self._grads = [vf.enable_flat_grad(self.grad_all[:vf.flat.numel()]),
                       rn.enable_flat_grad(self.grad_all[seg[0]:seg[0] + rn.flat.numel()]),
                       model.density.enable_flat_grad(self.grad_all[seg[0] + seg[1]:seg[0] + seg[1] + 3])]
        self._scratch = torch.zeros(_lib.SQNORM_SCRATCH_FLOATS, dtype=torch.float32, device=dev)
        self._masks = [vf.trainable_mask(), rn.trainable_mask(), None]
        self._m = [torch.zeros_like(f) for f in flats]
        self._v = [torch.zeros_like(f) for f in flats]
        self._step = torch.zeros(1, dtype=torch.float32, device=dev)
        self._sq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._flats = flats

    def zero_grad(self, set_to_none: bool = False) -> None:      # the gradients are persistent views: always zero in place
        self.grad_all.zero_()

    def _check_storage(self) -> None:
        """The arenas are re-flattened when someone moves the modules (.to(), load_state_dict(assign=True)); moments held
        for the old storage would silently update orphaned memory."""
        cur = [self._arenas[0].sync(), self._arenas[1].sync(), self._density.flat()]
        gcur = [self._arenas[0].grad_flat, self._arenas[1].grad_flat, getattr(self._density, "grad_flat", None)]
        for a, b in zip(cur, self._flats):
            if a.data_ptr() != b.data_ptr():
                raise RuntimeError("ArenaAdam: parameter storage moved after the optimizer was built (.to() / "
                                   "load_state_dict(assign=True)); call vfnerf_b200.optim.use_arena_optimizer(model) again")
        for a, b in zip(gcur, self._grads):
            if a is None or a.data_ptr() != b.data_ptr():
                raise RuntimeError("ArenaAdam: flat-gradient mode was switched off (another optimizer was installed with "
                                   "model.new_scheduler()/reset_scheduler()?); call use_arena_optimizer(model) again")

    # ---- checkpoint format: the reference's ---------------------------------------------------------------------------
    # VectorFieldNerf.save() stores optimizer.state_dict() and load() restores it (vector_field_nerf.py:178-214).  The
    # reference's optimizer is torch.optim.Adam over model.parameters() -- one group, the VF tensors listed twice
    # (:132-137) -- so that is the format emitted and accepted here: per-parameter exp_avg / exp_avg_sq / step, packed
    # out of / into the flat moment arenas.  A checkpoint written by the reference trainer resumes here and vice versa.
    def _param_slots(self):
        """id(Parameter) -> (arena index, offset, numel) for every trainable tensor."""
        where = {}
        for k, ar in enumerate(self._arenas):
            for _, t, off in ar.slots:
                if isinstance(t, torch.nn.Parameter):
                    where[id(t)] = (k, off, t.numel())
        for i, p in enumerate((self._density.beta, self._density.scale, self._density.mean)):
            where[id(p)] = (2, i, 1)
        return where

    def state_dict(self):
        where = self._param_slots()
        params = self.model.parameters()          # the reference's list, duplicates included
        index = {}
        for i, p in enumerate(params):
            index[id(p)] = i          # torch's own packing: a duplicated tensor is known by its LAST position
        g0 = self.param_groups[0]
        lr = g0["lr"]
        state = {}
        stepped = bool(self._step.item() > 0)
        if stepped:
            for p in params:
                i = index[id(p)]
                if i in state:
                    continue
                k, off, n = where[id(p)]
                state[i] = {"step": self._step.detach().clone().reshape(()),
                            "exp_avg": self._m[k][off:off + n].view(p.shape).clone(),
                            "exp_avg_sq": self._v[k][off:off + n].view(p.shape).clone()}
        group = {"lr": float(lr.item()) if isinstance(lr, torch.Tensor) else float(lr), "betas": tuple(g0["betas"]),
                 "eps": g0["eps"], "weight_decay": g0["weight_decay"], "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": [index[id(p)] for p in params]}
        if "initial_lr" in g0:
            group["initial_lr"] = g0["initial_lr"]
        return {"state": state, "param_groups": [group]}

    @torch.no_grad()
    def load_state_dict(self, sd) -> None:
        groups = sd["param_groups"]
        if len(groups) != 1:
            raise ValueError("ArenaAdam.load_state_dict: expected the reference's single parameter group")
        params = self.model.parameters()
        ids = groups[0]["params"]
        if len(ids) != len(params):
            raise ValueError(f"ArenaAdam.load_state_dict: checkpoint lists {len(ids)} parameters, the model has {len(params)}")
        where = self._param_slots()
        for m, v in zip(self._m, self._v):
            m.zero_(); v.zero_()
        steps = []
        done = set()
        for i, p in zip(ids, params):
            st = sd["state"].get(i, sd["state"].get(str(i)))
            if st is None or id(p) in done:
                continue
            done.add(id(p))
            k, off, n = where[id(p)]
            self._m[k][off:off + n].copy_(st["exp_avg"].reshape(-1).to(self._m[k].device, torch.float32))
            self._v[k][off:off + n].copy_(st["exp_avg_sq"].reshape(-1).to(self._v[k].device, torch.float32))
            steps.append(float(st["step"]))
        # the reference's Adam steps its duplicated VF entries twice per iteration: the iteration count is the smallest
        # per-parameter counter (colour net / density), which is what the single counter here means
        self._step.fill_(min([x for x in steps if x > 0], default=0.0))
        g0 = self.param_groups[0]
        lr = groups[0]["lr"]
        lr = float(lr.item()) if isinstance(lr, torch.Tensor) else float(lr)
        if isinstance(g0["lr"], torch.Tensor):
            g0["lr"].fill_(lr)
        else:
            g0["lr"] = torch.tensor(lr, device=self._step.device)
        for key in ("betas", "eps", "weight_decay", "initial_lr"):
            if key in groups[0]:
                g0[key] = tuple(groups[0][key]) if key == "betas" else groups[0][key]

    def grad_norm(self) -> torch.Tensor:
        """Global 2-norm of the gradient as clip_grad_norm_ returns it (after step(): of the unclipped gradient).
        Reproducible run to run: csrc/optim.cu sums per-block partials in a fixed order."""
        return self._sq.sqrt()

    @torch.no_grad()
    def step(self, closure=None, max_norm: Optional[float] = None):
        L = _lib.lib()
        self._check_storage()
        g0 = self.param_groups[0]
        lr = g0["lr"]
        if not isinstance(lr, torch.Tensor):                 # a scheduler replaced the tensor by a float
            lr = g0["lr"] = torch.tensor(float(lr), device=self._step.device)
        lr = lr.reshape(1).float()
        b1, b2 = g0["betas"]
        mn = self.max_norm if max_norm is None else max_norm
        mn = 0.0 if mn is None else float(mn)
        stream = torch.cuda.current_stream(self._step.device).cuda_stream
        self._step.add_(1.0)
        gs = float(self.grad_scale)
        if mn > 0:
            # running statistics and padding hold zeros in the gradient tensor: the norm over all of it is the parameters'
            self._sq.zero_()
            _lib.check(L.vfnerf_sqnorm_accumulate(self.grad_all.data_ptr(), self.grad_all.numel(), gs, self._sq.data_ptr(),
                                                  self._scratch.data_ptr(), stream), "vfnerf_sqnorm_accumulate")
        for p, g, m, v, mask in zip(self._flats, self._grads, self._m, self._v, self._masks):
            _lib.check(L.vfnerf_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _lib.ptr(mask), p.numel(),
                                          lr.data_ptr(), self._step.data_ptr(), float(b1), float(b2), float(g0["eps"]),
                                          float(g0["weight_decay"]), mn, self._sq.data_ptr(), gs, stream), "vfnerf_adam_step")
        for ar in self._arenas:
            ar.generation += 1         # the arenas changed behind torch's version counters
        return None


def use_arena_optimizer(model, max_norm: Optional[float] = None) -> ArenaAdam:
    """Replace model.optimizer / model.scheduler by ArenaAdam + the same ExponentialLR decay."""
    sc = model.config.scheduler_config
    old = model.optimizer.param_groups[0]
    lr = old["lr"]
    lr = float(lr.item()) if isinstance(lr, torch.Tensor) else float(lr)
    opt = ArenaAdam(model, lr=lr, weight_decay=old.get("weight_decay", 0.0), max_norm=max_norm)
    gamma = getattr(model.scheduler, "gamma", sc.lr_decay_factor ** (1. / sc.lr_decay_steps))
    model.optimizer = opt
    model.scheduler = torch.optim.lr_scheduler.ExponentialLR(opt, gamma)
    return opt
