"""Parameter owners of the hot path: the VF MLP, the colour MLP and the Laplace density.

The classes keep the reference's module structure so ``state_dict()`` has the reference's keys
(SURVEY.md §5 "Checkpoint / resume"): ``layers.{i}.0.{weight,bias}`` + ``layers.{i}.1.{weight,bias,
running_mean,running_var,num_batches_tracked}`` for hidden layers and ``layers.{n-1}.{weight,bias}``
for the last one; the density holds ``beta / scale / mean``.  They are *containers*: all arithmetic
runs in the CUDA library.  Every fp32 parameter and running statistic of a network is a view into
one flat "arena" tensor whose layout is described to the kernels by a ``vfnerf_mlp_desc``; the
gradient arena written by the backward has the same layout.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import _lib


class ParamArena:
    """Flat fp32 storage behind a module's parameters / running stats + its vfnerf_mlp_desc."""

    def __init__(self, module: nn.Module, layers: nn.ModuleList) -> None:
        self.module = module
        self.layers = layers
        self.flat: Optional[torch.Tensor] = None
        self.desc = _lib.MlpDesc()
        self.slots: List[Tuple[str, torch.Tensor, int]] = []   # (kind, tensor, offset)
        self.grad_flat: Optional[torch.Tensor] = None           # flat gradient arena (optim.ArenaAdam), else None
        self.dirty = False                                      # set by the owner's _apply / load_state_dict overrides
        self.generation = 0         # bumped by code that writes the arena behind torch's back (ArenaAdam, train-mode BN)
        self.rebuild()

    def _tensors(self):
        """(layer, kind, tensor) for every fp32 tensor, in a fixed order."""
        for i, layer in enumerate(self.layers):
            if isinstance(layer, nn.Sequential):
                lin, bn = layer[0], layer[1]
            else:
                lin, bn = layer, None
            yield i, "w", lin.weight
            yield i, "b", lin.bias
            if bn is not None:
                yield i, "gamma", bn.weight
                yield i, "beta", bn.bias
                yield i, "mean", bn.running_mean
                yield i, "var", bn.running_var

    def rebuild(self) -> None:
        tens = list(self._tensors())
        device = tens[0][2].device
        total = sum(t.numel() for _, _, t in tens)
        flat = torch.empty(total, dtype=torch.float32, device=device)
        d = self.desc
        d.n_layers = len(self.layers)
        for arr in (d.w_off, d.b_off, d.gamma_off, d.beta_off, d.mean_off, d.var_off):
            for i in range(_lib.MAX_LAYERS):
                arr[i] = -1
        off = 0
        self.slots = []
        with torch.no_grad():
            for i, kind, t in tens:
                n = t.numel()
                view = flat[off:off + n].view(t.shape)
                view.copy_(t.detach().to(torch.float32))
                t.data = view                       # the Parameter / buffer now aliases the arena
                getattr(d, f"{kind}_off")[i] = off
                if kind == "w":
                    d.out_dim[i], d.in_dim[i] = t.shape
                self.slots.append((kind, t, off))
                off += n
        d.arena_floats = total
        self.flat = flat
        if self.grad_flat is not None:          # storage was replaced (.to(), load_state_dict(assign=True)): follow it
            self.grad_flat = None
            self.enable_flat_grad()

    def enable_flat_grad(self, storage: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Give the gradients the arena's layout too: one zero-initialised flat tensor, every Parameter's ``.grad`` a
        view of it.  The backward of ops.render_call / ops.vf_query then accumulates straight into it (one add per
        network instead of one per parameter) and optim.ArenaAdam updates the whole network with two launches.
        ``storage``: a [arena_floats] slice of a larger buffer to use (optim.ArenaAdam keeps the gradients of both
        networks and the density in ONE tensor, so clipping reads it once and multi-GPU training all-reduces it once)."""
        if storage is not None:
            if storage.numel() != self.flat.numel() or storage.device != self.flat.device or storage.dtype != torch.float32:
                raise ValueError("enable_flat_grad: storage must be a float32 [arena_floats] tensor on the arena's device")
            self.grad_flat = storage
        elif self.grad_flat is None:
            self.grad_flat = torch.zeros_like(self.flat)
        for _, t, off in self.slots:
            if isinstance(t, nn.Parameter):
                t.grad = self.grad_flat[off:off + t.numel()].view(t.shape)
        return self.grad_flat

    def disable_flat_grad(self) -> None:
        """Back to per-parameter gradients (a torch optimizer took over: its zero_grad(set_to_none=True) would drop the
        views and the flat accumulation of ops.render_call would then feed nobody)."""
        self.grad_flat = None
        for _, t, _ in self.slots:
            if isinstance(t, nn.Parameter):
                t.grad = None

    def trainable_mask(self) -> torch.Tensor:
        """uint8 [arena_floats]: 1 on Parameter elements, 0 on BatchNorm running statistics."""
        m = torch.zeros(self.flat.numel(), dtype=torch.uint8, device=self.flat.device)
        for _, t, off in self.slots:
            if isinstance(t, nn.Parameter):
                m[off:off + t.numel()] = 1
        return m

    def sync(self) -> torch.Tensor:
        """Arena tensor, re-flattened first if someone replaced parameter / buffer storage (``.to()``,
        ``load_state_dict(assign=True)``, ...).  Looks the tensors up afresh: ``Module._apply`` swaps
        buffer objects."""
        base = self.flat.data_ptr()
        # fast path (this runs several times per render() call): storage can only move through Module._apply
        # (.to(), .cuda(), .float() ...) or load_state_dict(assign=True), both of which the owning module reports by
        # setting `dirty`; a spot check of the first and last slot catches manual `.data` swaps
        if not self.dirty:
            t0, tl, ol = self.slots[0][1], self.slots[-1][1], self.slots[-1][2]
            if t0.data_ptr() == base and tl.data_ptr() == base + 4 * ol:
                return self.flat
        self.dirty = False
        off = 0
        for _, _, t in self._tensors():
            if t.data_ptr() != base + 4 * off or t.dtype != torch.float32:
                self.rebuild()
                break
            off += t.numel()
        return self.flat

    def params(self) -> List[nn.Parameter]:
        return [t for _, t, _ in self.slots if isinstance(t, nn.Parameter)]

    def version(self):
        """Changes whenever the arena's contents may have: torch's version counters of every member tensor (in-place
        optimizer steps, load_state_dict, broadcasts), the storage address (re-flattening) and ``generation`` (writes by
        the library itself).  Writes through ``tensor.data`` are invisible to it.  Used to skip re-packing the weight images
        between CUDA-graph replays of render() (nerf._RenderGraph)."""
        self.sync()
        return (self.flat.data_ptr(), self.generation, self.flat._version, sum(t._version for _, t, _ in self.slots))

    def grad_views(self, grad_flat: torch.Tensor) -> List[torch.Tensor]:
        """Views of a gradient arena matching ``params()`` one to one."""
        return [grad_flat[off:off + t.numel()].view(t.shape)
                for _, t, off in self.slots if isinstance(t, nn.Parameter)]


class _ArenaOwner:
    """Mixin of the two network modules: tells the arena when parameter / buffer storage may have been replaced."""

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if getattr(self, "_arena", None) is not None:
            self._arena.dirty = True
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        if getattr(self, "_arena", None) is not None:
            self._arena.dirty = True
        return out


def _make_layers(dims: Sequence[Tuple[int, int]], batch_norm: bool, weight_norm: bool, xavier: bool,
                 bias_init: float) -> nn.ModuleList:
    if weight_norm:
        raise NotImplementedError("weight_norm=True is not on the accelerated path (the shipped config uses "
                                  "batch_norm, confs/vf_nerf.conf:19-20,32-33)")
    layers = nn.ModuleList()
    n = len(dims)
    for i, (fi, fo) in enumerate(dims):
        lin = nn.Linear(fi, fo)
        if xavier:
            nn.init.xavier_uniform_(lin.weight)
            nn.init.constant_(lin.bias, bias_init)
        if batch_norm and i < n - 1:
            layers.append(nn.Sequential(lin, nn.BatchNorm1d(fo)))
        else:
            layers.append(lin)
    return layers


class VectorFieldNetwork(_ArenaOwner, nn.Module):
    """Drop-in for models/vector_field/vector_field_network.py:14-208 (eval-mode forward).

    ``forward(points[P,3]) -> [P, 3 + feature_vector_dims]`` = tanh([v, feat]); in train() mode (BatchNorm batch
    statistics, :140-175) ``[P, 3 + feature_vector_dims + 9]`` with the flat autograd "Jacobian" appended
    (csrc/mlp_train.cu, fp32 layer-wise path).
    """

    def __init__(self, config) -> None:
        super().__init__()
        self.config = config
        from .synthetic import mlp_layer_dims
        self.multires = int(config.embedder_multires)
        if self.multires <= 0:
            raise NotImplementedError("embedder_multires must be > 0 on the accelerated path")
        emb = config.input_dims + 2 * config.input_dims * self.multires
        self.skip_connection_in = list(config.skip_connection_in or [])
        if len(self.skip_connection_in) > 1:
            raise NotImplementedError("at most one skip connection is supported")
        dims = mlp_layer_dims(emb, list(config.dimensions), config.output_dims + config.feature_vector_dims,
                              self.skip_connection_in, emb)
        self.num_layers = len(dims)
        self.layers = _make_layers(dims, config.batch_norm, config.weight_norm, config.xavier_init, config.bias_init)
        if not config.batch_norm:
            raise NotImplementedError("batch_norm=False nets are not on the accelerated path")
        self.precision = "fp32"
        self._arena: Optional[ParamArena] = None

    # reference API (vector_field_network.py:94-138)
    @property
    def init(self) -> str:
        return getattr(self.config, "init", "")

    @init.setter
    def init(self, value: str) -> None:
        self.config.init = value

    def load_init(self, init_path: str, device: torch.device = torch.device("cpu")) -> None:
        self.load_state_dict(torch.load(init_path, map_location=torch.device("cpu")))
        self.to(device)

    @property
    def skip_layer(self) -> int:
        return self.skip_connection_in[0] if self.skip_connection_in else -1

    def arena(self) -> ParamArena:
        if self._arena is None:
            self._arena = ParamArena(self, self.layers)
        self._arena.sync()
        return self._arena

    def get_outputs(self, input_tensor: torch.Tensor):
        out = self.forward(input_tensor)
        return out[:, :3], out[:, 3:]

    def forward(self, points: torch.Tensor) -> torch.Tensor:
        from .ops import vf_query, vf_query_train
        if self.training:
            # BatchNorm batch statistics + the autograd "Jacobian" (vector_field_network.py:140-175): [P, 3 + feat + 9]
            return vf_query_train(self, points)
        return vf_query(self, points)


class RenderingNetwork(_ArenaOwner, nn.Module):
    """Parameter owner for models/vector_field/rendering_network.py:13-108 (mode 'idr')."""

    def __init__(self, config) -> None:
        super().__init__()
        self.config = config
        if config.mode != "idr":
            raise NotImplementedError(f"rendering mode {config.mode!r}: only 'idr' is on the accelerated path")
        self.multires_view = int(config.embedder_multires)
        in_dim = 3 + (3 + 6 * self.multires_view) + 3 + config.feature_vector_dims
        dims = [in_dim] + list(config.dimensions) + [config.output_dims]
        pairs = [(dims[i], dims[i + 1]) for i in range(len(dims) - 1)]
        self.num_layers = len(pairs)
        self.layers = _make_layers(pairs, config.batch_norm, config.weight_norm, False, 0.0)
        self.mode = config.mode
        self._arena: Optional[ParamArena] = None

    def arena(self) -> ParamArena:
        if self._arena is None:
            self._arena = ParamArena(self, self.layers)
        self._arena.sync()
        return self._arena

    def forward(self, points, normals, view_dirs, feature_vectors):
        raise NotImplementedError("the colour MLP is evaluated inside VectorFieldNerf.render(); it has no "
                                  "stand-alone caller in the reference (SURVEY.md §1)")


class LaplaceDensity(nn.Module):
    """Parameter owner for models/helpers/density_functions.py:111-204.  The clamps of
    get_beta/get_scale/get_mean are applied inside the kernels; the getters exist because the
    trainer logs them (train/vector_field_nerf_train.py:286-288)."""

    def __init__(self, params_init: Dict[str, float], beta_bounds=(1e-6, 0.0006), scale_min: float = 1.0,
                 mean_bounds=(0.5, 1.0)) -> None:
        super().__init__()
        for k in ("beta", "scale", "mean"):
            if k not in params_init:
                raise NotImplementedError(f"density parameter {k!r} missing: the accelerated path expects "
                                          "beta, scale and mean (confs/vf_nerf.conf:1-11)")
        for name in params_init:              # registration order follows the dict, like the reference
            setattr(self, name, nn.Parameter(torch.tensor(float(params_init[name]))))
        self.beta_bounds = torch.tensor([float(b) for b in beta_bounds])
        self.scale_min = torch.tensor(float(scale_min))
        self.mean_bounds = torch.tensor([float(b) for b in mean_bounds])
        self._flat: Optional[torch.Tensor] = None

    def flat(self) -> torch.Tensor:
        """[beta, scale, mean] contiguous on the parameters' device; the Parameters alias it."""
        ps = (self.beta, self.scale, self.mean)
        f = self._flat
        if f is None or any(p.data_ptr() != f.data_ptr() + 4 * i for i, p in enumerate(ps)):
            f = torch.empty(3, dtype=torch.float32, device=self.beta.device)
            with torch.no_grad():
                for i, p in enumerate(ps):
                    f[i] = p.detach()
                    p.data = f[i]
            self._flat = f
            if getattr(self, "grad_flat", None) is not None:
                self.grad_flat = None
                self.enable_flat_grad()
        return f

    def enable_flat_grad(self, storage: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[d beta, d scale, d mean] as one tensor; the three Parameters' ``.grad`` are views of it (optim.ArenaAdam)."""
        self.flat()
        if storage is not None:
            self.grad_flat = storage
        elif getattr(self, "grad_flat", None) is None:
            self.grad_flat = torch.zeros(3, dtype=torch.float32, device=self.beta.device)
        for i, p in enumerate((self.beta, self.scale, self.mean)):
            p.grad = self.grad_flat[i]
        return self.grad_flat

    def disable_flat_grad(self) -> None:
        self.grad_flat = None
        for p in (self.beta, self.scale, self.mean):
            p.grad = None

    def get_beta(self) -> torch.Tensor:
        return torch.clamp(self.beta, float(self.beta_bounds[0]), float(self.beta_bounds[1]))

    def set_beta(self, beta: torch.Tensor) -> None:
        self.beta.data.copy_(beta)

    def get_scale(self) -> torch.Tensor:
        return torch.max(self.scale.abs(), self.scale_min.to(self.scale.device))

    def get_mean(self) -> torch.Tensor:
        return torch.clamp(self.mean, float(self.mean_bounds[0]), float(self.mean_bounds[1]))
