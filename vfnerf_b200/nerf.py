"""VectorFieldNerf -- the drop-in boundary.

Same constructor argument, attributes and methods as the reference facade
(models/nerf/vector_field_nerf.py:23-338): ``render``, ``parameters``, ``train/eval``, ``load/save``,
``to/cpu``, ``new_scheduler/reset_scheduler`` and the attributes callers touch (``ray_sampler``,
``fine_sampler``, ``vector_field_network``, ``fine_vector_field_network``, ``rendering_network``,
``density``, ``optimizer``, ``scheduler``, ``config``).  ``render()`` runs entirely in
libvfnerf_b200.so; this file only marshals arguments.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch

from . import _lib, ops
from .networks import LaplaceDensity, RenderingNetwork, VectorFieldNetwork
from .output import NerfOutput
from .samplers import RangeFineSampler, UniformSampler, cpu_generator_rand_


def c_is_nerf(config) -> bool:
    return getattr(config, "rendering", "volsdf") == "nerf"


class VectorFieldNerf:
    def __init__(self, config, precision: str = "fp32") -> None:
        """:param config: a VFNerfConfig (ours or the reference's own dataclass; only attributes are read).
        :param precision: "fp32" | "bf16" | "bf16x3" | "fp16f8" -- arithmetic of the two MLPs (include/vfnerf_b200.h)."""
        self.config = config
        rs = config.ray_sampler_config
        self.vector_field_network = VectorFieldNetwork(config.vf_net_config)
        if rs.fine_sampling():
            self.fine_vector_field_network = self.vector_field_network     # shared, vector_field_nerf.py:36
        self.rendering_network = RenderingNetwork(config.rendering_net_config)
        self.ray_sampler = UniformSampler(rs.n_samples, rs.near, rs.far, (not rs.perturb))
        if rs.fine_sampling():
            self.fine_sampler = RangeFineSampler(rs.n_importance, rs.near, rs.far, (not rs.perturb),
                                                 range=rs.fine_range, max_samples=rs.max_samples)
        self.density = LaplaceDensity(**config.density_config.todict())
        self.set_precision(precision)

        sc = config.scheduler_config
        self.optimizer = torch.optim.Adam(self.parameters(), lr=sc.lr, weight_decay=sc.weight_decay)
        self.scheduler = torch.optim.lr_scheduler.ExponentialLR(
            self.optimizer, sc.lr_decay_factor ** (1. / sc.lr_decay_steps))
        if getattr(config.cuda_config, "num_gpus", 1) > 1:
            # The reference wraps the nets in nn.DataParallel here (:70-75), a path that cannot run as
            # shipped.  Multi-GPU is one process per GPU: see vfnerf_b200.dist.
            pass
        self.to(config.cuda_config.device)
        self.return_ray_dirs = True
        # True: evaluate the VF MLP on the coarse points a second time among the merged points, literally like
        # vector_field_nerf.py:289-312; False (default): reuse the coarse sweep's results on the bf16 forward-only
        # path (bit-identical output, 25 % fewer FLOPs; include/vfnerf_b200.h VFNERF_FLAG_RECOMPUTE_COARSE)
        self.recompute_coarse = False
        # False (default): the uniform draws come from the global CPU generator in the reference's order and are copied to
        # the device, so torch.manual_seed reproduces the reference's sample positions bit for bit.  True: draw them on
        # the device (same distribution, a different random stream) -- saves ~100 us of host time per 1024-ray call.
        self.draws_on_device = False
        self._stage: dict = {}          # pinned staging rings of the CPU-generator draws, per device
        self.last_extras: dict = {}
        # Opt-in for chunked evaluation loops (evaluation/methods.py:510-530 calls render() once per 1024 rays): under
        # torch.no_grad() the whole forward is captured once per (ray count, sampler settings, precision) as a CUDA
        # graph and replayed -- per call the host then only copies the inputs into static buffers.  The NerfOutput of a
        # replayed call aliases static buffers that the NEXT render() call overwrites (the evaluation loop copies
        # rgb / depth to the host right away); off by default for that reason.  The weight images are re-tiled only when
        # the parameters changed since the previous replay, as seen by torch's version counters, load_state_dict and the
        # library's own writers (ParamArena.version); after writing parameters through ``tensor.data`` call
        # ``model.invalidate_replay()``.
        self.graph_replay = False
        self._graphs: dict = {}
        # Two paths of the reference's render() cannot execute upstream: white=True reads rgb before it is
        # assigned (vector_field_nerf.py:273-277) and rendering="nerf" passes nerf_volume_rendering its arguments swapped
        # (:271,312 vs utils/rendering.py:98).  By default both are rejected like upstream behaves; enable_reference_fix()
        # opts into their evident intent (include/vfnerf_b200.h: VFNERF_FLAG_WHITE_BG / VFNERF_FLAG_NERF_WEIGHTS).
        self.reference_fixes: set = set()

    # ---- module plumbing (vector_field_nerf.py:84-214) ------------------------------------------
    def set_precision(self, precision: str) -> None:
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}, got {precision!r}")
        self.precision = precision
        self.vector_field_network.precision = precision

    def _modules(self):
        return (self.vector_field_network, self.rendering_network, self.density)

    def cpu(self) -> None:
        for m in self._modules():
            m.cpu()

    def to(self, device: torch.device) -> None:
        for m in self._modules():
            m.to(device)

    def _new_schedule(self, num_steps: int) -> None:
        sc = self.config.scheduler_config
        self.scheduler = torch.optim.lr_scheduler.ExponentialLR(self.optimizer, sc.lr_decay_factor ** (1. / num_steps))
        # a plain torch optimizer takes over: leave flat-gradient mode (optim.ArenaAdam), or backward() would keep
        # accumulating into arenas nobody reads while p.grad stays None
        self.vector_field_network.arena().disable_flat_grad()
        self.rendering_network.arena().disable_flat_grad()
        self.density.disable_flat_grad()
        self.optimizer = torch.optim.Adam(self.parameters(), lr=sc.lr)

    def new_scheduler(self, num_steps: int) -> None:
        self._new_schedule(num_steps)

    def reset_scheduler(self, num_steps: Optional[int] = None) -> None:
        self._new_schedule(self.config.scheduler_config.lr_decay_steps if num_steps is None else num_steps)

    def parameters(self) -> List[torch.nn.Parameter]:
        # the VF parameters appear twice when fine sampling is on, exactly as in the reference (:132-137):
        # clip_grad_norm_ and Adam then see the duplicate list the reference trainer sees.
        params = list(self.vector_field_network.parameters()) + list(self.rendering_network.parameters()) + \
            list(self.density.parameters())
        if self.config.ray_sampler_config.fine_sampling():
            params += list(self.fine_vector_field_network.parameters())
        return params

    def train(self) -> None:
        if self.config.numerical_jacobian:
            self.vector_field_network.eval()
        else:
            self.vector_field_network.train()
        self.rendering_network.train()
        self.density.train()

    def eval(self) -> None:
        for m in self._modules():
            m.eval()

    def load(self, path: str) -> int:
        ckpt = torch.load(path, map_location=self.config.cuda_config.device)
        self.vector_field_network.load_state_dict(ckpt["vf_net"])
        self.rendering_network.load_state_dict(ckpt["rendering_net"])
        self.density.load_state_dict(ckpt["density"])
        epoch = ckpt["epoch"] + 1
        self.optimizer.load_state_dict(ckpt["optimizer"])
        self.scheduler.load_state_dict(ckpt["scheduler"])
        if self.config.ray_sampler_config.fine_sampling() and "fine_vf_net" in ckpt:
            self.fine_vector_field_network.load_state_dict(ckpt["fine_vf_net"])
        return epoch

    def save(self, epoch: int, path: str) -> None:
        state = {"vf_net": self.vector_field_network.state_dict(),
                 "rendering_net": self.rendering_network.state_dict(),
                 "density": self.density.state_dict(), "epoch": epoch,
                 "optimizer": self.optimizer.state_dict(), "scheduler": self.scheduler.state_dict()}
        if self.config.ray_sampler_config.fine_sampling():
            state["fine_vf_net"] = self.fine_vector_field_network.state_dict()
        torch.save(state, os.path.join(path, f"{epoch}.pth"))
        torch.save(state, os.path.join(path, "latest.pth"))

    # ---- host draws -> device ---------------------------------------------------------------------
    def _upload_draw(self, rows: int, cols: int, dev: torch.device) -> torch.Tensor:
        """``torch.rand([rows, cols])`` on the global CPU generator (the reference's stream, ray_sampler.py:138,292,297)
        made directly in pinned memory and copied with one DMA.  A pageable tensor would be staged by the driver at a
        few GB/s on the compute stream (16.8 MB per 65 536-ray chunk).  The DMA runs on a dedicated upload stream, so it
        overlaps the previous call's kernels; ring of four buffers per device, a buffer is reused only after the copy that
        read it has completed (event)."""
        if rows * cols < (1 << 18):
            # small draws (the reference's 1024-ray chunks: 256 KB): the plain pageable copy is cheaper than stream / event
            # bookkeeping on the host, which is what bounds calls of that size
            return torch.rand([rows, cols]).to(dev, non_blocking=True)
        ring = self._stage.get(str(dev))
        if ring is None:
            ring = self._stage[str(dev)] = {"bufs": [], "next": 0, "stream": torch.cuda.Stream(device=dev)}
        n = rows * cols
        if len(ring["bufs"]) < 4:
            ring["bufs"].append([torch.empty(max(n, 1 << 16), dtype=torch.float32).pin_memory(), None])
            slot = ring["bufs"][-1]
        else:
            slot = ring["bufs"][ring["next"]]
            ring["next"] = (ring["next"] + 1) % 4
        if slot[1] is not None:
            slot[1].synchronize()          # the copy runs on its own stream: done long before the buffer comes round again
        if slot[0].numel() < n:
            slot[0] = torch.empty(n, dtype=torch.float32).pin_memory()
        host = slot[0][:n].view(rows, cols)
        cpu_generator_rand_(host)      # == torch.rand([rows, cols], out=host), same stream and state (samplers.py)
        main = torch.cuda.current_stream(dev)
        with torch.cuda.stream(ring["stream"]):
            out = host.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(ring["stream"])
        main.wait_event(ev)                # render()'s kernels are stream-ordered after the upload
        out.record_stream(main)
        slot[1] = ev
        return out

    # ---- the hot path ---------------------------------------------------------------------------
    def enable_reference_fix(self, *names: str) -> None:
        """Opt into paths that are broken upstream, with their evident intent: "white_background" (render(white=True):
        rgb += 1 - sum of weights after the final composite) and/or "nerf_rendering" (config.rendering == "nerf":
        nerf_volume_rendering with its arguments in the function's own order)."""
        for n in names:
            if n not in ("white_background", "nerf_rendering"):
                raise ValueError(f"unknown fix {n!r}: expected 'white_background' or 'nerf_rendering'")
            self.reference_fixes.add(n)

    def _render_cfg(self, n_rays: int, pose_is_quat: bool, white: bool = False) -> _lib.RenderCfg:
        c, d = self.config, self.density
        if not c.ray_sampler_config.fine_sampling():
            # the reference dies with UnboundLocalError at vector_field_nerf.py:331 in this configuration
            raise UnboundLocalError("render() needs fine sampling (n_importance > 0), like the reference")
        if c.rendering not in ("volsdf", "nerf"):
            raise ValueError(f"rendering must be 'volsdf' or 'nerf', got {c.rendering!r}")
        if c.rendering == "nerf" and "nerf_rendering" not in self.reference_fixes:
            raise NotImplementedError('rendering="nerf" passes its arguments swapped in the reference (SURVEY.md §8a) and '
                                      "renders garbage there; call model.enable_reference_fix('nerf_rendering') for the "
                                      "corrected argument order")
        near, far = self.ray_sampler.near, self.ray_sampler.far
        if not isinstance(far, (int, float)) or not isinstance(near, (int, float)):
            near, far = float(near), float(far)

        cfg = _lib.RenderCfg()
        cfg.n_rays, cfg.n_coarse, cfg.n_fine = n_rays, self.ray_sampler.N_samples, self.fine_sampler.n_fine()
        cfg.perturb = 0 if self.ray_sampler.deterministic else 1
        cfg.pose_is_quat = 1 if pose_is_quat else 0
        cfg.window = int(len(c.cos_sim_weights))
        cfg.normalize = 1 if c.normalize_rendering else 0
        cfg.multires = self.vector_field_network.multires
        cfg.multires_view = self.rendering_network.multires_view
        cfg.skip_layer = self.vector_field_network.skip_layer
        cfg.precision = _lib.PRECISIONS[self.precision]
        cfg.flags = (_lib.FLAG_RECOMPUTE_COARSE if self.recompute_coarse else 0) | \
                    (_lib.FLAG_WHITE_BG if white else 0) | (_lib.FLAG_NERF_WEIGHTS if c.rendering == "nerf" else 0)
        cfg.near_, cfg.far_, cfg.fine_range = float(near), float(far), float(self.fine_sampler.range)
        # the fine sampler draws its fallback samples in ITS OWN [near, far] (ray_sampler.py:296-299)
        cfg.fine_near_, cfg.fine_far_ = float(self.fine_sampler.near), float(self.fine_sampler.far)
        cfg.dir_to_normal_th = float(c.dir_to_normal_th)
        cfg.beta_lo, cfg.beta_hi = float(d.beta_bounds[0]), float(d.beta_bounds[1])
        cfg.mean_lo, cfg.mean_hi = float(d.mean_bounds[0]), float(d.mean_bounds[1])
        cfg.scale_min = float(d.scale_min)
        cfg.bn_eps = 1e-5
        return cfg

    def render(self, pose: torch.Tensor, pixels: torch.Tensor, intrinsics: torch.Tensor, epoch: int,
               white: bool = False, *, draws=None, z_vals_override: Optional[torch.Tensor] = None,
               _workspace: Optional[torch.Tensor] = None, _weights_packed: bool = False) -> NerfOutput:
        """Same contract as vector_field_nerf.py:216-338.

        :param pose: [R,4,4] camera-to-world matrices or [R,7] (quaternion, translation), on the GPU.
        :param pixels: [R,2] (u, v).   :param intrinsics: [R,4,4].   :param epoch: unused on this path (the
            annealed cos_sim_weights are only read for their length).
        :param draws: optional (U1, U2, U3) uniform tensors replacing the CPU-generator draws (tests, and
            multi-GPU ray sharding where each rank takes its slice of the global draws).
        :param z_vals_override: optional [R,N] merged z values for the second pass (parity protocol).
        """
        if white and "white_background" not in self.reference_fixes:
            # vector_field_nerf.py:274-277 reads rgb_values_coarse before assignment
            raise UnboundLocalError("white=True is broken in the reference (SURVEY.md fact 2); call "
                                    "model.enable_reference_fix('white_background') for rgb += 1 - sum(weights)")
        train_mode = self.vector_field_network.training and self.rendering_network.training
        if self.vector_field_network.training != self.rendering_network.training or \
                (train_mode and getattr(self.config, "numerical_jacobian", False)):
            # config.numerical_jacobian: train() leaves the VF net in eval but switches the colour net to batch statistics
            # and asks for finite-difference directional derivatives (vector_field_nerf.py:84-101,262-263)
            raise NotImplementedError("render() with only one network in train() mode / numerical_jacobian=True is not "
                                      "built; model.train() (both networks, autograd Jacobian) and model.eval() are")
        if train_mode:
            if self.precision != "fp32":
                raise NotImplementedError("train() mode (batch-statistic BatchNorm, Jacobian, directional derivatives) "
                                          "runs on the fp32 layer-wise path: use precision='fp32', or model.eval() like "
                                          "the reference trainer does when the directional-derivative weight is 0 "
                                          "(train/vector_field_nerf_train.py:140-141)")
            if white or c_is_nerf(self.config):
                raise NotImplementedError("the white-background / nerf-rendering fixes are eval-mode options")
        if not getattr(self.config.rendering_net_config, "detach_normals", True):
            raise NotImplementedError("rendering_net_config.detach_normals=False: the backward kernels implement the shipped "
                                      "detach (rendering_network.py:76-77); gradients through the colour net's normal "
                                      "input are not built")
        pixels = ops._require_cuda("pixels", pixels)
        pose = ops._require_cuda("pose", pose)
        intrinsics = ops._require_cuda("intrinsics", intrinsics)
        dev = pixels.device
        R = pixels.shape[0]
        if pose.shape[0] != R or intrinsics.shape[0] != R:
            raise ValueError("pose, pixels and intrinsics must have one row per ray")
        quat = pose.dim() == 2 and pose.shape[1] == 7
        if self.graph_replay and z_vals_override is None and R > 0 and not white and not torch.is_grad_enabled() \
                and not train_mode:
            return self._render_replay(pose, pixels, intrinsics, quat, draws)
        cfg = self._render_cfg(R, quat, white)
        # host-side draws in the reference's order (ray_sampler.py:138, 292, 297), then H2D
        if draws is None and self.draws_on_device:
            nf = self.fine_sampler.n_fine()
            U1 = None if self.ray_sampler.deterministic else torch.rand(R, self.ray_sampler.N_samples, device=dev)
            U2 = None if self.fine_sampler.deterministic else torch.rand(R, nf, device=dev)
            U3 = torch.rand(R, nf, device=dev)
        elif draws is None:
            # same generator, same order and sizes as ray_sampler.draw / fine_sampler.draw (U1 -> U2 -> U3)
            nf = self.fine_sampler.n_fine()
            U1 = None if self.ray_sampler.deterministic else self._upload_draw(R, self.ray_sampler.N_samples, dev)
            U2 = None if self.fine_sampler.deterministic else self._upload_draw(R, nf, dev)
            U3 = self._upload_draw(R, nf, dev)
        else:
            U1, U2, U3 = draws
        # host linspace (bit-exact with the reference), uploaded once per (device, count): keeps render() free of
        # host->device copies when the draws are given on the device (CUDA-graph capture, graphed.py)
        key = (str(dev), self.ray_sampler.N_samples)
        if getattr(self, "_t_vals_key", None) != key:
            self._t_vals_dev = self.ray_sampler.t_vals().to(dev).float().contiguous()
            self._t_vals_key = key
        t_vals = self._t_vals_dev

        def dev_(t):
            return None if t is None else t.to(dev, non_blocking=True).float().contiguous()
        call = ops.RenderCall(cfg, self.vector_field_network, self.rendering_network, self.density,
                              pixels, pose, intrinsics, dev_(t_vals), dev_(U1), dev_(U2), dev_(U3),
                              z_override=dev_(z_vals_override), want_extras=True,
                              want_ray_dirs=self.return_ray_dirs, train=train_mode, workspace=_workspace,
                              weights_packed=_weights_packed)
        res = ops.render_call(call)
        rgb, depth, normals, colors, points, z_vals = res[:6]
        ray_dirs = res[6] if len(res) > 6 else None
        self.last_extras = call.extras
        return NerfOutput(points_coarse=points, points_fine=None, coarse_normals=normals,
                          coarse_rgb_values=rgb, coarse_depth_map=depth, fine_normals=None,
                          fine_rgb_values=None, fine_depth_map=None, z_vals=z_vals,
                          directional_derivtives=call.extras.get("directional_derivatives"), ray_dirs=ray_dirs,
                          coarse_colors=colors,
                          weights=call.extras.get("weights"))


    def invalidate_replay(self) -> None:
        """Forget the captured render() graphs (and with them the packed weight images they reuse)."""
        self._graphs.clear()

    # ---- CUDA-graph replay of the forward-only call (opt-in: self.graph_replay) ---------------------------------------
    def _replay_key(self, R: int, quat: bool, dev: torch.device):
        rs, fs, c = self.ray_sampler, self.fine_sampler, self.config
        return (R, quat, str(dev), rs.N_samples, fs.n_fine(), float(rs.near), float(rs.far), float(fs.range),
                bool(rs.deterministic), bool(fs.deterministic), self.precision, self.recompute_coarse,
                self.return_ray_dirs, float(c.dir_to_normal_th), len(c.cos_sim_weights), bool(c.normalize_rendering),
                c.rendering, float(fs.near), float(fs.far))

    def _render_replay(self, pose, pixels, intrinsics, quat: bool, draws) -> NerfOutput:
        dev = pixels.device
        R = pixels.shape[0]
        key = self._replay_key(R, quat, dev)
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= 8:                 # a caller that keeps changing shapes should not hoard graph memory
                self._graphs.clear()
            g = self._graphs[key] = _RenderGraph(self, R, quat, dev)
        return g.replay(pose, pixels, intrinsics, draws)


class _RenderGraph:
    """One captured forward-only render() for a fixed ray count and sampler configuration.  The weight images are
    re-packed from the parameter arenas inside the graph, so parameter updates between replays are picked up."""

    def __init__(self, model: VectorFieldNerf, R: int, quat: bool, dev: torch.device) -> None:
        self.model, self.R, self.dev = model, R, dev
        f32 = dict(dtype=torch.float32, device=dev)
        nc, nf = model.ray_sampler.N_samples, model.fine_sampler.n_fine()
        self.pose = torch.zeros((R, 7) if quat else (R, 4, 4), **f32)
        if not quat:
            self.pose[:] = torch.eye(4, **f32)
        else:
            self.pose[:, 0] = 1.0
        self.pixels = torch.zeros(R, 2, **f32)
        self.intrinsics = torch.eye(4, **f32).repeat(R, 1, 1)
        self.U = [None if model.ray_sampler.deterministic else torch.rand(R, nc, **f32),
                  None if model.fine_sampler.deterministic else torch.rand(R, nf, **f32),
                  torch.rand(R, nf, **f32)]
        # pinned staging of the CPU-generator draws: ring of 4 per tensor, reused after the copy that read it completed
        self._host = [[None if u is None else [torch.empty(u.shape, dtype=torch.float32).pin_memory(), None] for _ in range(4)]
                      for u in self.U]
        self._slot = 0
        # Two captures over ONE static workspace: `graph` re-tiles the weight images from the parameter arenas and renders,
        # `graph_packed` renders with the images the workspace already holds.  replay() picks the first whenever the
        # arenas' version changed since the last re-tiling (ParamArena.version) -- hundreds of chunks of one image then
        # pay for the re-tiling once.
        L = _lib.lib()
        cfg = model._render_cfg(R, quat)
        va, ra = model.vector_field_network.arena(), model.rendering_network.arena()
        import ctypes as C
        nbytes = L.vfnerf_render_workspace_bytes(C.byref(cfg), C.byref(va.desc), C.byref(ra.desc), 0)
        if nbytes < 0:
            _lib.check(1, "vfnerf_render_workspace_bytes")
        self.ws = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=dev)
        self.graph, self.graph_packed = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        self._packed_version = None
        was = model.graph_replay
        model.graph_replay = False
        try:
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s), torch.no_grad():
                for _ in range(2):      # lazy one-time work (kernel attributes, arenas, t_vals upload) outside the capture
                    model.render(self.pose, self.pixels, self.intrinsics, 0, draws=tuple(self.U), _workspace=self.ws)
            torch.cuda.current_stream(dev).wait_stream(s)
            torch.cuda.synchronize(dev)
            with torch.cuda.graph(self.graph), torch.no_grad():
                self.out = model.render(self.pose, self.pixels, self.intrinsics, 0, draws=tuple(self.U), _workspace=self.ws)
            self.extras = model.last_extras
            self.out_packed = self.extras_packed = None
            if model.precision != "fp32":
                with torch.cuda.graph(self.graph_packed), torch.no_grad():
                    self.out_packed = model.render(self.pose, self.pixels, self.intrinsics, 0, draws=tuple(self.U),
                                                   _workspace=self.ws, _weights_packed=True)
                self.extras_packed = model.last_extras
        finally:
            model.graph_replay = was

    def replay(self, pose, pixels, intrinsics, draws) -> NerfOutput:
        m = self.model
        srcs = [pose, pixels, intrinsics]
        if all(t.is_cuda and t.dtype == torch.float32 and t.device == self.dev for t in srcs):
            # one fused launch for the three input copies (a replayed 1024-ray call is a handful of launches in total)
            torch._foreach_copy_([self.pose, self.pixels, self.intrinsics],
                                 [pose.reshape(self.pose.shape), pixels.reshape(self.pixels.shape),
                                  intrinsics.reshape(self.intrinsics.shape)], non_blocking=True)
        else:
            self.pose.copy_(pose, non_blocking=True)
            self.pixels.copy_(pixels, non_blocking=True)
            self.intrinsics.copy_(intrinsics, non_blocking=True)
        if draws is not None:
            for dst, src in zip(self.U, draws):
                if dst is not None:
                    dst.copy_(src, non_blocking=True)
        elif m.draws_on_device:
            for dst in self.U:
                if dst is not None:
                    dst.uniform_()
        else:
            # the reference's stream: global CPU generator, order U1 -> U2 -> U3 (ray_sampler.py:138,292,297)
            k = self._slot
            self._slot = (k + 1) % 4
            for dst, ring in zip(self.U, self._host):
                if dst is None:
                    continue
                host, ev = ring[k]
                if ev is not None:
                    ev.synchronize()
                cpu_generator_rand_(host)
                dst.copy_(host, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.dev))
                ring[k][1] = ev
        version = (m.vector_field_network.arena().version(), m.rendering_network.arena().version())
        if self.out_packed is not None and version == self._packed_version:
            self.graph_packed.replay()
            m.last_extras = self.extras_packed
            return self.out_packed
        self.graph.replay()
        self._packed_version = version
        m.last_extras = self.extras
        return self.out
