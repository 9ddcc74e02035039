"""torch.autograd bridges between the parameter owners (networks.py) and the C ABI (_lib.py).

PyTorch is plumbing here: it owns device memory (outputs, workspaces, arenas), the current stream and
the autograd tape.  All arithmetic happens in libvfnerf_b200.so.  Tensors handed to the library must
be CUDA fp32; host tensors are rejected (no CPU path exists).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib

_NUM_LAUNCHES = {"count": 0}     # kernels enqueued by this process (bench.py reports it)


def _require_cuda(name: str, t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: vfnerf_b200 has no CPU fallback "
                           "(move inputs to the device first, like train/vector_field_nerf_train.py:172-174)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _on_device:
    """Makes `device` the current CUDA device around a library call.  The C ABI launches on the calling thread's current
    device (kernel attributes and the SM count are cached per device ordinal on the C side), while the stream and every
    pointer belong to the tensors' device: a model on cuda:1 in a process whose current device is 0 must switch.  Free
    when the device is already current (the one-process-per-GPU layout)."""
    __slots__ = ("idx", "prev")

    def __init__(self, device: torch.device) -> None:
        self.idx = device.index if device.index is not None else torch.cuda.current_device()
        self.prev = -1

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev >= 0:
            torch.cuda.set_device(self.prev)
        return False


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------------
# VF-only query (VectorFieldNetwork.__call__)
# ------------------------------------------------------------------------------------------------
class _VFQuery(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, points, n_cols, need_bwd, *params):
        L = _lib.lib()
        ar = net.arena()
        dev = points.device
        P = points.shape[0]
        Do = ar.desc.out_dim[ar.desc.n_layers - 1]
        n_cols = Do if n_cols is None else int(n_cols)
        prec = _lib.PRECISIONS[net.precision]
        nbytes = L.vfnerf_vf_workspace_bytes(C.byref(ar.desc), P, net.multires, int(need_bwd), prec)
        if nbytes < 0:
            _lib.check(1, "vfnerf_vf_workspace_bytes")
        ws = _workspace(nbytes, dev)
        out = torch.empty(P, n_cols, dtype=torch.float32, device=dev)
        with _on_device(dev):
            _lib.check(L.vfnerf_vf_fwd(C.byref(ar.desc), ar.flat.data_ptr(), net.multires, net.skip_layer, 1e-5,
                                       prec, points.data_ptr(), P, out.data_ptr(), n_cols, n_cols,
                                       ws.data_ptr(), ws.numel(), int(need_bwd), _stream_ptr(dev)), "vfnerf_vf_fwd")
        ctx.net, ctx.ws, ctx.P, ctx.n_cols, ctx.prec = net, ws, P, n_cols, prec
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, d_out):
        L = _lib.lib()
        net = ctx.net
        ar = net.arena()
        (out,) = ctx.saved_tensors
        dev = out.device
        d_out = d_out.contiguous().float()
        grad = torch.empty(ar.desc.arena_floats, dtype=torch.float32, device=dev)
        with _on_device(dev):
            _lib.check(L.vfnerf_vf_bwd(C.byref(ar.desc), ar.flat.data_ptr(), net.multires, net.skip_layer, 1e-5,
                                       ctx.prec, ctx.P, out.data_ptr(), ctx.n_cols, d_out.data_ptr(), ctx.n_cols,
                                       ctx.n_cols, grad.data_ptr(), 0, ctx.ws.data_ptr(), ctx.ws.numel(),
                                       _stream_ptr(dev)), "vfnerf_vf_bwd")
        ctx.ws = None
        if ar.grad_flat is not None:               # flat-gradient mode (optim.ArenaAdam): one accumulate per network
            ar.grad_flat.add_(grad)
            return (None, None, None, None) + (None,) * len(ar.params())
        return (None, None, None, None) + tuple(ar.grad_views(grad))


def vf_query(net, points: torch.Tensor, n_cols: Optional[int] = None) -> torch.Tensor:
    """tanh([v, feat]) of the VF MLP at ``points[P,3]`` (vector_field_network.py:177-208).  ``n_cols=3``
    evaluates only the vector (the marching-cubes query keeps just [:, :3], mc_utils.py:99-101)."""
    if points.dim() != 2 or points.shape[1] != 3:
        raise ValueError(f"points must be [P,3], got {tuple(points.shape)}")
    points = _require_cuda("points", points.detach())
    ar = net.arena()
    if ar.flat.device != points.device:
        raise RuntimeError(f"network parameters are on {ar.flat.device}, points on {points.device}")
    params = ar.params()
    need_bwd = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    return _VFQuery.apply(net, points, n_cols, need_bwd, *params)


BN_MOMENTUM = 0.1          # nn.BatchNorm1d default, vector_field_network.py:60 / rendering_network.py:52


def _bump_batches_tracked(net, times: int = 1) -> None:
    """nn.BatchNorm1d.forward in training mode advances num_batches_tracked; the C side updates the running statistics in
    the arena, the int64 counters (not arena members) are advanced here with one foreach launch."""
    counters = [layer[1].num_batches_tracked for layer in net.layers
                if isinstance(layer, torch.nn.Sequential) and layer[1].num_batches_tracked is not None]
    if counters:
        torch._foreach_add_(counters, times)


class _VFQueryTrain(torch.autograd.Function):
    """VectorFieldNetwork.forward in train() mode (vector_field_network.py:140-175): [y, flat "Jacobian"]."""

    @staticmethod
    def forward(ctx, net, points, need_bwd, *params):
        L = _lib.lib()
        ar = net.arena()
        dev = points.device
        P = points.shape[0]
        Do = ar.desc.out_dim[ar.desc.n_layers - 1]
        if net.precision != "fp32":
            raise NotImplementedError("train() mode (batch-statistic BatchNorm + Jacobian) runs on the fp32 layer-wise path: "
                                      "build the model with precision='fp32'")
        nbytes = L.vfnerf_vf_train_workspace_bytes(C.byref(ar.desc), P, net.multires)
        if nbytes < 0:
            _lib.check(1, "vfnerf_vf_train_workspace_bytes")
        ws = _workspace(nbytes, dev)
        out = torch.empty(P, Do + 9, dtype=torch.float32, device=dev)
        with _on_device(dev):
            _lib.check(L.vfnerf_vf_train_fwd(C.byref(ar.desc), ar.flat.data_ptr(), net.multires, net.skip_layer, 1e-5,
                                             BN_MOMENTUM, points.data_ptr(), P, out.data_ptr(), Do + 9, Do,
                                             out.data_ptr() + 4 * Do, Do + 9, ws.data_ptr(), ws.numel(), _stream_ptr(dev)),
                       "vfnerf_vf_train_fwd")
        _bump_batches_tracked(net)
        ar.generation += 1             # running statistics updated in the arena by the library
        ctx.net, ctx.ws, ctx.P, ctx.Do = net, (ws if need_bwd else None), P, Do
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, d_out):
        L = _lib.lib()
        net = ctx.net
        ar = net.arena()
        (out,) = ctx.saved_tensors
        dev = out.device
        Do = ctx.Do
        if ctx.ws is None:
            raise RuntimeError("backward through a train-mode VF query made under no_grad")
        # gradients enter through the tanh outputs; the Jacobian columns carry none on any path of the reference (the
        # trainer slices [:, :3], render() consumes the Jacobian under no_grad)
        d_out = d_out.contiguous().float()
        grad = torch.empty(ar.desc.arena_floats, dtype=torch.float32, device=dev)
        with _on_device(dev):
            _lib.check(L.vfnerf_vf_train_bwd(C.byref(ar.desc), ar.flat.data_ptr(), net.multires, net.skip_layer, ctx.P,
                                             out.data_ptr(), Do + 9, d_out.data_ptr(), Do + 9, Do, grad.data_ptr(), 0,
                                             ctx.ws.data_ptr(), ctx.ws.numel(), _stream_ptr(dev)), "vfnerf_vf_train_bwd")
        ctx.ws = None
        if ar.grad_flat is not None:
            ar.grad_flat.add_(grad)
            return (None, None, None) + (None,) * len(ar.params())
        return (None, None, None) + tuple(ar.grad_views(grad))


def vf_query_train(net, points: torch.Tensor) -> torch.Tensor:
    """[P, 3 + feat + 9]: tanh outputs followed by cat(d sum(y0)/dx, d sum(y1)/dx, d sum(y2)/dx), the train-mode return
    value of VectorFieldNetwork.forward (vector_field_network.py:146-173).  Normalises with the statistics of THIS batch
    and folds them into the running statistics."""
    if points.dim() != 2 or points.shape[1] != 3:
        raise ValueError(f"points must be [P,3], got {tuple(points.shape)}")
    points = _require_cuda("points", points.detach())
    ar = net.arena()
    if ar.flat.device != points.device:
        raise RuntimeError(f"network parameters are on {ar.flat.device}, points on {points.device}")
    params = ar.params()
    need_bwd = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    return _VFQueryTrain.apply(net, points, need_bwd, *params)


def mlp_points(vf_net, rn_net, points: torch.Tensor, ray_dirs: torch.Tensor, samples_per_ray: int,
               workspace: Optional[torch.Tensor] = None, repack: bool = True):
    """VF vectors and colours at ``points[P,3]`` seen along ``ray_dirs[P/samples_per_ray,3]`` -- the two-MLP evaluation
    of VectorFieldNerf.get_colors (vector_field_nerf.py:341-375) as ONE fused tcgen05 launch (forward only; bf16, or
    bf16x3 when ``vf_net.precision == "bf16x3"``).
    Returns (normals[P,3], colors[P,3], workspace)."""
    L = _lib.lib()
    points, ray_dirs = _require_cuda("points", points.detach()), _require_cuda("ray_dirs", ray_dirs.detach())
    va, ra = vf_net.arena(), rn_net.arena()
    dev = points.device
    P = points.shape[0]
    if workspace is None:
        nb = L.vfnerf_mlp_points_workspace_bytes(C.byref(va.desc), C.byref(ra.desc), vf_net.multires, rn_net.multires_view,
                                                 vf_net.skip_layer)
        if nb < 0:
            _lib.check(1, "vfnerf_mlp_points_workspace_bytes")
        workspace = _workspace(nb, dev)
    normals = torch.empty(P, 3, dtype=torch.float32, device=dev)
    colors = torch.empty(P, 3, dtype=torch.float32, device=dev)
    with _on_device(dev):
        _lib.check(L.vfnerf_mlp_points_fwd(C.byref(va.desc), va.flat.data_ptr(), C.byref(ra.desc), ra.flat.data_ptr(),
                                           vf_net.multires, rn_net.multires_view, vf_net.skip_layer, 1e-5,
                                           _lib.PRECISIONS[vf_net.precision] if vf_net.precision != "fp32" else _lib.PREC_BF16,
                                           points.data_ptr(), ray_dirs.data_ptr(), int(samples_per_ray), P, normals.data_ptr(),
                                           colors.data_ptr(), workspace.data_ptr(), workspace.numel(), int(repack),
                                           _stream_ptr(dev)), "vfnerf_mlp_points_fwd")
    return normals, colors, workspace


# ------------------------------------------------------------------------------------------------
# render()
# ------------------------------------------------------------------------------------------------
# test support: keep the last training workspace alive so the activation stash can be inspected
# (vfnerf_debug_stash_read); never set in production, the workspace is several KB per sample point
DEBUG_KEEP_WORKSPACE = False
_debug_last: dict = {}


def debug_stash_read(tensor: int) -> torch.Tensor:
    """Test support: reads one activation-stash tensor back through the TEST-ONLY library (_lib.build_debug())."""
    d = _debug_last
    if not d:
        raise RuntimeError("no workspace kept: set ops.DEBUG_KEEP_WORKSPACE = True before backward()")
    L = _lib.debug_lib()
    cfg = d["cfg"]
    n_cols = C.c_int(0)
    args = (C.byref(cfg), C.byref(d["vf"].desc), C.byref(d["rn"].desc), d["ws"].data_ptr(), int(tensor))
    _lib.check_debug(L.vfnerf_debug_stash_read(*args, None, C.byref(n_cols), None), "vfnerf_debug_stash_read")
    P = cfg.n_rays * (cfg.n_coarse + cfg.n_fine)
    out = torch.empty(P, n_cols.value, dtype=torch.float32, device=d["ws"].device)
    _lib.check_debug(L.vfnerf_debug_stash_read(*args, out.data_ptr(), C.byref(n_cols), _stream_ptr(out.device)),
                     "vfnerf_debug_stash_read")
    return out


class RenderCall:
    """Everything one render() call needs besides the parameters (built by nerf.VectorFieldNerf)."""

    def __init__(self, cfg: _lib.RenderCfg, vf_net, rn_net, density, uv, pose, intrinsics, t_vals, U1, U2, U3,
                 z_override=None, want_extras: bool = False, want_ray_dirs: bool = True, train: bool = False,
                 workspace: Optional[torch.Tensor] = None, weights_packed: bool = False):
        self.cfg, self.vf_net, self.rn_net, self.density = cfg, vf_net, rn_net, density
        self.uv, self.pose, self.intrinsics, self.t_vals = uv, pose, intrinsics, t_vals
        self.U1, self.U2, self.U3, self.z_override = U1, U2, U3, z_override
        self.want_extras, self.want_ray_dirs = want_extras, want_ray_dirs
        self.extras = {}
        self.need_bwd = False
        self.train = train          # batch-statistic BatchNorm + directional derivatives (csrc/mlp_train.cu)
        # a caller-owned workspace that outlives the call (CUDA-graph replay: the packed weight images stay in it), and
        # the promise that it already holds them (include/vfnerf_b200.h: VFNERF_FLAG_WEIGHTS_PACKED)
        self.workspace, self.weights_packed = workspace, weights_packed


class _Render(torch.autograd.Function):
    @staticmethod
    def forward(ctx, call: RenderCall, *params):
        L = _lib.lib()
        cfg = call.cfg
        vf_ar, rn_ar = call.vf_net.arena(), call.rn_net.arena()
        dflat = call.density.flat()
        dev = call.uv.device
        R, N = cfg.n_rays, cfg.n_coarse + cfg.n_fine
        need_bwd = call.need_bwd
        if call.z_override is not None:
            # a prescribed z has no coarse sweep to reuse; render_bwd derives the stash order from cfg alone
            cfg.flags |= _lib.FLAG_RECOMPUTE_COARSE
        if need_bwd and cfg.precision not in (_lib.PREC_FP32, _lib.PREC_BF16):
            raise RuntimeError("training (backward) runs on precision 'fp32' or 'bf16'")
        if call.train:
            nbytes = L.vfnerf_render_train_workspace_bytes(C.byref(cfg), C.byref(vf_ar.desc), C.byref(rn_ar.desc))
        else:
            nbytes = L.vfnerf_render_workspace_bytes(C.byref(cfg), C.byref(vf_ar.desc), C.byref(rn_ar.desc), int(need_bwd))
        if nbytes < 0:
            _lib.check(1, "vfnerf_render_workspace_bytes")
        if call.workspace is not None and not need_bwd and call.workspace.numel() >= nbytes:
            ws = call.workspace
            if call.weights_packed:
                cfg.flags |= _lib.FLAG_WEIGHTS_PACKED
        else:
            ws = _workspace(nbytes, dev)
        f32 = dict(dtype=torch.float32, device=dev)
        points = torch.empty(R, N, 3, **f32)
        normals = torch.empty(R, N, 3, **f32)
        rgb = torch.empty(R, 3, **f32)
        depth = torch.empty(R, 1, **f32)
        z_vals = torch.empty(R, N, **f32)
        colors = torch.empty(R * N, 3, **f32)
        ray_dirs = torch.empty(R * N, 3, **f32) if call.want_ray_dirs else None
        weights = torch.empty(R, N, **f32)
        z_c = torch.empty(R, cfg.n_coarse, **f32) if call.want_extras else None
        w_c = torch.empty(R, cfg.n_coarse, **f32) if call.want_extras else None
        out = _lib.RenderOut(points.data_ptr(), normals.data_ptr(), rgb.data_ptr(), depth.data_ptr(),
                             z_vals.data_ptr(), _lib.ptr(ray_dirs), colors.data_ptr(), weights.data_ptr(),
                             _lib.ptr(z_c), _lib.ptr(w_c))
        dd = None
        with _on_device(dev):
            if call.train:
                dd = torch.empty(4 * R * cfg.n_coarse, **f32)
                _lib.check(L.vfnerf_render_train_fwd(
                    C.byref(cfg), C.byref(vf_ar.desc), vf_ar.flat.data_ptr(), C.byref(rn_ar.desc), rn_ar.flat.data_ptr(),
                    dflat.data_ptr(), call.uv.data_ptr(), call.pose.data_ptr(), call.intrinsics.data_ptr(),
                    call.t_vals.data_ptr(), _lib.ptr(call.U1), _lib.ptr(call.U2), _lib.ptr(call.U3),
                    _lib.ptr(call.z_override), C.byref(out), dd.data_ptr(), BN_MOMENTUM, ws.data_ptr(), ws.numel(),
                    _stream_ptr(dev)), "vfnerf_render_train_fwd")
                _bump_batches_tracked(call.vf_net, 2)          # coarse pass + merged pass
                _bump_batches_tracked(call.rn_net, 1)
                vf_ar.generation += 1
                rn_ar.generation += 1
            else:
                _lib.check(L.vfnerf_render_fwd(
                    C.byref(cfg), C.byref(vf_ar.desc), vf_ar.flat.data_ptr(), C.byref(rn_ar.desc), rn_ar.flat.data_ptr(),
                    dflat.data_ptr(), call.uv.data_ptr(), call.pose.data_ptr(), call.intrinsics.data_ptr(),
                    call.t_vals.data_ptr(), _lib.ptr(call.U1), _lib.ptr(call.U2), _lib.ptr(call.U3),
                    _lib.ptr(call.z_override), C.byref(out), ws.data_ptr(), ws.numel(), int(need_bwd), _stream_ptr(dev)),
                    "vfnerf_render_fwd")
        call.extras = dict(weights=weights, z_coarse=z_c, weights_coarse=w_c, directional_derivatives=dd)
        if need_bwd:
            ctx.call, ctx.ws, ctx.out_struct = call, ws, out
            ctx.keep = (points, z_vals, colors, normals, weights, ray_dirs, z_c, w_c)
        if ray_dirs is not None:
            ctx.mark_non_differentiable(points, z_vals, ray_dirs)
            return rgb, depth, normals, colors, points, z_vals, ray_dirs
        ctx.mark_non_differentiable(points, z_vals)
        return rgb, depth, normals, colors, points, z_vals

    @staticmethod
    def backward(ctx, d_rgb, d_depth, d_normals, d_colors, *_):
        L = _lib.lib()
        call = ctx.call
        cfg = call.cfg
        vf_ar, rn_ar = call.vf_net.arena(), call.rn_net.arena()
        dflat = call.density.flat()
        dev = call.uv.device
        R, N = cfg.n_rays, cfg.n_coarse + cfg.n_fine
        f32 = dict(dtype=torch.float32, device=dev)

        def prep(g, shape):
            return None if g is None else g.contiguous().float().view(shape)
        d_rgb = prep(d_rgb, (R, 3)) if d_rgb is not None else torch.zeros(R, 3, **f32)
        d_depth = prep(d_depth, (R, 1)) if d_depth is not None else torch.zeros(R, 1, **f32)
        d_normals = prep(d_normals, (R, N, 3))
        d_colors = prep(d_colors, (R * N, 3))
        g_vf = torch.empty(vf_ar.desc.arena_floats, **f32)
        g_rn = torch.empty(rn_ar.desc.arena_floats, **f32)
        g_d = torch.empty(3, **f32)
        bwd = L.vfnerf_render_train_bwd if call.train else L.vfnerf_render_bwd
        with _on_device(dev):
            _lib.check(bwd(
                C.byref(cfg), C.byref(vf_ar.desc), vf_ar.flat.data_ptr(), C.byref(rn_ar.desc), rn_ar.flat.data_ptr(),
                dflat.data_ptr(), C.byref(ctx.out_struct), d_rgb.data_ptr(), d_depth.data_ptr(), _lib.ptr(d_normals),
                _lib.ptr(d_colors), g_vf.data_ptr(), g_rn.data_ptr(), g_d.data_ptr(), ctx.ws.data_ptr(),
                ctx.ws.numel(), _stream_ptr(dev)), "vfnerf_render_train_bwd" if call.train else "vfnerf_render_bwd")
        if DEBUG_KEEP_WORKSPACE:
            _debug_last.update(cfg=cfg, ws=ctx.ws, vf=vf_ar, rn=rn_ar)
        ctx.ws = ctx.keep = None
        if vf_ar.grad_flat is not None and rn_ar.grad_flat is not None and getattr(call.density, "grad_flat", None) is not None:
            # flat-gradient mode (optim.ArenaAdam): three accumulates instead of 55 per-parameter ones
            vf_ar.grad_flat.add_(g_vf)
            rn_ar.grad_flat.add_(g_rn)
            call.density.grad_flat.add_(g_d)
            return (None,) * (1 + len(vf_ar.params()) + len(rn_ar.params()) + 3)
        grads = vf_ar.grad_views(g_vf) + rn_ar.grad_views(g_rn) + [g_d[0], g_d[1], g_d[2]]
        return (None,) + tuple(grads)


def render_params(vf_net, rn_net, density) -> List[torch.nn.Parameter]:
    """Parameter order used by _Render (and by its backward's return value)."""
    return vf_net.arena().params() + rn_net.arena().params() + [density.beta, density.scale, density.mean]


def render_call(call: RenderCall) -> Tuple[torch.Tensor, ...]:
    params = render_params(call.vf_net, call.rn_net, call.density)
    call.density.flat()
    call.need_bwd = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    return _Render.apply(call, *params)
