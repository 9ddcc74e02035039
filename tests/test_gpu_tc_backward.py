"""GPU: training backward on the tensor cores (precision 'bf16': fused tcgen05 dgrad chain + MN-major wgrad GEMMs,
mlp_tc_bwd.cu) against the fp32 CUDA-core backward, which itself is pinned to the reference's gradients
(test_gpu_backward.py).  Same inputs, same sample positions (z_vals_override), same upstream gradients.

Two comparisons.  (1) Against an fp32 autograd emulation that rounds where the kernels round: tight (1.5e-2 relative
L2 per tensor) -- this is the implementation check.  (2) Against the fp32 path: the bf16 forward differs by a few
1e-3 in every pre-activation, which flips the ReLU gate of ~0.5 % of the units; a flipped gate changes that unit's
gradient contribution by 100 %, so per-tensor relative L2 differences of 2e-2 (last layers) to 1e-1 (first layers of
the chain) are the gradient of a slightly different function, not an error.  Bound asserted: 0.15."""
import numpy as np
import pytest
import torch

import vfn_testutil as U

pytestmark = pytest.mark.gpu
DEV = "cuda"
REL_L2 = 0.15


def _upstream(R, N):
    g = torch.Generator().manual_seed(0)
    return (torch.randn(R, 3, generator=g), torch.randn(R, 1, generator=g),
            torch.randn(R, N, 3, generator=g) * 0.01, torch.randn(R * N, 3, generator=g) * 0.01)


def _grads(case, z, state, precision, rays=None):
    model = U.make_model(case, state, DEV, precision=precision)
    draws = (U.t(z, "U1"), U.t(z, "U2"), U.t(z, "U3"))
    zr = U.t(z, "ref_z_vals")
    uv, pose, K = U.t(z, "uv"), U.t(z, "pose"), U.t(z, "K")
    if rays is not None:
        uv, pose, K, zr = uv[:rays], pose[:rays], K[:rays], zr[:rays]
        draws = tuple(d[:rays] for d in draws)
    out = model.render(pose.to(DEV), uv.to(DEV), K.to(DEV), 0, draws=draws, z_vals_override=zr)
    R, N = zr.shape
    c_rgb, c_dep, c_nrm, c_col = (c.to(DEV) for c in _upstream(R, N))
    model.optimizer.zero_grad()
    loss = (out.coarse_rgb_values * c_rgb).sum() + (out.coarse_depth_map * c_dep).sum() + \
        (out.coarse_normals * c_nrm).sum() + (out.coarse_colors * c_col).sum()
    loss.backward()
    g = {}
    for prefix, net in (("vf.", model.vector_field_network), ("rn.", model.rendering_network)):
        for k, p in net.named_parameters():
            g[prefix + k] = p.grad.detach().cpu().clone()
    for k, p in model.density.named_parameters():
        g["density." + k] = p.grad.detach().cpu().clone()
    return g, out


def _compare(ga, gb, tol):
    worst = ("", 0.0)
    for k, a in ga.items():
        b = gb[k]
        assert torch.isfinite(b).all(), k
        if k.startswith("density."):
            continue
        rel = ((a - b).norm() / (a.norm() + 1e-20)).item()
        if rel > worst[1]:
            worst = (k, rel)
    return worst


class _RoundBF16(torch.autograd.Function):
    """bf16 rounding with a straight-through gradient: the kernels round activations when they store them and
    differentiate as if they had not."""
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


def _q(x):
    return _RoundBF16.apply(x)


def _emulated_grads(state, points, view_dirs, c_nrm, c_col, eps=1e-5, skip=4):
    """fp32 autograd through a torch model that rounds exactly where the tensor-core path rounds: BatchNorm folded
    into bf16 weights (1/sqrt(2) of the skip connection folded into the producing layer), bf16 activations, fp32
    accumulation, fp32 bias; the two 3-wide output layers (VF vector rows, colour) are fp32 dot products of the
    UNROUNDED previous activation with fp32 weights (TcStep::dot, csrc/mlp_tc.cuh).  What is left between this and the
    kernels is the bf16 rounding of dL/dY and accumulation order."""
    def leaves(sd):
        return {k: (v.to(DEV).clone().requires_grad_(True) if "running" not in k and v.is_floating_point() else v.to(DEV))
                for k, v in sd.items()}
    vf, rn = leaves(state["vf_net"]), leaves(state["rendering_net"])

    def layer(sd, i, x, last, post=1.0, x32=None, n32=0):
        if last:
            W, b = sd[f"layers.{i}.weight"], sd[f"layers.{i}.bias"]
            y = x @ _q(W).T + b
            if n32:
                y = torch.cat([x32 @ W[:n32].T + b[:n32], y[:, n32:]], 1)
            return y
        W, b = sd[f"layers.{i}.0.weight"], sd[f"layers.{i}.0.bias"]
        sc = sd[f"layers.{i}.1.weight"] / torch.sqrt(sd[f"layers.{i}.1.running_var"] + eps)
        sh = (b - sd[f"layers.{i}.1.running_mean"]) * sc + sd[f"layers.{i}.1.bias"]
        return x @ _q(W * sc[:, None] * post).T + sh * post

    L = 1 + sum(1 for k in vf if k.endswith(".0.weight"))
    emb = U.O.embed(points, 6)
    x = emb                                       # hi + lo split: effectively unrounded
    inv = 1.0 / np.sqrt(2.0)
    for i in range(L):
        last = i == L - 1
        if i == skip:
            x = torch.cat([x, _q(emb * inv)], 1)
        y = layer(vf, i, x, last, post=inv if i == skip - 1 else 1.0, x32=x32 if last else None, n32=3 if last else 0)
        if last:
            v, feat = torch.tanh(y[:, :3]), _q(torch.tanh(_q(y[:, 3:])))
        else:
            x32 = torch.relu(y)
            x = _q(x32)
    Lr = 1 + sum(1 for k in rn if k.endswith(".0.weight"))
    x = torch.cat([_q(points), _q(U.O.embed(view_dirs, 4)), _q(v.detach()), feat], 1)
    for i in range(Lr):
        last = i == Lr - 1
        y = layer(rn, i, x, last, x32=x32 if last else None, n32=3 if last else 0)
        if last:
            x = y
        else:
            x32 = torch.relu(y)
            x = _q(x32)
    colors = torch.sigmoid(x)
    ((v * c_nrm).sum() + (colors * c_col).sum()).backward()
    g = {}
    for prefix, sd in (("vf.", vf), ("rn.", rn)):
        for k, t in sd.items():
            if t.requires_grad:
                g[prefix + k] = t.grad.detach().cpu()
    return g, v.detach(), colors.detach()


@pytest.mark.parametrize("beta_shift", [0.0, 4.0])
def test_bf16_backward_matches_rounding_emulation(built_lib, beta_shift):
    """Implementation check proper: same rounding points, so ReLU gates agree up to the hardware tanh approximation
    of the feature vector (1 bf16 ulp), which still flips a few gates of the colour net.  beta_shift = 4 moves every
    BatchNorm shift up so that (almost) all units are active: the chain becomes linear, gate flips disappear and
    what remains is bf16 rounding of dL/dY and of the saturated sigmoid outputs."""
    case, z = U.load_golden("full_det")
    st = U.S.synthetic_state(0, vf_gain=1.0, center_output=False)
    if beta_shift:
        for net in ("vf_net", "rendering_net"):
            for k in st[net]:
                if k.endswith(".1.bias"):
                    st[net][k] = st[net][k] + beta_shift
    model = U.make_model(case, st, DEV, precision="bf16")
    draws = (U.t(z, "U1"), U.t(z, "U2"), U.t(z, "U3"))
    zr = U.t(z, "ref_z_vals")
    out = model.render(U.t(z, "pose").to(DEV), U.t(z, "uv").to(DEV), U.t(z, "K").to(DEV), 0, draws=draws, z_vals_override=zr)
    R, N = zr.shape
    _, _, c_nrm, c_col = (c.to(DEV) for c in _upstream(R, N))
    model.optimizer.zero_grad()
    ((out.coarse_normals * c_nrm).sum() + (out.coarse_colors * c_col).sum()).backward()
    geo = U.O.ray_geometry(U.t(z, "uv"), U.t(z, "pose"), U.t(z, "K"))
    ray_dirs = geo[1]
    view = ray_dirs.reshape(R, 1, 3).expand(R, N, 3).reshape(-1, 3).to(DEV)
    ge, v, colors = _emulated_grads(st, out.points_coarse.detach().reshape(-1, 3), view, c_nrm.reshape(-1, 3), c_col)
    print(f"forward: v max diff {(v - out.coarse_normals.reshape(-1, 3)).abs().max().item():.2e}, "
          f"colours max diff {(colors - out.coarse_colors).abs().max().item():.2e}")
    worst = ("", 0.0)
    for prefix, net in (("vf.", model.vector_field_network), ("rn.", model.rendering_network)):
        for k, p in net.named_parameters():
            a, b = ge[prefix + k], p.grad.detach().cpu()
            rel = ((a - b).norm() / (a.norm() + 1e-20)).item()
            print(f"{prefix + k:40s} |g|={a.norm().item():.3e} rel={rel:.2e}")
            if rel > worst[1]:
                worst = (prefix + k, rel)
    # beta_shift: activations of magnitude ~5 put many colours into sigmoid saturation, where c(1-c) of the emulation
    # and of the kernels differ by more than the gradient averaging removes; the tight check is the stage-by-stage test
    assert worst[1] <= (0.06 if beta_shift else 0.15), worst


def test_bf16_backward_matches_fp32_on_the_tame_model(built_lib):
    case, z = U.load_golden("full_det")
    st = U.S.synthetic_state(0, vf_gain=1.0, center_output=False)
    g32, _ = _grads(case, z, st, "fp32")
    g16, _ = _grads(case, z, st, "bf16")
    k, rel = _compare(g32, g16, REL_L2)
    for name in sorted(g32):
        a, b = g32[name], g16[name]
        print(f"{name:40s} |g32|={a.norm().item():.3e} rel={((a - b).norm() / (a.norm() + 1e-20)).item():.2e}")
    assert rel <= REL_L2, (k, rel)


def test_bf16_backward_matches_fp32_on_the_bending_model(built_lib):
    """The golden (centred, gain 2) model: the density is non-degenerate, so the vector rows of the VF output layer
    and the density parameters receive gradient through the VolSDF weights."""
    case, z = U.load_golden("full_det")
    st = U.case_state(case, z)
    g32, _ = _grads(case, z, st, "fp32")
    g16, _ = _grads(case, z, st, "bf16")
    k, rel = _compare(g32, g16, 0.0)
    print(f"worst tensor {k}: rel L2 {rel:.2e}")
    # the forward on this model already differs by ~2e-2 in the normals (test_gpu_tc.py) and the density term
    # amplifies it: this is the gradient of a measurably different function, reported and only loosely bounded
    assert rel <= 1.5, (k, rel)
    for name in ("beta", "scale", "mean"):
        a, b = g32["density." + name].item(), g16["density." + name].item()
        print(f"density.{name}: fp32 {a:.4e} bf16 {b:.4e}")


def test_bf16_backward_ragged_point_count(built_lib):
    """A ray count whose R*N is not a multiple of the 256-point cluster tile (odd tile count: a phantom tile in the pair)."""
    case, z = U.load_golden("full_det")
    st = U.S.synthetic_state(0, vf_gain=1.0, center_output=False)
    g32, _ = _grads(case, z, st, "fp32", rays=3)
    g16, _ = _grads(case, z, st, "bf16", rays=3)
    k, rel = _compare(g32, g16, REL_L2)
    assert rel <= REL_L2, (k, rel)


def test_bf16_training_step_runs(built_lib):
    """render -> loss -> backward -> clip -> Adam.step on the tensor-core path."""
    case, z = U.load_golden("full_perturb")
    model = U.make_model(case, U.case_state(case, z), DEV, precision="bf16")
    draws = (U.t(z, "U1"), U.t(z, "U2"), U.t(z, "U3"))
    out = model.render(U.t(z, "pose").to(DEV), U.t(z, "uv").to(DEV), U.t(z, "K").to(DEV), 0, draws=draws)
    loss = U.O.vf_loss(out.coarse_rgb_values, out.coarse_depth_map, out.coarse_normals.reshape(-1, 3),
                       U.t(z, "rgb_gt").to(DEV), U.t(z, "depth_gt").to(DEV), U.LOSS_W, 0.5)
    model.optimizer.zero_grad()
    loss.backward()
    before = model.vector_field_network.layers[3][0].weight.detach().clone()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
    model.optimizer.step()
    assert torch.isfinite(loss) and not torch.equal(before, model.vector_field_network.layers[3][0].weight.detach())


@pytest.mark.parametrize("n_rays", [32, 601, 602])
def test_bf16_backward_stage_by_stage(built_lib, debug_lib, n_rays):
    """Tight check of every stage of the tensor-core backward on its OWN inputs: the activation stash is read back
    (vfnerf_debug_stash_read) and each step is recomputed in float64 from the tensors the kernels actually consumed --
    forward stash consistency, every dgrad step (transposed bf16 weights, ReLU / tanh gates), every weight-gradient
    GEMM and the BatchNorm chain rule of the finalize kernel.  No ReLU-gate ambiguity enters, so the bounds are bf16
    rounding of single values (dgrad outputs) and fp32 accumulation order (parameter gradients: 2e-4).
    32 rays: the golden rays with the reference's sample positions (one tile per CTA).  601 rays: synthetic rays,
    own sampler, an odd number of 128-point tiles and several tiles per CTA (persistent loops, barrier phases).
    602 rays: the coarse block ends on a tile boundary, so the forward reuses the coarse sweep (two stash-writing launches,
    301 + 301 tiles, points stashed in evaluation order; vfnerf_debug_stash_read hands the rows back in merged order)."""
    from vfnerf_b200 import ops
    case, z = U.load_golden("full_det")
    st = U.case_state(case, z)                       # the bending model: density terms are alive
    model = U.make_model(case, st, DEV, precision="bf16")
    if n_rays == 32:
        draws = (U.t(z, "U1"), U.t(z, "U2"), U.t(z, "U3"))
        zr = U.t(z, "ref_z_vals")
        uv, pose, K = U.t(z, "uv"), U.t(z, "pose"), U.t(z, "K")
    else:
        uv, pose, K = U.S.synthetic_rays(n_rays, seed=case["seed"], start=case["start"], stride=case["stride"])
        draws, zr = U.S.synthetic_draws(n_rays, case["n_coarse"], case["n_fine"]), None
    R, N = n_rays, case["n_coarse"] + case["n_fine"]
    ops.DEBUG_KEEP_WORKSPACE = True
    try:
        out = model.render(pose.to(DEV), uv.to(DEV), K.to(DEV), 0, draws=draws, z_vals_override=zr)
        c_rgb, c_dep, c_nrm, c_col = (c.to(DEV) for c in _upstream(R, N))
        model.optimizer.zero_grad()
        ((out.coarse_rgb_values * c_rgb).sum() + (out.coarse_depth_map * c_dep).sum() +
         (out.coarse_normals * c_nrm).sum() + (out.coarse_colors * c_col).sum()).backward()
        torch.cuda.synchronize()
        Lv, Lr = 9, 5
        n_y = Lv + Lr - 1
        T = [ops.debug_stash_read(i).double() for i in range(2 * n_y + 3)]
        d3 = ops.debug_stash_read(1000).double()
    finally:
        ops.DEBUG_KEEP_WORKSPACE = False
        ops._debug_last.clear()
    Ys, Yfeat, Yc = T[:Lv - 1], T[Lv - 1], T[Lv:n_y]
    emb0, skipx, aux = T[n_y], T[n_y + 1], T[n_y + 2]
    D = T[n_y + 3:]
    Ds, Dfeat, Dc = D[:Lv - 1], D[Lv - 1], D[Lv:n_y]
    dcol, dv = d3[:, :3], d3[:, 3:]
    q = lambda x: x.float().bfloat16().double()
    inv = 1.0 / np.sqrt(2.0)
    eps = 1e-5

    def fold(sd, i, post=1.0):
        W, b = sd[f"layers.{i}.0.weight"].double().to(DEV), sd[f"layers.{i}.0.bias"].double().to(DEV)
        g, beta = sd[f"layers.{i}.1.weight"].double().to(DEV), sd[f"layers.{i}.1.bias"].double().to(DEV)
        m, var = sd[f"layers.{i}.1.running_mean"].double().to(DEV), sd[f"layers.{i}.1.running_var"].double().to(DEV)
        istd = 1.0 / torch.sqrt(var.float() + eps).double()
        sc = g * istd
        return dict(W=W, b=b, m=m, istd=istd, sc=sc, post=post, Wq=q((W.float() * sc.float()[:, None] * np.float32(post))),
                    sh=((b - m) * sc + beta) * post)
    vf, rn = st["vf_net"], st["rendering_net"]
    FV = [fold(vf, i, inv if i == 3 else 1.0) for i in range(Lv - 1)]
    FC = [fold(rn, i) for i in range(Lr - 1)]
    W8, b8 = vf["layers.8.weight"].double().to(DEV), vf["layers.8.bias"].double().to(DEV)
    W4c = rn["layers.4.weight"].double().to(DEV)
    emb = emb0[:, :39] + emb0[:, 48:87]
    n_prev = FV[3]["W"].shape[0]                      # 217: outputs of the layer before the skip
    Xs = [emb] + [Ys[l - 1] for l in range(1, Lv - 1)]
    Xs[4] = torch.cat([Ys[3][:, :n_prev], skipx[:, :39]], 1)
    small = torch.cat([aux[:, 8:11], aux[:, 11:38], aux[:, 0:3]], 1)       # reference order: p, embed(view), n
    Xc = [torch.cat([small, Yfeat], 1)] + [Yc[l - 1] for l in range(1, Lr - 1)]

    def close(a, b, what, frac=0.999):
        ok = (a - b).abs() <= 0.012 * b.abs() + 1e-30 + 2e-3 * b.abs().mean()
        f = ok.double().mean().item()
        print(f"{what:28s} within one bf16 step: {f:.5f}")
        assert f >= frac, (what, f)

    # A. the forward stash holds what the next layer consumed
    for l in range(Lv - 1):
        n = FV[l]["W"].shape[0]
        close(Ys[l][:, :n], q(torch.relu(Xs[l] @ FV[l]["Wq"].T + FV[l]["sh"])), f"stash Y_s{l}")
    for l in range(Lr - 1):
        close(Yc[l], q(torch.relu(Xc[l] @ FC[l]["Wq"].T + FC[l]["sh"])), f"stash Y_c{l}")
    # B. dgrad chain, one step at a time on the kernels' own inputs
    gate = lambda d, y: d * (y > 0)
    close(Dc[3], q(gate(dcol @ q(W4c), Yc[3])), "dgrad D_c3")
    for l in (3, 2, 1):
        close(Dc[l - 1], q(gate(Dc[l] @ FC[l]["Wq"], Yc[l - 1])), f"dgrad D_c{l - 1}")
    close(Dfeat, q((Dc[0] @ FC[0]["Wq"][:, 33:]) * (1 - Yfeat * Yfeat)), "dgrad D_feat")
    close(Ds[7], q(gate(Dfeat @ q(W8[3:]) + dv @ q(W8[:3]), Ys[7])), "dgrad D_s7")
    for l in range(7, 0, -1):
        n = FV[l - 1]["W"].shape[0]
        nl = FV[l]["W"].shape[0]
        close(Ds[l - 1][:, :n], q(gate(Ds[l][:, :nl] @ FV[l]["Wq"][:, :n], Ys[l - 1][:, :n])), f"dgrad D_s{l - 1}")
    # C. parameter gradients from the stashed tensors
    worst = ("", 0.0)

    def chk(name, got, exp):
        nonlocal worst
        rel = ((got.double().cpu() - exp.cpu()).norm() / (exp.norm().cpu() + 1e-30)).item()
        if rel > worst[1]:
            worst = (name, rel)
        assert rel <= 2e-4, (name, rel)

    def bn_layer(prefix, net, i, F, Dl, Xl):
        n = F["W"].shape[0]
        Dl = Dl[:, :n]
        G, s = Dl.T @ Xl, Dl.sum(0)
        P = dict(net.named_parameters())
        chk(f"{prefix}{i}.0.weight", P[f"layers.{i}.0.weight"].grad, F["sc"][:, None] * F["post"] * G)
        chk(f"{prefix}{i}.0.bias", P[f"layers.{i}.0.bias"].grad, F["sc"] * F["post"] * s)
        chk(f"{prefix}{i}.1.weight", P[f"layers.{i}.1.weight"].grad,
            F["istd"] * F["post"] * ((F["W"] * G).sum(1) + (F["b"] - F["m"]) * s))
        chk(f"{prefix}{i}.1.bias", P[f"layers.{i}.1.bias"].grad, F["post"] * s)
    for l in range(Lv - 1):
        bn_layer("vf.", model.vector_field_network, l, FV[l], Ds[l], Xs[l])
    for l in range(Lr - 1):
        bn_layer("rn.", model.rendering_network, l, FC[l], Dc[l], Xc[l])
    Pv, Pc = dict(model.vector_field_network.named_parameters()), dict(model.rendering_network.named_parameters())
    chk("vf.8.weight", Pv["layers.8.weight"].grad, torch.cat([dv.T @ Ys[7], Dfeat.T @ Ys[7]], 0))
    chk("vf.8.bias", Pv["layers.8.bias"].grad, torch.cat([dv.sum(0), Dfeat.sum(0)]))
    chk("rn.4.weight", Pc["layers.4.weight"].grad, dcol.T @ Yc[3])
    chk("rn.4.bias", Pc["layers.4.bias"].grad, dcol.sum(0))
    print(f"worst parameter-gradient deviation: {worst[0]} {worst[1]:.2e}")


@pytest.mark.parametrize("n_cols,P", [(None, 4096), (3, 4096), (None, 128 * 300 + 77)])
def test_bf16_vf_query_backward(built_lib, n_cols, P):
    """The VF-only module call with gradients (supervision points of the trainer, train/vector_field_nerf_train.py:
    191-216) on the tensor cores: forward with activation stash, VF-only dgrad chain, the shared wgrad / finalize
    kernels.  Compared with the fp32 path on the tame model (same bound and reasoning as the render() gradients)."""
    from vfnerf_b200 import ops
    case, z = U.load_golden("full_det")
    st = U.S.synthetic_state(0, vf_gain=1.0, center_output=False)
    g = torch.Generator().manual_seed(11)
    pts = ((torch.rand(P, 3, generator=g) - 0.5) * 6).to(DEV)
    cols = 259 if n_cols is None else n_cols
    c = (torch.randn(P, cols, generator=g) * 0.01).to(DEV)
    grads = {}
    for prec in ("fp32", "bf16"):
        model = U.make_model(case, st, DEV, precision=prec)
        out = ops.vf_query(model.vector_field_network, pts, n_cols=n_cols)
        model.optimizer.zero_grad()
        (out * c).sum().backward()
        grads[prec] = {k: p.grad.detach().cpu().clone() for k, p in model.vector_field_network.named_parameters()}
    worst = ("", 0.0)
    for k, a in grads["fp32"].items():
        b = grads["bf16"][k]
        assert torch.isfinite(b).all(), k
        rel = ((a - b).norm() / (a.norm() + 1e-20)).item()
        if rel > worst[1]:
            worst = (k, rel)
    print(f"n_cols={n_cols} P={P}: worst tensor {worst[0]} rel L2 {worst[1]:.2e}")
    assert worst[1] <= REL_L2, worst


def test_bf16_render_plus_supervision_accumulates(built_lib):
    """The trainer's step with supervision points: render() gradients and VF-only gradients of extra points add up in
    .grad through autograd, exactly like two separate backward() calls."""
    from vfnerf_b200 import ops
    case, z = U.load_golden("full_det")
    st = U.S.synthetic_state(0, vf_gain=1.0, center_output=False)
    g16, out = _grads(case, z, st, "bf16")
    model = U.make_model(case, st, DEV, precision="bf16")
    draws = (U.t(z, "U1"), U.t(z, "U2"), U.t(z, "U3"))
    zr = U.t(z, "ref_z_vals")
    o = model.render(U.t(z, "pose").to(DEV), U.t(z, "uv").to(DEV), U.t(z, "K").to(DEV), 0, draws=draws, z_vals_override=zr)
    R, N = zr.shape
    c_rgb, c_dep, c_nrm, c_col = (c.to(DEV) for c in _upstream(R, N))
    gsup = torch.Generator().manual_seed(5)
    pts = ((torch.rand(3000, 3, generator=gsup) - 0.5) * 6).to(DEV)
    tgt = torch.nn.functional.normalize(torch.randn(3000, 3, generator=gsup), dim=1).to(DEV)
    v = model.vector_field_network(pts)[:, :3]
    loss = (o.coarse_rgb_values * c_rgb).sum() + (o.coarse_depth_map * c_dep).sum() + (o.coarse_normals * c_nrm).sum() + \
        (o.coarse_colors * c_col).sum() + ((v - tgt) ** 2).mean()
    model.optimizer.zero_grad()
    loss.backward()
    # the supervision term alone
    m2 = U.make_model(case, st, DEV, precision="bf16")
    v2 = m2.vector_field_network(pts)[:, :3]
    m2.optimizer.zero_grad()
    ((v2 - tgt) ** 2).mean().backward()
    for (k, p), (_, p2) in zip(model.vector_field_network.named_parameters(), m2.vector_field_network.named_parameters()):
        want = g16["vf." + k] + p2.grad.cpu()
        assert ((p.grad.cpu() - want).norm() <= 1e-4 * want.norm() + 1e-9), k


@pytest.mark.parametrize("R", [64, 602])
def test_bf16_backward_with_reused_coarse_sweep_matches_literal_schedule(built_lib, R):
    """Training forward that evaluates every unique point once (coarse sweep with stash + fine candidates with stash,
    gradients scattered into evaluation order) vs the literal schedule of vector_field_nerf.py:252-312 (VF-only coarse
    sweep, both MLPs on all merged points).  Forward outputs are bit-identical; the per-point gradient contributions are
    identical too, only the order in which the weight-gradient GEMMs and column sums add them differs (fp32)."""
    case, z = U.load_golden("full_perturb")
    st = U.case_state(case, z)
    uv, pose, K = U.S.synthetic_rays(R, seed=0, start=3, stride=797)
    draws = U.S.synthetic_draws(R, case["n_coarse"], case["n_fine"], seed=17)
    N = case["n_coarse"] + case["n_fine"]
    c_rgb, c_dep, c_nrm, c_col = (c.to(DEV) for c in _upstream(R, N))
    res = {}
    for literal in (False, True):
        model = U.make_model(case, st, DEV, precision="bf16")
        model.recompute_coarse = literal
        out = model.render(pose.to(DEV), uv.to(DEV), K.to(DEV), 0, draws=draws)
        model.optimizer.zero_grad()
        ((out.coarse_rgb_values * c_rgb).sum() + (out.coarse_depth_map * c_dep).sum() +
         (out.coarse_normals * c_nrm).sum() + (out.coarse_colors * c_col).sum()).backward()
        g = {}
        for prefix, net in (("vf.", model.vector_field_network), ("rn.", model.rendering_network), ("density.", model.density)):
            for k, p in net.named_parameters():
                g[prefix + k] = p.grad.detach().clone()
        res[literal] = (out, g)
    for f in ("z_vals", "points_coarse", "coarse_normals", "coarse_colors", "coarse_rgb_values", "coarse_depth_map"):
        assert torch.equal(getattr(res[False][0], f), getattr(res[True][0], f)), f
    worst = ("", 0.0)
    for k, a in res[True][1].items():
        b = res[False][1][k]
        assert torch.isfinite(b).all(), k
        rel = ((a - b).norm() / (a.norm() + 1e-20)).item()
        if rel > worst[1]:
            worst = (k, rel)
    print(f"R={R}: worst relative L2 difference of a parameter gradient, reuse vs literal: {worst[1]:.2e} ({worst[0]})")
    assert worst[1] <= 2e-4
