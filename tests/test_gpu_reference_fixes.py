"""GPU: the two render() paths that cannot execute upstream, behind their explicit opt-ins:
white background (vector_field_nerf.py:325-329) and rendering="nerf" with nerf_volume_rendering's own argument order
(utils/rendering.py:98-119).  Expected values: the oracle's render() with the same two corrections; the weight function
itself is pinned to the live reference function by tests/golden/volume_weights.npz (test_gpu_volume_weights.py)."""
import pytest
import torch

import vfn_testutil as U

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(name, rendering, normalize=True):
    case, z = U.load_golden(name)
    st = U.case_state(case, z)
    cfgm = U.make_config(case, DEV)
    cfgm.rendering = rendering
    cfgm.normalize_rendering = normalize
    from vfnerf_b200 import VectorFieldNerf
    model = VectorFieldNerf(cfgm)
    model.vector_field_network.load_state_dict(st["vf_net"])
    model.rendering_network.load_state_dict(st["rendering_net"])
    model.density.load_state_dict(st["density"])
    model.ray_sampler.far = model.fine_sampler.far = case["far"]
    model.eval()
    inputs = (U.t(z, "pose").to(DEV), U.t(z, "uv").to(DEV), U.t(z, "K").to(DEV), 0)
    draws = (U.t(z, "U1"), U.t(z, "U2"), U.t(z, "U3"))
    return case, z, st, model, inputs, draws


def test_default_behaviour_is_the_references(built_lib):
    case, z, st, model, inputs, draws = _setup("small_det", "nerf")
    with pytest.raises(NotImplementedError):
        model.render(*inputs, draws=draws)
    case, z, st, model, inputs, draws = _setup("small_det", "volsdf")
    with pytest.raises(UnboundLocalError):
        model.render(*inputs, white=True, draws=draws)
    with pytest.raises(ValueError):
        model.enable_reference_fix("everything")


@pytest.mark.parametrize("name", ["small_det", "small_perturb"])
@pytest.mark.parametrize("rendering,white", [("volsdf", True), ("nerf", False), ("nerf", True)])
def test_opt_in_paths_match_the_corrected_oracle_forward_and_backward(built_lib, name, rendering, white):
    # white background on un-normalised weights: with normalize_rendering the weights sum to 1 - 1e-5 and the term vanishes
    normalize = not white
    case, z, st, model, inputs, draws = _setup(name, rendering, normalize)
    model.enable_reference_fix("white_background", "nerf_rendering")
    ocfg = dict(U.oracle_cfg(case), rendering=rendering, white=white, normalize=normalize)
    vf = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in st["vf_net"].items()}
    rn = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in st["rendering_net"].items()}
    dens = {k: v.clone().requires_grad_(True) for k, v in st["density"].items()}
    ora = U.O.render(vf, rn, dens, ocfg, U.t(z, "uv"), U.t(z, "pose"), U.t(z, "K"), U.t(z, "t_vals"), *draws)
    out = model.render(*inputs, white=white, draws=draws, z_vals_override=ora["z_vals"].detach())
    free = model.render(*inputs, white=white, draws=draws)
    same = (free.z_vals.cpu() == ora["z_vals"]).all(dim=1)
    assert same.float().mean().item() >= 0.9                    # the coarse pass uses the same weight function
    ok = U.discontinuity_guard({k: (v.detach() if torch.is_tensor(v) else v) for k, v in ora.items()}, case)
    assert (out.weights.cpu() - ora["weights"].detach())[ok].abs().max().item() <= 1e-4
    assert (out.coarse_rgb_values.detach().cpu() - ora["rgb"].detach())[ok].abs().max().item() <= 1e-3
    assert (out.coarse_depth_map.detach().cpu() - ora["depth"].detach())[ok].abs().max().item() <= 1e-3
    if white:
        assert (out.coarse_rgb_values.detach() >= 0).all()
        bg = 1 - out.weights.sum(1, keepdim=True)
        assert bg.abs().max().item() > 1e-3                      # the term is alive on this model
    # gradients of a loss that reaches every output, rays on a discontinuity masked out on both sides
    m = ok.float()
    g = torch.Generator().manual_seed(5)
    c_rgb, c_dep = torch.randn(ora["rgb"].shape, generator=g), torch.randn(ora["depth"].shape, generator=g)
    (ora["rgb"] * c_rgb * m[:, None]).sum().add((ora["depth"] * c_dep * m[:, None]).sum()).backward()
    model.optimizer.zero_grad()
    md = m.to(DEV)
    ((out.coarse_rgb_values * c_rgb.to(DEV) * md[:, None]).sum() + (out.coarse_depth_map * c_dep.to(DEV) * md[:, None]).sum()).backward()
    worst = 0.0
    for k, p in model.vector_field_network.named_parameters():
        a, b = vf[k].grad, p.grad.cpu()
        worst = max(worst, ((a - b).abs().max() / (a.abs().max() + 1e-12)).item())
    for k, p in model.rendering_network.named_parameters():
        a, b = rn[k].grad, p.grad.cpu()
        worst = max(worst, ((a - b).abs().max() / (a.abs().max() + 1e-12)).item())
    for k in ("beta", "scale", "mean"):
        a, b = dens[k].grad, getattr(model.density, k).grad.cpu()
        if a.abs().item() > 1e-8:
            worst = max(worst, ((a - b).abs() / a.abs()).item())
    print(f"[{name} {rendering} white={white}] worst gradient deviation (relative to each tensor's max) {worst:.2e}")
    assert worst <= 5e-3


def test_nerf_weights_with_the_split_precision_chain(built_lib):
    """The opt-in composes with the tensor-core modes: same flags reach the same density / composite kernels."""
    case, z = U.load_golden("full_det")
    st = U.case_state(case, z)
    cfgm = U.make_config(case, DEV)
    cfgm.rendering = "nerf"
    from vfnerf_b200 import VectorFieldNerf
    model = VectorFieldNerf(cfgm, precision="bf16x3")
    model.vector_field_network.load_state_dict(st["vf_net"]); model.rendering_network.load_state_dict(st["rendering_net"])
    model.density.load_state_dict(st["density"])
    model.ray_sampler.far = model.fine_sampler.far = case["far"]
    model.eval()
    model.enable_reference_fix("nerf_rendering", "white_background")
    draws = (U.t(z, "U1"), U.t(z, "U2"), U.t(z, "U3"))
    with torch.no_grad():
        ora = U.O.render(st["vf_net"], st["rendering_net"], st["density"], dict(U.oracle_cfg(case), rendering="nerf", white=True),
                         U.t(z, "uv"), U.t(z, "pose"), U.t(z, "K"), U.t(z, "t_vals"), *draws)
        out = model.render(U.t(z, "pose").to(DEV), U.t(z, "uv").to(DEV), U.t(z, "K").to(DEV), 0, white=True, draws=draws,
                           z_vals_override=ora["z_vals"])
    ok = U.discontinuity_guard(ora, case)
    assert (out.coarse_rgb_values.cpu() - ora["rgb"])[ok].abs().max().item() <= 1e-3
    assert (out.coarse_depth_map.cpu() - ora["depth"])[ok].abs().max().item() <= 2.5e-3
