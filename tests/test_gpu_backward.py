"""GPU: training-step gradients against the goldens produced by the reference's
own render() -> VFLoss -> backward on CPU, and against autograd through the oracle."""
import numpy as np
import pytest
import torch

import vfn_testutil as U

pytestmark = pytest.mark.gpu
DEV = "cuda"
GRAD_RTOL = 2e-3     # max abs deviation relative to the tensor's largest gradient entry (fp32 path)


def _loss(out, z):
    return U.O.vf_loss(out.coarse_rgb_values, out.coarse_depth_map, out.coarse_normals.reshape(-1, 3),
                       U.t(z, "rgb_gt").to(DEV), U.t(z, "depth_gt").to(DEV), U.LOSS_W, 0.5)


def _run(name):
    case, z = U.load_golden(name)
    st = U.case_state(case, z)
    model = U.make_model(case, st, DEV)
    draws = (U.t(z, "U1"), U.t(z, "U2"), U.t(z, "U3"))
    out = model.render(U.t(z, "pose").to(DEV), U.t(z, "uv").to(DEV), U.t(z, "K").to(DEV), 0, draws=draws,
                       z_vals_override=U.t(z, "ref_z_vals"))
    loss = _loss(out, z)
    model.optimizer.zero_grad()
    loss.backward()
    return case, z, model, loss


@pytest.mark.parametrize("name", ["small_det", "small_perturb"])
def test_gradients_match_reference_small(built_lib, name):
    case, z, model, loss = _run(name)
    assert abs(loss.item() - float(z["ref_loss"])) <= 1e-4
    worst = 0.0
    for prefix, net in (("g_vf.", model.vector_field_network), ("g_rn.", model.rendering_network)):
        for k, p in net.named_parameters():
            g = z[prefix + k]
            assert p.grad is not None, k
            rel = np.abs(p.grad.cpu().numpy() - g).max() / (np.abs(g).max() + 1e-12)
            worst = max(worst, rel)
            assert rel <= GRAD_RTOL, (k, rel)
    for k, p in model.density.named_parameters():
        g = float(z["g_density." + k])
        assert abs(p.grad.item() - g) <= GRAD_RTOL * abs(g) + 1e-6, (k, p.grad.item(), g)
    print(f"[{name}] worst relative gradient deviation {worst:.2e}")


@pytest.mark.parametrize("name", ["full_det", "full_perturb"])
def test_gradients_match_reference_full(built_lib, name):
    """Full-size nets: the fixture keeps per-tensor gradient norms and 64 strided samples per tensor."""
    case, z, model, loss = _run(name)
    assert abs(loss.item() - float(z["ref_loss"])) <= 1e-4
    for prefix, net in (("g_vf.", model.vector_field_network), ("g_rn.", model.rendering_network)):
        for k, p in net.named_parameters():
            g = p.grad.cpu().numpy()
            n_ref = float(z["n_" + prefix + k])
            assert abs(np.linalg.norm(g) - n_ref) <= 2e-3 * n_ref + 1e-7, (k, np.linalg.norm(g), n_ref)
            s_ref = z["s_" + prefix + k]
            s = g.reshape(-1)[:: max(1, g.size // 64)][:64]
            assert np.abs(s - s_ref).max() <= GRAD_RTOL * (np.abs(g).max() + 1e-12) + 1e-9, k
    for k, p in model.density.named_parameters():
        g = float(z["g_density." + k])
        assert abs(p.grad.item() - g) <= GRAD_RTOL * abs(g) + 1e-6, (k, p.grad.item(), g)


def test_gradients_with_upstream_on_every_output(built_lib):
    """All four differentiable outputs (rgb, depth, normals, colors) carry upstream gradient, perturbed
    density parameters inside their bounds; compared with autograd through the oracle."""
    case, z = U.load_golden("small_perturb")
    st = U.case_state(case, z)
    st["density"] = {"beta": torch.tensor(0.35), "scale": torch.tensor(-40.0), "mean": torch.tensor(0.8)}
    model = U.make_model(case, st, DEV)
    draws = (U.t(z, "U1"), U.t(z, "U2"), U.t(z, "U3"))
    zr = U.t(z, "ref_z_vals")
    out = model.render(U.t(z, "pose").to(DEV), U.t(z, "uv").to(DEV), U.t(z, "K").to(DEV), 0, draws=draws, z_vals_override=zr)
    g = torch.Generator().manual_seed(0)
    R, N = zr.shape
    c_rgb, c_dep = torch.randn(R, 3, generator=g), torch.randn(R, 1, generator=g)
    c_nrm, c_col = torch.randn(R, N, 3, generator=g) * 0.01, torch.randn(R * N, 3, generator=g) * 0.01

    def scalar(o_rgb, o_dep, o_nrm, o_col, dev):
        return (o_rgb * c_rgb.to(dev)).sum() + (o_dep * c_dep.to(dev)).sum() + (o_nrm * c_nrm.to(dev)).sum() + \
            (o_col * c_col.to(dev)).sum()
    model.optimizer.zero_grad()
    scalar(out.coarse_rgb_values, out.coarse_depth_map, out.coarse_normals, out.coarse_colors, DEV).backward()

    req = lambda sd: {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
                      for k, v in sd.items()}
    vf, rn = req(st["vf_net"]), req(st["rendering_net"])
    dn = {k: v.clone().requires_grad_(True) for k, v in st["density"].items()}
    ora = U.O.render(vf, rn, dn, U.oracle_cfg(case), U.t(z, "uv"), U.t(z, "pose"), U.t(z, "K"), U.t(z, "t_vals"),
                     *draws, z_vals_override=zr)
    scalar(ora["rgb"], ora["depth"], ora["normals"], ora["colors"], "cpu").backward()
    for net, sd in ((model.vector_field_network, vf), (model.rendering_network, rn)):
        for k, p in net.named_parameters():
            ref = sd[k].grad
            rel = (p.grad.cpu() - ref).abs().max().item() / (ref.abs().max().item() + 1e-12)
            assert rel <= GRAD_RTOL, (k, rel)
    for k, p in model.density.named_parameters():
        ref = dn[k].grad.item()
        assert abs(p.grad.item() - ref) <= GRAD_RTOL * abs(ref) + 1e-6, (k, p.grad.item(), ref)


def test_density_clamps_block_gradient(built_lib):
    """mean outside [0.6, 1.0] and |scale| below scale_min: torch.clamp / torch.max pass no gradient."""
    case, z = U.load_golden("small_det")
    st = U.case_state(case, z)
    st["density"] = {"beta": torch.tensor(0.5), "scale": torch.tensor(0.5), "mean": torch.tensor(0.3)}
    model = U.make_model(case, st, DEV)
    draws = (U.t(z, "U1"), U.t(z, "U2"), U.t(z, "U3"))
    out = model.render(U.t(z, "pose").to(DEV), U.t(z, "uv").to(DEV), U.t(z, "K").to(DEV), 0, draws=draws)
    out.coarse_depth_map.sum().backward()
    assert model.density.mean.grad.item() == 0.0 and model.density.scale.grad.item() == 0.0


def test_one_training_step_like_the_reference_trainer(built_lib):
    """render -> loss -> backward -> clip_grad_norm_(parameters()) -> Adam.step, the sequence of
    train/vector_field_nerf_train.py:177-260, runs and changes the parameters."""
    case, z, model, loss = _run("small_perturb")
    before = model.vector_field_network.layers[0][0].weight.detach().clone()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
    model.optimizer.step()
    model.scheduler.step()
    after = model.vector_field_network.layers[0][0].weight.detach()
    assert not torch.equal(before, after)
    # the arena still aliases the parameters after the in-place optimizer update
    ar = model.vector_field_network.arena()
    assert after.data_ptr() == ar.flat.data_ptr() + 4 * ar.desc.w_off[0]
