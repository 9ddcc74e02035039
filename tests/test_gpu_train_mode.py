"""GPU: the train-mode path (csrc/mlp_train.cu) against goldens written by the LIVE reference in
model.train() (tests/golden/make_golden_train.py): BatchNorm batch statistics in both passes of the VF net and in the
colour net, running-statistics update, the autograd "Jacobian" (column sums through the batch statistics), directional
derivatives, and the backward through the batch statistics."""
import numpy as np
import pytest
import torch

import vfn_testutil as U

pytestmark = pytest.mark.gpu
DEV = "cuda"
LOSS_W = dict(U.LOSS_W, directional_derivatives=0.05)


def _state(case, z):
    """Weights of a train golden: Linear weights from the fixture (small) or the seed (full); BatchNorm tensors -- randomised
    affine parameters and running statistics -- always from the fixture."""
    st = U.case_state(case, z) if any(k.startswith("w_vf.layers.0.0.") for k in z.files) else \
        U.S.synthetic_state(case["seed"], case["vf_hidden"], case["feat"], case["rn_hidden"], vf_gain=case["vf_gain"])
    st = {k: dict(v) for k, v in st.items()}
    for tag, net in (("w_vf.", "vf_net"), ("w_rn.", "rendering_net")):
        for k in z.files:
            if k.startswith(tag):
                st[net][k[len(tag):]] = torch.from_numpy(z[k])
    return st


def _model(name):
    case, z = U.load_golden(name)
    st = _state(case, z)
    model = U.make_model(case, st, DEV)
    model.train()
    assert model.vector_field_network.training and model.rendering_network.training
    return case, z, st, model


def _rel(a, b, floor=1e-4):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), floor))


def _rel_q(a, b, q=0.999):
    """Like _rel on the q-quantile of the absolute error.  Derivatives of a ReLU net are discontinuous where a
    pre-activation crosses zero: among ~2M hidden units a few sit within rounding noise of 0, take the other branch than
    the reference's CPU arithmetic, and move THEIR sample's Jacobian by a finite amount (outputs are unaffected: the unit
    contributes ~0 either way).  The bulk must agree tightly, the outliers loosely.  (In train mode a flipped unit also
    shifts the batch means of the BatchNorm backward, i.e. EVERY sample's Jacobian by O(1/batch): with the 512-point
    coarse batch of train_full that is the 1e-3 level, hence the 5e-3 bound on the derivative outputs there.)"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.quantile(np.abs(a - b), q) / max(np.abs(b).max(), 1e-4))


@pytest.mark.parametrize("name", ["train_small", "train_full"])
def test_train_mode_render_matches_reference(built_lib, name):
    case, z, st, model = _model(name)
    R, Nc = case["n_rays"], case["n_coarse"]
    draws = (U.t(z, "U1"), U.t(z, "U2"), U.t(z, "U3"))
    model.optimizer.zero_grad()
    out = model.render(U.t(z, "pose").to(DEV), U.t(z, "uv").to(DEV), U.t(z, "K").to(DEV), 0, draws=draws)
    same = (out.z_vals.cpu() == U.t(z, "ref_z_vals")).all(dim=1)
    assert same.float().mean().item() >= 0.9, "sample placement diverges from the reference"
    ok = same
    d = {"normals": (out.coarse_normals.detach().cpu() - U.t(z, "ref_normals"))[ok].abs().max().item(),
         "rgb": (out.coarse_rgb_values.detach().cpu() - U.t(z, "ref_rgb"))[ok].abs().max().item(),
         "depth": (out.coarse_depth_map.detach().cpu() - U.t(z, "ref_depth"))[ok].abs().max().item(),
         "colors": (out.coarse_colors.detach().cpu().reshape(R, -1, 3) - U.t(z, "ref_colors").reshape(R, -1, 3))[ok].abs().max().item()}
    dd = out.directional_derivtives
    assert dd is not None and dd.shape == (4 * R * Nc,) and not dd.requires_grad
    d["dir_derivs(rel, 99.9%)"] = _rel_q(dd.cpu().numpy(), z["ref_dir_derivs"])
    dd_max = _rel(dd.cpu().numpy(), z["ref_dir_derivs"])
    print(name, d, "dir_derivs max rel", dd_max)
    dd_q = d.pop("dir_derivs(rel, 99.9%)")
    assert max(d.values()) <= 1e-3 and dd_q <= 5e-3 and dd_max <= 2e-2, (d, dd_q, dd_max)
    assert torch.equal(dd[:2 * R * Nc], dd[2 * R * Nc:])          # vector_field_nerf.py:305 duplicates the coarse block
    # running statistics after the call: VF net folded two batches (coarse pass, merged pass), colour net one
    for tag, net in (("after_vf.", model.vector_field_network), ("after_rn.", model.rendering_network)):
        sd = net.state_dict()
        for k in z.files:
            if k.startswith(tag):
                key = k[len(tag):]
                if "num_batches" in key:
                    assert int(sd[key]) == int(z[k]), (key, int(sd[key]), int(z[k]))
                else:
                    assert _rel(sd[key].cpu().numpy(), z[k]) <= 1e-4, key
    # the trainer's loss with a directional-derivative weight, and the backward through the batch statistics
    loss = U.O.vf_loss(out.coarse_rgb_values, out.coarse_depth_map, out.coarse_normals.reshape(-1, 3),
                       U.t(z, "rgb_gt").to(DEV), U.t(z, "depth_gt").to(DEV), LOSS_W, 0.5) + \
        LOSS_W["directional_derivatives"] * dd.mean()
    assert abs(loss.item() - float(z["ref_loss"])) <= 2e-3 * abs(float(z["ref_loss"]))
    loss.backward()
    worst, checked = 0.0, 0
    for tag, net in (("g_vf.", model.vector_field_network), ("g_rn.", model.rendering_network)):
        for k, p in net.named_parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all(), k
            g = p.grad.cpu().numpy()
            if tag + k in z.files:
                worst = max(worst, _rel(g, z[tag + k]))
                checked += 1
            else:                                       # full-size weight matrices: norm + 256 strided samples
                assert abs(np.linalg.norm(g) - float(z["n_" + tag + k])) <= 5e-3 * float(z["n_" + tag + k]), k
                s = g.reshape(-1)[:: max(1, g.size // 256)][:256]
                worst = max(worst, _rel(s, z["s_" + tag + k], floor=1e-3 * float(np.abs(g).max())))
                checked += 1
    for k, p in model.density.named_parameters():
        worst = max(worst, abs(p.grad.item() - float(z["g_density." + k])) / max(abs(float(z["g_density." + k])), 1e-4))
    print(name, "worst relative gradient deviation", worst, "over", checked, "tensors")
    assert worst <= 1e-2, worst


@pytest.mark.parametrize("name", ["train_small", "train_full"])
def test_train_mode_vf_query_matches_reference(built_lib, name):
    """The trainer's call on its supervision points (train/vector_field_nerf_train.py:191,204,217) in train mode:
    [y, Jacobian] and the gradients of an MSE on the first three columns."""
    case, z, st, model = _model(name)
    net = model.vector_field_network
    pts, gt = U.t(z, "sup_pts").to(DEV), U.t(z, "sup_gt").to(DEV)
    Do = 3 + case["feat"]
    model.optimizer.zero_grad()
    q = net(pts)
    assert q.shape == (pts.shape[0], Do + 9)
    ref = U.t(z, "ref_sup_out")
    mine = q.detach().cpu()
    if ref.shape[1] != Do + 9:                                # full case: 8 output columns + the 9 Jacobian columns
        mine = mine[:, list(range(8)) + list(range(Do, Do + 9))]
    ny = ref.shape[1] - 9
    assert (mine[:, :ny] - ref[:, :ny]).abs().max().item() <= 1e-3
    assert _rel_q(mine[:, ny:].numpy(), ref[:, ny:].numpy()) <= 5e-3 and _rel(mine[:, ny:].numpy(), ref[:, ny:].numpy()) <= 2e-2
    ((q[:, :3] - gt) ** 2).mean().backward()
    worst = 0.0
    for k, p in net.named_parameters():
        g = p.grad.cpu().numpy()
        if "gq_vf." + k in z.files:
            worst = max(worst, _rel(g, z["gq_vf." + k]))
        else:
            s = g.reshape(-1)[:: max(1, g.size // 256)][:256]
            worst = max(worst, _rel(s, z["s_gq_vf." + k], floor=1e-3 * float(np.abs(g).max())))
    print(name, "VF-only train call: worst relative gradient deviation", worst)
    assert worst <= 1e-2, worst
    sd = net.state_dict()
    for k in z.files:
        if k.startswith("afterq_vf."):
            key = k[len("afterq_vf."):]
            if "num_batches" in key:
                assert int(sd[key]) == int(z[k])
            else:
                assert _rel(sd[key].cpu().numpy(), z[k]) <= 1e-4, key


def test_train_mode_follows_the_oracle_on_a_larger_batch(built_lib):
    """1024 rays x (32 + 32) samples: several slabs per column reduction (the goldens fit in one or two)."""
    case, z, st, model = _model("train_small")
    case = dict(case, n_rays=640, n_coarse=40, n_fine=24)
    model = U.make_model(case, st, DEV)
    model.train()
    R = case["n_rays"]
    uv, pose, K = U.S.synthetic_rays(R, seed=11, start=100, stride=331)
    U1, U2, U3 = U.S.synthetic_draws(R, case["n_coarse"], case["n_fine"], seed=5)
    t_vals = torch.linspace(0., 1., steps=case["n_coarse"])
    with torch.no_grad():
        out = model.render(pose.to(DEV), uv.to(DEV), K.to(DEV), 0, draws=(U1, U2, U3))
        ora = U.O.render_train(st["vf_net"], st["rendering_net"], st["density"], U.oracle_cfg(case), uv, pose, K, t_vals,
                               U1, U2, U3)
    same = (out.z_vals.cpu() == ora["z_vals"]).all(dim=1)
    assert same.float().mean().item() >= 0.9
    ok = same & U.discontinuity_guard(ora, case)
    assert (out.coarse_normals.cpu() - ora["normals"])[ok].abs().max().item() <= 1e-3
    assert (out.coarse_rgb_values.cpu() - ora["rgb"])[ok].abs().max().item() <= 1e-3
    assert (out.coarse_depth_map.cpu() - ora["depth"])[ok].abs().max().item() <= 2e-3
    assert _rel_q(out.directional_derivtives.cpu().numpy(), ora["directional_derivatives"].numpy()) <= 2e-3
    sd = model.vector_field_network.state_dict()
    for k, v in ora["vf_sd_after"].items():
        if "running" in k:
            assert _rel(sd[k].cpu().numpy(), v.numpy()) <= 1e-4, k
