"""GPU: supervision-point helpers of the trainer on the device (csrc/supervision.cu, vfnerf_b200/functions.py) against
golden vectors written by the live reference functions (models/helpers/functions.py:75-157, sampler.py:160-193)."""
import os

import numpy as np
import pytest
import torch

import vfn_testutil as U
from vfnerf_b200 import functions as VF

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(U.GOLDEN_DIR, "supervision.npz"))


def test_sphere_samplers_match_reference(built_lib, gold):
    z = gold
    c = torch.from_numpy(z["centroid"])
    far, radius = float(z["far"]), float(z["radius"])
    n = z["border_phi"].shape[0]
    p, g = VF.sample_border_points(far / 2 - radius, far / 2, n, c, DEV, draws=(z["border_phi"], z["border_cos"], z["border_u"]))
    # float64 sphere arithmetic like numpy's, one fp32 rounding at the end: identical up to libm's last ulp
    assert (p.cpu() - torch.from_numpy(z["border_points"])).abs().max().item() <= 5e-7
    assert (g.cpu() - torch.from_numpy(z["border_gt"])).abs().max().item() <= 1e-6
    assert (p.cpu() == torch.from_numpy(z["border_points"])).float().mean().item() >= 0.999
    p, g = VF.sample_center_points(c, radius, n, DEV, draws=(z["center_phi"], z["center_cos"], z["center_u"]))
    assert (p.cpu() - torch.from_numpy(z["center_points"])).abs().max().item() <= 5e-7
    assert (g.cpu() - torch.from_numpy(z["center_gt"])).abs().max().item() <= 2e-6
    # empty request, and the numpy stream: seeding numpy reproduces the reference's own draws (same call order and sizes)
    e, _ = VF.sample_center_points(c, radius, 0, DEV)
    assert e.shape == (0, 3)
    np.random.seed(11)
    a, _ = VF.sample_border_points(1.0, 2.0, 257, c, DEV)
    np.random.seed(11)
    phi, cs, u = np.random.uniform(0, 2 * np.pi, 257), np.random.uniform(-1, 1, 257), np.random.uniform(0, 1, 257)
    from oracle import supervision_oracle as SO
    b, _ = SO.sample_border_points(1.0, 2.0, c, phi, cs, u)
    assert (a.cpu() - b).abs().max().item() <= 5e-7
    # device draws: same distribution -- every point inside its shell
    d, gt = VF.sample_border_points(1.0, 2.0, 100000, c, DEV, on_device=True)
    r = (d.cpu() - c).norm(dim=1)
    assert (r >= 1.0 - 1e-5).all() and (r <= 2.0 + 1e-5).all() and abs((r - 1.0).pow(3).mean().item() - 0.5) < 0.02       # r = cbrt(u) (r_max - r_min) + r_min, sampler.py:177
    assert (gt.norm(dim=1) - 1).abs().max().item() <= 1e-5


def test_ray_sample_selection_matches_reference_and_keeps_autograd(built_lib, gold):
    z = gold
    c = torch.from_numpy(z["centroid"])
    far, radius = float(z["far"]), float(z["radius"])
    pts = torch.from_numpy(z["ray_points"]).to(DEV)
    nrm = torch.from_numpy(z["ray_normals"]).to(DEV).requires_grad_(True)
    n, g = VF.get_border_indices_and_gt(pts, nrm, far, radius, c.to(DEV))
    assert torch.equal(n.detach().cpu(), torch.from_numpy(z["sel_border_normals"]))          # same samples, same order
    assert (g.cpu() - torch.from_numpy(z["sel_border_gt"])).abs().max().item() <= 1e-6
    n2, g2 = VF.get_center_indices_and_gt(pts, nrm, c.to(DEV), radius)
    assert torch.equal(n2.detach().cpu(), torch.from_numpy(z["sel_center_normals"]))
    assert (g2.cpu() - torch.from_numpy(z["sel_center_gt"])).abs().max().item() <= 2e-6
    # the selected vectors are still part of the graph: the MSE supervision term reaches the right rows only
    ((n - g) ** 2).mean().backward()
    touched = (nrm.grad.abs().sum(dim=2) > 0)
    want = (pts.cpu() - c).norm(dim=2) > (far / 2 - radius)
    assert not (touched.cpu() & ~want).any() and want.sum().item() == n.shape[0]


def test_supervised_training_step_like_the_reference_trainer(built_lib):
    """train/vector_field_nerf_train.py:177-216 with the device helpers: render, border + centre supervision through the
    VF-only module call, VFLoss with the supervision term, backward."""
    from vfnerf_b200.losses import VFLoss
    import types
    case, z = U.load_golden("small_perturb")
    model = U.make_model(case, U.case_state(case, z), DEV)
    uv, pose, K = (U.t(z, k).to(DEV) for k in ("uv", "pose", "K"))
    out = model.render(pose, uv, K, 0)
    c = torch.zeros(3, device=DEV)
    P = out.points_coarse.shape[0] * out.points_coarse.shape[1]
    sup, gt = VF.get_center_indices_and_gt(out.points_coarse, out.coarse_normals, c, 1.5)
    bp, bgt = VF.sample_border_points(5.0, 6.0, P // 10, c, DEV)
    cp, cgt = VF.sample_center_points(c, 0.5, P // 10, DEV)
    sup = torch.cat([sup, model.vector_field_network(bp)[:, :3], model.vector_field_network(cp)[:, :3]], 0)
    gt = torch.cat([gt, bgt, cgt], 0)
    loss_mod = VFLoss(types.SimpleNamespace(norm_smaller_than_one_start=11000, depth_loss_clamp=0.5, directional_derivatives_start=100),
                      types.SimpleNamespace(rgb=2.0, depth=0.5, unit_norm=0.1, supervision=1.0, norm_smaller_than_one=0.1,
                                            directional_derivatives=0.0))
    loss = loss_mod({"rgb": out.coarse_rgb_values, "depth": out.coarse_depth_map, "normals": out.coarse_normals.reshape(-1, 3),
                     "supervised_normals": sup, "directional_derivatives": None},
                    {"rgb": U.t(z, "rgb_gt").to(DEV), "depth": U.t(z, "depth_gt").to(DEV), "supervised_normals": gt}, 0)[0]
    ref = U.O.vf_loss(out.coarse_rgb_values, out.coarse_depth_map, out.coarse_normals.reshape(-1, 3), U.t(z, "rgb_gt").to(DEV),
                      U.t(z, "depth_gt").to(DEV), U.LOSS_W, 0.5, supervised=sup, supervised_gt=gt)
    assert abs(loss.item() - ref.item()) <= 1e-5 * max(1.0, abs(ref.item()))
    model.optimizer.zero_grad()
    loss.backward()
    g = model.vector_field_network.layers[0][0].weight.grad
    assert g is not None and torch.isfinite(g).all() and g.abs().max().item() > 0
