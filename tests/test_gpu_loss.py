"""GPU: the fused VFLoss (csrc/loss.cu, vfnerf_b200/losses.py) against the oracle's restatement of the reference's
VFLoss.forward (models/losses/vf_loss.py:34-87), values and gradients, and against the reference's own loss value stored
in the golden fixtures."""
import types

import pytest
import torch

import vfn_testutil as U
from vfnerf_b200.losses import VFLoss

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _mk(weights=None, clamp=0.5, lt1_start=11000, dd_start=100, sync=True):
    w = dict(U.LOSS_W) if weights is None else weights
    return VFLoss(types.SimpleNamespace(norm_smaller_than_one_start=lt1_start, depth_loss_clamp=clamp,
                                        directional_derivatives_start=dd_start), types.SimpleNamespace(**w), sync=sync)


@pytest.mark.parametrize("n_sup,epoch,with_depth", [(0, 0, True), (777, 0, True), (777, 20000, True), (5, 20000, False)])
def test_fused_loss_matches_oracle(built_lib, n_sup, epoch, with_depth):
    g = torch.Generator().manual_seed(n_sup + epoch)
    R, N = 257, 33
    rgb, rgb_gt = torch.rand(R, 3, generator=g), torch.rand(R, 3, generator=g)
    depth, depth_gt = torch.rand(R, 1, generator=g) * 3, torch.rand(R, 1, generator=g) * 3
    normals = torch.randn(R * N, 3, generator=g) * 0.8
    normals[5] = 0.0                                         # zero vector: torch.norm's subgradient is 0
    sup, sup_gt = torch.randn(n_sup, 3, generator=g), torch.randn(n_sup, 3, generator=g)

    def leaf(t, dev):
        return t.clone().to(dev).requires_grad_(True)
    # oracle (CPU, autograd)
    a = [leaf(t, "cpu") for t in (rgb, depth, normals, sup)]
    ref = U.O.vf_loss(a[0], a[1], a[2], rgb_gt, depth_gt, U.LOSS_W if with_depth else dict(U.LOSS_W, depth=0.0), 0.5,
                      supervised=a[3] if n_sup else None, supervised_gt=sup_gt, epoch=epoch)
    ref.backward()
    # fused (GPU)
    b = [leaf(t, DEV) for t in (rgb, depth, normals, sup)]
    loss_fn = _mk()
    pred = {"rgb": b[0], "depth": b[1], "normals": b[2], "supervised_normals": b[3], "directional_derivatives": None}
    gt = {"rgb": rgb_gt.to(DEV), "depth": depth_gt.to(DEV) if with_depth else torch.empty(0, device=DEV),
          "supervised_normals": sup_gt.to(DEV)}
    loss, terms = loss_fn(pred, gt, epoch)
    (loss * 1.7).backward()                                  # a non-unit upstream gradient
    assert abs(loss.item() - ref.item()) <= 2e-6 * max(1.0, abs(ref.item()))
    assert set(terms) == {"rgb_loss", "depth_loss", "unit_norm_loss", "supervision_loss", "norm_smaller_than_one_loss",
                          "directional_derivatives_loss"} and all(isinstance(v, float) for v in terms.values())
    assert (terms["norm_smaller_than_one_loss"] > 0) == (epoch >= 11000)
    for x, y, name in zip(a, b, ("rgb", "depth", "normals", "sup")):
        if name == "sup" and n_sup == 0:
            continue
        want = x.grad * 1.7 if x.grad is not None else torch.zeros_like(x)
        got = y.grad.cpu() if y.grad is not None else torch.zeros_like(x)
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-9), name


def test_fused_loss_reproduces_the_reference_loss_value(built_lib):
    """render() on the fp32 path + fused VFLoss == the loss the reference computed on CPU (golden fixture)."""
    case, z = U.load_golden("full_det")
    model = U.make_model(case, U.case_state(case, z), DEV)
    draws = (U.t(z, "U1"), U.t(z, "U2"), U.t(z, "U3"))
    out = model.render(U.t(z, "pose").to(DEV), U.t(z, "uv").to(DEV), U.t(z, "K").to(DEV), 0, draws=draws,
                       z_vals_override=U.t(z, "ref_z_vals"))
    loss, _ = _mk(sync=False)({"rgb": out.coarse_rgb_values, "depth": out.coarse_depth_map,
                               "normals": out.coarse_normals.reshape(-1, 3), "supervised_normals": torch.empty(0, 3, device=DEV),
                               "directional_derivatives": None},
                              {"rgb": U.t(z, "rgb_gt").to(DEV), "depth": U.t(z, "depth_gt").to(DEV),
                               "supervised_normals": torch.empty(0, device=DEV)}, 0)
    assert abs(loss.item() - float(z["ref_loss"])) <= 1e-4
    model.optimizer.zero_grad()
    loss.backward()
    k = "layers.3.0.weight"
    g = dict(model.vector_field_network.named_parameters())[k].grad.cpu().numpy()
    n_ref = float(z["n_g_vf." + k])
    assert abs(float((g ** 2).sum() ** 0.5) - n_ref) <= 2e-3 * n_ref
