"""GPU: the reference trainer's inner loop (tests/trainer_harness.py = train/vector_field_nerf_train.py:169-260) driven
against vfnerf_b200 unchanged -- its own optimizer / scheduler objects, clip_grad_norm_ over model.parameters() (VF tensors
listed twice), numpy-seeded supervision points, dict-in / (loss, dict)-out VFLoss -- in both modes the trainer uses:
eval() when the directional-derivative weight is 0 (:140-141), train() otherwise."""
import types
import warnings

import numpy as np
import pytest
import torch

import trainer_harness as H
import vfn_testutil as U
from vfnerf_b200 import functions as VF
from vfnerf_b200.losses import VFLoss

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")


class _Dataset:
    """The four members of the reference datasets the loop body touches."""

    def __init__(self, init, far, centroid, white=False):
        self._init, self._far, self._c, self.white_bkgd = init, far, centroid, white

    def get_vf_init_method(self):
        return (self._init, "")

    def get_bounds(self):
        return (0.0, self._far)

    def get_centroid(self, device):
        return self._c.to(device)


def _trainer(case_name, precision, init, dd_weight, train_mode):
    case, z = U.load_golden(case_name)
    model = U.make_model(case, U.case_state(case, z), DEV, precision=precision)
    model.config.border_supervision, model.config.center_supervision = True, True
    w = types.SimpleNamespace(rgb=2.0, depth=0.5, unit_norm=0.1, supervision=1.0, norm_smaller_than_one=0.1,
                              directional_derivatives=dd_weight)
    loss = VFLoss(types.SimpleNamespace(norm_smaller_than_one_start=11000, depth_loss_clamp=0.5,
                                        directional_derivatives_start=0), w)
    cfg = types.SimpleNamespace(vf_nerf_config=model.config,
                                dataset_config=types.SimpleNamespace(dataset_name="replica", border_radius=0.5))
    model.config.cuda_config.device = DEV
    tr = types.SimpleNamespace(model=model, loss=loss, functions=VF, config=cfg,
                               dataset=_Dataset(init, case["far"], torch.zeros(3)))
    if train_mode:
        model.train()                       # the trainer only calls eval() when the dd weight is 0 (:140-141)
    else:
        model.eval()
    R = case["n_rays"]
    data = {"uv": U.t(z, "uv")[None], "intrinsics": U.t(z, "K")[None], "pose": U.t(z, "pose")[None],
            "rgb": U.t(z, "rgb_gt")[None], "depth": U.t(z, "depth_gt")[None]}
    assert data["uv"].shape == (1, R, 2)
    return case, z, tr, data


@pytest.mark.parametrize("precision,init,dd,train_mode", [("fp32", "exterior", 0.0, False), ("bf16", "center", 0.0, False),
                                                          ("fp32", "exterior", 0.05, True), ("fp32", "center", 0.05, True)])
def test_reference_trainer_loop_runs_unchanged(built_lib, precision, init, dd, train_mode):
    # the tensor-core path is built for the shipped 256-wide nets: the bf16 leg runs the full-size golden model
    case, z, tr, data = _trainer("small_perturb" if precision == "fp32" else "full_perturb", precision, init, dd, train_mode)
    torch.manual_seed(5); np.random.seed(5)
    before = [p.detach().clone() for p in tr.model.vector_field_network.parameters()]
    lr0 = tr.model.optimizer.param_groups[0]["lr"]
    losses = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # torch warns about the duplicated parameters, as it does upstream
        for it in range(4):
            loss, terms, out = H.train_step(tr, data, epoch=it)
            losses.append(loss.item())
            assert set(terms) == {"rgb_loss", "depth_loss", "unit_norm_loss", "supervision_loss",
                                  "norm_smaller_than_one_loss", "directional_derivatives_loss"}
            assert terms["supervision_loss"] > 0
            assert (out.directional_derivtives is not None) == train_mode
            if train_mode:
                assert terms["directional_derivatives_loss"] > 0
    assert all(np.isfinite(losses)), losses
    after = list(tr.model.vector_field_network.parameters())
    assert any(not torch.equal(a, b) for a, b in zip(after, before))
    assert tr.model.optimizer.param_groups[0]["lr"] < lr0                   # ExponentialLR stepped
    if train_mode:
        bn = tr.model.vector_field_network.layers[0][1]
        # per iteration: two render() passes + the supervision-point calls of the branch taken (:191 / :204 + :217)
        per_it = 2 + (1 if init == "center" else 2)
        assert int(bn.num_batches_tracked) == 4 * per_it
    # the rgb + depth terms respond to training: the data term of the last step is below the first one's
    # (4 Adam steps at lr 5e-4 on 24 rays: a weak but deterministic check)
    assert losses[-1] <= losses[0] * 1.05


def test_first_step_matches_the_oracle(built_lib):
    """Same loop body, first iteration, eval mode, fp32: the loss equals the oracle's (render -> supervision helpers ->
    VFLoss) when both consume the same generators (torch CPU generator for the sampler draws in the order U1, U2, U3;
    numpy's for the sphere samplers in the order phi, cos_theta, u)."""
    from oracle import supervision_oracle as SO
    case, z, tr, data = _trainer("small_perturb", "fp32", "exterior", 0.0, False)
    R, Nc, nf = case["n_rays"], case["n_coarse"], min(case["n_fine"], case["max_samples"])
    torch.manual_seed(9); np.random.seed(9)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        loss, terms, out = H.train_step(tr, data, epoch=0)
    # the oracle, consuming the generators identically
    torch.manual_seed(9); np.random.seed(9)
    U1, U2, U3 = torch.rand([R, Nc]), torch.rand([R, nf]), torch.rand([R, nf])
    st = U.case_state(case, z)
    with torch.no_grad():
        ora = U.O.render(st["vf_net"], st["rendering_net"], st["density"], U.oracle_cfg(case), U.t(z, "uv"), U.t(z, "pose"),
                         U.t(z, "K"), U.t(z, "t_vals"), U1, U2, U3)
        assert torch.equal(out.z_vals.cpu(), ora["z_vals"])
        n_extra = (R * ora["z_vals"].shape[1]) // 10
        c = torch.zeros(3)
        far, rad = case["far"], 0.5
        draw = lambda: (np.random.uniform(0.0, 2.0 * np.pi, n_extra), np.random.uniform(-1.0, 1.0, n_extra),      # noqa: E731
                        np.random.uniform(0.0, 1.0, n_extra))
        bp, bgt = SO.sample_border_points(far - 5 * rad, far, c, *draw())
        sup = [U.O.vf_network(st["vf_net"], bp)[:, :3]]
        gt = [bgt]
        n_sel, g_sel = SO.get_center_indices_and_gt(ora["points"], ora["normals"], c, rad)
        cp, cgt = SO.sample_center_points(c, rad, *draw())
        sup += [n_sel, U.O.vf_network(st["vf_net"], cp)[:, :3]]
        gt += [g_sel, cgt]
        ref = U.O.vf_loss(ora["rgb"], ora["depth"], ora["normals"].reshape(-1, 3), U.t(z, "rgb_gt"), U.t(z, "depth_gt"),
                          U.LOSS_W, 0.5, supervised=torch.cat(sup, 0), supervised_gt=torch.cat(gt, 0))
    assert abs(loss.item() - ref.item()) <= 2e-4 * max(1.0, abs(ref.item())), (loss.item(), ref.item())
