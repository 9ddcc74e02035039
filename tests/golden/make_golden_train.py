"""Golden vectors of the TRAIN-MODE path from the LIVE reference.

    python tests/golden/make_golden_train.py          (build container only: needs /root/reference)

``model.train()`` puts both networks into BatchNorm batch-statistics mode and makes the VF net return the autograd
"Jacobian" next to its outputs (vector_field_network.py:140-175); render() then also returns directional derivatives
(vector_field_nerf.py:264-270,301-305,476-498).  This script runs the unmodified reference that way on CPU -- render(),
the trainer's loss (with a non-zero directional-derivative weight) and backward, and the VF-only call the trainer makes on
its supervision points (train/vector_field_nerf_train.py:191,204,217) -- checks the oracle restatement
(oracle/render_oracle.py: render_train & co) against it, and writes inputs + reference outputs to train_small.npz /
train_full.npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G                                     # noqa: E402  (sets sys.path for the reference + repo)
from make_golden import O, S                                # noqa: E402
from models.losses.vf_loss import VFLoss                    # noqa: E402
from config_parser.vf_nerf_config import VFLossConfig, VFLossWeights   # noqa: E402

CASES = {
    "train_small": dict(seed=5, vf_hidden=(64,) * 8, feat=32, rn_hidden=(64,) * 4, n_rays=24, n_coarse=24,
                        n_fine=20, max_samples=100, perturb=True, near=0.25, far=5.0, fine_range=0.4,
                        window=11, dir_to_normal_th=-2.0, vf_gain=2.0, start=4321, stride=7919, n_sup=200),
    # shipped network shape, few rays (CPU autograd through 8 x 256 batch-stat layers three times)
    "train_full": dict(seed=1, vf_hidden=(256,) * 8, feat=256, rn_hidden=(256,) * 4, n_rays=8, n_coarse=64,
                       n_fine=64, max_samples=100, perturb=True, near=0.0, far=6.0, fine_range=0.3,
                       window=11, dir_to_normal_th=-2.0, vf_gain=2.0, start=999, stride=25013, n_sup=256),
}
LOSS_W = dict(G.LOSS_W, directional_derivatives=0.05)


def randomised_bn(st, seed):
    """Non-default affine parameters and running statistics, so that their update and their gradients are exercised."""
    g = torch.Generator().manual_seed(seed)
    for net in ("vf_net", "rendering_net"):
        for k in list(st[net]):
            v = st[net][k]
            if k.endswith(".1.weight"):
                st[net][k] = 0.75 + 0.5 * torch.rand(v.shape, generator=g)
            elif k.endswith(".1.bias"):
                st[net][k] = 0.2 * torch.randn(v.shape, generator=g)
            elif k.endswith("running_mean"):
                st[net][k] = 0.1 * torch.randn(v.shape, generator=g)
            elif k.endswith("running_var"):
                st[net][k] = 0.5 + torch.rand(v.shape, generator=g)
    return st


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    for name, case in CASES.items():
        model, st = G.make_ref_model(case)
        st = randomised_bn(st, 77 + case["seed"])
        model.vector_field_network.load_state_dict(st["vf_net"])
        model.rendering_network.load_state_dict(st["rendering_net"])
        model.train()
        assert model.vector_field_network.training and model.rendering_network.training
        R = case["n_rays"]
        uv, pose, K = S.synthetic_rays(R, seed=case["seed"], start=case["start"], stride=case["stride"])
        nf = min(case["n_fine"], case["max_samples"])
        U1, U2, U3 = S.synthetic_draws(R, case["n_coarse"], nf, seed=1234 + case["seed"])
        t_vals = torch.linspace(0., 1., steps=case["n_coarse"])
        gen = np.random.default_rng(55 + case["seed"])
        rgb_gt = torch.from_numpy(gen.random((R, 3), dtype=np.float32))
        depth_gt = torch.from_numpy(gen.random((R, 1), dtype=np.float32)) * case["far"]

        # ---- the reference: render (train mode) -> VFLoss -> backward
        for p in model.parameters():
            p.grad = None
        ref = G.run_reference(model, case, uv, pose, K, U1, U2, U3, with_grad=True)
        assert ref.directional_derivtives is not None and not ref.directional_derivtives.requires_grad
        assert ref.directional_derivtives.shape == (4 * R * case["n_coarse"],)
        lossmod = VFLoss(VFLossConfig(norm_smaller_than_one_start=11000, depth_loss_clamp=0.5, directional_derivatives_start=0),
                         VFLossWeights(**LOSS_W))
        pred = {"rgb": ref.coarse_rgb_values, "depth": ref.coarse_depth_map, "normals": ref.coarse_normals.reshape(-1, 3),
                "supervised_normals": torch.empty(0, 3), "directional_derivatives": ref.directional_derivtives}
        loss, terms = lossmod(pred, {"rgb": rgb_gt, "depth": depth_gt, "supervised_normals": torch.empty(0)}, 0)
        loss.backward()
        grads = {}
        for k, p in model.vector_field_network.named_parameters():
            grads["g_vf." + k] = p.grad.numpy().copy()
        for k, p in model.rendering_network.named_parameters():
            grads["g_rn." + k] = p.grad.numpy().copy()
        for k, p in model.density.named_parameters():
            grads["g_density." + k] = (p.grad if p.grad is not None else torch.zeros(())).numpy().copy()
        vf_after = {k: v.detach().clone() for k, v in model.vector_field_network.state_dict().items()}
        rn_after = {k: v.detach().clone() for k, v in model.rendering_network.state_dict().items()}

        # ---- the oracle on the same inputs
        ocfg = G.oracle_cfg(case)
        req = lambda sd: {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())  # noqa: E731
                          for k, v in sd.items()}
        vf_sd, rn_sd = req(st["vf_net"]), req(st["rendering_net"])
        dn = {k: v.clone().requires_grad_(True) for k, v in st["density"].items()}
        mine = O.render_train(vf_sd, rn_sd, dn, ocfg, uv, pose, K, t_vals, U1, U2, U3)
        same_z = (mine["z_vals"] == ref.z_vals).all(dim=1).float().mean().item()
        dev = {
            "normals": (mine["normals"] - ref.coarse_normals).abs().max().item(),
            "rgb": (mine["rgb"] - ref.coarse_rgb_values).abs().max().item(),
            "depth": (mine["depth"] - ref.coarse_depth_map).abs().max().item(),
            "colors": (mine["colors"] - ref.coarse_colors).abs().max().item(),
            "dir_deriv(rel)": ((mine["directional_derivatives"] - ref.directional_derivtives).abs().max() /
                               ref.directional_derivtives.abs().max()).item(),
        }
        for tag, after, mine_after in (("vf", vf_after, mine["vf_sd_after"]), ("rn", rn_after, mine["rn_sd_after"])):
            for k, v in after.items():
                if "running" in k:
                    dev[f"{tag}.running"] = max(dev.get(f"{tag}.running", 0.0),
                                                ((mine_after[k] - v).abs().max() / (v.abs().max() + 1e-12)).item())
                elif "num_batches" in k:
                    assert int(mine_after[k]) == int(v), (k, mine_after[k], v)
        print(f"[{name}] z_vals identical on {100 * same_z:.0f}% of rays; oracle-vs-reference max abs dev {dev}; sigma>0 on "
              f"{100 * (mine['sigma'] > 0).float().mean().item():.2f}% samples; dir-derivative mean "
              f"{ref.directional_derivtives.mean().item():.4f}")
        assert same_z == 1.0 and max(dev.values()) <= 2e-4, (name, dev)
        ol = O.vf_loss(mine["rgb"], mine["depth"], mine["normals"].reshape(-1, 3), rgb_gt, depth_gt, LOSS_W, 0.5) + \
            LOSS_W["directional_derivatives"] * mine["directional_derivatives"].mean()
        ol.backward()
        assert abs(ol.item() - loss.item()) < 2e-5, (ol.item(), loss.item())
        worst = 0.0
        for tag, sd in (("g_vf.", vf_sd), ("g_rn.", rn_sd)):
            for k, v in sd.items():
                if v.requires_grad:
                    g = grads[tag + k]
                    gm = np.zeros_like(g) if v.grad is None else v.grad.numpy()
                    # Linear biases in front of a batch-statistics BatchNorm have a mathematically zero gradient: rounding noise
                    scale = max(np.abs(g).max(), 1e-4)
                    worst = max(worst, np.abs(gm - g).max() / scale)
                    if os.environ.get("VERBOSE"):
                        print(f"   {tag + k}: ref max {np.abs(g).max():.3e}  dev {np.abs(gm - g).max():.3e}")
        print(f"[{name}] loss {loss.item():.6f} (dd term {terms['directional_derivatives_loss']:.5f}); oracle-vs-reference "
              f"worst relative grad dev {worst:.2e}")
        assert worst < 5e-3, (name, worst)

        # ---- the VF-only call of the trainer on supervision points, train mode: [y, jacobian] and grads of an MSE on [:, :3]
        gsup = torch.Generator().manual_seed(91 + case["seed"])
        sup_pts = (torch.rand(case["n_sup"], 3, generator=gsup) - 0.5) * 4.0
        sup_gt = torch.nn.functional.normalize(torch.randn(case["n_sup"], 3, generator=gsup), dim=1)
        model.vector_field_network.load_state_dict(st["vf_net"])
        for p in model.parameters():
            p.grad = None
        q = model.vector_field_network(sup_pts)
        Do = 3 + case["feat"]
        assert q.shape == (case["n_sup"], Do + 9)
        ((q[:, :3] - sup_gt) ** 2).mean().backward()
        q_grads = {"gq_vf." + k: p.grad.numpy().copy() for k, p in model.vector_field_network.named_parameters()}
        q_after = {k: v.detach().clone() for k, v in model.vector_field_network.state_dict().items()}
        stq: dict = {}
        yq, jq = O.vf_network_train_with_jacobian({k: v.detach() for k, v in st["vf_net"].items()}, sup_pts, 6, (4,), stq)
        dq = {"y": (yq - q[:, :Do]).abs().max().item(),
              "jac(rel)": ((jq - q[:, Do:]).abs().max() / q[:, Do:].abs().max()).item()}
        print(f"[{name}] VF-only train call: oracle-vs-reference {dq}")
        assert max(dq.values()) <= 2e-4, dq

        fx = dict(
            case=np.array(repr(case)), uv=uv.numpy(), pose=pose.numpy(), K=K.numpy(), t_vals=t_vals.numpy(),
            U1=U1.numpy(), U2=U2.numpy(), U3=U3.numpy(), rgb_gt=rgb_gt.numpy(), depth_gt=depth_gt.numpy(),
            ref_z_vals=ref.z_vals.numpy(), ref_normals=ref.coarse_normals.detach().numpy(),
            ref_rgb=ref.coarse_rgb_values.detach().numpy(), ref_depth=ref.coarse_depth_map.detach().numpy(),
            ref_colors=ref.coarse_colors.detach().numpy(), ref_dir_derivs=ref.directional_derivtives.numpy(),
            ref_loss=np.float32(loss.item()), sup_pts=sup_pts.detach().numpy(), sup_gt=sup_gt.numpy(),
            ref_sup_out=q.detach().numpy() if name == "train_small" else q.detach().numpy()[:, list(range(8)) + list(range(Do, Do + 9))],
        )
        small = name == "train_small"
        for tag, sd in (("w_vf.", st["vf_net"]), ("w_rn.", st["rendering_net"])):
            for k, v in sd.items():
                if small or ".1." in k:          # full-size Linear weights are regenerated from the seed; BN tensors are kept
                    fx[tag + k] = v.numpy()
        for tag, sd in (("after_vf.", vf_after), ("after_rn.", rn_after), ("afterq_vf.", q_after)):
            for k, v in sd.items():
                if "running" in k or "num_batches" in k:
                    fx[tag + k] = v.numpy()
        for k, g in {**grads, **q_grads}.items():
            if small or k.startswith("g_density") or g.ndim == 1 or g.size <= 1024:
                fx[k] = g
            else:
                fx["n_" + k] = np.float32(np.linalg.norm(g))
                fx["s_" + k] = g.reshape(-1)[:: max(1, g.size // 256)][:256].copy()
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **fx)
        print(f"[{name}] wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
