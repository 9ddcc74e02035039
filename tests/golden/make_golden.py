"""Generate the golden vectors under tests/golden/ from the LIVE reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It (1) imports the unmodified reference, (2) loads the synthetic weights of
``vfnerf_b200.synthetic`` into it, (3) patches ``torch.rand`` inside the reference's sampler module
so the three uniform draws are the supplied tensors (the reference has no injection point,
ray_sampler.py:138,292,297), (4) runs ``VectorFieldNerf.render()`` on CPU, (5) checks the oracle
restatement (oracle/render_oracle.py) against it -- bit-exact for directions / z_vals / points,
<= 1e-4 abs elsewhere (fp32 op-order noise of BatchNorm and sums) -- and (6) writes inputs + reference outputs as .npz fixtures.
The fixtures are what pins the oracle on machines where the reference is absent.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VFNERF_REF", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import render_oracle as O                       # noqa: E402
from vfnerf_b200 import synthetic as S                      # noqa: E402

import models.samplers.ray_sampler as ref_sampler           # noqa: E402
from config_parser.vf_nerf_config import (CudaConfig, DensityConfig, RaySamplerConfig,   # noqa: E402
                                          RenderingNetConfig, SchedulerConfig, VFNerfConfig, VFNetConfig)
from models.nerf.vector_field_nerf import VectorFieldNerf   # noqa: E402
from models.losses.vf_loss import VFLoss                    # noqa: E402
from config_parser.vf_nerf_config import VFLossConfig, VFLossWeights   # noqa: E402


def make_ref_model(case):
    cfg = VFNerfConfig(
        VFNetConfig(input_dims=3, output_dims=3, dimensions=list(case["vf_hidden"]),
                    feature_vector_dims=case["feat"], embedder_multires=6, weight_norm=False,
                    batch_norm=True, skip_connection_in=[4], bias_init=0.0, dropout=False,
                    dropout_probability=0.2, xavier_init=False, init=""),
        RenderingNetConfig(output_dims=3, dimensions=list(case["rn_hidden"]),
                           feature_vector_dims=case["feat"], weight_norm=False, batch_norm=True,
                           mode="idr", embedder_multires=4, detach_normals=True),
        RaySamplerConfig(n_samples=case["n_coarse"], n_importance=case["n_fine"], rays_per_batch=1024,
                         perturb=case["perturb"], near=case["near"], far=case["far"],
                         fine_range=case["fine_range"], increase_every=50, max_samples=case["max_samples"]),
        CudaConfig(device=torch.device("cpu"), num_gpus=0),
        SchedulerConfig(lr=5e-4, lr_decay_factor=0.1, clip_norm=0.5, weight_decay=0.0),
        DensityConfig(beta_bounds=[1e-4, 1e9], mean_bounds=[0.6, 1.0], scale_min=1.0,
                      params_init={"beta": 0.5, "scale": 100.0, "mean": 0.7}, cutoff=-2.0),
        cos_sim_weights=[0.09] * case["window"], cos_sim_weights_anneal="hard", anneal_start=700,
        anneal_end=1400, rendering="volsdf", normalize_rendering=True,
        dir_to_normal_th=case["dir_to_normal_th"], numerical_jacobian=False)
    model = VectorFieldNerf(cfg)
    st = S.synthetic_state(case["seed"], case["vf_hidden"], case["feat"], case["rn_hidden"],
                           vf_gain=case["vf_gain"])
    model.vector_field_network.load_state_dict(st["vf_net"])
    model.rendering_network.load_state_dict(st["rendering_net"])
    model.density.load_state_dict(st["density"])
    model.eval()
    return model, st


class _Draws:
    """Stand-in for torch.rand inside models.samplers.ray_sampler: hands out the queued tensors."""

    def __init__(self, queue):
        self.queue = list(queue)

    def rand(self, *shape, **kw):
        t = self.queue.pop(0)
        want = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
        assert tuple(t.shape) == want, (t.shape, want)
        return t.clone()


def run_reference(model, case, uv, pose, K, U1, U2, U3, with_grad=False, targets=None):
    queue = [U1, U2, U3] if case["perturb"] else [U3]
    real_torch = ref_sampler.torch

    class _TorchProxy:
        def __getattr__(self, name):
            return getattr(real_torch, name)
    proxy = _TorchProxy()
    proxy.rand = _Draws(queue).rand
    ref_sampler.torch = proxy
    try:
        if with_grad:
            out = model.render(pose, uv, K, epoch=0)
        else:
            with torch.no_grad():
                out = model.render(pose, uv, K, epoch=0)
    finally:
        ref_sampler.torch = real_torch
    return out


def oracle_cfg(case):
    return dict(n_coarse=case["n_coarse"], n_fine=min(case["n_fine"], case["max_samples"]),
                near=case["near"], far=case["far"], fine_range=case["fine_range"],
                perturb=case["perturb"], window=case["window"],
                dir_to_normal_th=case["dir_to_normal_th"], normalize=True,
                beta_bounds=(1e-4, 1e9), scale_min=1.0, mean_bounds=(0.6, 1.0),
                multires=6, multires_view=4, skip_in=(4,))


CASES = {
    # tiny network: weights are stored in the fixture; generic-shape path of the kernels
    "small_det": dict(seed=3, vf_hidden=(64,) * 8, feat=32, rn_hidden=(64,) * 4, n_rays=24, n_coarse=24,
                      n_fine=20, max_samples=100, perturb=False, near=0.0, far=6.0, fine_range=0.3,
                      window=11, dir_to_normal_th=-0.2, vf_gain=2.0, start=1234, stride=7919),
    "small_perturb": dict(seed=4, vf_hidden=(64,) * 8, feat=32, rn_hidden=(64,) * 4, n_rays=24, n_coarse=32,
                          n_fine=40, max_samples=36, perturb=True, near=0.25, far=5.0, fine_range=0.5,
                          window=11, dir_to_normal_th=-2.0, vf_gain=2.0, start=99, stride=4099),
    # shipped network shape (confs/vf_nerf.conf:13-37) at BASELINE's 64+64 samples
    "full_det": dict(seed=0, vf_hidden=(256,) * 8, feat=256, rn_hidden=(256,) * 4, n_rays=32, n_coarse=64,
                     n_fine=64, max_samples=100, perturb=False, near=0.0, far=6.0, fine_range=0.3,
                     window=11, dir_to_normal_th=-0.2, vf_gain=2.0, start=40000, stride=25013),
    "full_perturb": dict(seed=0, vf_hidden=(256,) * 8, feat=256, rn_hidden=(256,) * 4, n_rays=32, n_coarse=64,
                         n_fine=64, max_samples=100, perturb=True, near=0.0, far=6.0, fine_range=0.3,
                         window=11, dir_to_normal_th=-2.0, vf_gain=2.0, start=777, stride=25013),
}

LOSS_W = dict(rgb=2.0, depth=0.5, unit_norm=0.1, supervision=1.0, norm_smaller_than_one=0.1,
              directional_derivatives=0.0)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    for name, case in CASES.items():
        model, st = make_ref_model(case)
        R = case["n_rays"]
        uv, pose, K = S.synthetic_rays(R, seed=case["seed"], start=case["start"], stride=case["stride"])
        nf = min(case["n_fine"], case["max_samples"])
        U1, U2, U3 = S.synthetic_draws(R, case["n_coarse"], nf, seed=1234 + case["seed"])
        t_vals = torch.linspace(0., 1., steps=case["n_coarse"])

        ref = run_reference(model, case, uv, pose, K, U1, U2, U3)
        ocfg = oracle_cfg(case)
        with torch.no_grad():
            mine = O.render(st["vf_net"], st["rendering_net"], st["density"], ocfg, uv, pose, K, t_vals,
                            U1, U2, U3)

        # --- pin the oracle against the live reference
        assert torch.equal(mine["z_vals"], ref.z_vals), f"{name}: z_vals not bit-exact"
        assert torch.equal(mine["points"], ref.points_coarse), f"{name}: points not bit-exact"
        dev = {
            "normals": (mine["normals"] - ref.coarse_normals).abs().max().item(),
            "rgb": (mine["rgb"] - ref.coarse_rgb_values).abs().max().item(),
            "depth": (mine["depth"] - ref.coarse_depth_map).abs().max().item(),
            "colors": (mine["colors"] - ref.coarse_colors).abs().max().item(),
            "ray_dirs": (mine["rep_ray_dirs"] - ref.ray_dirs).abs().max().item(),
        }
        frac = (mine["sigma"] > 0).float().mean().item()
        print(f"[{name}] oracle-vs-reference max abs dev: {dev}; sigma>0 on {100 * frac:.2f}% samples; "
              f"rgb mean {ref.coarse_rgb_values.mean().item():.4f} depth mean {ref.coarse_depth_map.mean().item():.4f}")
        assert max(dev.values()) <= 1e-4, (name, dev)
        assert frac > 0.002, f"{name}: degenerate density"

        # --- gradient golden: reference render -> reference VFLoss -> backward
        gen = np.random.default_rng(55 + case["seed"])
        rgb_gt = torch.from_numpy(gen.random((R, 3), dtype=np.float32))
        depth_gt = torch.from_numpy(gen.random((R, 1), dtype=np.float32)) * case["far"]
        for p in model.parameters():
            p.grad = None
        out_g = run_reference(model, case, uv, pose, K, U1, U2, U3, with_grad=True)
        lossmod = VFLoss(VFLossConfig(norm_smaller_than_one_start=11000, depth_loss_clamp=0.5),
                         VFLossWeights(**LOSS_W))
        pred = {"rgb": out_g.coarse_rgb_values, "depth": out_g.coarse_depth_map,
                "normals": out_g.coarse_normals.reshape(-1, 3),
                "supervised_normals": torch.empty(0, 3), "directional_derivatives": None}
        loss, _ = lossmod(pred, {"rgb": rgb_gt, "depth": depth_gt, "supervised_normals": torch.empty(0)}, 0)
        loss.backward()
        grads = {}
        for k, p in model.vector_field_network.named_parameters():
            grads["g_vf." + k] = p.grad.numpy().copy()
        for k, p in model.rendering_network.named_parameters():
            grads["g_rn." + k] = p.grad.numpy().copy()
        for k, p in model.density.named_parameters():
            grads["g_density." + k] = (p.grad if p.grad is not None else torch.zeros(())).numpy().copy()

        # oracle gradient check (autograd through the restatement)
        vf_sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
                 for k, v in st["vf_net"].items()}
        rn_sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
                 for k, v in st["rendering_net"].items()}
        dn = {k: v.clone().requires_grad_(True) for k, v in st["density"].items()}
        og = O.render(vf_sd, rn_sd, dn, ocfg, uv, pose, K, t_vals, U1, U2, U3)
        ol = O.vf_loss(og["rgb"], og["depth"], og["normals"].reshape(-1, 3), rgb_gt, depth_gt, LOSS_W, 0.5)
        ol.backward()
        assert abs(ol.item() - loss.item()) < 1e-5, (ol.item(), loss.item())
        worst = 0.0
        for k, v in vf_sd.items():
            if isinstance(v, torch.Tensor) and v.requires_grad:
                g = grads["g_vf." + k]
                worst = max(worst, np.abs(v.grad.numpy() - g).max() / (np.abs(g).max() + 1e-12))
        for k, v in rn_sd.items():
            if isinstance(v, torch.Tensor) and v.requires_grad:
                g = grads["g_rn." + k]
                worst = max(worst, np.abs(v.grad.numpy() - g).max() / (np.abs(g).max() + 1e-12))
        for k, v in dn.items():
            g = grads["g_density." + k]
            worst = max(worst, abs(v.grad.item() - float(g)) / (abs(float(g)) + 1e-12))
        print(f"[{name}] loss {loss.item():.6f}; oracle-vs-reference worst relative grad dev {worst:.2e}")
        assert worst < 2e-3, (name, worst)

        fx = dict(
            case=np.array(repr(case)),
            uv=uv.numpy(), pose=pose.numpy(), K=K.numpy(), t_vals=t_vals.numpy(),
            U1=U1.numpy(), U2=U2.numpy(), U3=U3.numpy(),
            ref_z_vals=ref.z_vals.numpy(), ref_points=ref.points_coarse.numpy(),
            ref_normals=ref.coarse_normals.numpy(), ref_rgb=ref.coarse_rgb_values.numpy(),
            ref_depth=ref.coarse_depth_map.numpy(), ref_colors=ref.coarse_colors.numpy(),
            ref_ray_dirs=ref.ray_dirs.numpy(),
            rgb_gt=rgb_gt.numpy(), depth_gt=depth_gt.numpy(), ref_loss=np.float32(loss.item()),
        )
        if name.startswith("small"):
            for k, v in st["vf_net"].items():
                fx["w_vf." + k] = v.numpy()
            for k, v in st["rendering_net"].items():
                fx["w_rn." + k] = v.numpy()
            fx.update(grads)
        else:
            # full-size weights are regenerated from the seed (vfnerf_b200.synthetic); keep only the
            # density-parameter grads and per-tensor gradient norms to bound the fixture size
            for k, g in grads.items():
                if k.startswith("g_density"):
                    fx[k] = g
                else:
                    fx["n_" + k] = np.float32(np.linalg.norm(g))
                    fx["s_" + k] = g.reshape(-1)[:: max(1, g.size // 64)][:64].copy()
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **fx)
        print(f"[{name}] wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
