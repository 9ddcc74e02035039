"""Shared helpers of the test-suite: golden loading, model construction from a golden case, and the
oracle configuration.  The oracle (oracle/render_oracle.py) is imported here -- tests are one of the
three places allowed to."""
import ast
import os

import numpy as np
import torch

from oracle import render_oracle as O           # noqa: F401  (re-exported)
from vfnerf_b200 import synthetic as S
from vfnerf_b200.config import (CudaConfig, DensityConfig, RaySamplerConfig, RenderingNetConfig,
                                SchedulerConfig, VFNerfConfig, VFNetConfig)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOSS_W = dict(rgb=2.0, depth=0.5, unit_norm=0.1, supervision=1.0, norm_smaller_than_one=0.1,
              directional_derivatives=0.0)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    case = ast.literal_eval(str(z["case"]))
    return case, z


def case_state(case, z):
    """Weights of a golden case: stored in the fixture for the small nets, regenerated from the seed
    (vfnerf_b200.synthetic) for the full-size ones."""
    if any(k.startswith("w_vf.") for k in z.files):
        vf = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w_vf.")}
        rn = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w_rn.")}
        dens = {"beta": torch.tensor(0.5), "scale": torch.tensor(100.0), "mean": torch.tensor(0.7)}
        return {"vf_net": vf, "rendering_net": rn, "density": dens}
    return S.synthetic_state(case["seed"], case["vf_hidden"], case["feat"], case["rn_hidden"],
                             vf_gain=case["vf_gain"])


def oracle_cfg(case):
    return dict(n_coarse=case["n_coarse"], n_fine=min(case["n_fine"], case["max_samples"]),
                near=case["near"], far=case["far"], fine_range=case["fine_range"],
                perturb=case["perturb"], window=case["window"],
                dir_to_normal_th=case["dir_to_normal_th"], normalize=True,
                beta_bounds=(1e-4, 1e9), scale_min=1.0, mean_bounds=(0.6, 1.0),
                multires=6, multires_view=4, skip_in=(4,))


make_config, make_model = S.make_config, S.make_model      # (live in the package: bench.py builds its models with them too)


def t(z, key):
    return torch.from_numpy(np.asarray(z[key]))


def discontinuity_guard(oracle_out, case, eps=2e-4):
    """Rays whose density sits on a discontinuity of the reference's own definition and therefore
    amplify fp32 rounding noise arbitrarily: a sample whose windowed cosine is within eps of the
    relu threshold (c = 0.5) or -- when the direction mask can fire -- of the mask threshold c = 0.
    Returns a bool [R] mask of rays that are safe to compare at tight tolerance."""
    c = oracle_out["cosw"]
    risky = (c - 0.5).abs() < eps
    if case["dir_to_normal_th"] > -1.0:
        risky |= c.abs() < eps
    return ~risky.any(dim=1)


def reference_grid_points(res, scale, translation, centroid):
    """Restatement of the grid construction of evaluation/methods.py:194-208 (z fastest, fp32 op order)."""
    idx = torch.arange(0, res ** 3, 1, dtype=torch.long)
    samples = torch.zeros(res ** 3, 3)
    samples[:, 2] = idx % res
    samples[:, 1] = (idx // res) % res
    samples[:, 0] = ((idx // res) // res) % res
    vs = scale * 2.0 / (res - 1)
    for c in range(3):
        samples[:, c] = (samples[:, c] * vs) + (-scale) + translation[c] + centroid[c]
    return samples
