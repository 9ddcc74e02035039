"""CPU oracle for the supervision-point helpers of the reference trainer.  TEST INFRASTRUCTURE ONLY (see
oracle/render_oracle.py for the rules: tests/, smoke() and bench.py's cpu_baseline leg may import it; the product never).

Restates, in numpy float64 / torch-CPU fp32 like the reference:
    models/samplers/sampler.py:160-193     SphereSampler.sample, with the three uniform draws as explicit arguments
    models/helpers/functions.py:75-97      get_border_indices_and_gt
    models/helpers/functions.py:99-130     sample_border_points / sample_center_points
    models/helpers/functions.py:132-154    get_center_indices_and_gt
Parity pin: tests/golden/make_golden_supervision.py runs the live reference functions (numpy seeded, draws captured) and
asserts this file reproduces them bit for bit before writing tests/golden/supervision.npz."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def sphere_sample(phi: np.ndarray, cos_theta: np.ndarray, u: np.ndarray, r_max: float, r_min: float) -> np.ndarray:
    theta = np.arccos(cos_theta)                                  # sampler.py:174
    r = np.cbrt(u) * (r_max - r_min) + r_min                      # :177
    x = r * np.sin(theta) * np.cos(phi)
    y = r * np.sin(theta) * np.sin(phi)
    z = r * np.cos(theta)
    return np.stack((x, y, z), axis=1)


def sample_border_points(r_min, r_max, centroid: torch.Tensor, phi, cos_theta, u):
    points = torch.from_numpy(sphere_sample(phi, cos_theta, u, r_max, r_min)).float() + centroid     # functions.py:109
    return points, F.normalize(centroid - points, dim=1)


def sample_center_points(centroid: torch.Tensor, radius, phi, cos_theta, u):
    points = torch.from_numpy(sphere_sample(phi, cos_theta, u, radius, 0.0)).float() + centroid      # functions.py:125
    return points, F.normalize(points - centroid, dim=1)


def get_border_indices_and_gt(points, normals, far, radius, centroid):
    distances = torch.norm(points - centroid, dim=2)
    cond = distances > (far / 2 - radius)
    return normals[cond], F.normalize(centroid - points[cond], dim=1)


def get_center_indices_and_gt(points, normals, centroid, radius):
    distances = torch.norm(points - centroid, dim=2)
    cond = distances < radius
    return normals[cond], F.normalize(points[cond] - centroid, dim=1)
