"""GPU: the VF-only entry (VectorFieldNetwork.__call__, the marching-cubes grid query of
evaluation/utils/mc_utils.py:88-104 and the supervision-point queries of the trainer)."""
import ctypes as C

import numpy as np
import pytest
import torch

import vfn_testutil as U
from vfnerf_b200 import _lib
from vfnerf_b200.grid_query import get_set_predictions, grid_query

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("name,P", [("small_det", 1000), ("full_det", 4096), ("full_det", 1)])
def test_vf_module_call_matches_oracle(built_lib, name, P):
    case, z = U.load_golden(name)
    st = U.case_state(case, z)
    model = U.make_model(case, st, DEV)
    g = torch.Generator().manual_seed(P)
    pts = (torch.rand(P, 3, generator=g) - 0.5) * 8
    with torch.no_grad():
        ref = U.O.vf_network(st["vf_net"], pts)
        out = model.vector_field_network(pts.to(DEV))
    assert out.shape == ref.shape
    assert (out.cpu() - ref).abs().max().item() <= 1e-3     # fp32 path tolerance (north_star)
    assert (out.cpu() - ref).abs().max().item() <= 1e-4     # and in practice much tighter


def test_vf_module_call_backward_matches_oracle(built_lib):
    """Supervision-point path of the trainer: MSE on vf(points)[:, :3] (vf_loss.py:56-59)."""
    case, z = U.load_golden("small_det")
    st = U.case_state(case, z)
    model = U.make_model(case, st, DEV)
    g = torch.Generator().manual_seed(3)
    pts = (torch.rand(777, 3, generator=g) - 0.5) * 6
    gt = torch.nn.functional.normalize(-pts, dim=1)
    model.optimizer.zero_grad()
    pred = model.vector_field_network(pts.to(DEV))[:, :3]
    torch.mean((pred - gt.to(DEV)) ** 2).backward()
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
          for k, v in st["vf_net"].items()}
    torch.mean((U.O.vf_network(sd, pts)[:, :3] - gt) ** 2).backward()
    for k, p in model.vector_field_network.named_parameters():
        ref = sd[k].grad
        rel = (p.grad.cpu() - ref).abs().max().item() / (ref.abs().max().item() + 1e-12)
        assert rel <= 2e-3, (k, rel)


def test_get_set_predictions_drop_in(built_lib):
    """Same signature and result layout as mc_utils.get_set_predictions (CPU samples in, CPU [P,3] out)."""
    case, z = U.load_golden("full_det")
    st = U.case_state(case, z)
    model = U.make_model(case, st, DEV)
    g = torch.Generator().manual_seed(1)
    samples = (torch.rand(2500, 3, generator=g) - 0.5) * 4
    pred = get_set_predictions(model.vector_field_network, samples, 1000, torch.device(DEV))
    assert pred.device.type == "cpu" and pred.shape == (2500, 3)
    with torch.no_grad():
        ref = U.O.vf_network(st["vf_net"], samples)[:, :3]
    assert (pred - ref).abs().max().item() <= 1e-4


def test_grid_query_generates_reference_coordinates(built_lib):
    """In-kernel grid coordinates follow evaluation/methods.py:194-208 bit for bit."""
    case, z = U.load_golden("full_det")
    st = U.case_state(case, z)
    model = U.make_model(case, st, DEV)
    res, scale = 24, 1.3
    translation, centroid = torch.tensor([0.65, -0.65, 0.65]), torch.tensor([0.1, 0.2, -0.3])
    # restatement of the reference's grid construction
    idx = torch.arange(0, res ** 3, 1, dtype=torch.long)
    samples = torch.zeros(res ** 3, 3)
    samples[:, 2] = idx % res
    samples[:, 1] = (idx // res) % res
    samples[:, 0] = ((idx // res) // res) % res
    vs = scale * 2.0 / (res - 1)
    for c in range(3):
        samples[:, c] = (samples[:, c] * vs) + (-scale) + translation[c] + centroid[c]
    pred, pts = grid_query(model.vector_field_network, res, scale, translation, centroid, return_points=True)
    assert torch.equal(pts.cpu(), samples)
    with torch.no_grad():
        ref = U.O.vf_network(st["vf_net"], samples)[:, :3]
    assert pred.shape == (res ** 3, 3)
    assert (pred.cpu() - ref).abs().max().item() <= 1e-4
