"""GPU, bench scale: a size-independent property of the whole path.  One 65 536-ray chunk of the bench image rendered in ONE
call, in one call again, and in 64 calls of the reference's 1024 rays (evaluation/methods.py:516-530) -- different tile ->
cluster assignment, different timing of every hand-off inside the fused tcgen05 kernel -- must agree BIT FOR BIT in every
output field.  A race in the activation tile (the aux columns of the split-precision tile alias remainder columns, the
weight ring, the accumulator hand-off) shows up here as a run-to-run or chunk-to-chunk difference long before it is large
enough to break a tolerance."""
import pytest
import torch

import vfn_testutil as U

pytestmark = pytest.mark.gpu
FIELDS = ("coarse_rgb_values", "coarse_depth_map", "coarse_normals", "coarse_colors", "z_vals", "weights", "points_coarse")


@pytest.mark.parametrize("precision", ["fp16f8", "bf16x3", "bf16"])
def test_one_call_equals_sixty_four_reference_sized_calls(built_lib, precision):
    dev, R = "cuda", 65536
    case, z = U.load_golden("full_det")
    uv, pose, K = (t.to(dev) for t in U.S.synthetic_rays(R, seed=0, start=0, stride=1))
    g = torch.Generator(device=dev).manual_seed(3)
    draws = tuple(torch.rand(R, n, device=dev, generator=g) for n in (case["n_coarse"], case["n_fine"], case["n_fine"]))
    m = U.make_model(case, U.case_state(case, z), dev, precision=precision)
    with torch.no_grad():
        a = m.render(pose, uv, K, 0, draws=draws)
        a = {f: getattr(a, f).clone() for f in FIELDS}
        b = m.render(pose, uv, K, 0, draws=draws)
        b = {f: getattr(b, f).clone() for f in FIELDS}
        parts = []
        for i in range(0, R, 1024):
            o = m.render(pose[i:i + 1024], uv[i:i + 1024], K[i:i + 1024], 0, draws=tuple(d[i:i + 1024] for d in draws))
            parts.append({f: getattr(o, f).clone() for f in FIELDS})
    torch.cuda.synchronize()
    assert (a["weights"] > 0).float().mean().item() > 0.01, "degenerate image: nothing would be compared"
    for f in FIELDS:
        assert torch.equal(a[f], b[f]), f"{f}: two identical calls differ"
        cat = torch.cat([p[f] for p in parts])
        assert torch.equal(a[f].reshape(cat.shape), cat), f"{f}: one 65536-ray call != 64 calls of 1024 rays"
