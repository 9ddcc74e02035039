"""GPU: opt-in CUDA-graph replay of the forward-only render() (model.graph_replay) -- the calling pattern of the reference's
chunked evaluation loop (evaluation/methods.py:510-530: one render() per 1024 rays under torch.no_grad())."""
import pytest
import torch

import vfn_testutil as U

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("precision", ["bf16x3", "bf16", "fp32"])
def test_replay_equals_eager_call(built_lib, precision):
    case, z = U.load_golden("full_perturb" if precision != "fp32" else "small_perturb")
    st = U.case_state(case, z)
    model = U.make_model(case, st, DEV, precision=precision)
    uv, pose, K = (U.t(z, k).to(DEV) for k in ("uv", "pose", "K"))
    draws = tuple(U.t(z, k).to(DEV) for k in ("U1", "U2", "U3"))
    with torch.no_grad():
        eager = model.render(pose, uv, K, 0, draws=draws)
        model.graph_replay = True
        for _ in range(3):                              # first call captures, the others replay
            rep = model.render(pose, uv, K, 0, draws=draws)
            for f in ("z_vals", "points_coarse", "coarse_normals", "coarse_colors", "coarse_rgb_values", "coarse_depth_map"):
                assert torch.equal(getattr(rep, f), getattr(eager, f)), f
        assert len(model._graphs) == 1
        # a different ray count is a different graph; the first one stays valid
        h = uv.shape[0] // 2
        half = model.render(pose[:h], uv[:h], K[:h], 0, draws=tuple(d[:h] for d in draws))
        assert torch.equal(half.coarse_rgb_values, eager.coarse_rgb_values[:h]) and len(model._graphs) == 2
        # parameter updates between replays are picked up (the weight images are re-packed inside the graph)
        for p in model.rendering_network.parameters():
            p.mul_(0.5)
        model.graph_replay = False
        eager2 = model.render(pose, uv, K, 0, draws=draws)
        model.graph_replay = True
        rep2 = model.render(pose, uv, K, 0, draws=draws)
        assert torch.equal(rep2.coarse_rgb_values, eager2.coarse_rgb_values)
        assert not torch.equal(eager2.coarse_rgb_values, eager.coarse_rgb_values)
        # ... and only then: with unchanged parameters the replay reuses the packed weight images (second capture)
        g = next(iter(v for k, v in model._graphs.items() if k[0] == uv.shape[0]))
        rep3 = model.render(pose, uv, K, 0, draws=draws)
        assert torch.equal(rep3.coarse_rgb_values, eager2.coarse_rgb_values)
        if precision != "fp32":
            assert rep3.coarse_rgb_values.data_ptr() == g.out_packed.coarse_rgb_values.data_ptr()
        # writes through .data are invisible to the version counters: the documented way out is invalidate_replay()
        for p in model.rendering_network.parameters():
            p.data.mul_(2.0)
        model.invalidate_replay()
        rep4 = model.render(pose, uv, K, 0, draws=draws)
        assert torch.equal(rep4.coarse_rgb_values, eager.coarse_rgb_values)


def test_replay_draws_follow_the_cpu_generator(built_lib):
    """Without explicit draws a replayed call consumes the global CPU generator exactly like the eager call (and like
    the reference: U3 is drawn even when sampling is deterministic, ray_sampler.py:297)."""
    case, z = U.load_golden("full_det")
    model = U.make_model(case, U.case_state(case, z), DEV, precision="bf16x3")
    uv, pose, K = (U.t(z, k).to(DEV) for k in ("uv", "pose", "K"))
    with torch.no_grad():
        torch.manual_seed(77)
        a1 = model.render(pose, uv, K, 0).coarse_rgb_values.clone()
        a2 = model.render(pose, uv, K, 0).coarse_rgb_values.clone()
        model.graph_replay = True
        torch.manual_seed(77)
        b1 = model.render(pose, uv, K, 0).coarse_rgb_values.clone()
        b2 = model.render(pose, uv, K, 0).coarse_rgb_values.clone()
    assert torch.equal(a1, b1) and torch.equal(a2, b2)


def test_replay_is_skipped_when_gradients_are_needed(built_lib):
    case, z = U.load_golden("small_perturb")
    model = U.make_model(case, U.case_state(case, z), DEV)
    model.graph_replay = True
    uv, pose, K = (U.t(z, k).to(DEV) for k in ("uv", "pose", "K"))
    out = model.render(pose, uv, K, 0)
    assert out.coarse_rgb_values.requires_grad and not model._graphs
