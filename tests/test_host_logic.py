"""CPU: host-side mirror of the reference interface -- state_dict keys, parameter list quirks,
draw order, error behaviour, and the refusal to run without CUDA."""
import pytest
import torch

import vfn_testutil as U
from vfnerf_b200 import VectorFieldNerf, VFNerfConfig
from vfnerf_b200.samplers import RangeFineSampler, UniformSampler


@pytest.fixture(scope="module")
def model():
    case, z = U.load_golden("full_det")
    return U.make_model(case, U.case_state(case, z), "cpu")


def test_state_dict_keys_match_reference(model):
    # key pattern of the reference's checkpoints (SURVEY.md §5 "Checkpoint / resume")
    want_vf = []
    for i in range(8):
        want_vf += [f"layers.{i}.0.weight", f"layers.{i}.0.bias", f"layers.{i}.1.weight", f"layers.{i}.1.bias",
                    f"layers.{i}.1.running_mean", f"layers.{i}.1.running_var", f"layers.{i}.1.num_batches_tracked"]
    want_vf += ["layers.8.weight", "layers.8.bias"]
    assert list(model.vector_field_network.state_dict().keys()) == want_vf
    rn = list(model.rendering_network.state_dict().keys())
    assert rn[0] == "layers.0.0.weight" and rn[-2:] == ["layers.4.weight", "layers.4.bias"] and len(rn) == 30
    assert sorted(model.density.state_dict().keys()) == ["beta", "mean", "scale"]
    sd = model.vector_field_network.state_dict()
    assert sd["layers.0.0.weight"].shape == (256, 39)
    assert sd["layers.3.0.weight"].shape == (217, 256)      # narrowed before the skip layer
    assert sd["layers.8.weight"].shape == (259, 256)
    assert model.rendering_network.state_dict()["layers.0.0.weight"].shape == (256, 289)


def test_parameter_list_has_the_reference_duplicates(model):
    # 34 VF + 18 colour + 3 density + 34 VF again (vector_field_nerf.py:132-137)
    assert len(model.parameters()) == 89
    assert model.fine_vector_field_network is model.vector_field_network


def test_arena_aliases_parameters_and_survives_to_and_load(model):
    ar = model.vector_field_network.arena()
    w = model.vector_field_network.layers[0][0].weight
    assert w.data_ptr() == ar.flat.data_ptr() + 4 * ar.desc.w_off[0]
    with torch.no_grad():
        w[0, 0] = 123.0
    assert ar.flat[ar.desc.w_off[0]].item() == 123.0
    sd = {k: v.clone() for k, v in model.vector_field_network.state_dict().items()}
    sd["layers.0.0.weight"][0, 0] = -7.0
    model.vector_field_network.load_state_dict(sd)
    ar = model.vector_field_network.arena()
    assert ar.flat[ar.desc.w_off[0]].item() == -7.0
    model.vector_field_network.double().float()              # storage replaced -> arena rebuilt
    ar2 = model.vector_field_network.arena()
    rm = model.vector_field_network.layers[0][1].running_mean
    assert rm.data_ptr() == ar2.flat.data_ptr() + 4 * ar2.desc.mean_off[0]
    assert ar2.flat[ar2.desc.w_off[0]].item() == -7.0
    assert ar2.desc.in_dim[4] == 256 and ar2.desc.out_dim[3] == 217 and ar2.desc.n_layers == 9


def test_no_cpu_fallback(model):
    uv, pose, K = U.S.synthetic_rays(4)
    with pytest.raises(RuntimeError, match="CUDA"):
        model.render(pose, uv, K, 0)
    with pytest.raises(RuntimeError, match="CUDA"):
        model.vector_field_network(torch.zeros(4, 3))


def test_reference_error_behaviour(model):
    uv, pose, K = U.S.synthetic_rays(4)
    with pytest.raises(UnboundLocalError):
        model.render(pose, uv, K, 0, white=True)
    with pytest.raises(ValueError):
        VFNerfConfig(rendering="splat")
    with pytest.raises(ValueError):
        VFNerfConfig(cos_sim_weights_anneal="anneal_fine")
    model.train()          # train mode is built (csrc/mlp_train.cu) and, like everything else, CUDA-only
    try:
        with pytest.raises(RuntimeError, match="CUDA"):
            model.vector_field_network(torch.zeros(4, 3))
        model.set_precision("bf16")
        with pytest.raises(NotImplementedError):       # batch statistics run on the fp32 layer-wise path only
            model.render(pose, uv, K, 0)
    finally:
        model.set_precision("fp32")
        model.eval()


def test_draw_order_matches_reference_generator_consumption():
    """U1 -> U2 -> U3 when perturbing; only U3 when deterministic (ray_sampler.py:138,292,297)."""
    cs, fs = UniformSampler(8, 0.0, 1.0, deterministic=False), RangeFineSampler(6, 0.0, 1.0, False, 0.3, 100)
    torch.manual_seed(7)
    a1 = torch.rand([3, 8]); a2 = torch.rand([3, 6]); a3 = torch.rand((3, 6))
    torch.manual_seed(7)
    U1 = cs.draw(3); U2, U3 = fs.draw(3)
    assert torch.equal(U1, a1) and torch.equal(U2, a2) and torch.equal(U3, a3)
    cs.deterministic = fs.deterministic = True
    torch.manual_seed(7)
    b3 = torch.rand((3, 6))
    torch.manual_seed(7)
    assert cs.draw(3) is None
    U2, U3 = fs.draw(3)
    assert U2 is None and torch.equal(U3, b3)
    fs.N_samples = 500
    assert fs.n_fine() == 100                      # min(max_samples, N_samples), ray_sampler.py:276


def test_sampler_attributes_are_read_at_call_time(model):
    model.ray_sampler.near, model.ray_sampler.far = 0.5, 4.0
    model.fine_sampler.N_samples = 70
    cfg = model._render_cfg(16, False)
    assert (cfg.near_, cfg.far_, cfg.n_fine, cfg.n_coarse) == (0.5, 4.0, 70, 64)
    assert cfg.window == 11 and cfg.skip_layer == 4 and cfg.multires == 6 and cfg.multires_view == 4
    model.ray_sampler.near, model.ray_sampler.far = 0.0, 6.0
    model.fine_sampler.N_samples = 64


def test_device_side_trainer_pieces_refuse_cpu(model):
    """losses.VFLoss, optim.ArenaAdam and graphed.GraphedTrainStep are CUDA-only like the rest of the package."""
    import types
    from vfnerf_b200 import graphed, optim
    from vfnerf_b200.losses import VFLoss
    loss = VFLoss(types.SimpleNamespace(norm_smaller_than_one_start=11000, depth_loss_clamp=0.5, directional_derivatives_start=100),
                  types.SimpleNamespace(rgb=2.0, depth=0.5, unit_norm=0.1, supervision=1.0, norm_smaller_than_one=0.1,
                                        directional_derivatives=0.0))
    pred = {"rgb": torch.rand(4, 3), "depth": torch.rand(4, 1), "normals": torch.rand(8, 3), "supervised_normals": None,
            "directional_derivatives": None}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        loss(pred, {"rgb": torch.rand(4, 3), "depth": torch.rand(4, 1)}, 0)
    with pytest.raises(RuntimeError, match="CUDA"):
        optim.ArenaAdam(model)
    with pytest.raises(RuntimeError, match="CUDA"):
        graphed.GraphedTrainStep(model, lambda out: out, 4, {})


def test_flat_gradient_mode_aliases_parameter_grads(model):
    """ParamArena.enable_flat_grad: every Parameter's .grad is a view of one flat tensor in arena layout."""
    ar = model.rendering_network.arena()
    g = ar.enable_flat_grad()
    assert g.shape == ar.flat.shape and float(g.abs().sum()) == 0.0
    w = model.rendering_network.layers[1][0].weight
    w.grad.fill_(3.0)
    off = ar.desc.w_off[1]
    assert torch.equal(g[off:off + w.numel()], torch.full((w.numel(),), 3.0))
    mask = ar.trainable_mask()
    n_params = sum(p.numel() for p in model.rendering_network.parameters())
    assert int(mask.sum()) == n_params and mask.numel() == ar.flat.numel() and n_params < mask.numel()
    ar.grad_flat = None
    for p in model.rendering_network.parameters():
        p.grad = None


def test_fast_cpu_generator_stream_is_torch_rand():
    """csrc/host_rng.cu: the vectorised MT19937 behind the large sampler draws reproduces torch.rand on the global CPU
    generator bit for bit -- values AND the generator state afterwards -- from arbitrary stream positions, so seeding
    torch.manual_seed reproduces the reference's sample positions exactly as before."""
    import vfn_testutil  # noqa: F401  (path setup)
    from vfnerf_b200 import _lib, samplers as S
    _lib.build()
    assert S._fast_rng_self_check()
    for seed, skip, n in ((0, 0, 1 << 16), (5, 3, 70001), (9, 623, 4 * 65536 + 5), (11, 624, 1 << 18)):
        torch.manual_seed(seed)
        if skip:
            torch.rand(skip)
        st = torch.get_rng_state()
        want = torch.rand(n)
        st_after = torch.get_rng_state()
        torch.set_rng_state(st)
        got = S.cpu_generator_rand_(torch.empty(n))
        assert torch.equal(got, want)
        assert torch.equal(torch.get_rng_state(), st_after)
        assert torch.equal(torch.rand(9), (torch.set_rng_state(st_after), torch.rand(9))[1])
    # small or non-contiguous tensors take torch.rand itself
    torch.manual_seed(1)
    a = torch.rand(10)
    torch.manual_seed(1)
    assert torch.equal(S.cpu_generator_rand_(torch.empty(10)), a)


def test_schedule_flag_and_arena_spot_check(model):
    """recompute_coarse selects the literal two-evaluation schedule through cfg.flags; the O(1) arena validation still
    notices storage that was swapped behind the module's back (first / last tensor), not only Module._apply."""
    from vfnerf_b200 import _lib
    assert model._render_cfg(8, False).flags == 0
    model.recompute_coarse = True
    assert model._render_cfg(8, False).flags == _lib.FLAG_RECOMPUTE_COARSE
    model.recompute_coarse = False
    net = model.rendering_network
    ar = net.arena()
    flat0 = ar.flat
    w = net.layers[0][0].weight
    w.data = w.data.clone()                    # manual swap of the first tensor: no _apply, no load_state_dict
    ar2 = net.arena()
    assert ar2.flat is not flat0
    assert w.data_ptr() == ar2.flat.data_ptr() + 4 * ar2.desc.w_off[0]
    net.load_state_dict({k: v.clone() for k, v in net.state_dict().items()})     # marks dirty, aliasing kept
    assert net.arena().flat is ar2.flat


def test_render_rejects_what_it_does_not_implement():
    """Configurations the reference handles differently from this path must raise, not silently diverge."""
    import pytest
    import torch
    from vfnerf_b200 import synthetic as S
    case = dict(seed=0, vf_hidden=(32, 32), feat=16, rn_hidden=(32,), n_coarse=8, n_fine=8, max_samples=100, perturb=False,
                near=0.0, far=6.0, fine_range=0.3, window=11, dir_to_normal_th=-0.2, vf_gain=2.0)
    try:
        model = S.make_model(case, S.synthetic_state(0, (32, 32), 16, (32,), vf_gain=2.0), torch.device("cpu"))
    except Exception:
        pytest.skip("model construction needs the layer shapes of the synthetic helper")
    pose, uv, K = torch.zeros(1, 4, 4), torch.zeros(1, 2), torch.zeros(1, 4, 4)
    model.rendering_network.train()
    with pytest.raises(NotImplementedError):
        model.render(pose, uv, K, 0)
    model.eval()
    # the fine sampler's own range reaches the kernels (z_add fallback samples, ray_sampler.py:296-299)
    model.fine_sampler.near, model.fine_sampler.far = 0.25, 7.0
    cfg = model._render_cfg(1, False)
    assert (cfg.fine_near_, cfg.fine_far_) == (0.25, 7.0) and (cfg.near_, cfg.far_) == (0.0, 6.0)
