"""GPU: ArenaAdam (csrc/optim.cu, vfnerf_b200/optim.py) == torch's clip_grad_norm_ + Adam.step on every parameter once
(train/vector_field_nerf_train.py:252-258), and its use inside the eager and the CUDA-graph training step."""
import pytest
import torch

import vfn_testutil as U
from vfnerf_b200 import graphed, optim

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _unique_params(model):
    return list({id(p): p for p in model.parameters()}.values())


@pytest.mark.parametrize("max_norm,weight_decay", [(0.5, 0.0), (None, 0.0), (0.05, 0.01)])
def test_arena_adam_matches_torch_adam(built_lib, max_norm, weight_decay):
    case, z = U.load_golden("small_det")
    st = U.case_state(case, z)
    ref = U.make_model(case, st, DEV)
    new = U.make_model(case, st, DEV)
    params = _unique_params(ref)
    topt = torch.optim.Adam(params, lr=5e-4, weight_decay=weight_decay)
    aopt = optim.ArenaAdam(new, lr=5e-4, weight_decay=weight_decay, max_norm=max_norm)
    bn_before = new.vector_field_network.layers[0][1].running_var.clone()
    g = torch.Generator().manual_seed(0)
    new_params = _unique_params(new)
    for it in range(6):
        scale = 10.0 ** (it - 3)                       # gradient norms from far below to far above the clip threshold
        aopt.zero_grad()
        for p, q in zip(params, new_params):
            gr = (torch.randn(p.shape, generator=g) * scale).to(DEV)
            p.grad = gr.clone()
            q.grad.add_(gr)                            # the persistent flat-gradient views
        if max_norm is not None:
            tn = torch.nn.utils.clip_grad_norm_(params, max_norm)
        topt.step()
        aopt.step()
        if max_norm is not None:
            assert abs(aopt.grad_norm().item() - tn.item()) <= 1e-4 * tn.item()
    for p, q in zip(params, new_params):
        assert torch.allclose(p, q, rtol=2e-5, atol=2e-7), (p - q).abs().max()
    # BatchNorm running statistics share the arena but are not parameters: untouched even with weight decay
    assert torch.equal(new.vector_field_network.layers[0][1].running_var, bn_before)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_training_with_arena_optimizer_eager_and_graphed(built_lib, precision):
    case, z = U.load_golden("full_det")
    st = U.case_state(case, z)
    R = 64
    uv, pose, K = (t.to(DEV) for t in U.S.synthetic_rays(R, seed=0, start=case["start"], stride=case["stride"]))
    draws = U.S.synthetic_draws(R, case["n_coarse"], case["n_fine"])
    rgb_gt, dep_gt = torch.rand(R, 3, device=DEV), torch.rand(R, 1, device=DEV) * case["far"]

    def loss_fn(out, rgb_gt, depth_gt):
        nrm = torch.norm(out.coarse_normals.reshape(-1, 3), dim=1)
        return 2.0 * (out.coarse_rgb_values - rgb_gt).abs().mean() + \
            0.5 * (out.coarse_depth_map - depth_gt).abs().clamp(max=0.5).mean() + 0.1 * torch.mean((nrm - 1) ** 2)

    # reference sequence with torch's optimizer on unique parameters
    ref = U.make_model(case, st, DEV, precision=precision)
    params = _unique_params(ref)
    topt = torch.optim.Adam(params, lr=5e-4)
    out = ref.render(pose, uv, K, 0, draws=draws)
    l_ref = loss_fn(out, rgb_gt, dep_gt)
    topt.zero_grad()
    l_ref.backward()
    torch.nn.utils.clip_grad_norm_(params, 0.5)
    g_ref = [p.grad.clone() for p in params]
    w0 = ref.rendering_network.layers[1][0].weight.detach().clone()
    topt.step()

    # eager with ArenaAdam
    m = U.make_model(case, st, DEV, precision=precision)
    optim.use_arena_optimizer(m, max_norm=0.5)
    out = m.render(pose, uv, K, 0, draws=draws)
    loss = loss_fn(out, rgb_gt, dep_gt)
    m.optimizer.zero_grad()
    loss.backward()
    m.optimizer.step()
    m.scheduler.step()
    assert loss.item() == l_ref.item()
    for gr, p in zip(g_ref, _unique_params(m)):          # ArenaAdam leaves the clipped gradient in .grad, like clip_grad_norm_
        assert (gr - p.grad).norm() <= 1e-4 * gr.norm() + 1e-12
    lr = 5e-4
    we, wa = ref.rendering_network.layers[1][0].weight, m.rendering_network.layers[1][0].weight
    assert (we - w0).abs().max().item() > 0.5 * lr and ((we - wa).abs() > 0.5 * lr).float().mean().item() < 0.02

    # graphed with ArenaAdam
    gm = U.make_model(case, st, DEV, precision=precision)
    optim.use_arena_optimizer(gm)
    step = graphed.GraphedTrainStep(gm, loss_fn, R, dict(rgb_gt=rgb_gt, depth_gt=dep_gt), clip_norm=0.5, given_draws=True)
    l0 = step(pose, uv, K, draws=draws, rgb_gt=rgb_gt, depth_gt=dep_gt).item()
    wg = gm.rendering_network.layers[1][0].weight
    assert l0 == l_ref.item()
    assert ((we - wg).abs() > 0.5 * lr).float().mean().item() < 0.02
    l1 = step(pose, uv, K, draws=draws, rgb_gt=rgb_gt, depth_gt=dep_gt).item()
    assert l1 == l1 and l1 != l0


def _rand_grads(params, gen, scale=1e-2):
    for p in params:
        g = (torch.randn(p.shape, generator=gen) * scale).to(DEV)
        if p.grad is None:
            p.grad = g
        else:
            p.grad.copy_(g)


def test_arena_adam_checkpoint_round_trip_and_reference_format(built_lib, tmp_path):
    """model.save()/load() with ArenaAdam installed (vector_field_nerf.py:178-214): the optimizer state is emitted in the
    reference's torch.optim.Adam format (one group over model.parameters(), VF entries duplicated), so (a) a resumed run
    continues bit for bit, and (b) a checkpoint written with the reference's own optimizer loads."""
    case, z = U.load_golden("small_det")
    st = U.case_state(case, z)
    gen = torch.Generator().manual_seed(3)
    a = U.make_model(case, st, DEV)
    optim.use_arena_optimizer(a, max_norm=0.5)
    for _ in range(3):
        a.optimizer.zero_grad()
        _rand_grads(_unique_params(a), gen)
        a.optimizer.step()
    a.save(7, str(tmp_path))
    sd = a.optimizer.state_dict()
    n_par = len(a.parameters())
    assert len(sd["param_groups"]) == 1 and len(sd["param_groups"][0]["params"]) == n_par
    assert len(sd["state"]) == len(_unique_params(a))             # duplicates share one state entry, like torch
    k0 = sd["param_groups"][0]["params"][0]                       # the (duplicated) first VF tensor, torch's numbering
    assert float(sd["state"][k0]["step"]) == 3.0
    # (a) resume into a fresh model + fresh ArenaAdam, then take the same step in both
    b = U.make_model(case, st, DEV)
    optim.use_arena_optimizer(b, max_norm=0.5)
    assert b.load(str(tmp_path / "latest.pth")) == 8
    g2 = torch.Generator().manual_seed(11)
    grads = [(torch.randn(p.shape, generator=g2) * 1e-2).to(DEV) for p in _unique_params(a)]
    for m in (a, b):
        m.optimizer.zero_grad()
        for p, g in zip(_unique_params(m), grads):
            p.grad.copy_(g)
        m.optimizer.step()
    for p, q in zip(_unique_params(a), _unique_params(b)):
        assert torch.equal(p, q)
    # (b) the reference trainer's optimizer state (torch Adam over the duplicated list) -> ArenaAdam
    c = U.make_model(case, st, DEV)                                # its default optimizer IS the reference's
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(2):
            c.optimizer.zero_grad()
            _rand_grads(_unique_params(c), gen)
            c.optimizer.step()
    ref_sd = c.optimizer.state_dict()
    d = U.make_model(case, st, DEV)
    optim.use_arena_optimizer(d)
    d.optimizer.load_state_dict(ref_sd)
    new_sd = d.optimizer.state_dict()
    assert new_sd["param_groups"][0]["params"] == ref_sd["param_groups"][0]["params"]     # same numbering as torch's
    assert set(new_sd["state"]) == set(ref_sd["state"])
    for key in ref_sd["state"]:
        assert torch.equal(new_sd["state"][key]["exp_avg"], ref_sd["state"][key]["exp_avg"])
        assert torch.equal(new_sd["state"][key]["exp_avg_sq"], ref_sd["state"][key]["exp_avg_sq"])
    assert float(d.optimizer._step.item()) == 2.0                  # iteration count, not the doubled VF counter


def test_leaving_flat_gradient_mode_when_a_torch_optimizer_takes_over(built_lib):
    """new_scheduler()/reset_scheduler() install a plain torch Adam (vector_field_nerf.py:103-130): gradients must go back
    to p.grad, or clip + step would silently do nothing."""
    case, z = U.load_golden("small_perturb")
    model = U.make_model(case, U.case_state(case, z), DEV)
    opt = optim.use_arena_optimizer(model)
    model.reset_scheduler(100)
    assert isinstance(model.optimizer, torch.optim.Adam) and model.vector_field_network.arena().grad_flat is None
    uv, pose, K = (U.t(z, k).to(DEV) for k in ("uv", "pose", "K"))
    model.optimizer.zero_grad()
    out = model.render(pose, uv, K, 0)
    (out.coarse_rgb_values.sum() + (out.coarse_normals ** 2).sum()).backward()
    w = model.vector_field_network.layers[0][0].weight
    assert w.grad is not None and w.grad.abs().sum().item() > 0
    before = w.detach().clone()
    model.optimizer.step()
    assert not torch.equal(before, w.detach())
    with pytest.raises(RuntimeError):
        opt.step()                                  # the orphaned ArenaAdam refuses instead of updating nothing
