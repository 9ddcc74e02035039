"""CPU: the oracle restatement against the golden vectors generated from the live reference
(tests/golden/make_golden.py).  This is what pins the oracle on machines without /root/reference."""
import numpy as np
import pytest
import torch

import vfn_testutil as U

CASES = ["small_det", "small_perturb", "full_det", "full_perturb"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    case, z = U.load_golden(name)
    st = U.case_state(case, z)
    with torch.no_grad():
        out = U.O.render(st["vf_net"], st["rendering_net"], st["density"], U.oracle_cfg(case),
                         U.t(z, "uv"), U.t(z, "pose"), U.t(z, "K"), U.t(z, "t_vals"),
                         U.t(z, "U1"), U.t(z, "U2"), U.t(z, "U3"))
    # sample positions: bit-exact
    assert torch.equal(out["z_vals"], U.t(z, "ref_z_vals"))
    assert torch.equal(out["points"], U.t(z, "ref_points"))
    # everything else: fp32 op-order noise only (tolerance 1e-4 abs, see make_golden.py)
    for key, ref in (("normals", "ref_normals"), ("rgb", "ref_rgb"), ("depth", "ref_depth"),
                     ("colors", "ref_colors"), ("rep_ray_dirs", "ref_ray_dirs")):
        dev = (out[key] - U.t(z, ref)).abs().max().item()
        assert dev <= 1e-4, (name, key, dev)
    # the synthetic model must not be degenerate
    assert (out["sigma"] > 0).float().mean().item() > 0.002
    assert U.t(z, "ref_rgb").abs().max().item() > 0.05


@pytest.mark.parametrize("name", ["small_det", "small_perturb"])
def test_oracle_gradients_match_reference_golden(name):
    case, z = U.load_golden(name)
    st = U.case_state(case, z)
    req = lambda sd: {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
                      for k, v in sd.items()}
    vf, rn = req(st["vf_net"]), req(st["rendering_net"])
    dn = {k: v.clone().requires_grad_(True) for k, v in st["density"].items()}
    out = U.O.render(vf, rn, dn, U.oracle_cfg(case), U.t(z, "uv"), U.t(z, "pose"), U.t(z, "K"), U.t(z, "t_vals"),
                     U.t(z, "U1"), U.t(z, "U2"), U.t(z, "U3"))
    loss = U.O.vf_loss(out["rgb"], out["depth"], out["normals"].reshape(-1, 3), U.t(z, "rgb_gt"),
                       U.t(z, "depth_gt"), U.LOSS_W, 0.5)
    assert abs(loss.item() - float(z["ref_loss"])) < 1e-5
    loss.backward()
    for prefix, sd in (("g_vf.", vf), ("g_rn.", rn)):
        for k, v in sd.items():
            if isinstance(v, torch.Tensor) and v.requires_grad:
                g = z[prefix + k]
                rel = np.abs(v.grad.numpy() - g).max() / (np.abs(g).max() + 1e-12)
                assert rel < 2e-3, (k, rel)
    for k, v in dn.items():
        g = float(z["g_density." + k])
        assert abs(v.grad.item() - g) <= 2e-3 * abs(g) + 1e-7, (k, v.grad.item(), g)


def test_fine_sampler_edge_cases():
    """argmax ties / all-zero weights select the uniform z_add branch; rows are sorted; U3 is consumed
    even when deterministic (ray_sampler.py:297)."""
    R, Nc, Nf = 5, 16, 8
    zc = torch.linspace(0, 6, Nc).repeat(R, 1)
    w = torch.zeros(R, Nc)
    w[1, 0] = 1.0          # argmax 0 -> z_add branch
    w[2, 3] = 0.5
    w[2, 9] = 0.5          # tie -> first index (3)
    w[3, Nc - 1] = 1.0
    w[4, 5] = 0.2
    U3 = torch.rand(R, Nf)
    z = U.O.fine_z_vals(zc, w, 0.0, 6.0, 0.3, Nf, False, None, U3)
    assert z.shape == (R, Nc + Nf)
    assert torch.all(z[:, 1:] >= z[:, :-1])
    for r in (0, 1):       # all-zero row and argmax==0 row use U3*(far-near)+near
        want = torch.sort(torch.cat([zc[r], U3[r] * 6.0 + 0.0]))[0]
        assert torch.equal(z[r], want)
    centre = zc[2, 3]
    assert ((z[2] - centre).abs() <= 0.3 + 1e-6).sum().item() >= Nf


def test_window_cosine_closed_form():
    """c[j] for 7 <= j < N-1-7 is the mean over 5 back / 6 forward partners."""
    torch.manual_seed(1)
    n = torch.randn(3, 40, 3)
    c = U.O.window_cosine(n, 11)
    u = n / n.norm(dim=-1, keepdim=True)
    j = 12
    partners = list(range(j - 5, j)) + list(range(j + 1, j + 7))
    want = sum((u[:, j] * u[:, k]).sum(-1) for k in partners) / 11.0
    assert torch.allclose(c[:, j], want, atol=2e-6)
    assert torch.allclose(c[:, 2], (u[:, 2] * u[:, 3]).sum(-1), atol=1e-6)
    assert c.shape == (3, 39)


@pytest.mark.parametrize("tag", ["det", "rand", "ragged", "tiny"])
def test_pdf_sampler_oracle_matches_reference_golden(tag):
    """FineSampler.sample_pdf / get_z_vals (ray_sampler.py:163-237): the oracle restatement reproduces the live
    reference's outputs bit for bit on CPU (fixture from tests/golden/make_golden_pdf.py)."""
    import os
    z = np.load(os.path.join(U.GOLDEN_DIR, "pdf_sampler.npz"))
    g = {k: torch.from_numpy(z[f"{tag}.{k}"]) for k in ("z_c", "w_c", "u", "ref_z", "ref_samples")}
    assert torch.equal(U.O.pdf_fine_z_vals(g["z_c"], g["w_c"], g["u"]), g["ref_z"])
    mid = .5 * (g["z_c"][..., 1:] + g["z_c"][..., :-1])
    assert torch.equal(U.O.sample_pdf(mid, g["w_c"][..., 1:-1], g["u"]), g["ref_samples"])
    assert g["ref_z"].shape[1] == g["z_c"].shape[1] + g["u"].shape[-1]


@pytest.mark.parametrize("tag,N", [("n24", 24), ("n33", 33)])
def test_mc_oracle_matches_reference_golden(tag, N):
    """extract_divergence / unify_direction / make_comb_format / block-ordered compaction (evaluation/utils/mc_utils.py,
    evaluation/methods.py:209-278): the per-cell oracle restatement equals the live reference's outputs exactly
    (fixture from tests/golden/make_golden_mc.py)."""
    import os
    from oracle import mc_oracle as MO
    z = np.load(os.path.join(U.GOLDEN_DIR, "mc_preprocess.npz"))
    g = {k: torch.from_numpy(z[f"{tag}.{k}"]) for k in ("pred", "div", "choice", "cells", "comb", "udf")}
    div = MO.extract_divergence(g["pred"], N)
    assert torch.equal(div, g["div"])
    assert torch.equal(MO.unify_direction(div, g["pred"], N).to(torch.uint8), g["choice"])
    cells, comb, udf = MO.mc_preprocess(g["pred"], N)
    assert torch.equal(cells.int(), g["cells"]) and torch.equal(comb, g["comb"]) and torch.equal(udf, g["udf"])
    assert g["cells"].shape[0] > 1000


@pytest.mark.parametrize("tag", ["n128", "n200", "n2"])
def test_weight_function_oracles_match_reference_golden(tag):
    """nerf_volume_rendering / volsdf_volume_rendering (utils/rendering.py:98-148): oracle == live reference, bit for bit
    (fixture from tests/golden/make_golden_weights.py)."""
    import os
    z = np.load(os.path.join(U.GOLDEN_DIR, "volume_weights.npz"))
    zz, sg = torch.from_numpy(z[f"{tag}.z"]), torch.from_numpy(z[f"{tag}.sigma"])
    for norm in (0, 1):
        assert torch.equal(U.O.nerf_weights(sg, zz, bool(norm)), torch.from_numpy(z[f"{tag}.nerf.{norm}"]))
        assert torch.equal(U.O.volsdf_weights(zz, sg, bool(norm)), torch.from_numpy(z[f"{tag}.volsdf.{norm}"]))


def test_smooth_oracle_matches_reference_golden():
    """smooth_vf (guassian_smoothing.py:81-97) and the smooth_after variant of the mesh preprocessing: the oracle
    restatements vs outputs of the live reference (tests/golden/make_golden_smooth.py)."""
    import os
    from oracle import mc_oracle as MO
    z = np.load(os.path.join(U.GOLDEN_DIR, "mc_smooth.npz"))
    pred = torch.from_numpy(z["pred"])
    N = round(pred.shape[0] ** (1 / 3))
    for k, sigma in ((3, 1.0), (9, 2.0)):
        got = MO.smooth_vf(pred.reshape(N, N, N, 3), k, sigma)
        assert (got - torch.from_numpy(z[f"smooth_k{k}"])).abs().max().item() <= 2e-6
    cells, comb, udf = MO.mc_preprocess_smooth_after(pred, N)
    assert torch.equal(cells.int(), torch.from_numpy(z["after.cells"]))
    assert torch.equal(comb, torch.from_numpy(z["after.comb"]))
    assert (udf - torch.from_numpy(z["after.udf"])).abs().max().item() <= 2e-6


def test_supervision_oracle_reproduces_reference_golden():
    """oracle/supervision_oracle.py vs vectors written by the live reference functions (make_golden_supervision.py)."""
    import os
    import numpy as np
    from oracle import supervision_oracle as SO
    z = np.load(os.path.join(U.GOLDEN_DIR, "supervision.npz"))
    c = torch.from_numpy(z["centroid"])
    far, radius = float(z["far"]), float(z["radius"])
    p, g = SO.sample_border_points(far / 2 - radius, far / 2, c, z["border_phi"], z["border_cos"], z["border_u"])
    assert torch.equal(p, torch.from_numpy(z["border_points"])) and torch.equal(g, torch.from_numpy(z["border_gt"]))
    p, g = SO.sample_center_points(c, radius, z["center_phi"], z["center_cos"], z["center_u"])
    assert torch.equal(p, torch.from_numpy(z["center_points"])) and torch.equal(g, torch.from_numpy(z["center_gt"]))
    pts, nrm = torch.from_numpy(z["ray_points"]), torch.from_numpy(z["ray_normals"])
    n, g = SO.get_border_indices_and_gt(pts, nrm, far, radius, c)
    assert torch.equal(n, torch.from_numpy(z["sel_border_normals"])) and torch.equal(g, torch.from_numpy(z["sel_border_gt"]))
    n, g = SO.get_center_indices_and_gt(pts, nrm, c, radius)
    assert torch.equal(n, torch.from_numpy(z["sel_center_normals"])) and torch.equal(g, torch.from_numpy(z["sel_center_gt"]))
    # shell / ball membership of the sampled points (size-independent property)
    d = (torch.from_numpy(z["border_points"]) - c).norm(dim=1)
    assert (d >= far / 2 - radius - 1e-5).all() and (d <= far / 2 + 1e-5).all()
    assert ((torch.from_numpy(z["center_points"]) - c).norm(dim=1) <= radius + 1e-6).all()


def test_train_mode_oracle_matches_reference_golden():
    """Train mode: batch-statistic BatchNorm, Jacobian of column sums, directional derivatives,
    running-statistics update and gradients of the restatement against what the LIVE reference produced in model.train()
    (tests/golden/make_golden_train.py)."""
    case, z = U.load_golden("train_small")
    vf0 = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w_vf.")}
    rn0 = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w_rn.")}
    req = lambda sd: {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)   # noqa: E731
                      for k, v in sd.items()}
    vf, rn = req(vf0), req(rn0)
    dn = {"beta": torch.tensor(0.5, requires_grad=True), "scale": torch.tensor(100.0, requires_grad=True),
          "mean": torch.tensor(0.7, requires_grad=True)}
    out = U.O.render_train(vf, rn, dn, U.oracle_cfg(case), U.t(z, "uv"), U.t(z, "pose"), U.t(z, "K"), U.t(z, "t_vals"),
                           U.t(z, "U1"), U.t(z, "U2"), U.t(z, "U3"))
    assert torch.equal(out["z_vals"], U.t(z, "ref_z_vals"))
    for key, ref in (("normals", "ref_normals"), ("rgb", "ref_rgb"), ("depth", "ref_depth"), ("colors", "ref_colors")):
        assert (out[key].detach() - U.t(z, ref)).abs().max().item() <= 2e-4, key
    dd, dd_ref = out["directional_derivatives"], U.t(z, "ref_dir_derivs")
    assert dd.shape == dd_ref.shape and ((dd - dd_ref).abs().max() / dd_ref.abs().max()).item() <= 2e-4
    for tag, after in (("after_vf.", out["vf_sd_after"]), ("after_rn.", out["rn_sd_after"])):
        for k in z.files:
            if k.startswith(tag):
                ref = torch.from_numpy(z[k])
                if "num_batches" in k:
                    assert int(after[k[len(tag):]]) == int(ref)
                else:
                    assert ((after[k[len(tag):]] - ref).abs().max() / ref.abs().max()).item() <= 1e-5, k
    w = dict(U.LOSS_W, directional_derivatives=0.05)
    loss = U.O.vf_loss(out["rgb"], out["depth"], out["normals"].reshape(-1, 3), U.t(z, "rgb_gt"), U.t(z, "depth_gt"), w, 0.5) + \
        0.05 * dd.mean()
    assert abs(loss.item() - float(z["ref_loss"])) < 2e-5
    loss.backward()
    for prefix, sd in (("g_vf.", vf), ("g_rn.", rn)):
        for k, v in sd.items():
            if isinstance(v, torch.Tensor) and v.requires_grad:
                g = z[prefix + k]
                got = np.zeros_like(g) if v.grad is None else v.grad.numpy()
                # (Linear biases in front of a batch-statistics BatchNorm: mathematically zero, 1e-8 noise on both sides)
                assert np.abs(got - g).max() / max(np.abs(g).max(), 1e-4) < 2e-3, k
