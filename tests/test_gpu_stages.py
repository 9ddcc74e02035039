"""GPU: per-kernel parity against the oracle, through the C-ABI stage entry points (one test per
SURVEY.md §8(a) row).  Sampler rows are bit-exact; floating-point rows state their tolerance."""
import ctypes as C

import pytest
import torch

import vfn_testutil as U
from vfnerf_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


_ALIVE = []


def keep(t):
    """device copy of a host tensor that stays referenced until the test module is torn down (a
    temporary would be freed -- and its memory reused -- before the asynchronous kernel reads it)"""
    d = t.to(DEV).contiguous()
    _ALIVE.append(d)
    if len(_ALIVE) > 64:
        torch.cuda.synchronize()
        del _ALIVE[:32]
    return d.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _cfg(R, window=11, th=-0.2, normalize=1):
    cfg = _lib.RenderCfg()
    cfg.n_rays, cfg.window, cfg.normalize, cfg.dir_to_normal_th = R, window, normalize, th
    cfg.beta_lo, cfg.beta_hi, cfg.mean_lo, cfg.mean_hi, cfg.scale_min = 1e-4, 1e9, 0.6, 1.0, 1.0
    return cfg


@pytest.mark.parametrize("R", [1, 33, 4096])
def test_a1_ray_geometry_bit_exact(built_lib, R):
    uv, pose, K = U.S.synthetic_rays(R, seed=5, start=17, stride=613)
    K = K.clone(); K[:, 0, 1] = 0.37            # exercise the skew terms
    d_ref, rd_ref, cam_ref = U.O.ray_geometry(uv, pose, K)
    d, rd, cam = (torch.empty(R, 3, device=DEV) for _ in range(3))
    _lib.check(built_lib.vfnerf_ray_geometry(R, 0, keep(uv), keep(pose), keep(K),
                                             d.data_ptr(), rd.data_ptr(), cam.data_ptr(), _stream()), "ray_geometry")
    assert torch.equal(d.cpu(), d_ref)               # feeds the bit-exact sample positions
    assert torch.equal(cam.cpu(), cam_ref)
    assert (rd.cpu() - rd_ref).abs().max().item() <= 2e-7


def test_a1_quaternion_pose(built_lib):
    R = 64
    uv, pose, K = U.S.synthetic_rays(R, seed=2)
    g = torch.Generator().manual_seed(3)
    q = torch.randn(R, 4, generator=g); t = torch.randn(R, 3, generator=g)
    pose7 = torch.cat([q, t], dim=1)
    d_ref, rd_ref, cam_ref = U.O.ray_geometry(uv, pose7, K)
    d, rd, cam = (torch.empty(R, 3, device=DEV) for _ in range(3))
    _lib.check(built_lib.vfnerf_ray_geometry(R, 1, keep(uv), keep(pose7), keep(K),
                                             d.data_ptr(), rd.data_ptr(), cam.data_ptr(), _stream()), "ray_geometry")
    # quat_to_rot is .cuda()-only in the reference (pinhole_model.py:22): tolerance, not bit-exact
    assert (d.cpu() - d_ref).abs().max().item() <= 1e-5
    assert torch.equal(cam.cpu(), cam_ref)


@pytest.mark.parametrize("perturb", [0, 1])
@pytest.mark.parametrize("R,Nc,near,far", [(7, 64, 0.0, 6.0), (1024, 100, 0.25, 4.7), (3, 2, 0.0, 1.0)])
def test_a2_coarse_sampler_bit_exact(built_lib, perturb, R, Nc, near, far):
    uv, pose, K = U.S.synthetic_rays(R, seed=1, start=5, stride=977)
    d, _, cam = U.O.ray_geometry(uv, pose, K)
    U1, _, _ = U.S.synthetic_draws(R, Nc, 4, seed=9)
    t_vals = torch.linspace(0., 1., steps=Nc)
    z_ref = U.O.coarse_z_vals(R, near, far, t_vals, bool(perturb), U1)
    p_ref = U.O.sample_points(cam, z_ref, d)
    z = torch.empty(R, Nc, device=DEV); p = torch.empty(R, Nc, 3, device=DEV)
    _lib.check(built_lib.vfnerf_coarse_sample(R, Nc, near, far, perturb, keep(t_vals), keep(U1),
                                              keep(d), keep(cam), z.data_ptr(), p.data_ptr(),
                                              _stream()), "coarse_sample")
    assert torch.equal(z.cpu(), z_ref)
    assert torch.equal(p.cpu(), p_ref)


@pytest.mark.parametrize("perturb", [0, 1])
@pytest.mark.parametrize("R,Nc,Nf,near,far,rng", [(64, 64, 64, 0.0, 6.0, 0.3), (257, 100, 100, 0.25, 4.7, 0.5),
                                                  (5, 24, 20, 0.0, 6.0, 0.3), (9, 130, 126, 0.0, 2.0, 0.1)])
def test_a7_fine_sampler_bit_exact(built_lib, perturb, R, Nc, Nf, near, far, rng):
    uv, pose, K = U.S.synthetic_rays(R, seed=4, start=3, stride=1013)
    d, _, cam = U.O.ray_geometry(uv, pose, K)
    U1, U2, U3 = U.S.synthetic_draws(R, Nc, Nf, seed=11)
    t_vals = torch.linspace(0., 1., steps=Nc)
    z_c = U.O.coarse_z_vals(R, near, far, t_vals, bool(perturb), U1)
    g = torch.Generator().manual_seed(5)
    w_c = torch.rand(R, Nc, generator=g) ** 8
    w_c[0] = 0.0                                   # all-zero weights -> argmax 0 -> z_add branch
    if R > 2:
        w_c[1] = 0.0; w_c[1, 0] = 1.0              # argmax exactly 0
        w_c[2] = 0.0; w_c[2, 5] = 0.5; w_c[2, 11] = 0.5   # tie -> first index
    z_ref = U.O.fine_z_vals(z_c, w_c, near, far, rng, Nf, bool(perturb), U2, U3)
    p_ref = U.O.sample_points(cam, z_ref, d)
    z = torch.empty(R, Nc + Nf, device=DEV); p = torch.empty(R, Nc + Nf, 3, device=DEV)
    _lib.check(built_lib.vfnerf_fine_sample(R, Nc, Nf, near, far, rng, perturb, keep(z_c), keep(w_c),
                                            keep(U2), keep(U3), keep(d),
                                            keep(cam), z.data_ptr(), p.data_ptr(), _stream()), "fine_sample")
    assert torch.equal(z.cpu(), z_ref)
    assert torch.equal(p.cpu(), p_ref)
    assert torch.all(z[:, 1:] >= z[:, :-1])       # sortedness (size-independent property)


@pytest.mark.parametrize("R,N,window,th", [(64, 128, 11, -0.2), (300, 200, 11, -2.0), (17, 44, 11, -0.2),
                                           (8, 16, 11, -0.2), (33, 135, 7, 0.3), (5, 256, 11, -0.2)])
def test_a4_a5_a6_density_and_weights(built_lib, R, N, window, th):
    g = torch.Generator().manual_seed(N)
    # smooth-ish random field so that windowed cosines cover (-1, 1)
    base = torch.randn(R, N // 4 + 2, 3, generator=g)
    normals = torch.nn.functional.interpolate(base.permute(0, 2, 1), size=N, mode="linear").permute(0, 2, 1).contiguous()
    normals = torch.tanh(2 * normals)
    rd = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=1)
    z = torch.sort(torch.rand(R, N, generator=g) * 6, dim=1)[0]
    beta, scale, mean = torch.tensor(0.5), torch.tensor(100.0), torch.tensor(0.7)
    b, s, m = U.O.effective_density_params(beta, scale, mean, (1e-4, 1e9), 1.0, (0.6, 1.0))
    sig_ref, c_ref = U.O.get_density(normals, rd, b, s, m, window, th)
    w_ref = U.O.volsdf_weights(z, sig_ref, True)
    cfg = _cfg(R, window, th)
    dp = torch.stack([beta, scale, mean]).to(DEV)
    c = torch.empty(R, N - 1, device=DEV); sig = torch.empty(R, N, device=DEV); w = torch.empty(R, N, device=DEV)
    _lib.check(built_lib.vfnerf_density_weights(C.byref(cfg), N, dp.data_ptr(), keep(normals), 3,
                                                keep(rd), keep(z), c.data_ptr(), sig.data_ptr(),
                                                w.data_ptr(), _stream()), "density_weights")
    assert (c.cpu() - c_ref).abs().max().item() <= 2e-6           # fp32 cosine arithmetic
    # sigma = 100 * laplace(c): 2e-6 in c is 2e-4 in sigma away from the mask discontinuity
    safe = ~(((c_ref.abs() < 1e-5) & (th > -1)).any(dim=1))
    assert (sig.cpu() - sig_ref)[safe].abs().max().item() <= 1e-3
    assert (w.cpu() - w_ref)[safe].abs().max().item() <= 1e-4      # tolerance 1e-4 abs on weights in [0,1]
    assert (sig_ref > 0).float().mean().item() > 0.01
    assert (w.sum(dim=1).cpu()[safe] <= 1.0 + 1e-5).all()


def test_a9_composite(built_lib):
    R, N = 513, 128
    g = torch.Generator().manual_seed(0)
    w = torch.rand(R, N, generator=g); w = w / w.sum(1, keepdim=True)
    col = torch.rand(R * N, 3, generator=g); z = torch.rand(R, N, generator=g) * 6
    rgb = torch.empty(R, 3, device=DEV); dep = torch.empty(R, 1, device=DEV)
    _lib.check(built_lib.vfnerf_composite(R, N, keep(w), keep(col), keep(z),
                                          rgb.data_ptr(), dep.data_ptr(), _stream()), "composite")
    rgb_ref = (w.unsqueeze(-1) * col.reshape(R, N, 3)).sum(1)
    dep_ref = (w * z).sum(1, keepdim=True)
    assert (rgb.cpu() - rgb_ref).abs().max().item() <= 1e-6
    assert (dep.cpu() - dep_ref).abs().max().item() <= 5e-6


def test_empty_batches_are_noops(built_lib):
    e = torch.empty(0, device=DEV)
    assert built_lib.vfnerf_ray_geometry(0, 0, e.data_ptr(), e.data_ptr(), e.data_ptr(), e.data_ptr(), e.data_ptr(),
                                         e.data_ptr(), _stream()) == 0
    assert built_lib.vfnerf_composite(0, 8, e.data_ptr(), e.data_ptr(), e.data_ptr(), e.data_ptr(), e.data_ptr(), _stream()) == 0


def test_sample_limit_is_an_error_not_a_crash(built_lib):
    e = torch.zeros(4, device=DEV)
    st = built_lib.vfnerf_fine_sample(1, 200, 100, 0.0, 1.0, 0.3, 0, e.data_ptr(), e.data_ptr(), e.data_ptr(), e.data_ptr(),
                                      e.data_ptr(), e.data_ptr(), e.data_ptr(), e.data_ptr(), _stream())
    assert st != 0 and b"exceeds" in built_lib.vfnerf_last_error()


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_fused_per_ray_kernels_equal_the_stage_kernels(built_lib, precision):
    """render() runs the per-ray stages as three fused launches (csrc/render_fused.cu: ray head, coarse-to-fine, tail);
    they share their device code with the stage entry points, so feeding the stage kernels render()'s own intermediate
    results must reproduce its outputs: weights bit for bit, the composite to the last ulp of its FMA chain."""
    case, z = U.load_golden("small_perturb" if precision == "fp32" else "full_perturb")
    model = U.make_model(case, U.case_state(case, z), DEV, precision=precision)
    uv, pose, K = (U.t(z, k).to(DEV) for k in ("uv", "pose", "K"))
    draws = tuple(U.t(z, k).to(DEV) for k in ("U1", "U2", "U3"))
    with torch.no_grad():
        out = model.render(pose, uv, K, 0, draws=draws)
    ex = model.last_extras
    R, N = out.z_vals.shape
    Nc = case["n_coarse"]
    # a1 + a2: the head's coarse z values against the two stage kernels
    d, rd, cam = (torch.empty(R, 3, device=DEV) for _ in range(3))
    _lib.check(built_lib.vfnerf_ray_geometry(R, 0, uv.data_ptr(), pose.data_ptr(), K.data_ptr(), d.data_ptr(), rd.data_ptr(),
                                             cam.data_ptr(), _stream()), "ray_geometry")
    zc, pc = torch.empty(R, Nc, device=DEV), torch.empty(R, Nc, 3, device=DEV)
    t_vals = U.t(z, "t_vals").to(DEV)
    _lib.check(built_lib.vfnerf_coarse_sample(R, Nc, case["near"], case["far"], 1, t_vals.data_ptr(), draws[0].data_ptr(),
                                              d.data_ptr(), cam.data_ptr(), zc.data_ptr(), pc.data_ptr(), _stream()), "coarse")
    assert torch.equal(zc, ex["z_coarse"])
    # a7: the fine sampler stage on render()'s coarse weights reproduces the merged z values and points
    zf, pf = torch.empty(R, N, device=DEV), torch.empty(R, N, 3, device=DEV)
    _lib.check(built_lib.vfnerf_fine_sample(R, Nc, N - Nc, case["near"], case["far"], case["fine_range"], 1, zc.data_ptr(),
                                            ex["weights_coarse"].data_ptr(), draws[1].data_ptr(), draws[2].data_ptr(),
                                            d.data_ptr(), cam.data_ptr(), zf.data_ptr(), pf.data_ptr(), _stream()), "fine")
    assert torch.equal(zf, out.z_vals) and torch.equal(pf, out.points_coarse)
    # a4-a6: the density stage on render()'s merged vectors reproduces its weights
    cfg = model._render_cfg(R, False)
    w = torch.empty(R, N, device=DEV)
    nrm = out.coarse_normals.contiguous()
    _lib.check(built_lib.vfnerf_density_weights(C.byref(cfg), N, model.density.flat().data_ptr(), nrm.data_ptr(), 3,
                                                rd.data_ptr(), out.z_vals.data_ptr(), None, None, w.data_ptr(), _stream()),
               "density_weights")
    assert torch.equal(w, ex["weights"])
    # a9
    rgb, dep = torch.empty(R, 3, device=DEV), torch.empty(R, 1, device=DEV)
    _lib.check(built_lib.vfnerf_composite(R, N, w.data_ptr(), out.coarse_colors.data_ptr(), out.z_vals.data_ptr(),
                                          rgb.data_ptr(), dep.data_ptr(), _stream()), "composite")
    assert (rgb - out.coarse_rgb_values).abs().max().item() <= 1e-6
    assert (dep - out.coarse_depth_map).abs().max().item() <= 1e-6
