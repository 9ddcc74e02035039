"""GPU: tensor-core (tcgen05 / TMEM) building blocks."""
import pytest
import torch

import vfn_testutil as U  # noqa: F401
from vfnerf_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("N,K", [(256, 256), (16, 256), (256, 48), (224, 256), (128, 64), (64, 16)])
def test_umma_descriptor_conventions(debug_lib, N, K):
    g = torch.Generator().manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g).to(DEV)
    B = torch.randn(N, K, generator=g).to(DEV)
    D = torch.zeros(128, N, device=DEV)
    _lib.check_debug(debug_lib.vfnerf_debug_umma_gemm(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, 0,
                                                torch.cuda.current_stream().cuda_stream), "debug_umma_gemm")
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ B.bfloat16().float().T
    err = (D - ref).abs().max().item()
    print(f"N={N} K={K}: max abs err {err:.3e} (ref magnitude {ref.abs().max().item():.1f})")
    assert err <= 1e-3 * max(1.0, ref.abs().max().item())      # fp32 accumulation-order noise only


def _tame_state():
    """A bf16-friendly model: default-like init (gain 1, no output centring): a plain-bf16 chain can meet north_star's
    5e-3 on its per-point outputs.  Its DENSITY is degenerate, so render()-level rgb / depth claims
    are never made on it: those live in tests/test_gpu_x3.py, on the golden model, for the mode that meets them."""
    return U.S.synthetic_state(0, vf_gain=1.0, center_output=False)


# Every expected value below comes from the CPU oracle (oracle/render_oracle.py, pinned to the live reference by
# tests/golden/make_golden.py) or from the reference goldens -- never from this repo's own fp32 kernels.

@pytest.mark.parametrize("P", [1, 127, 128, 129, 5000, 300 * 128 + 17])
def test_tc_vf_query_matches_oracle(built_lib, P):
    """bf16 tcgen05 chain vs the oracle's VF MLP on the same points (tail tiles, multi-tile CTAs)."""
    from vfnerf_b200.ops import vf_query
    case, z = U.load_golden("full_det")
    st = _tame_state()
    model = U.make_model(case, st, DEV, precision="bf16")
    g = torch.Generator().manual_seed(P)
    pts = (torch.rand(P, 3, generator=g) - 0.5) * 8
    with torch.no_grad():
        ref = U.O.vf_network(st["vf_net"], pts)
        out = model.vector_field_network(pts.to(DEV)).cpu()
        v3 = vf_query(model.vector_field_network, pts.to(DEV), n_cols=3).cpu()
    err = (out - ref).abs()
    print(f"P={P}: bf16 vs oracle  v max {err[:, :3].max().item():.2e} mean {err[:, :3].mean().item():.2e} | "
          f"feat max {err[:, 3:].max().item():.2e} mean {err[:, 3:].mean().item():.2e}")
    assert out.shape == ref.shape and torch.isfinite(out).all()
    assert torch.equal(v3, out[:, :3])                    # V_ONLY and VF_FULL programs agree on the vector
    assert err.max().item() <= 5e-3                       # BASELINE.json north_star: 5e-3 abs on the bf16 path


def test_tc_bf16_on_the_golden_model_is_outside_the_tolerance_and_says_so(built_lib):
    """The golden (non-degenerate, centred, gain 2) model amplifies operand rounding ~60x: plain bf16 operands give
    normals that differ from the REFERENCE by up to ~0.13 -- outside north_star's 5e-3.  Plain bf16 therefore does not
    claim parity on this model (precision="bf16x3" does: tests/test_gpu_x3.py).  This test pins the size of the gap
    against the reference goldens so that it cannot grow silently.  (The colour net stays plain bf16 in both modes and
    is not the problem: with a split-precision VF chain in front of it the colours are within 2e-4, test_gpu_x3.py.)"""
    case, z = U.load_golden("full_det")
    st = U.case_state(case, z)
    model = U.make_model(case, st, DEV, precision="bf16")
    uv, pose, K = (U.t(z, k).to(DEV) for k in ("uv", "pose", "K"))
    draws = (U.t(z, "U1"), U.t(z, "U2"), U.t(z, "U3"))
    with torch.no_grad():
        out = model.render(pose, uv, K, 0, draws=draws, z_vals_override=U.t(z, "ref_z_vals"))
    dn = (out.coarse_normals.cpu() - U.t(z, "ref_normals")).abs()
    dc = (out.coarse_colors.cpu() - U.t(z, "ref_colors")).abs()
    print(f"bf16 vs reference golden (bending model): normals max {dn.max().item():.2e} mean {dn.mean().item():.2e} | "
          f"colours max {dc.max().item():.2e} mean {dc.mean().item():.2e}")
    assert torch.equal(out.points_coarse.cpu(), U.t(z, "ref_points"))
    assert dn.mean().item() <= 3e-2 and dn.max().item() <= 0.3          # the documented envelope of the fast mode
    assert dc.mean().item() <= 1e-2


def test_tc_render_per_point_outputs_match_oracle(built_lib):
    """Fused VF + colour launch inside render() (plain bf16) vs the ORACLE's render on the bf16-friendly model, second
    pass conditioned on the oracle's z: per-point normals and colours within 5e-3; sample positions bit-exact."""
    case, z = U.load_golden("full_perturb")
    st = _tame_state()
    model = U.make_model(case, st, DEV, precision="bf16")
    R = 256
    uv, pose, K = U.S.synthetic_rays(R, seed=0, start=0, stride=797)
    draws = U.S.synthetic_draws(R, 64, 64, seed=99)
    with torch.no_grad():
        ora = U.O.render(st["vf_net"], st["rendering_net"], st["density"], U.oracle_cfg(case), uv, pose, K,
                         torch.linspace(0., 1., 64), *draws)
        out = model.render(pose.to(DEV), uv.to(DEV), K.to(DEV), 0, draws=draws, z_vals_override=ora["z_vals"])
    N = 128
    dn = (out.coarse_normals.cpu() - ora["normals"]).abs()
    dc = (out.coarse_colors.cpu() - ora["colors"]).abs()
    print(f"render bf16 vs oracle: normals max {dn.max().item():.2e} | colours max {dc.max().item():.2e}")
    assert torch.equal(out.points_coarse.cpu(), ora["points"])
    assert dn.max().item() <= 5e-3 and dc.max().item() <= 5e-3
    assert out.ray_dirs.shape == (R * N, 3) and (out.ray_dirs.cpu() - ora["rep_ray_dirs"]).abs().max().item() <= 1e-6


def test_tc_grid_query_matches_oracle(built_lib):
    from vfnerf_b200.grid_query import grid_query
    case, z = U.load_golden("full_det")
    st = _tame_state()
    model = U.make_model(case, st, DEV, precision="bf16")
    res = 40
    tr, ce = torch.tensor([0.5, -0.5, 0.5]), torch.tensor([0.1, 0.0, -0.2])
    with torch.no_grad():
        ref = U.O.vf_network(st["vf_net"], U.reference_grid_points(res, 1.0, tr, ce))[:, :3]
    out = grid_query(model.vector_field_network, res, 1.0, tr, ce, chunk=10000)
    slab = grid_query(model.vector_field_network, res, 1.0, tr, ce, i0=res * res * 7, n_points=res * res * 3)
    assert (out.cpu() - ref).abs().max().item() <= 5e-3
    assert torch.equal(slab, out[res * res * 7: res * res * 10])      # z-slab partition == slice of the whole grid


@pytest.mark.parametrize("N,K", [(256, 256), (32, 256), (256, 64), (224, 96)])
def test_umma_2cta_conventions(debug_lib, N, K):
    """cta_group::2: A rows split across the CTA pair, B rows (output channels) split in halves, M = 256."""
    g = torch.Generator().manual_seed(N + K)
    A = torch.randn(256, K, generator=g).to(DEV)
    B = torch.randn(N, K, generator=g).to(DEV)
    D = torch.zeros(256, N, device=DEV)
    _lib.check_debug(debug_lib.vfnerf_debug_umma2_gemm(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K,
                                                 torch.cuda.current_stream().cuda_stream), "debug_umma2_gemm")
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ B.bfloat16().float().T
    err = (D - ref).abs().max().item()
    print(f"2-CTA N={N} K={K}: max abs err {err:.3e}")
    assert err <= 1e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("a_fmt,b_fmt", [(2, 2), (0, 1), (1, 0), (0, 0)])
@pytest.mark.parametrize("N,K", [(256, 256), (224, 64), (64, 32)])
def test_umma_fp16_and_8bit_operand_conventions(debug_lib, N, K, a_fmt, b_fmt):
    """The operand kinds of the fp16 + fp8-remainder chain (precision "fp16f8"): kind::f16 with fp16 operands (format 2
    here) and kind::f8f6f4 with e4m3 (0) / e5m2 (1) operands, K = 32 per instruction, 16 columns per 16-byte unit of the
    K-slab.  Products of 8-bit operands are exact in fp32, so the result must equal the fp64 product of the rounded
    operands up to fp32 accumulation."""
    def rnd(x, fmt):
        return x.half().float() if fmt == 2 else x.to(torch.float8_e5m2 if fmt else torch.float8_e4m3fn).float()
    g = torch.Generator().manual_seed(N + K + 7 * a_fmt + b_fmt)
    A = torch.randn(256, K, generator=g).to(DEV)
    B = (torch.randn(N, K, generator=g) * 0.5).to(DEV)
    D = torch.zeros(256, N, device=DEV)
    _lib.check_debug(debug_lib.vfnerf_debug_umma2_alt_gemm(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, a_fmt, b_fmt,
                                                     torch.cuda.current_stream().cuda_stream), "debug_umma2_alt_gemm")
    torch.cuda.synchronize()
    ref = rnd(A, a_fmt).double() @ rnd(B, b_fmt).double().T
    err = (D.double() - ref).abs().max().item()
    assert err <= 5e-6 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("N,K", [(256, 128), (48, 128), (256, 64), (96, 16)])
def test_umma_mn_major_conventions(debug_lib, N, K):
    """Both operands MN-major (reduction index = row of the stashed [points, channels] tiles): the wgrad GEMM."""
    g = torch.Generator().manual_seed(3 * N + K)
    At = torch.randn(K, 128, generator=g).to(DEV)
    Bt = torch.randn(K, N, generator=g).to(DEV)
    ref = At.bfloat16().float().T @ Bt.bfloat16().float()
    errs = []
    for variant in (0, 1):
        D = torch.zeros(128, N, device=DEV)
        _lib.check_debug(debug_lib.vfnerf_debug_umma_mn_gemm(At.data_ptr(), Bt.data_ptr(), D.data_ptr(), N, K, variant,
                                                       torch.cuda.current_stream().cuda_stream), "debug_umma_mn_gemm")
        torch.cuda.synchronize()
        errs.append((D - ref).abs().max().item())
    print(f"MN-major N={N} K={K}: max abs err variant0 {errs[0]:.3e} variant1 {errs[1]:.3e}")
    assert errs[0] <= 1e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("R,n_coarse,n_fine", [(3, 40, 24), (77, 64, 36), (1, 64, 64)])
def test_tc_render_ragged_sample_counts(built_lib, R, n_coarse, n_fine):
    """Sample counts for which R*N is not a multiple of the 128-point tile (partial last tile, odd tile counts, the
    caller-mutable sampler attributes of SURVEY.md §8b): bf16 forward vs the ORACLE, and a bf16 backward that stays
    within the gate-flip bound of the fp32 gradients (which test_gpu_backward.py pins to the reference's gradients)."""
    case, z = U.load_golden("full_perturb")
    case = dict(case, n_coarse=n_coarse, n_fine=n_fine, max_samples=100)
    st = _tame_state()
    uv, pose, K = U.S.synthetic_rays(R, seed=0, start=11, stride=797)
    draws = U.S.synthetic_draws(R, n_coarse, n_fine, seed=7)
    a = (pose.to(DEV), uv.to(DEV), K.to(DEV), 0)
    with torch.no_grad():
        ora = U.O.render(st["vf_net"], st["rendering_net"], st["density"], U.oracle_cfg(case), uv, pose, K,
                         torch.linspace(0., 1., n_coarse), *draws)
    grads = {}
    outs = {}
    for prec in ("fp32", "bf16"):
        model = U.make_model(case, st, DEV, precision=prec)
        out = model.render(*a, draws=draws, z_vals_override=ora["z_vals"])
        model.optimizer.zero_grad()
        (out.coarse_rgb_values.sum() + out.coarse_depth_map.sum() + (out.coarse_normals ** 2).sum() * 0.01).backward()
        outs[prec] = out
        grads[prec] = [p.grad.detach().clone() for p in model.rendering_network.parameters()] + \
                      [p.grad.detach().clone() for p in model.vector_field_network.parameters()]
    N = n_coarse + n_fine
    o = outs["bf16"]
    assert o.coarse_normals.shape == (R, N, 3)
    assert torch.equal(o.points_coarse.cpu(), ora["points"])
    assert (o.coarse_normals.detach().cpu() - ora["normals"]).abs().max().item() <= 5e-3
    assert (o.coarse_colors.detach().cpu() - ora["colors"]).abs().max().item() <= 5e-3
    for ga, gb in zip(grads["fp32"], grads["bf16"]):
        assert torch.isfinite(gb).all()
        assert (ga - gb).norm() <= 0.2 * ga.norm() + 1e-7


@pytest.mark.parametrize("R,n_coarse,n_fine,perturb", [(512, 64, 64, True), (333, 64, 64, False), (77, 100, 30, True),
                                                        (5, 40, 24, False), (1, 20, 16, True)])
def test_tc_render_reusing_coarse_results_is_bit_identical(built_lib, R, n_coarse, n_fine, perturb):
    """bf16 forward-only render(): by default the coarse sweep runs both MLPs and the merged pass evaluates only the
    fine candidates (include/vfnerf_b200.h, VFNERF_FLAG_RECOMPUTE_COARSE).  Every output field must equal, bit for
    bit, the literal schedule of vector_field_nerf.py:252-312 (VF on the coarse points, then both MLPs on all merged
    points): same points, and the fused chain is a pure function of (point, ray direction)."""
    case, z = U.load_golden("full_perturb" if perturb else "full_det")
    case = dict(case, n_coarse=n_coarse, n_fine=n_fine, max_samples=100, perturb=perturb)
    model = U.make_model(case, U.case_state(case, z), DEV, precision="bf16")
    uv, pose, K = U.S.synthetic_rays(R, seed=0, start=5, stride=797)
    draws = U.S.synthetic_draws(R, n_coarse, n_fine, seed=21)
    a = (pose.to(DEV), uv.to(DEV), K.to(DEV), 0)
    with torch.no_grad():
        fast = model.render(*a, draws=draws)
        w_fast = model.last_extras["weights"].clone()
        model.recompute_coarse = True
        lit = model.render(*a, draws=draws)
        w_lit = model.last_extras["weights"]
    for f in ("points_coarse", "z_vals", "coarse_normals", "coarse_colors", "coarse_rgb_values", "coarse_depth_map",
              "ray_dirs"):
        assert torch.equal(getattr(fast, f), getattr(lit, f)), f
    assert torch.equal(w_fast, w_lit)
    assert (fast.z_vals[:, 1:] >= fast.z_vals[:, :-1]).all()
    assert fast.coarse_rgb_values.abs().sum().item() > 0 or R < 8
