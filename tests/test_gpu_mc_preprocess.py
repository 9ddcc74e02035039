"""GPU: marching-cubes preprocessing on the device (csrc/mc_preprocess.cu) through the C ABI vs the
golden vectors produced by the reference's own extract_divergence / unify_direction / make_comb_format chain
(tests/golden/make_golden_mc.py) and vs the oracle restatement on fresh grids."""
import os

import numpy as np
import pytest
import torch

import vfn_testutil as U
from oracle import mc_oracle as MO

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GUARD = 1e-4        # cells whose divergence is this close to the reference's -0.5 threshold may legitimately flip


def _golden(tag):
    z = np.load(os.path.join(U.GOLDEN_DIR, "mc_preprocess.npz"))
    return {k: torch.from_numpy(z[f"{tag}.{k}"]) for k in ("pred", "div", "choice", "cells", "comb", "udf", "raw_div")}


def _check(pred, N, ref_div, ref_choice, ref_cells, ref_comb, ref_udf, ref_raw):
    from vfnerf_b200 import mc_utils as M
    flag, raw = M.extract_divergence(pred.to(DEV), N, return_raw=True)
    flag, raw = flag.cpu(), raw.cpu()
    # raw divergence of the interior cells the reference meshes: fp32 op-order noise only
    H = 2 * (N // 2)
    lim = min(N - 1, H)
    err = (raw[:lim, :lim, :lim] - ref_raw[:lim, :lim, :lim]).abs().max().item()
    assert err <= 2e-5, err
    safe = torch.ones(N, N, N, dtype=torch.bool)
    safe[:N - 1, :N - 1, :N - 1] = (ref_raw + 0.5).abs() > GUARD
    inside = torch.zeros(N, N, N, dtype=torch.bool)
    inside[:H, :H, :H] = True
    assert torch.equal(flag[safe & inside], ref_div[safe & inside])
    # side choices of the surface cells (bit-exact arithmetic order; near-ties of the 64-way argmax may flip: bounded)
    ch = M.unify_direction(None, torch.nn.functional.normalize(pred, dim=1).reshape(N, N, N, 3).permute(3, 0, 1, 2).to(DEV), N)
    ch = ch.cpu().reshape(N, N, N, 8)
    surf = (ref_div == 1) & safe & inside
    same = (ch[surf] == ref_choice[surf].long()).all(dim=1).float().mean().item()
    assert same >= 0.995, same
    # the fused, compacted outputs
    cells, comb, udf = (t.cpu() for t in M.mc_preprocess(pred.to(DEV), N))
    key = lambda c: (c[:, 0].long() * N + c[:, 1].long()) * N + c[:, 2].long()
    ka, kb = key(cells), key(ref_cells)
    common = np.intersect1d(ka.numpy(), kb.numpy())
    unsafe_cells = int((~safe & inside).sum())
    assert len(ka) - len(common) <= unsafe_cells and len(kb) - len(common) <= unsafe_cells
    # identical order (block order of methods.py:186-192) on the common cells
    ia = torch.from_numpy(np.isin(ka.numpy(), common)); ib = torch.from_numpy(np.isin(kb.numpy(), common))
    assert torch.equal(ka[ia], kb[ib])
    agree = (comb[ia] == ref_comb[ib]).all(dim=1).float().mean().item()
    assert agree >= 0.995, agree
    assert (udf[ia] - ref_udf[ib]).abs().max().item() <= 1e-6
    return err, same, agree, len(ka), len(kb)


@pytest.mark.parametrize("tag,N", [("n24", 24), ("n33", 33)])
def test_mc_preprocess_matches_reference_golden(built_lib, tag, N):
    g = _golden(tag)
    err, same, agree, ma, mb = _check(g["pred"], N, g["div"], g["choice"], g["cells"], g["comb"], g["udf"], g["raw_div"])
    print(f"{tag}: raw divergence max err {err:.1e}; surface cells with all 8 side choices equal {same:.4f}; emitted {ma} "
          f"(reference {mb}); rows with all 28 flags equal {agree:.4f}")


def test_mc_preprocess_matches_oracle_on_a_larger_grid(built_lib):
    N = 64
    pred = U.S.synthetic_vector_grid(N, seed=5)
    div = MO.extract_divergence(pred, N)
    choice = MO.unify_direction(div, pred, N)
    cells, comb, udf = MO.mc_preprocess(pred, N)
    err, same, agree, ma, mb = _check(pred, N, div, choice, cells.int(), comb, udf, MO.divergence(pred, N))
    print(f"N=64: raw divergence max err {err:.1e}; choices equal {same:.4f}; emitted {ma} (oracle {mb}); rows equal {agree:.4f}")
    assert mb > 1000


def test_mc_preprocess_edge_cases(built_lib):
    from vfnerf_b200 import mc_utils as M
    # a constant field has no surface: nothing is emitted
    N = 16
    pred = torch.tensor([0.3, -0.2, 0.9]).repeat(N ** 3, 1)
    cells, comb, udf = M.mc_preprocess(pred.to(DEV), N)
    assert cells.shape == (0, 3) and comb.shape == (0, 28) and udf.shape == (0, 28, 2)
    # zero vectors normalise to zero (F.normalize eps), divergence 0, no surface
    cells, _, _ = M.mc_preprocess(torch.zeros(N ** 3, 3, device=DEV), N)
    assert cells.shape[0] == 0
    with pytest.raises(RuntimeError):
        M.mc_preprocess(pred, N)                       # host tensor: no CPU path
    with pytest.raises(ValueError):
        M.mc_preprocess(pred[:-1].to(DEV), N)


def test_smooth_vf_and_smooth_after_chain_match_reference_golden(built_lib):
    """smooth_vf as three 1-D passes (csrc/mc_preprocess.cu) vs the reference's k^3-tap conv3d, and the smooth_after
    variant of the chain (divergence of the raw field, sides and norms of the k = 9 smoothed one) through
    mc_preprocess(..., surface=...)."""
    from vfnerf_b200 import mc_utils as M
    z = np.load(os.path.join(U.GOLDEN_DIR, "mc_smooth.npz"))
    pred = torch.from_numpy(z["pred"])
    N = round(pred.shape[0] ** (1 / 3))
    for k, sigma in ((3, 1.0), (9, 2.0)):
        got = M.smooth_vf(pred.reshape(N, N, N, 3).to(DEV), k=k, sigma=sigma).cpu()
        err = (got - torch.from_numpy(z[f"smooth_k{k}"])).abs().max().item()
        print(f"smooth_vf k={k}: max abs err {err:.1e}")
        assert err <= 2e-6
    div = M.extract_divergence(pred.to(DEV), N)
    sm = M.smooth_vf(pred.reshape(N, N, N, 3).to(DEV), k=9, sigma=2.0).reshape(N ** 3, 3)
    cells, comb, udf = (t.cpu() for t in M.mc_preprocess(sm, N, surface=div))
    ref_cells, ref_comb, ref_udf = (torch.from_numpy(z[f"after.{k}"]) for k in ("cells", "comb", "udf"))
    assert torch.equal(cells, ref_cells)
    agree = (comb == ref_comb).all(dim=1).float().mean().item()
    assert agree >= 0.995, agree
    assert (udf - ref_udf).abs().max().item() <= 2e-6
    with pytest.raises(ValueError):
        M.smooth_vf(pred.to(DEV), 3, 1.0)              # wants [N,N,N,3]
