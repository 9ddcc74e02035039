"""Golden vectors of smooth_vf (evaluation/utils/guassian_smoothing.py:81-97) and of the smooth_after variant of the mesh
preprocessing (evaluation/methods.py:209-278) from the LIVE reference.  Run in the build container only:

    python tests/golden/make_golden_smooth.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VFNERF_REF", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import mc_oracle as MO                          # noqa: E402
from vfnerf_b200 import synthetic as S                      # noqa: E402
import evaluation.utils.mc_utils as ref_mc                  # noqa: E402
from evaluation.utils.guassian_smoothing import smooth_vf   # noqa: E402


def main():
    N = 20
    pred = S.synthetic_vector_grid(N, seed=3)
    out = {"pred": pred.numpy()}
    for k, sigma in ((3, 1.0), (9, 2.0)):
        ref = smooth_vf(pred.reshape(N, N, N, 3), k=k, sigma=sigma)
        mine = MO.smooth_vf(pred.reshape(N, N, N, 3), k, sigma)
        err = (ref - mine).abs().max().item()
        assert err <= 2e-6, err                  # conv3d's summation order vs the explicit sum
        out[f"smooth_k{k}"] = ref.contiguous().numpy()
        print(f"smooth_vf k={k} sigma={sigma}: oracle vs reference max abs err {err:.1e}")
    # smooth_after chain, verbatim call sequence of methods.py:209-278
    inc = np.array([[0, 0, 0], [0, 1, 0], [1, 1, 0], [1, 0, 0], [0, 0, 1], [0, 1, 1], [1, 1, 1], [1, 0, 1]])
    sel = np.moveaxis(np.mgrid[: N // 2, : N // 2, : N // 2], 0, -1).reshape(-1, 3)
    sel = (sel[:, None] * 2 + inc[None]).reshape(-1, 3)
    div = ref_mc.extract_divergence(pred, N)
    sm = smooth_vf(pred.reshape(N, N, N, 3), k=9, sigma=2).reshape(N ** 3, 3)
    norms = torch.norm(sm.clone(), dim=1)
    vt = F.normalize(sm, dim=1).reshape(N, N, N, 3)
    choice = ref_mc.unify_direction(div, vt.permute(3, 0, 1, 2), N=N)
    comb, nrm = ref_mc.make_comb_format(choice, norms, N)
    comb = comb.reshape(N, N, N, 28)[sel[:, 0], sel[:, 1], sel[:, 2]]
    nrm = nrm.reshape(N, N, N, 28, 2)[sel[:, 0], sel[:, 1], sel[:, 2]]
    mask = comb.sum(-1) > 0
    cells = torch.from_numpy(sel)[mask]
    c2, k2, u2 = MO.mc_preprocess_smooth_after(pred, N)
    assert torch.equal(c2, cells) and torch.equal(k2, comb[mask]) and (u2 - nrm[mask]).abs().max().item() <= 2e-6
    out["after.cells"] = cells.int().numpy()
    out["after.comb"] = comb[mask].numpy()
    out["after.udf"] = nrm[mask].numpy()
    print(f"smooth_after chain: {cells.shape[0]} cells emitted; oracle == reference (norms to 2e-6)")
    np.savez_compressed(os.path.join(HERE, "mc_smooth.npz"), **out)


if __name__ == "__main__":
    main()
