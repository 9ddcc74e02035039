"""Golden vectors of the marching-cubes preprocessing (evaluation/utils/mc_utils.py, evaluation/methods.py:209-278) from
the LIVE reference.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_mc.py

A seeded synthetic vector field (unit vectors pointing at a sphere of radius 0.6 and at a plane, tanh-squashed, plus
noise) on an N^3 grid goes through the reference's own extract_divergence -> unify_direction -> make_comb_format ->
block-ordered compaction; the oracle restatement (oracle/mc_oracle.py) is asserted equal, and inputs + reference
outputs are written to tests/golden/mc_preprocess.npz.
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


def reference_chain(pred, N):
    """methods.py:176-192 and :209-278 with the default flags, verbatim call sequence."""
    inc = np.array([[0, 0, 0], [0, 1, 0], [1, 1, 0], [1, 0, 0], [0, 0, 1], [0, 1, 1], [1, 1, 1], [1, 0, 1]])
    sel = np.mgrid[: int(N / 2), : int(N / 2), : int(N / 2)]
    sel = np.moveaxis(sel, 0, -1).reshape(-1, 3)
    sel = (sel[:, None] * 2 + inc[None]).reshape(-1, 3)
    div = ref_mc.extract_divergence(pred, N)
    norms = torch.norm(pred.clone(), dim=1)
    vt = F.normalize(pred, dim=1).reshape(N, N, N, 3)
    choice = ref_mc.unify_direction(div, vt.permute(3, 0, 1, 2), N=N)
    comb, nrm = ref_mc.make_comb_format(choice, norms, N)
    comb = comb.reshape(N, N, N, 28)[sel[:, 0], sel[:, 1], sel[:, 2]]
    nrm = nrm.reshape(N, N, N, 28, 2)[sel[:, 0], sel[:, 1], sel[:, 2]]
    mask = comb.sum(-1) > 0
    return div, choice.reshape(N, N, N, 8), torch.from_numpy(sel)[mask], comb[mask], nrm[mask]


def main():
    out = {}
    for tag, N in (("n24", 24), ("n33", 33)):
        pred = S.synthetic_vector_grid(N, seed=N)
        div, choice, cells, comb, udf = reference_chain(pred, N)
        assert torch.equal(MO.extract_divergence(pred, N), div), tag
        assert torch.equal(MO.unify_direction(div, pred, N), choice), tag
        c2, k2, u2 = MO.mc_preprocess(pred, N)
        assert torch.equal(c2, cells) and torch.equal(k2, comb) and torch.equal(u2, udf), tag
        raw = MO.divergence(pred, N)
        print(f"{tag}: N={N} surface cells {int(div.sum())}, emitted {cells.shape[0]}; cells within 1e-4 of the threshold: "
              f"{int(((raw + 0.5).abs() < 1e-4).sum())}; oracle == reference")
        for k, v in (("pred", pred), ("div", div), ("choice", choice.to(torch.uint8)), ("cells", cells.int()), ("comb", comb),
                     ("udf", udf), ("raw_div", raw)):
            out[f"{tag}.{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "mc_preprocess.npz"), **out)


if __name__ == "__main__":
    main()
