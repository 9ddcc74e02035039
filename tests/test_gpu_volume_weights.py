"""GPU: the stand-alone weight functions (utils/rendering.py:98-148) through the C ABI vs goldens from the live reference."""
import os

import numpy as np
import pytest
import torch

import vfn_testutil as U

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("tag", ["n128", "n200", "n2"])
def test_volume_weights_match_reference_golden(built_lib, tag):
    from vfnerf_b200 import rendering as VR
    z = np.load(os.path.join(U.GOLDEN_DIR, "volume_weights.npz"))
    zz, sg = torch.from_numpy(z[f"{tag}.z"]).to(DEV), torch.from_numpy(z[f"{tag}.sigma"]).to(DEV)
    for norm in (0, 1):
        a = VR.nerf_volume_rendering(sg, zz, normalize=bool(norm)).cpu()
        b = VR.volsdf_volume_rendering(zz, sg, normalize=bool(norm)).cpu()
        ea = (a - torch.from_numpy(z[f"{tag}.nerf.{norm}"])).abs().max().item()
        eb = (b - torch.from_numpy(z[f"{tag}.volsdf.{norm}"])).abs().max().item()
        ma = (a - torch.from_numpy(z[f"{tag}.nerf.{norm}"])).abs().mean().item()
        print(f"{tag} normalize={norm}: nerf max abs err {ea:.1e} (mean {ma:.1e}), volsdf {eb:.1e}")
        # unnormalised weights agree to one fp32 ulp (6e-8: scan order, expf).  normalize=True divides by (sum w + 1e-5):
        # on nearly empty rays (sum w ~ 1e-4) that one ulp becomes ~1e-4 of the normalised weight -- loose max, tight mean
        assert eb <= 2e-6 and ea <= 2e-3 and ma <= 2e-6
    with pytest.raises(RuntimeError):
        VR.nerf_volume_rendering(sg.cpu(), zz.cpu())
