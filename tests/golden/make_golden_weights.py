"""Golden vectors of the stand-alone weight functions of utils/rendering.py (nerf_volume_rendering :98-119,
volsdf_volume_rendering :122-148) from the LIVE reference.  Run in the build container only:

    python tests/golden/make_golden_weights.py

Seeded densities (mostly zero with a few dense runs, like the Laplace density of a surface crossing) and sorted z values;
the oracle restatements (oracle/render_oracle.py: nerf_weights, volsdf_weights) are asserted bit-equal to the reference.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VFNERF_REF", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import render_oracle as O                       # noqa: E402
import utils.rendering as ref_r                             # noqa: E402


def main():
    g = torch.Generator().manual_seed(77)
    out = {}
    for tag, (R, N) in {"n128": (64, 128), "n200": (17, 200), "n2": (5, 2)}.items():
        z = torch.sort(torch.rand(R, N, generator=g) * 6.0, dim=-1)[0]
        sigma = torch.rand(R, N, generator=g) ** 6 * 80.0
        sigma[torch.rand(R, N, generator=g) < 0.6] = 0.0
        sigma[::5] = 0.0                                     # empty rays
        for norm in (False, True):
            a = ref_r.nerf_volume_rendering(sigma, z, normalize=norm)
            b = ref_r.volsdf_volume_rendering(z, sigma, normalize=norm)
            assert torch.equal(O.nerf_weights(sigma, z, norm), a), (tag, norm)
            assert torch.equal(O.volsdf_weights(z, sigma, norm), b), (tag, norm)
            out[f"{tag}.nerf.{int(norm)}"] = a.numpy()
            out[f"{tag}.volsdf.{int(norm)}"] = b.numpy()
        out[f"{tag}.z"] = z.numpy()
        out[f"{tag}.sigma"] = sigma.numpy()
        print(f"{tag}: R={R} N={N}: oracle == reference (bit-exact)")
    np.savez_compressed(os.path.join(HERE, "volume_weights.npz"), **out)


if __name__ == "__main__":
    main()
