"""Golden vectors of the inverse-CDF fine sampler (FineSampler, models/samplers/ray_sampler.py:145-237) from the LIVE
reference.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_pdf.py

The reference class is instantiated unmodified; its ``torch.rand`` draw (:183) is replaced by a supplied tensor by
patching ``torch.rand`` inside the sampler module, exactly like make_golden.py does for render().  The oracle
restatement (oracle/render_oracle.py: sample_pdf, pdf_fine_z_vals) is asserted bit-equal to the reference on CPU and the
inputs + reference outputs are written to tests/golden/pdf_sampler.npz.
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
import models.samplers.ray_sampler as ref_sampler           # noqa: E402


def weights_like_render(R, Nc, gen):
    """Compositing-weight shaped rows: a few sharp peaks over a tiny floor, some all-zero rows, some flat rows."""
    w = torch.rand(R, Nc, generator=gen) ** 8
    w[::7] = 0.0
    w[1::11] = 0.25
    peak = torch.randint(0, Nc, (R,), generator=gen)
    w[torch.arange(R), peak] += (torch.arange(R) % 3 == 0).float() * 5.0
    return w


def main():
    gen = torch.Generator().manual_seed(2024)
    out = {}
    for tag, (R, Nc, Nf, det) in {"det": (96, 64, 64, True), "rand": (96, 64, 64, False), "ragged": (33, 100, 30, False),
                                  "tiny": (4, 3, 5, False)}.items():
        near, far = 0.0, 6.0
        t = torch.linspace(0., 1., Nc)
        z_c = (near * (1 - t) + far * t).repeat(R, 1)
        # stratified coarse samples, like a perturbed UniformSampler
        mids = .5 * (z_c[:, 1:] + z_c[:, :-1])
        upper, lower = torch.cat([mids, z_c[:, -1:]], -1), torch.cat([z_c[:, :1], mids], -1)
        z_c = lower + (upper - lower) * torch.rand(R, Nc, generator=gen)
        w_c = weights_like_render(R, Nc, gen)
        u_draw = torch.rand(R, Nf, generator=gen)
        sampler = ref_sampler.FineSampler(Nf, deterministic=det)
        real_rand = torch.rand
        ref_sampler.torch.rand = lambda *a, **k: u_draw.clone()
        try:
            ref_z = sampler.get_z_vals(None, None, torch.device("cpu"), coarse_z_vals=z_c, coarse_weights=w_c)
            ref_s = sampler.sample_pdf(.5 * (z_c[..., 1:] + z_c[..., :-1]), w_c[..., 1:-1])
        finally:
            ref_sampler.torch.rand = real_rand
        u = torch.linspace(0., 1., steps=Nf) if det else u_draw
        assert torch.equal(O.pdf_fine_z_vals(z_c, w_c, u), ref_z), tag
        assert torch.equal(O.sample_pdf(.5 * (z_c[..., 1:] + z_c[..., :-1]), w_c[..., 1:-1], u), ref_s), tag
        for k, v in (("z_c", z_c), ("w_c", w_c), ("u", u), ("ref_z", ref_z), ("ref_samples", ref_s)):
            out[f"{tag}.{k}"] = v.numpy()
        print(f"{tag}: R={R} Nc={Nc} Nf={Nf} deterministic={det}: oracle == reference (bit-exact)")
    np.savez_compressed(os.path.join(HERE, "pdf_sampler.npz"), **out)


if __name__ == "__main__":
    main()
