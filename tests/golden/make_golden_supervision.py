"""Golden vectors for the supervision-point helpers from the LIVE reference (run in the build container, where
/root/reference exists):  python tests/golden/make_golden_supervision.py  ->  tests/golden/supervision.npz

The reference draws with numpy's global generator inside SphereSampler.sample; np.random.uniform is wrapped to capture
the three draws of every call so the oracle and the kernels can be fed the same numbers."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from models.helpers import functions as RF      # noqa: E402  (the reference)
from oracle import supervision_oracle as SO     # noqa: E402


def capture(fn, *args):
    rec = []
    real = np.random.uniform

    def spy(lo, hi, n):
        v = real(lo, hi, n)
        rec.append(v.copy())
        return v
    np.random.uniform = spy
    try:
        out = fn(*args)
    finally:
        np.random.uniform = real
    assert len(rec) == 3
    return out, rec


def main():
    np.random.seed(20260)
    centroid = torch.tensor([0.15, -0.3, 0.45])
    far, radius, n = 6.0, 0.35, 4099
    out = {"centroid": centroid.numpy(), "far": far, "radius": radius}
    (bp, bg), d = capture(RF.sample_border_points, far / 2 - radius, far / 2, n, centroid)
    op, og = SO.sample_border_points(far / 2 - radius, far / 2, centroid, *d)
    assert torch.equal(op, bp) and torch.equal(og, bg)
    out.update(border_phi=d[0], border_cos=d[1], border_u=d[2], border_points=bp.numpy(), border_gt=bg.numpy())
    (cp, cg), d = capture(RF.sample_center_points, centroid, radius, n)
    op, og = SO.sample_center_points(centroid, radius, *d)
    assert torch.equal(op, cp) and torch.equal(og, cg)
    out.update(center_phi=d[0], center_cos=d[1], center_u=d[2], center_points=cp.numpy(), center_gt=cg.numpy())
    # ray samples: points spread so that both conditions select a few per cent
    g = torch.Generator().manual_seed(7)
    pts = centroid + torch.randn(64, 48, 3, generator=g) * 1.4
    pts[:, :6] = centroid + torch.randn(64, 6, 3, generator=g) * 0.25
    nrm = torch.tanh(torch.randn(64, 48, 3, generator=g))
    bn, bgt = RF.get_border_indices_and_gt(pts, nrm, far, radius, centroid)
    on, ogt = SO.get_border_indices_and_gt(pts, nrm, far, radius, centroid)
    assert torch.equal(bn, on) and torch.equal(bgt, ogt) and 20 < bn.shape[0] < pts.shape[0] * pts.shape[1]
    cn, cgt = RF.get_center_indices_and_gt(pts, nrm, centroid, radius)
    on, ogt = SO.get_center_indices_and_gt(pts, nrm, centroid, radius)
    assert torch.equal(cn, on) and torch.equal(cgt, ogt) and cn.shape[0] > 20
    out.update(ray_points=pts.numpy(), ray_normals=nrm.numpy(), sel_border_normals=bn.numpy(), sel_border_gt=bgt.numpy(),
               sel_center_normals=cn.numpy(), sel_center_gt=cgt.numpy())
    np.savez_compressed(os.path.join(HERE, "supervision.npz"), **out)
    print(f"wrote supervision.npz: {n} sphere points per sampler, {bn.shape[0]} border / {cn.shape[0]} centre ray samples; "
          "oracle == reference bit for bit")


if __name__ == "__main__":
    main()
