"""CPU oracle for the marching-cubes preprocessing that consumes the dense VF grid query.  TEST INFRASTRUCTURE ONLY
(same rules as render_oracle.py: imported by tests/ and bench.py's checker legs, never by vfnerf_b200/).

Clean-room restatement, as per-cell gathers in fp32 torch-CPU arithmetic, of what the reference computes between the
grid query and its CPU marching cubes (evaluation/methods.py:209-278 with the default flags alternative = smooth_after =
smooth_all = False):

    evaluation/utils/mc_utils.py:34-85     extract_divergence   (normalise, 2x2x2 corner-direction conv3d, d|d| sum, <= -0.5)
    evaluation/methods.py:219-223          norms = |v|, vt = normalize(v)
    evaluation/utils/mc_utils.py:107-166   unify_direction      (most opposite corner pair, nearer-of-the-two choice)
    evaluation/utils/mc_utils.py:169-223   make_comb_format     (28 corner pairs: different side flag, the two norms)
    evaluation/methods.py:186-192,260-278  block-ordered cell list, mask = any pair differs, compaction

Every quantity of a cell (i, j, k) depends only on the vectors at its 8 corners (i+a, j+b, k+c), so the chain is
restated cell by cell.  Corner order is the reference's `inc` table (methods.py:176-186 == the selection filters of
mc_utils.py:111-121).  Parity pin: tests/golden/make_golden_mc.py runs the unmodified reference functions from
/root/reference on a seeded synthetic field and asserts this file reproduces them (flags, choices, pairs exactly).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

INC = torch.tensor([[0, 0, 0], [0, 1, 0], [1, 1, 0], [1, 0, 0], [0, 0, 1], [0, 1, 1], [1, 1, 1], [1, 0, 1]])
PAIRS = [(i, j) for i in range(7) for j in range(i + 1, 8)]          # mc_utils.py:205-211
FACE_AREA = math.sqrt(3.0) / 4.0                                     # mc_utils.py:71
SHAPE_VOLUME = math.sqrt(2.0) / 3.0                                  # mc_utils.py:72
THRESHOLD = -0.5                                                     # mc_utils.py:37


def _corners(grid: torch.Tensor, order: torch.Tensor) -> torch.Tensor:
    """grid [N,N,N,C] -> [N-1,N-1,N-1,8,C]: the 8 corners of every interior cell in `order` (rows of (a,b,c))."""
    N = grid.shape[0]
    return torch.stack([grid[a:N - 1 + a, b:N - 1 + b, c:N - 1 + c] for a, b, c in order.tolist()], dim=3)


def divergence(pred: torch.Tensor, N: int) -> torch.Tensor:
    """Raw divergence of every interior cell [N-1,N-1,N-1] (mc_utils.py:40-76).  Output channel m = 4a + 2b + c of the
    reference's conv3d picks corner (a, b, c) and projects its unit vector on normalize((2a-1, 2b-1, 2c-1))."""
    u = F.normalize(pred, dim=1).reshape(N, N, N, 3)
    order = torch.tensor([[(m >> 2) & 1, (m >> 1) & 1, m & 1] for m in range(8)])
    uc = _corners(u, order)                                               # [.., 8, 3]
    f = F.normalize(order.float() * 2.0 - 1.0, dim=1)                     # mc_utils.py:46-60
    d = (uc[..., 0] * f[:, 0] + uc[..., 1] * f[:, 1]) + uc[..., 2] * f[:, 2]
    terms = d * torch.abs(d) * FACE_AREA                                  # mc_utils.py:75
    acc = terms[..., 0]
    for m in range(1, 8):
        acc = acc + terms[..., m]
    return acc / SHAPE_VOLUME


def extract_divergence(pred: torch.Tensor, N: int) -> torch.Tensor:
    """[N,N,N] float 0/1: 1 where the divergence is <= -0.5 (boundary cells are 0), mc_utils.py:78-85."""
    out = torch.zeros(N, N, N)
    out[:-1, :-1, :-1] = (divergence(pred, N) <= THRESHOLD).float()
    return out


def unify_direction(div_grid: torch.Tensor, pred: torch.Tensor, N: int) -> torch.Tensor:
    """[N,N,N,8] int64 side choice of the 8 corners (`inc` order) of every surface cell, 0 elsewhere (mc_utils.py:107-166)."""
    u = F.normalize(pred, dim=1).reshape(N, N, N, 3)
    uc = _corners(u, INC)                                                 # [.., 8, 3]
    x, y, z = uc[..., 0], uc[..., 1], uc[..., 2]
    dist = 1.0 - ((x.unsqueeze(-1) * x.unsqueeze(-2) + y.unsqueeze(-1) * y.unsqueeze(-2)) + z.unsqueeze(-1) * z.unsqueeze(-2))
    ext = torch.argmax(dist.reshape(*dist.shape[:3], 64), dim=-1)         # first index on ties
    first, second = ext // 8, ext % 8
    g = lambda i: torch.gather(uc, 3, i[..., None, None].expand(*i.shape, 1, 3))
    d1 = torch.norm(g(first) - uc, dim=-1)
    d2 = torch.norm(g(second) - uc, dim=-1)
    choice = torch.argmin(torch.stack((d1, d2), dim=-1), dim=-1)          # 0 on ties
    out = torch.zeros(N, N, N, 8, dtype=torch.int64)
    interior = out[:-1, :-1, :-1]
    surf = div_grid[:-1, :-1, :-1] == 1
    interior[surf] = choice[surf]
    return out


def comb_format(choice: torch.Tensor, pred: torch.Tensor, N: int):
    """different_side [N,N,N,28] float and the pair norms [N,N,N,28,2] (mc_utils.py:169-223); corners beyond the grid
    contribute norm 0 (the conv3d's zero padding)."""
    nrm = torch.zeros(N + 1, N + 1, N + 1)
    nrm[:N, :N, :N] = torch.norm(pred, dim=1).reshape(N, N, N)
    nc = torch.stack([nrm[a:N + a, b:N + b, c:N + c] for a, b, c in INC.tolist()], dim=3)      # [N,N,N,8]
    i0 = torch.tensor([p[0] for p in PAIRS])
    i1 = torch.tensor([p[1] for p in PAIRS])
    diff = (choice[..., i0] != choice[..., i1]).float()
    return diff, torch.stack((nc[..., i0], nc[..., i1]), dim=-1)


def block_order_cells(N: int) -> torch.Tensor:
    """[8 * (N//2)^3, 3] cell coordinates in the order the reference hands cells to its marching cubes
    (methods.py:186-192): 2x2x2 blocks in C order, the 8 cells of a block in `inc` order."""
    h = N // 2
    b = torch.stack(torch.meshgrid(torch.arange(h), torch.arange(h), torch.arange(h), indexing="ij"), dim=-1).reshape(-1, 3)
    return (b[:, None] * 2 + INC[None]).reshape(-1, 3)


def mc_preprocess(pred: torch.Tensor, N: int):
    """The whole chain -> (cells [M,3] int64, comb_values [M,28], udf [M,28,2]) exactly as the reference passes them to
    contrastive_marching_cubes (methods.py:260-278, before its flattening reshape)."""
    div = extract_divergence(pred, N)
    choice = unify_direction(div, pred, N)
    diff, nrm = comb_format(choice, pred, N)
    cells = block_order_cells(N)
    c = diff[cells[:, 0], cells[:, 1], cells[:, 2]]
    n = nrm[cells[:, 0], cells[:, 1], cells[:, 2]]
    mask = c.sum(-1) > 0
    return cells[mask], c[mask], n[mask]


def smooth_vf(vf: torch.Tensor, k: int = 3, sigma: float = 1.0) -> torch.Tensor:
    """evaluation/utils/guassian_smoothing.py:81-97 (GaussianSmoothing :24-72): replicate-padded depthwise k^3 gaussian of a
    [N,N,N,3] grid, restated as the explicit weighted sum of shifted copies.  The kernel is the reference's: the product
    over axes of 1/(sigma sqrt(2 pi)) exp(-((i - mean) / (2 sigma))^2), divided by its sum, all in fp32."""
    ax = torch.arange(k, dtype=torch.float32)
    mean = (k - 1) / 2
    g1 = 1 / (sigma * math.sqrt(2 * math.pi)) * torch.exp(-(((ax - mean) / (2 * sigma)) ** 2))
    kern = g1[:, None, None] * g1[None, :, None] * g1[None, None, :]
    kern = kern / torch.sum(kern)
    h = k // 2
    N = vf.shape[0]
    pad = F.pad(vf.permute(3, 0, 1, 2).unsqueeze(0), (h, h, h, h, h, h), mode="replicate")[0]    # [3, N+2h, ...]
    out = torch.zeros(3, N, N, N)
    for a in range(k):
        for b in range(k):
            for c in range(k):
                out = out + kern[a, b, c] * pad[:, a:a + N, b:b + N, c:c + N]
    return out.permute(1, 2, 3, 0)


def mc_preprocess_smooth_after(pred: torch.Tensor, N: int):
    """methods.py:209-278 with smooth_after=True: divergence of the raw field, sides / norms of the k = 9, sigma = 2
    smoothed one."""
    div = extract_divergence(pred, N)
    sm = smooth_vf(pred.reshape(N, N, N, 3), 9, 2.0).reshape(N ** 3, 3)
    choice = unify_direction(div, sm, N)
    diff, nrm = comb_format(choice, sm, N)
    cells = block_order_cells(N)
    c = diff[cells[:, 0], cells[:, 1], cells[:, 2]]
    n = nrm[cells[:, 0], cells[:, 1], cells[:, 2]]
    mask = c.sum(-1) > 0
    return cells[mask], c[mask], n[mask]
