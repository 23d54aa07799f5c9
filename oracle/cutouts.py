"""ORACLE (test infrastructure, not product): CPU fp32 restatement of MakeCutouts.forward (main.py:212-229)
with the default augmentation list ('Af','Pe','Ji','Er') (main.py:164-165,171-172,177-178,181-182,189-190) and
the CLIP normalisation (main.py:631-632,797).

Third-party boundary: the warps / colour jitter / erasing are `kornia==0.5.10` (requirements.txt:9; absent) — NOT pinned by
kornia; the affine and perspective inverse maps + bilinear convention, the erase rectangle and the HSV hue shift are cross-checked
against torchvision's affine / perspective / erase / adjust_hue (tests/test_oracle_golden.py).  Their *sampling* is host policy; the *arithmetic* is restated here on EXPLICIT
parameters: per-cutout inverse homographies for RandomAffine (border padding) and
RandomPerspective (zero padding), per-cutout saturation factor / hue shift (HSV round trip as in
kornia.color.rgb_to_hsv / hsv_to_rgb), one erase rectangle for the whole batch (same_on_batch=True), and the
additive noise tensor (main.py:223-225).  Pixel centres sit at integer coordinates (align_corners=True).
The pooling line `(av_pool(x) + max_pool(x)) / 2` (main.py:218) is plain torch and identical to the reference.
"""
import math

import torch
import torch.nn.functional as F

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)    # main.py:81
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)    # main.py:82
TWO_PI = 2.0 * math.pi


def warp(img, hinv, padding_mode):
    """img (N,3,H,W); hinv (N,3,3) maps output pixel (x,y,1) -> source pixel coords."""
    N, _, H, W = img.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=img.dtype, device=img.device), torch.arange(W, dtype=img.dtype, device=img.device),
                            indexing="ij")
    pts = torch.stack([xs, ys, torch.ones_like(xs)], dim=-1).reshape(1, H * W, 3)
    src = pts @ hinv.transpose(1, 2)
    sx = src[..., 0] / src[..., 2]
    sy = src[..., 1] / src[..., 2]
    grid = torch.stack([2 * sx / (W - 1) - 1, 2 * sy / (H - 1) - 1], dim=-1).reshape(N, H, W, 2)
    return F.grid_sample(img, grid, mode="bilinear", padding_mode=padding_mode, align_corners=True)


def rgb_to_hsv(img, eps=1e-6):
    r, g, b = img[:, 0], img[:, 1], img[:, 2]
    maxc, argmax = img.max(dim=1)
    minc = img.min(dim=1).values
    v = maxc
    deltac = maxc - minc
    s = deltac / (v + eps)
    deltac = torch.where(deltac == 0, torch.ones_like(deltac), deltac)
    rc, gc, bc = maxc - r, maxc - g, maxc - b
    h = torch.stack([bc - gc, 2.0 * deltac + rc - bc, 4.0 * deltac + gc - rc], dim=1)
    h = torch.gather(h, 1, argmax[:, None])[:, 0] / deltac
    h = (h / 6.0) % 1.0
    return TWO_PI * h, s, v


def hsv_to_rgb(h, s, v):
    h6 = h / TWO_PI * 6.0
    hi = torch.floor(h6) % 6
    f = (h6 % 6) - hi
    p = v * (1 - s)
    q = v * (1 - f * s)
    t = v * (1 - (1 - f) * s)
    hi = hi.long()
    r = torch.stack([v, q, p, p, t, v], dim=1)
    g = torch.stack([t, v, v, q, p, p], dim=1)
    b = torch.stack([p, p, t, v, v, q], dim=1)
    idx = hi[:, None]
    return torch.stack([torch.gather(r, 1, idx)[:, 0], torch.gather(g, 1, idx)[:, 0], torch.gather(b, 1, idx)[:, 0]], 1)


def color_jitter(img, sat, hue):
    h, s, v = rgb_to_hsv(img)
    s = (s * sat[:, None, None]).clamp(0, 1)
    h = torch.fmod(h + hue[:, None, None] + TWO_PI, TWO_PI)
    return hsv_to_rgb(h, s, v)


def make_cutouts(x, cutn, params, cut_size=224, normalize=True):
    """x (B,3,H,W) in [0,1] -> (cutn*B,3,cut,cut), cutout-major order k*B+j (main.py:219)."""
    pooled = (F.adaptive_avg_pool2d(x, cut_size) + F.adaptive_max_pool2d(x, cut_size)) / 2
    batch = pooled.repeat(cutn, 1, 1, 1)
    batch = warp(batch, params["affine_inv"], "border")
    batch = warp(batch, params["persp_inv"], "zeros")
    batch = color_jitter(batch, params["sat"], params["hue"])
    x0, y0, x1, y1 = [int(v) for v in params["erase"]]
    if x1 > x0 and y1 > y0:
        mask = torch.ones_like(batch)
        mask[:, :, y0:y1, x0:x1] = 0
        batch = batch * mask
    batch = batch + params["noise"]
    if normalize:
        mean = torch.tensor(CLIP_MEAN, device=batch.device).view(1, 3, 1, 1)
        std = torch.tensor(CLIP_STD, device=batch.device).view(1, 3, 1, 1)
        batch = (batch - mean) / std
    return batch
