"""MakeCutouts (main.py:154-229) + CLIP normalisation (main.py:797), B200-native.

Same constructor surface as the reference class (cut_size, cutn, cut_pow, pool_size, interp_size, augs, pool,
interpolate) and the same output ordering (cutout-major, k*B + j, because train() builds the targets with
`.repeat(cutn, 1)`, main.py:801-805).  Randomness is sampled on the host (`sample_params`) — the policy of kornia's
RandomAffine / RandomPerspective / ColorJitter / RandomErasing with the reference's arguments —
and handed to the kernels as explicit tensors, so the oracle and the CUDA path can be driven by identical draws.
Arithmetic: libffvc_sm100.so (ffvc_cutout_*): pooled image -> affine warp (border) -> perspective warp (zeros)
+ hue/sat jitter + erase + noise + normalise -> patch-major bf16 for the ViT.
"""
import math

import torch
from torch import nn

from .ops import BF16, F32, call

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)    # main.py:81
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)    # main.py:82


def _persp_coeffs(src, dst):
    """Batched 3x3 homographies mapping the 4 points src -> dst; src, dst: (n,4,2) float64.  General 8x8 solve (kept as the
    reference form for tests; `_quad_to_square` is the closed form sample_params uses)."""
    n = src.shape[0]
    x, y = src[..., 0], src[..., 1]
    u, v = dst[..., 0], dst[..., 1]
    z, o = torch.zeros_like(x), torch.ones_like(x)
    r1 = torch.stack([x, y, o, z, z, z, -u * x, -u * y], dim=-1)
    r2 = torch.stack([z, z, z, x, y, o, -v * x, -v * y], dim=-1)
    A = torch.stack([r1, r2], dim=2).reshape(n, 8, 8)
    b = torch.stack([u, v], dim=2).reshape(n, 8, 1)
    h = torch.linalg.solve(A, b)[..., 0]
    return torch.cat([h, torch.ones(n, 1, dtype=h.dtype)], dim=1).view(n, 3, 3)


def _quad_to_square(quad, q):
    """Batched homographies mapping the quadrilateral quad[:, 0..3] (corners in the order (0,0), (q,0), (q,q), (0,q) of the
    square they correspond to) back onto that square — the inverse map of RandomPerspective.  Closed form (Heckbert's
    square -> quad projective map, inverted by the adjugate) instead of 512 LAPACK solves per step: the batched
    torch.linalg.solve took 60 ms for 512 cutouts on 8 host cores, i.e. half a train step of host time in the e2e path."""
    x0, y0 = quad[:, 0, 0], quad[:, 0, 1]
    x1, y1 = quad[:, 1, 0], quad[:, 1, 1]
    x2, y2 = quad[:, 2, 0], quad[:, 2, 1]
    x3, y3 = quad[:, 3, 0], quad[:, 3, 1]
    dx1, dx2, sx = x1 - x2, x3 - x2, x0 - x1 + x2 - x3
    dy1, dy2, sy = y1 - y2, y3 - y2, y0 - y1 + y2 - y3
    den = dx1 * dy2 - dy1 * dx2
    g = (sx * dy2 - sy * dx2) / den
    h = (dx1 * sy - dy1 * sx) / den
    # unit square (u, v) -> quad, then u = X / q, v = Y / q
    a, b, c = (x1 - x0 + g * x1) / q, (x3 - x0 + h * x3) / q, x0
    d, e, f = (y1 - y0 + g * y1) / q, (y3 - y0 + h * y3) / q, y0
    g, h = g / q, h / q
    one = torch.ones_like(a)
    # adjugate of [[a b c], [d e f], [g h 1]] (the inverse up to scale), normalised to [2][2] = 1
    A = torch.stack([e * one - f * h, c * h - b * one, b * f - c * e,
                     f * g - d * one, a * one - c * g, c * d - a * f,
                     d * h - e * g, b * g - a * h, a * e - b * d], dim=-1)
    A = A / A[:, 8:9]
    return A.view(-1, 3, 3)


def sample_params(n, cut_size, generator=None, augs=("Af", "Pe", "Ji", "Er"), noise_fac=0.1, with_noise=True):
    """Draw the augmentation parameters for n cutouts (host side, vectorised, fp32 CPU tensors).

    Af: RandomAffine(degrees=15, translate=0.1, p=0.7, padding_mode='border')   main.py:181-182
    Pe: RandomPerspective(distortion_scale=0.7, p=0.7)                           main.py:177-178
    Ji: ColorJitter(hue=0.1, saturation=0.1, p=0.7)                              main.py:171-172
    Er: RandomErasing((.1,.4), (.3,1/.3), same_on_batch=True, p=0.7)             main.py:189-190
    Matrices are INVERSE maps (output pixel -> source pixel), pixel centres at integers.
    """
    g = generator
    P = cut_size
    f64 = torch.float64

    def U(lo, hi, *shape):
        return torch.rand(*shape, generator=g, dtype=f64) * (hi - lo) + lo

    eye = torch.eye(3, dtype=f64).expand(n, 3, 3)
    aff, per = eye.clone(), eye.clone()
    sat = torch.ones(n, dtype=f64)
    hue = torch.zeros(n, dtype=f64)
    erase = [0, 0, 0, 0]
    c = (P - 1) / 2.0
    if "Af" in augs:
        apply = torch.rand(n, generator=g) < 0.7
        ang = U(-15.0, 15.0, n) * math.pi / 180.0
        tx, ty = U(-0.1 * P, 0.1 * P, n), U(-0.1 * P, 0.1 * P, n)
        ca, sa = torch.cos(ang), torch.sin(ang)
        # forward map = rotate about the centre then translate; inverse: p -> R(-a)(p - c - t) + c
        inv = torch.zeros(n, 3, 3, dtype=f64)
        inv[:, 0, 0], inv[:, 0, 1], inv[:, 0, 2] = ca, sa, c - ca * (c + tx) - sa * (c + ty)
        inv[:, 1, 0], inv[:, 1, 1], inv[:, 1, 2] = -sa, ca, c + sa * (c + tx) - ca * (c + ty)
        inv[:, 2, 2] = 1.0
        aff = torch.where(apply[:, None, None], inv, aff)
    if "Pe" in augs:
        apply = torch.rand(n, generator=g) < 0.7
        hw = 0.7 * (P - 1) / 2.0
        r = U(0.0, 1.0, n, 8) * hw
        q = float(P - 1)
        dst = torch.stack([torch.stack([r[:, 0], r[:, 1]], -1), torch.stack([q - r[:, 2], r[:, 3]], -1),
                           torch.stack([q - r[:, 4], q - r[:, 5]], -1), torch.stack([r[:, 6], q - r[:, 7]], -1)], dim=1)
        inv = _quad_to_square(dst, q)          # maps output (dst) pixels back to the source square (== _persp_coeffs(dst, src))
        per = torch.where(apply[:, None, None], inv, per)
    if "Ji" in augs:
        apply = torch.rand(n, generator=g) < 0.7
        s = U(0.9, 1.1, n)
        h = U(-0.1, 0.1, n) * 2.0 * math.pi
        sat = torch.where(apply, s, sat)
        hue = torch.where(apply, h, hue)
    if "Er" in augs:
        r = U(0.0, 1.0, 5).tolist()
        apply = r[0] < 0.7
        area = (0.1 + 0.3 * r[1]) * P * P
        ratio = math.exp(math.log(0.3) + r[2] * (math.log(1 / 0.3) - math.log(0.3)))
        eh = min(P, int(round(math.sqrt(area * ratio))))
        ew = min(P, int(round(math.sqrt(area / ratio))))
        x0 = int(r[3] * (P - ew + 1))
        y0 = int(r[4] * (P - eh + 1))
        if apply and eh > 0 and ew > 0:
            erase = [x0, y0, x0 + ew, y0 + eh]
    out = dict(affine_inv=aff.float().contiguous(), persp_inv=per.float().contiguous(), sat=sat.float(), hue=hue.float(),
               erase=erase)
    if with_noise:
        out["facs"] = (torch.rand(n, generator=g) * noise_fac).float()
        out["noise_raw"] = torch.randn(n, 3, P, P, generator=g)
        out["noise"] = out["facs"].view(n, 1, 1, 1) * out["noise_raw"]
    return out


class CutoutEngine:
    """forward: image [B,H,W,3] fp32 NHWC in [0,1] -> patches [cutn*B][g*g][3*ps*ps] bf16 (normalised);
    backward: d(patches) -> d(image)."""

    def __init__(self, cut_size, cutn, patch, device, mean=CLIP_MEAN, std=CLIP_STD):
        self.P, self.cutn, self.patch, self.dev = cut_size, cutn, patch, device
        import ctypes as C
        self._mean = (C.c_float * 3)(*mean)
        self._std = (C.c_float * 3)(*std)
        self._C = C

    def forward(self, img, prm, want_image=False):
        B, H, W, _ = img.shape
        P, N = self.P, self.cutn * B
        dev = self.dev
        pooled = torch.empty(B, P, P, 3, device=dev, dtype=F32)
        call("cutout_pool_fwd", img, pooled, B, H, W, P)
        cut1 = torch.empty(N, P, P, 3, device=dev, dtype=F32)
        call("cutout_warp_fwd", pooled, prm["affine_inv"], cut1, N, B, P, 1)
        g = P // self.patch
        patches = torch.empty(N, g * g, 3 * self.patch * self.patch, device=dev, dtype=BF16)
        out_img = torch.empty(N, 3, P, P, device=dev, dtype=F32) if want_image else None
        erase = prm["erase"]
        call("cutout_final_fwd", cut1, prm["persp_inv"], prm["sat"], prm["hue"], prm["noise_raw"], prm["facs"],
             erase, self._C.addressof(self._mean), self._C.addressof(self._std), patches, out_img, N, P,
             self.patch)
        saved = dict(img=img, cut1=cut1, prm=prm, B=B, H=H, W=W, erase=erase)
        return patches, saved, out_img

    def backward(self, saved, dpatches):
        B, H, W = saved["B"], saved["H"], saved["W"]
        P, N = self.P, self.cutn * B
        prm = saved["prm"]
        dev = self.dev
        # the two scatter stages accumulate in 64-bit fixed point (x 2^40): reproducible d(image), see include/ffvc.h
        dcut1 = torch.empty(N, P, P, 3, device=dev, dtype=torch.int64)
        call("cutout_final_bwd", saved["cut1"], prm["persp_inv"], prm["sat"], prm["hue"], saved["erase"],
             self._C.addressof(self._mean), self._C.addressof(self._std), dpatches, dcut1, N, P, self.patch)
        dpooled = torch.empty(B, P, P, 3, device=dev, dtype=torch.int64)
        call("cutout_warp_bwd", dcut1, prm["affine_inv"], dpooled, N, B, P, 1)
        dimg = torch.empty(B, H, W, 3, device=dev, dtype=F32)
        call("cutout_pool_bwd", saved["img"], dpooled, dimg, B, H, W, P, 1)
        return dimg


def params_to_device(prm, device):
    out = {}
    for k, v in prm.items():
        if k == "noise":
            continue
        if k == "erase":
            v = torch.tensor([int(t) for t in v], dtype=torch.int32)
        out[k] = v.to(device).contiguous() if isinstance(v, torch.Tensor) else v
    return out


class MakeCutouts(nn.Module):
    """Drop-in for main.py:154-229.  Returns the cutouts as an (cutn*B, 3, cut, cut) fp32 tensor, already
    normalised when `normalize=True` (train() applies (x-mean)/std right after, main.py:797 — here it is fused)."""

    def __init__(self, cut_size, cutn, cut_pow=1., pool_size=None, interp_size=None, augs=None, pool=True,
                 interpolate=False, normalize=False, patch=32):
        super().__init__()
        if not pool or interpolate:
            raise NotImplementedError("B200 path implements the default pool=True, interpolate=False configuration")
        if pool_size not in (None, cut_size) or interp_size not in (None, cut_size):
            raise NotImplementedError("pool_size / interp_size other than cut_size")
        self.cut_size, self.cutn = cut_size, cutn
        self.augs = tuple(augs) if augs else ("Af", "Pe", "Ji", "Er")
        for a in self.augs:
            if a not in ("Af", "Pe", "Ji", "Er"):
                raise NotImplementedError("augmentation code %r (SURVEY App. A.3: out of scope)" % a)
        self.noise_fac = 0.1
        self.normalize = normalize
        self.patch = patch
        self.generator = None
        self.next_params = None     # tests can inject explicit parameters

    def forward(self, input):
        B = input.shape[0]
        prm = self.next_params or sample_params(self.cutn * B, self.cut_size, self.generator, self.augs, self.noise_fac)
        self.next_params = None
        return _CutoutFn.apply(self, input, params_to_device(prm, input.device))


class _CutoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, prm):
        mean, std = (CLIP_MEAN, CLIP_STD) if mod.normalize else ((0., 0., 0.), (1., 1., 1.))
        eng = CutoutEngine(mod.cut_size, mod.cutn, mod.patch, x.device, mean, std)
        img = x.permute(0, 2, 3, 1).contiguous().float()
        _, saved, out = eng.forward(img, prm, want_image=True)
        ctx.eng, ctx.saved = eng, saved
        return out

    @staticmethod
    def backward(ctx, g):
        eng = ctx.eng
        N, _, P, _ = g.shape
        ps = eng.patch
        gg = P // ps
        dp = g.reshape(N, 3, gg, ps, gg, ps).permute(0, 2, 4, 1, 3, 5).reshape(N, gg * gg, 3 * ps * ps).contiguous().to(BF16)
        dimg = eng.backward(ctx.saved, dp)
        return None, dimg.permute(0, 3, 1, 2), None
