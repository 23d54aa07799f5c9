"""LPIPS-VGG16 diversity term of the train step (main.py:532-537,776-791), B200-native.

`LpipsVGG16` holds the 13 VGG16 convolutions under taming's `vgg16` slice names (slice1.0 ... slice5.28) so the weights of
`LPIPS().net` load; `DiversityEngine.forward_backward(xr, repeat, bs, coef)` returns the diversity value and ACCUMULATES
-coef * d(div)/d(xr) into the image gradient (the loss is `dists - diversity_coef * div`, main.py:831).
Arithmetic: libffvc_sm100.so — 3x3 convs as tcgen05 implicit GEMMs with a fused ReLU epilogue (halo-reuse kernel on the wide
layers, im2col K=27 form for the 3-channel first layer), 2x2 max-pool, and a fused normalize_tensor + pairwise-difference
kernel per tap; backward = dgrad convs with the ReLU mask applied in the epilogue.
Mode 'between_same_prompts' (the default, main.py:695; needs repeat >= 2 to be non-zero) and mode 'all' (main.py:783-787:
the same expression with repeat = batch, bs = 1 — `forward_backward(xr, N, 1, ...)`).
"""
import ctypes as C

import torch
from torch import nn

from . import ops
from .cutouts import CLIP_MEAN, CLIP_STD
from .ops import BF16, F32, call

SLICES = [(1, [0, 2]), (2, [5, 7]), (3, [10, 12, 14]), (4, [17, 19, 21]), (5, [24, 26, 28])]
CH = {0: (3, 64), 2: (64, 64), 5: (64, 128), 7: (128, 128), 10: (128, 256), 12: (256, 256), 14: (256, 256),
      17: (256, 512), 19: (512, 512), 21: (512, 512), 24: (512, 512), 26: (512, 512), 28: (512, 512)}


class LpipsVGG16(nn.Module):
    def __init__(self):
        super().__init__()
        for s, idxs in SLICES:
            sl = nn.Sequential()
            for i in idxs:
                sl.add_module(str(i), nn.Conv2d(CH[i][0], CH[i][1], 3, padding=1))
            setattr(self, "slice%d" % s, sl)
        self._engine = None

    def engine(self):
        if self._engine is None or self._engine.ptr != self.slice1[0].weight.data_ptr():
            self._engine = DiversityEngine(self)
        return self._engine

    def forward(self, x):
        """The call the reference's loop makes, `lpips.net((xr - mean) / std)` (main.py:778): (N, 3, H, W) fp32 -> the five
        tap tensors (N, c, h, w) fp32, differentiable w.r.t. x."""
        return list(_TapsFn.apply(self, x))


class _TapsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, x):
        eng = net.engine()
        N = x.shape[0]
        taps, saved = eng.features(x.permute(0, 2, 3, 1).contiguous().float())
        ctx.eng, ctx.saved = eng, saved
        return tuple(f.float().view(N, h, w, c).permute(0, 3, 1, 2) for _, f, h, w, c in taps)

    @staticmethod
    def backward(ctx, *grads):
        eng, saved = ctx.eng, ctx.saved
        N, H, W = saved["N"], saved["H"], saved["W"]
        dtap = {}
        for (s, f, h, w, c), g in zip(saved["taps"], grads):
            if g is None:
                g = torch.zeros(N, c, h, w, device=f.device)
            dtap[s] = g.permute(0, 2, 3, 1).contiguous().to(BF16).view(N * h * w, c)
        dxn = eng.features_bwd(saved, dtap)
        return None, dxn.view(N, H, W, 3).permute(0, 3, 1, 2)


class LPIPS(nn.Module):
    """what `train()` builds for the diversity term (main.py:532-537): `LPIPS().net` is the tap network; the reference never
    uses the linear heads or the scaling layer."""

    def __init__(self):
        super().__init__()
        self.net = LpipsVGG16()

    def load_from_pretrained(self, path=None):
        """taming downloads vgg.pth into its cache; here the weights come from an explicit file when one is given"""
        if path:
            sd = torch.load(path, map_location="cpu", weights_only=False)
            self.net.load_state_dict({k[len("net."):]: v for k, v in sd.items() if k.startswith("net.slice")}, strict=False)
        return self


def normalize_tensor(x, eps=1e-10):
    """taming.modules.losses.lpips.normalize_tensor (imported at main.py:31, used at :780,:784): plain torch, as there"""
    return x / (torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True)) + eps)


class DiversityEngine:
    def __init__(self, net):
        self.dev = net.slice1[0].weight.device
        ops.require_cuda(self.dev, "the LPIPS diversity term")
        self.ptr = net.slice1[0].weight.data_ptr()
        self.pk = {}
        for k, v in net.state_dict().items():
            if k.endswith(".weight"):
                name = k[:-7]
                co, ci = v.shape[0], v.shape[1]
                if ci == 3:                                  # first layer: im2col form, K = 27 padded to 32
                    wp = torch.zeros(co, 32, device=self.dev, dtype=F32)
                    wp[:, :27] = v.permute(0, 2, 3, 1).reshape(co, 27)
                    self.pk[name + ".w"] = wp.to(BF16)
                    # dgrad 64 -> 3: tap-by-tap GEMM with N = 3, weights [3][flipped tap][64]
                    self.pk[name + ".wT"] = v.flip(2, 3).permute(1, 2, 3, 0).reshape(ci, 9 * co).contiguous().to(BF16)
                else:
                    self.pk[name + ".w"] = v.permute(0, 2, 3, 1).reshape(co, 9 * ci).contiguous().to(BF16)
                    self.pk[name + ".wT"] = v.flip(2, 3).permute(1, 2, 3, 0).reshape(ci, 9 * co).contiguous().to(BF16)
                self.pk[name + ".b"] = net.state_dict()[name + ".bias"].float().contiguous()
        self._mean = (C.c_float * 3)(*CLIP_MEAN)
        self._std = (C.c_float * 3)(*CLIP_STD)

    def _new(self, *shape, dtype=BF16):
        return torch.empty(*shape, device=self.dev, dtype=dtype)

    def _conv(self, x, name, N, H, W, cin, cout, aux=None, relu=True, dgrad=False):
        """3x3 conv (+bias, +ReLU) or its dgrad (x = dy, masked by relu'(aux) where aux = the layer's INPUT activation)."""
        wkey = name + (".wT" if dgrad else ".w")
        bias = None if dgrad else self.pk[name + ".b"]
        out = self._new(N * H * W, cout)
        act = ops.ACT_RELU if (relu and not dgrad) else ops.ACT_NONE
        mul = ops.ACT_RELU if aux is not None else ops.ACT_NONE
        if W % 128 == 0 and H % 2 == 0 and cin % 64 == 0 and cout <= 128:
            call("conv3x3_halo", x, self.pk[wkey], out, N, H, W, cin, cout, cout, bias, None, aux, mul, act, 0)
        else:
            ops.gemm(x, self.pk[wkey], out, N * H * W, cout, 9 * cin, a_mode=ops.CONV3X3, conv=(N, H, W, cin), bias=bias,
                     act=act, aux=aux, mul_mode=mul)
        return out

    def features(self, xn):
        """xn: [N, H, W, 3] fp32 NHWC, already normalised.  Returns (taps, saved): taps = [(slice, feats [N*h*w, c] bf16, h, w, c)] at
        relu1_2, relu2_2, relu3_3, relu4_3, relu5_3; every post-ReLU activation is kept (it is both the next layer's input and
        the ReLU mask of the backward)."""
        N, H, W, _ = xn.shape
        acts = {}
        col = self._new(N * H * W, 32)
        call("im2col3x3_cin3", xn, col, N, H, W)
        h = self._new(N * H * W, 64)
        ops.gemm(col, self.pk["slice1.0.w"], h, N * H * W, 64, 32, bias=self.pk["slice1.0.b"], act=ops.ACT_RELU)
        del col
        acts[0] = h
        taps, pools = [], {}
        ch, hh, ww = 64, H, W
        for s, idxs in SLICES:
            if s > 1:
                p = self._new(N * (hh // 2) * (ww // 2), ch)
                call("maxpool2x2_fwd", h, p, N, hh, ww, ch)
                pools[s] = (h, hh, ww, ch)
                h, hh, ww = p, hh // 2, ww // 2
                acts["pool%d" % s] = h
            for i in idxs:
                if i == 0:
                    continue
                cin, cout = CH[i]
                h = self._conv(h, "slice%d.%d" % (s, i), N, hh, ww, cin, cout)
                acts[i] = h
                ch = cout
            taps.append((s, h, hh, ww, ch))
        return taps, dict(N=N, H=H, W=W, acts=acts, pools=pools, taps=taps)

    def features_bwd(self, saved, dtap):
        """dtap: {slice: gradient w.r.t. that tap, [N*h*w, c] bf16} -> gradient w.r.t. xn, [N*H*W, 3] fp32 (frozen net: dgrad only)."""
        N, H, W, acts, pools, taps = saved["N"], saved["H"], saved["W"], saved["acts"], saved["pools"], saved["taps"]
        g = None
        for s, idxs in reversed(SLICES):
            _, f, th, tw, tc = taps[s - 1]
            if g is None:
                g = dtap[s]
            else:
                tmp = self._new(N * th * tw, tc)
                call("add_bf16", g, dtap[s], tmp, g.numel())
                g = tmp
            # g = gradient w.r.t. the slice's output (post-ReLU of its last conv)
            for i in reversed(idxs):
                cin, cout = CH[i]
                if i == 0:
                    break
                g = self._relu_mask(g, acts[i])          # ReLU of conv i (mask = its post-activation output)
                g = self._conv(g, "slice%d.%d" % (s, i), N, th, tw, cout, cin, dgrad=True)
            if s > 1:
                src, sh, sw, sc = pools[s]
                gp = self._new(N * sh * sw, sc)
                call("maxpool2x2_bwd", src, g, gp, N, sh, sw, sc)
                g = gp
        # first layer: ReLU mask, then 64 -> 3 dgrad (fp32 out)
        g = self._relu_mask(g, acts[0])
        dxn = self._new(N * H * W, 3, dtype=F32)
        ops.gemm(g, self.pk["slice1.0.wT"], dxn, N * H * W, 3, 9 * 64, a_mode=ops.CONV3X3, conv=(N, H, W, 64))
        return dxn

    def forward_backward(self, xr, repeat, bs, coef, dimg_accum, loss_accum):
        """xr: [repeat*bs, H, W, 3] fp32 NHWC in [0,1].  loss_accum[0] += -coef * div;  dimg_accum += -coef * d(div)/d(xr)."""
        N, H, W, _ = xr.shape
        assert N == repeat * bs
        xn = self._new(N, H, W, 3, dtype=F32)
        call("normalize3_fwd", xr, xn, N * H * W * 3, C.addressof(self._mean), C.addressof(self._std))
        taps, saved = self.features(xn)
        # ---- per-tap diversity + its gradient w.r.t. the tap activation
        dtap = {}
        for s, f, th, tw, tc in taps:
            d = self._new(N * th * tw, tc)
            call("diversity_tap", f, loss_accum, d, repeat, bs, th * tw, tc, -coef)
            dtap[s] = d
        # ---- backward through the VGG stack, then the normalisation's 1/std, accumulated into d(xr)
        dxn = self.features_bwd(saved, dtap)
        call("normalize3_bwd", dxn, dimg_accum, N * H * W * 3, C.addressof(self._std))

    def _relu_mask(self, g, post):
        """g * (post > 0)"""
        out = self._new(*g.shape)
        call("relu_mask", g, post, out, g.numel())
        return out
