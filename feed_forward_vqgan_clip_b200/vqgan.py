"""VQGAN f16 decoder + the reference's VQ glue, B200-native.

Drop-in for what `load_vqgan_model(...)` returns (main.py:84-103) as far as train() uses it:
`.quantize.embedding.weight`, `.decode(z_q)`, `.eval()`, `.requires_grad_(False)`, `.to(device)`, and a
`state_dict` with taming's key names so `vqgan_imagenet_f16_16384.ckpt` loads (`init_from_ckpt`).
`clamp_with_grad`, `vector_quantize`, `synth` mirror main.py:118-143.

All arithmetic is in libffvc_sm100.so: 3x3 / 1x1 convolutions are tcgen05 implicit GEMMs on NHWC bf16 (TMA box
per filter tap, zero fill = padding), GroupNorm+swish and nearest-2x upsample are vectorised HBM kernels, the
single-head attention is batched tcgen05 GEMMs + a row softmax, VQ is an fp32 distance search.  The decoder is
frozen: backward is dgrad only (flipped/transposed filter packs are prepared once).
Architecture restated from taming-transformers — see oracle/vqgan.py for the CPU restatement.
"""
import math

import os

import torch
from torch import nn

from . import ops
from .ops import BF16, F32, call

F16_16384 = dict(ch=128, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, attn_resolutions=(16,), resolution=256,
                 z_channels=256, out_ch=3, embed_dim=256, n_embed=16384)


def _decoder_layout(cfg):
    ch, ch_mult, nrb = cfg["ch"], cfg["ch_mult"], cfg["num_res_blocks"]
    nres = len(ch_mult)
    block_in = ch * ch_mult[-1]
    curr = cfg["resolution"] // 2 ** (nres - 1)
    levels = []
    for i_level in reversed(range(nres)):
        block_out = ch * ch_mult[i_level]
        blocks = []
        for _ in range(nrb + 1):
            blocks.append((block_in, block_out))
            block_in = block_out
        levels.append((i_level, blocks, curr in cfg["attn_resolutions"], i_level != 0))
        if i_level != 0:
            curr *= 2
    return levels


# ---- parameter containers with taming's module / key names
class _Res(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(32, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.nin_shortcut = nn.Conv2d(cin, cout, 1)


class _Attn(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.norm = nn.GroupNorm(32, c, eps=1e-6)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(c, c, 1) for _ in range(4))


class _Up(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)


class _DecoderParams(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        block_in = cfg["ch"] * cfg["ch_mult"][-1]
        self.conv_in = nn.Conv2d(cfg["z_channels"], block_in, 3, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = _Res(block_in, block_in)
        self.mid.attn_1 = _Attn(block_in)
        self.mid.block_2 = _Res(block_in, block_in)
        ups = {}
        last = block_in
        for i_level, blocks, has_attn, has_up in _decoder_layout(cfg):
            up = nn.Module()
            up.block = nn.ModuleList([_Res(a, b) for a, b in blocks])
            up.attn = nn.ModuleList([_Attn(b) for _, b in blocks] if has_attn else [])
            last = blocks[-1][1]
            if has_up:
                up.upsample = _Up(last)
            ups[i_level] = up
        self.up = nn.ModuleList([ups[i] for i in range(len(cfg["ch_mult"]))])
        self.norm_out = nn.GroupNorm(32, last, eps=1e-6)
        self.conv_out = nn.Conv2d(last, cfg["out_ch"], 3, padding=1)


class VQModel(nn.Module):
    """Decoder half of taming's VQModel (the encoder / loss are never used by the train step; main.py:102)."""

    def __init__(self, cfg=F16_16384):
        super().__init__()
        self.cfg = dict(cfg)
        self.quantize = nn.Module()
        self.quantize.embedding = nn.Embedding(cfg["n_embed"], cfg["embed_dim"])
        self.post_quant_conv = nn.Conv2d(cfg["embed_dim"], cfg["z_channels"], 1)
        self.decoder = _DecoderParams(cfg)
        self.loss = None
        self._engine = None

    def init_from_ckpt(self, path):
        sd = torch.load(path, map_location="cpu", weights_only=False)   # Lightning checkpoints carry non-tensor objects
        sd = sd.get("state_dict", sd)
        own = self.state_dict()
        self.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
        self._engine = None

    def engine(self):
        if self._engine is None or not self._engine.valid():
            self._engine = DecoderEngine(self)
        return self._engine

    def decode(self, z_q):
        """z_q (B, embed_dim, S, S) fp32 -> (B, 3, 16S, 16S) fp32 in ~[-1, 1]; differentiable w.r.t. z_q."""
        return _DecodeFn.apply(self, z_q)


class _DecodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, z_q):
        eng = model.engine()
        B, C, S, _ = z_q.shape
        zq = z_q.permute(0, 2, 3, 1).contiguous().to(BF16)
        img, tape = eng.forward(zq, post=False)
        ctx.eng, ctx.tape, ctx.shape = eng, tape, (B, C, S)
        H = img.shape[1]
        return img.view(B, H, H, 3).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        B, C, S = ctx.shape
        gd = g.permute(0, 2, 3, 1).contiguous().float()
        dzq = ctx.eng.backward(ctx.tape, gd, post=False)
        return None, dzq.view(B, S, S, C).permute(0, 3, 1, 2).float()


class DecoderEngine:
    """Explicit forward / backward over NHWC bf16 activations with a tape of backward closures."""

    def __init__(self, model):
        self.model = model
        self.cfg = model.cfg
        self.dev = model.post_quant_conv.weight.device
        ops.require_cuda(self.dev, "the VQGAN decoder")
        self._ptr = model.post_quant_conv.weight.data_ptr()
        self._ver = tuple(p._version for p in model.parameters())      # an in-place load_state_dict invalidates the packed weights
        self.layout = _decoder_layout(self.cfg)
        self.pk = {}
        sd = {k: v.detach() for k, v in model.state_dict().items()}
        for k, v in sd.items():
            if k.endswith(".weight") and v.dim() == 4:
                name = k[:-7]
                co, ci, kh, kw = v.shape
                if kh == 3 and ci % 64 == 0:
                    self.pk[name + ".w"] = v.permute(0, 2, 3, 1).reshape(co, 9 * ci).contiguous().to(BF16)
                    if co % 64 == 0:   # dgrad pack: [ci][flipped tap][co]
                        self.pk[name + ".wT"] = v.flip(2, 3).permute(1, 2, 3, 0).reshape(ci, 9 * co).contiguous().to(BF16)
                    else:              # conv_out: forward as a 1x1 GEMM into 9 * co tap columns (+ ffvc_conv_taps_gather) ...
                        if 9 * co <= 32:
                            wv = torch.zeros(32, ci, device=v.device, dtype=F32)      # row tap * co_n + co = w[co][:, kh, kw]
                            wv[:9 * co] = v.permute(2, 3, 0, 1).reshape(9 * co, ci)
                            self.pk[name + ".wv"] = wv.to(BF16)
                        # ... dgrad runs on the tiny-Cin SIMT kernel, fp32 [ci][flipped tap][co]
                        wt = v.flip(2, 3).permute(1, 2, 3, 0).reshape(ci, 9 * co).contiguous().float()
                        self.pk[name + ".wT32"] = wt
                        wp = torch.zeros(ci, 32, device=wt.device, dtype=F32)      # K padded 27 -> 32 for the GEMM form
                        wp[:, :9 * co] = wt
                        self.pk[name + ".wTp"] = wp.to(BF16)
                elif kh == 1:
                    self.pk[name + ".w"] = v.reshape(co, ci).contiguous().to(BF16)
                self.pk[name + ".b"] = sd[name + ".bias"].float().contiguous()
            elif k.endswith(".weight") and v.dim() == 1:
                name = k[:-7]
                self.pk[name + ".g"] = v.float().contiguous()
                self.pk[name + ".be"] = sd[name + ".bias"].float().contiguous()
        cb = sd["quantize.embedding.weight"].float().contiguous()
        self.codebook = cb
        self.codeT = cb.t().contiguous()
        self.cnorm = torch.empty(cb.shape[0], device=self.dev, dtype=F32)
        call("rownorm2", cb, self.cnorm, cb.shape[0], cb.shape[1])
        self._ws = None
        self._epi_stats = {}            # data_ptr of a conv output -> (mean, rstd, numel) from that conv's epilogue

    def valid(self):
        return (self.model.post_quant_conv.weight.data_ptr() == self._ptr
                and tuple(p._version for p in self.model.parameters()) == self._ver)

    def _new(self, *shape, dtype=BF16):
        return torch.empty(*shape, device=self.dev, dtype=dtype)

    def _gn_ws(self, N, HW, tag=0):
        """statistics workspace: folded sums + one partial per CTA / conv tile (include/ffvc.h: ffvc_groupnorm_ws_doubles).
        tag 1: a second buffer for the backward sums that a dgrad conv leaves for the following groupnorm_bwd_apply."""
        from . import _lib
        n = int(_lib.load().ffvc_groupnorm_ws_doubles(N, HW, 32))
        if self._ws is None:
            self._ws = {}
        if tag not in self._ws or self._ws[tag].numel() < n:
            self._ws[tag] = torch.empty(max(n, 4096), device=self.dev, dtype=torch.float64)
        return self._ws[tag]

    # ---------------------------------------------------------------- primitive ops (forward + backward closure)
    CONV_OUT_TAPS = os.environ.get("FFVC_CONV_OUT_TAPS", "1") == "1"   # conv_out as 1x1 GEMM + tap gather (0: implicit-GEMM conv)
    USE_HALO = True    # shared-memory halo reuse for the wide (W % 128 == 0), <= 128-output-channel 3x3 convs

    def _halo_ok(self, H, W, cin, cout, out_f32=False):
        return (self.USE_HALO and W % 128 == 0 and H % 2 == 0 and cin % 64 == 0 and cout <= 128
                and (cout % 8 == 0 or out_f32))

    # GroupNorm statistics from the epilogue of the conv that writes the tensor (ffvc_conv3x3_halo_gn): every forward
    # 128-channel halo conv of the decoder feeds a Normalize (norm2 of its block, norm1 of the next block, norm_out), so its
    # (mean, rstd) are produced on the fly and `gn` skips the statistics pass.  FFVC_GN_EPI_STATS=0 selects the separate pass.
    GN_EPI_STATS = os.environ.get("FFVC_GN_EPI_STATS", "1") == "1"    # measured +0.9 % prompts/s (profiles/r01_ab_kernels.md)

    # GroupNorm apply + swish fused into the consuming conv (ffvc_conv3x3_halo_xf): the wide layers never materialise the normalised
    # tensor (-17 GB of HBM traffic per step at config #2).  Built, parity-tested — and measured SLOWER on the B200: the halo conv is
    # bound by the shared-memory port (128 x 128 x 16 MMAs), and the in-place transform of the halo tile takes its share of that
    # port: 1465 us per 256 x 256 layer against 982 us conv + 395 us apply pass; step 114.1 ms against 111.5 ms
    # (profiles/r02_gn_apply_fusion.md).  Off by default; FFVC_GN_FUSE_APPLY=1 selects it.
    GN_FUSE_APPLY = os.environ.get("FFVC_GN_FUSE_APPLY", "0") == "1"

    def gn_stats(self, x, N, HW, C):
        """(mean, rstd) of a Normalize's input: from the epilogue of the conv that wrote x when it left them, else one statistics pass"""
        st = self._epi_stats.pop(x.data_ptr(), None)
        if st is not None and st[2] == N * HW * C:
            return st[0], st[1]
        mean, rstd = self._new(N * 32, dtype=F32), self._new(N * 32, dtype=F32)
        call("groupnorm_stats", x, self._gn_ws(N, HW), mean, rstd, N, HW, C, 32, 1e-6)
        return mean, rstd

    def conv3_of_gn(self, x, st, nname, name, N, H, W, cin, cout, res=None, gn_next=True):
        """conv(swish(GroupNorm(x))) with the apply fused into the conv's operand path; st = (mean, rstd) of x"""
        out = self._new(N * H * W, cout)
        ws = None
        if gn_next and self.GN_EPI_STATS and cout == 128:
            ws = self._gn_ws(N, H * W)
        call("conv3x3_halo_xf", x, self.pk[name + ".w"], out, N, H, W, cin, cout, cout, self.pk[name + ".b"], res, st[0], st[1],
             self.pk[nname + ".g"], self.pk[nname + ".be"], 32, ws)
        if ws is not None:
            mean, rstd = self._new(N * 32, dtype=F32), self._new(N * 32, dtype=F32)
            call("groupnorm_finalize", ws, mean, rstd, N, H * W, cout, 32, 1e-6)
            self._epi_stats[out.data_ptr()] = (mean, rstd, N * H * W * cout)
        return out

    def conv3(self, x, name, N, H, W, cin, cout, res=None, out_f32=False, gn_next=True):
        """gn_next: the output goes straight into a Normalize (true for every decoder conv but the last one of a level
        that is followed by Upsample, and conv_out)"""
        out = self._new(N * H * W, cout, dtype=F32 if out_f32 else BF16)
        if gn_next and self.GN_EPI_STATS and cout == 128 and not out_f32 and self._halo_ok(H, W, cin, cout):
            ws = self._gn_ws(N, H * W)
            mean, rstd = self._new(N * 32, dtype=F32), self._new(N * 32, dtype=F32)
            call("conv3x3_halo_gn", x, self.pk[name + ".w"], out, N, H, W, cin, cout, cout, self.pk[name + ".b"], res, ws)
            call("groupnorm_finalize", ws, mean, rstd, N, H * W, cout, 32, 1e-6)
            self._epi_stats[out.data_ptr()] = (mean, rstd, N * H * W * cout)
            return out
        if self._halo_ok(H, W, cin, cout, out_f32):      # incl. the 128 -> 3 conv_out (fp32 image, scalar stores)
            call("conv3x3_halo", x, self.pk[name + ".w"], out, N, H, W, cin, cout, cout, self.pk[name + ".b"], res, None, 0, 0,
                 int(out_f32))
            return out
        ops.gemm(x, self.pk[name + ".w"], out, N * H * W, cout, 9 * cin, a_mode=ops.CONV3X3, conv=(N, H, W, cin),
                 bias=self.pk[name + ".b"], res=res)
        return out

    # Backward statistics of a Normalize + swish from the epilogue of the dgrad conv that produces its output gradient
    # (ffvc_conv3x3_halo_gnbwd): `gnb` = (x, (mean, rstd), norm name) of the Normalize in front of this conv.  Returns
    # (dx, sums) — sums is None when the fused form does not apply and gn_bwd runs its own statistics pass.
    GN_EPI_BWD = os.environ.get("FFVC_GN_EPI_BWD", "1") == "1"      # measured +1.1 % prompts/s (profiles/r01_ab_kernels.md)

    def conv3_dgrad(self, dy, name, N, H, W, cin, cout, res=None, gnb=None):
        dx = self._new(N * H * W, cin)
        if gnb is not None:
            if self.GN_EPI_BWD and cin == 128 and res is None and self._halo_ok(H, W, cout, cin):
                x, st, nname = gnb
                sums = self._gn_ws(N, H * W, tag=1)
                call("conv3x3_halo_gnbwd", dy, self.pk[name + ".wT"], dx, N, H, W, cout, cin, cin, None, x, st[0], st[1],
                     self.pk[nname + ".g"], self.pk[nname + ".be"], sums)
                return dx, sums
            return self.conv3_dgrad(dy, name, N, H, W, cin, cout, res=res), None
        if self._halo_ok(H, W, cout, cin):
            call("conv3x3_halo", dy, self.pk[name + ".wT"], dx, N, H, W, cout, cin, cin, None, res, None, 0, 0, 0)
            return dx
        ops.gemm(dy, self.pk[name + ".wT"], dx, N * H * W, cin, 9 * cout, a_mode=ops.CONV3X3, conv=(N, H, W, cout), res=res)
        return dx

    def conv1(self, x, name, M, cin, cout, res=None):
        out = self._new(M, cout)
        ops.gemm(x, self.pk[name + ".w"], out, M, cout, cin, bias=self.pk[name + ".b"], res=res)
        return out

    def conv1_dgrad(self, dy, name, M, cin, cout, res=None):
        dx = self._new(M, cin)
        ops.linear_dgrad(dy, self.pk[name + ".w"], dx, M, cout, cin, res=res)
        return dx

    # GroupNorm: two-pass kernels (statistics, apply).  The single-kernel L2-resident forms (ffvc_groupnorm_fused_*) are
    # selectable per tensor size through GN_FUSED_MIN (elements per sample); measured on B200 they lose to the two-pass
    # kernels (one 512-thread CTA per SM keeps too few register-staged loads in flight), so they are off by default.
    GN_FUSED_MIN = int(os.environ["FFVC_GN_FUSED_MIN"]) if os.environ.get("FFVC_GN_FUSED_MIN") else None

    def _gn_fused(self, HW, C):
        return self.GN_FUSED_MIN is not None and HW * C >= self.GN_FUSED_MIN and 512 % (C // 8) == 0

    def gn(self, x, name, N, HW, C, swish):
        mean, rstd = self._new(N * 32, dtype=F32), self._new(N * 32, dtype=F32)
        y = self._new(N * HW, C)
        if self._gn_fused(HW, C):
            call("groupnorm_fused_fwd", x, self.pk[name + ".g"], self.pk[name + ".be"], y, mean, rstd, self._gn_ws(N, HW), N, HW, C,
                 32, int(swish), 1e-6)
            return y, (mean, rstd)
        st = self._epi_stats.pop(x.data_ptr(), None)
        if st is not None and st[2] == N * HW * C:       # statistics came with the conv that wrote x
            mean, rstd = st[0], st[1]
        else:
            call("groupnorm_stats", x, self._gn_ws(N, HW), mean, rstd, N, HW, C, 32, 1e-6)
        call("groupnorm_apply", x, mean, rstd, self.pk[name + ".g"], self.pk[name + ".be"], y, N, HW, C, 32, int(swish))
        return y, (mean, rstd)

    def gn_bwd(self, dy, x, stats, name, N, HW, C, swish, add=None, sums=None):
        dx = self._new(N * HW, C)
        if sums is not None:          # (sum g, sum g * xhat) came with the dgrad conv that wrote dy
            call("groupnorm_bwd_apply", dy, x, stats[0], stats[1], self.pk[name + ".g"], self.pk[name + ".be"], sums, add, dx,
                 N, HW, C, 32, int(swish))
            return dx
        call("groupnorm_fused_bwd" if self._gn_fused(HW, C) else "groupnorm_bwd", dy, x, stats[0], stats[1], self.pk[name + ".g"],
             self.pk[name + ".be"], self._gn_ws(N, HW), add, dx, N, HW, C, 32, int(swish))
        return dx

    def resblock(self, x, name, N, H, W, cin, cout, tape, gn_next=True):
        HW = H * W
        if self.GN_FUSE_APPLY and self._halo_ok(H, W, cin, cout) and not self._gn_fused(HW, cin):
            st1 = self.gn_stats(x, N, HW, cin)
            h1 = self.conv3_of_gn(x, st1, name + ".norm1", name + ".conv1", N, H, W, cin, cout)
        else:
            a1, st1 = self.gn(x, name + ".norm1", N, HW, cin, True)
            h1 = self.conv3(a1, name + ".conv1", N, H, W, cin, cout)
            del a1
        short = x if cin == cout else self.conv1(x, name + ".nin_shortcut", N * HW, cin, cout)
        if self.GN_FUSE_APPLY and self._halo_ok(H, W, cout, cout) and not self._gn_fused(HW, cout):
            st2 = self.gn_stats(h1, N, HW, cout)
            out = self.conv3_of_gn(h1, st2, name + ".norm2", name + ".conv2", N, H, W, cout, cout, res=short, gn_next=gn_next)
        else:
            a2, st2 = self.gn(h1, name + ".norm2", N, HW, cout, True)
            out = self.conv3(a2, name + ".conv2", N, H, W, cout, cout, res=short, gn_next=gn_next)
            del a2
        del short

        def bwd(d):
            d2, s2 = self.conv3_dgrad(d, name + ".conv2", N, H, W, cout, cout, gnb=(h1, st2, name + ".norm2"))
            dh1 = self.gn_bwd(d2, h1, st2, name + ".norm2", N, HW, cout, True, sums=s2)
            d1, s1 = self.conv3_dgrad(dh1, name + ".conv1", N, H, W, cin, cout, gnb=(x, st1, name + ".norm1"))
            ds = d if cin == cout else self.conv1_dgrad(d, name + ".nin_shortcut", N * HW, cin, cout)
            return self.gn_bwd(d1, x, st1, name + ".norm1", N, HW, cin, True, add=ds, sums=s1)

        tape.append(bwd)
        return out

    def attn(self, x, name, N, H, W, C, tape):
        HW = H * W
        M = N * HW
        scale = float(C) ** -0.5
        hn, st = self.gn(x, name + ".norm", N, HW, C, False)
        q = self.conv1(hn, name + ".q", M, C, C)
        k = self.conv1(hn, name + ".k", M, C, C)
        v = self.conv1(hn, name + ".v", M, C, C)
        del hn
        S = self._new(N, HW, HW, dtype=F32)
        ops.gemm(q, k, S, HW, HW, C, a_role=ops.ROLE_OUT, a_bs=HW * C, b_role=ops.ROLE_OUT, b_bs=HW * C, batch=N,
                 out_bs=HW * HW, alpha=scale)
        P = self._new(N, HW, HW)
        call("softmax_fwd", S, P, N * HW, HW, HW)
        del S
        O = self._new(M, C)
        ops.gemm(P, v, O, HW, C, HW, a_role=ops.ROLE_OUT, a_bs=HW * HW, b_mode=ops.MNMAJOR, b_ld=C, b_role=ops.ROLE_OUT,
                 b_bs=HW * C, batch=N, out_bs=HW * C)
        out = self.conv1(O, name + ".proj_out", M, C, C, res=x)
        del O

        def bwd(d):
            dO = self.conv1_dgrad(d, name + ".proj_out", M, C, C)
            dP = self._new(N, HW, HW, dtype=F32)
            ops.gemm(dO, v, dP, HW, HW, C, a_role=ops.ROLE_OUT, a_bs=HW * C, b_role=ops.ROLE_OUT, b_bs=HW * C, batch=N,
                     out_bs=HW * HW)
            dV = self._new(M, C)   # dV[j,c] = sum_i P[i,j] dO[i,c]
            ops.gemm(P, dO, dV, HW, C, HW, a_mode=ops.MNMAJOR, a_ld=HW, a_role=ops.ROLE_OUT, a_bs=HW * HW,
                     b_mode=ops.MNMAJOR, b_ld=C, b_role=ops.ROLE_OUT, b_bs=HW * C, batch=N, out_bs=HW * C)
            dS = self._new(N, HW, HW)
            call("softmax_bwd", P, dP, dS, N * HW, HW, HW, scale)
            del dP
            dQ = self._new(M, C)   # dQ[i,c] = sum_j dS[i,j] k[j,c]
            ops.gemm(dS, k, dQ, HW, C, HW, a_role=ops.ROLE_OUT, a_bs=HW * HW, b_mode=ops.MNMAJOR, b_ld=C,
                     b_role=ops.ROLE_OUT, b_bs=HW * C, batch=N, out_bs=HW * C)
            dK = self._new(M, C)   # dK[j,c] = sum_i dS[i,j] q[i,c]
            ops.gemm(dS, q, dK, HW, C, HW, a_mode=ops.MNMAJOR, a_ld=HW, a_role=ops.ROLE_OUT, a_bs=HW * HW,
                     b_mode=ops.MNMAJOR, b_ld=C, b_role=ops.ROLE_OUT, b_bs=HW * C, batch=N, out_bs=HW * C)
            dh = self.conv1_dgrad(dQ, name + ".q", M, C, C)
            dh = self.conv1_dgrad(dK, name + ".k", M, C, C, res=dh)
            dh = self.conv1_dgrad(dV, name + ".v", M, C, C, res=dh)
            return self.gn_bwd(dh, x, st, name + ".norm", N, HW, C, False, add=d)

        tape.append(bwd)
        return out

    # ---------------------------------------------------------------- whole decoder
    def forward(self, zq, post=True):
        """zq: [B, S, S, embed_dim] bf16 NHWC.  Returns (image [B, 16S, 16S, 3] fp32 NHWC, tape).
        post=True applies (x+1)/2 and clamp_with_grad(0,1) (main.py:142)."""
        cfg = self.cfg
        N, S = zq.shape[0], zq.shape[1]
        tape = []
        self._epi_stats.clear()
        H = W = S
        E, Z = cfg["embed_dim"], cfg["z_channels"]
        block_in = cfg["ch"] * cfg["ch_mult"][-1]
        h = self.conv1(zq.reshape(N * H * W, E), "post_quant_conv", N * H * W, E, Z)
        tape.append(lambda d, M=N * H * W: self.conv1_dgrad(d, "post_quant_conv", M, E, Z))
        h = self.conv3(h, "decoder.conv_in", N, H, W, Z, block_in)
        tape.append(lambda d, H=H, W=W: self.conv3_dgrad(d, "decoder.conv_in", N, H, W, Z, block_in))
        h = self.resblock(h, "decoder.mid.block_1", N, H, W, block_in, block_in, tape)
        h = self.attn(h, "decoder.mid.attn_1", N, H, W, block_in, tape)
        h = self.resblock(h, "decoder.mid.block_2", N, H, W, block_in, block_in, tape)
        c = block_in
        for i_level, blocks, has_attn, has_up in self.layout:
            for j, (cin, cout) in enumerate(blocks):
                last = j == len(blocks) - 1
                h = self.resblock(h, "decoder.up.%d.block.%d" % (i_level, j), N, H, W, cin, cout, tape,
                                  gn_next=has_attn or not (last and has_up))     # the level's last block feeds Upsample, not a Normalize
                if has_attn:
                    h = self.attn(h, "decoder.up.%d.attn.%d" % (i_level, j), N, H, W, cout, tape)
                c = cout
            if has_up:
                up = self._new(N * 4 * H * W, c)
                call("upsample2x_fwd", h, up, N, H, W, c)
                name = "decoder.up.%d.upsample.conv" % i_level
                H2, W2 = 2 * H, 2 * W
                h = self.conv3(up, name, N, H2, W2, c, c)
                del up

                def up_bwd(d, name=name, H=H, W=W, c=c):
                    du = self.conv3_dgrad(d, name, N, 2 * H, 2 * W, c, c)
                    dx = self._new(N * H * W, c)
                    call("upsample2x_bwd", du, dx, N, H, W, c)
                    return dx

                tape.append(up_bwd)
                H, W = H2, W2
        HW = H * W
        x_last = h
        a, st = self.gn(h, "decoder.norm_out", N, HW, c, True)
        co = cfg["out_ch"]
        if self.CONV_OUT_TAPS and ("decoder.conv_out.wv" in self.pk):
            # every pixel through the tensor core once (N = 9 * co tap columns), then the nine shifted taps + bias (+ post) in one pass
            taps = self._new(N * HW, 32, dtype=F32)
            ops.gemm(a, self.pk["decoder.conv_out.wv"], taps, N * HW, 32, c)
            del a
            dimg = self._new(N * HW, co, dtype=F32)
            img = self._new(N * HW, co, dtype=F32) if post else dimg
            call("conv_taps_gather", taps, self.pk["decoder.conv_out.b"], dimg, img if post else None, N, H, W, co)
            del taps
        else:
            dimg = self.conv3(a, "decoder.conv_out", N, H, W, c, co, out_f32=True)   # [N*HW, 3] fp32
            del a
            if post:
                img = self._new(N * HW, 3, dtype=F32)
                call("image_post_fwd", dimg, img, N * HW * 3)
            else:
                img = dimg

        def out_bwd(g, H=H, W=W, c=c):
            # g: [N*HW, 3] fp32 gradient w.r.t. the returned image
            if post:
                gd = self._new(N * HW, 3, dtype=F32)
                call("image_post_bwd", g, dimg, gd, N * HW * 3)
            else:
                gd = g
            col = self._new(N * HW, 32)
            call("im2col3x3_cin3", gd, col, N, H, W)
            da = self._new(N * HW, c)
            ops.gemm(col, self.pk["decoder.conv_out.wTp"], da, N * HW, c, 32)
            return self.gn_bwd(da, x_last, st, "decoder.norm_out", N, HW, c, True)

        tape.append(out_bwd)
        return img.view(N, H, W, 3), tape

    def backward(self, tape, g, post=True):
        """g: gradient w.r.t. the image [B, H, W, 3] fp32 -> gradient w.r.t. zq [B*S*S, embed_dim] bf16."""
        d = g.reshape(-1, 3)
        for fn in reversed(tape):
            d = fn(d)
        return d

    # ---------------------------------------------------------------- VQ (clamp + nearest code), main.py:763,134-138
    VQ_TENSOR_CORE = True    # False: the fp32 SIMT search (ffvc_vq_nearest)
    csplit = None

    def quantize(self, z_tok, lo, hi):
        """z_tok: [P, C] fp32 token-major latent.  Returns (zq bf16 [P,C], idx int32 [P], z_clamped fp32)."""
        P, C = z_tok.shape
        idx = self._new(P, dtype=torch.int32)
        zq = self._new(P, C)
        zc = self._new(P, C, dtype=F32)
        if self.VQ_TENSOR_CORE and C % 8 == 0:
            # distance search as one tcgen05 GEMM over a bf16 hi/lo split (K = 3C) with an arg-min epilogue
            if self.csplit is None:
                ncodes = self.codebook.shape[0]
                self.csplit = self._new(ncodes, 3 * C)
                self.cnorm_tc = self._new(ncodes, dtype=F32)
                call("vq_prepare_codebook", self.codebook, self.csplit, self.cnorm_tc, ncodes, C)
            zsplit = self._new(P, 3 * C)
            keys = self._new(P, dtype=torch.int64)
            call("vq_nearest_tc", z_tok, self.codebook, self.csplit, self.cnorm_tc, zsplit, keys, idx, zq, None, zc, P, C,
                 self.codebook.shape[0], float(lo), float(hi))
            return zq, idx, zc
        call("vq_nearest", z_tok, self.codebook, self.codeT, self.cnorm, idx, zq, None, zc, P, C, self.codebook.shape[0],
             float(lo), float(hi))
        return zq, idx, zc


# -------------------------------------------------------------------------------- reference glue (main.py:105-143)
class _ClampWithGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, lo, hi):
        ctx.lo, ctx.hi = float(lo), float(hi)
        ctx.save_for_backward(x)
        return x.clamp(ctx.lo, ctx.hi)

    @staticmethod
    def backward(ctx, g):
        x, = ctx.saved_tensors
        gx = torch.empty(x.shape, device=x.device, dtype=F32)        # contiguous: x may carry permuted strides
        call("clamp_bwd", g.contiguous().float(), x.contiguous().float(), gx, x.numel(), ctx.lo, ctx.hi)
        return gx, None, None


clamp_with_grad = _ClampWithGrad.apply


class _VQFn(torch.autograd.Function):
    """vector_quantize with the straight-through gradient of ReplaceGrad (main.py:105-116,134-138)."""

    @staticmethod
    def forward(ctx, x, codebook, codeT, cnorm):
        shp = x.shape
        zf = x.reshape(-1, shp[-1]).contiguous().float()
        P, C = zf.shape
        idx = torch.empty(P, device=x.device, dtype=torch.int32)
        zq = torch.empty(P, C, device=x.device, dtype=F32)
        call("vq_nearest", zf, codebook, codeT, cnorm, idx, None, zq, None, P, C, codebook.shape[0], -3.0e38, 3.0e38)
        ctx.mark_non_differentiable(idx)
        return zq.view(shp), idx.view(shp[:-1])

    @staticmethod
    def backward(ctx, g, _gi):
        return g, None, None, None


def vector_quantize(x, codebook):
    """main.py:134-138 `vector_quantize(x, codebook)`: x (..., C), codebook (ncodes, C) -> the nearest codes with the
    straight-through gradient.  `codebook` may be the tensor the reference passes (`model.quantize.embedding.weight`) or the
    VQModel itself, whose engine already holds the transposed copy and the code norms the search kernel wants."""
    if hasattr(codebook, "engine"):
        eng = codebook.engine()
        cb, cbT, cn = eng.codebook, eng.codeT, eng.cnorm
    else:
        cb = codebook.detach().float().contiguous()
        cbT = cb.t().contiguous()
        cn = torch.empty(cb.shape[0], device=cb.device, dtype=F32)
        call("rownorm2", cb, cn, cb.shape[0], cb.shape[1])
    return _VQFn.apply(x, cb, cbT, cn)[0]


def synth(model, z):
    z_q = vector_quantize(z.movedim(1, 3), model).movedim(3, 1)
    return clamp_with_grad(model.decode(z_q).add(1).div(2), 0, 1)
