"""CLIP ViT image encoder (frozen perceptor), B200-native.

Drop-in for what `load_clip_model(...)` returns (main.py:1308-1333) as far as train() uses it on the benchmark
path: `.encode_image(x: (N,3,R,R)) -> (N, output_dim)` differentiable w.r.t. x (main.py:799), `.logit_scale`,
`.visual` with the OpenAI-CLIP `state_dict` key names so real ViT-B/32 weights load.
`encode_text` belongs to SURVEY §8(f) "next" (the benchmark feeds pre-computed embeddings, main.py:733).

Arithmetic (libffvc_sm100.so): patch embedding / QKV / out-proj / MLP are tcgen05 GEMMs with fused bias /
QuickGELU (or exact GELU for OpenCLIP ViT-B-32) / residual epilogues; LayerNorm and the 50-token attention are
warp-level kernels.  Frozen weights => backward is dgrad only.  Follows the in-tree twin cloob.py:170-255.
"""
import torch
from torch import nn

from . import ops
from .ops import BF16, F32, call

VIT_B32 = dict(input_resolution=224, patch_size=32, width=768, layers=12, heads=12, output_dim=512)


class _Block(nn.Module):                      # parameter container, keys as cloob.py:184-196
    def __init__(self, w):
        super().__init__()
        self.attn = nn.MultiheadAttention(w, max(w // 64, 1))
        self.ln_1 = nn.LayerNorm(w)
        self.mlp = nn.Sequential()
        self.mlp.add_module("c_fc", nn.Linear(w, 4 * w))
        self.mlp.add_module("gelu", nn.Identity())
        self.mlp.add_module("c_proj", nn.Linear(4 * w, w))
        self.ln_2 = nn.LayerNorm(w)


class VisualTransformer(nn.Module):
    def __init__(self, input_resolution=224, patch_size=32, width=768, layers=12, heads=12, output_dim=512,
                 act="quick_gelu"):
        super().__init__()
        if width // heads != 64:
            raise NotImplementedError("head_dim must be 64")
        self.cfg = dict(input_resolution=input_resolution, patch_size=patch_size, width=width, layers=layers, heads=heads,
                        output_dim=output_dim)
        self.act = act
        self.conv1 = nn.Conv2d(3, width, patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = nn.Module()
        self.transformer.resblocks = nn.Sequential(*[_Block(width) for _ in range(layers)])
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        self._engine = None

    def engine(self):
        if self._engine is None or not self._engine.valid():
            self._engine = ClipEngine(self)
        return self._engine

    def forward(self, x):
        return _EncodeFn.apply(self, x)


class CLIP(nn.Module):
    """The slice of clip.model.CLIP the train step touches: encode_image (differentiable), encode_text (forward only),
    logit_scale.  The text tower's parameters sit at the top level like OpenAI CLIP's state_dict (token_embedding.weight,
    positional_embedding, transformer.*, ln_final.*, text_projection)."""

    def __init__(self, cfg=VIT_B32, act="quick_gelu", text_cfg=None):
        super().__init__()
        self.visual = VisualTransformer(act=act, **cfg)
        self.logit_scale = nn.Parameter(torch.ones([]) * 4.6052)
        self.text = None
        if text_cfg is not None:
            from .clip_text import TextTransformer
            self.text = TextTransformer(act=act, **text_cfg)

    def encode_image(self, image):
        return self.visual(image)

    def encode_text(self, text):
        if self.text is None:
            raise NotImplementedError("this CLIP was built without a text tower (pass text_cfg=clip_text.TEXT_B32)")
        return self.text(text)


class _EncodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vis, x):
        eng = vis.engine()
        N, _, R, _ = x.shape
        ps = eng.patch
        g = R // ps
        # NCHW fp32 -> patch-major bf16 [N][g*g][3*ps*ps] (plumbing for the generic entry; the fused train step
        # gets this layout directly from the cutout kernel)
        patches = x.reshape(N, 3, g, ps, g, ps).permute(0, 2, 4, 1, 3, 5).reshape(N, g * g, 3 * ps * ps).contiguous().to(BF16)
        emb, saved = eng.forward(patches)
        ctx.eng, ctx.saved, ctx.geom = eng, saved, (N, R, g, ps)
        return emb

    @staticmethod
    def backward(ctx, g_emb):
        N, R, g, ps = ctx.geom
        dpatch = ctx.eng.backward(ctx.saved, g_emb.contiguous().float())
        dx = dpatch.float().view(N, g, g, 3, ps, ps).permute(0, 3, 1, 4, 2, 5).reshape(N, 3, R, R)
        return None, dx


class ClipEngine:
    def __init__(self, vis):
        self.vis = vis
        cfg = vis.cfg
        self.dev = vis.proj.device
        ops.require_cuda(self.dev, "the CLIP image encoder")
        self._ptr = vis.proj.data_ptr()
        self._ver = tuple(p._version for p in vis.parameters())        # an in-place load_state_dict invalidates the packed weights
        self.W, self.L, self.Hh, self.E = cfg["width"], cfg["layers"], cfg["heads"], cfg["output_dim"]
        self.patch = cfg["patch_size"]
        self.G = cfg["input_resolution"] // self.patch
        self.T = self.G * self.G + 1
        self.act = ops.ACT_QUICKGELU if vis.act == "quick_gelu" else ops.ACT_GELU
        sd = {k: v.detach() for k, v in vis.state_dict().items()}
        self.sd32 = {k: v.float().contiguous() for k, v in sd.items() if v.dim() == 1}
        self.w = {}
        self.w["conv1"] = sd["conv1.weight"].reshape(self.W, -1).contiguous().to(BF16)
        self.w["proj"] = sd["proj"].contiguous().to(BF16)                    # [W][E]
        self.pos = sd["positional_embedding"].float().contiguous()
        self.cls = sd["class_embedding"].float().contiguous()
        for l in range(self.L):
            p = "transformer.resblocks.%d." % l
            self.w[p + "in"] = sd[p + "attn.in_proj_weight"].contiguous().to(BF16)
            self.w[p + "out"] = sd[p + "attn.out_proj.weight"].contiguous().to(BF16)
            self.w[p + "fc"] = sd[p + "mlp.c_fc.weight"].contiguous().to(BF16)
            self.w[p + "pj"] = sd[p + "mlp.c_proj.weight"].contiguous().to(BF16)

    def valid(self):
        return self.vis.proj.data_ptr() == self._ptr and tuple(p._version for p in self.vis.parameters()) == self._ver

    def _new(self, *shape, dtype=BF16):
        return torch.empty(*shape, device=self.dev, dtype=dtype)

    def _ln(self, x, name, rows):
        y, mu, rs = self._new(rows, self.W), self._new(rows, dtype=F32), self._new(rows, dtype=F32)
        call("layernorm_fwd", x, self.sd32[name + ".weight"], self.sd32[name + ".bias"], y, mu, rs, rows, self.W, 1e-5)
        return y, (mu, rs)

    def _ln_bwd(self, dy, x, st, name, rows, add=None):
        dx = self._new(rows, self.W)
        call("layernorm_bwd", dy, x, self.sd32[name + ".weight"], st[0], st[1], add, dx, None, None, rows, self.W)
        return dx

    # ---- multi-head attention.  Short sequences take the fused per-(sequence, head) kernel; the general form runs as
    # per (sequence, head) tcgen05 GEMMs batched through 4-D tensor maps (inner batch = head: 64-element column slice of
    # the fused qkv rows; outer batch = sequence), T padded to 64.
    TP = 64

    FUSED_ATTN = True   # tests switch it off to exercise the general (batched tcgen05 GEMM) attention path

    def _fused_attn(self):
        """T <= 64 tokens, 64-wide heads (every ViT-B/32 at 224 x 224): one CTA per (sequence, head) on mma.sync, scores and
        probabilities never leave the SM and nothing but qkv is saved for the backward (ffvc_mha_small_*)."""
        return self.FUSED_ATTN and self.T <= 64 and self.W == 64 * self.Hh

    def _attn_fwd(self, qkv, N):
        W, T, Hh, TP = self.W, self.T, self.Hh, self.TP
        if self._fused_attn():
            a = self._new(N * T, W)
            call("mha_small_fwd", qkv, a, N, T, Hh, 64, 0.125)
            return a, None
        if self.FUSED_ATTN and self.W == 64 * self.Hh:      # more than 64 tokens: tiled attention, saves the row log-sum-exp
            a = self._new(N * T, W)
            lse = self._new(N, Hh, T, dtype=F32)
            call("mha_flash_fwd", qkv, a, lse, N, T, Hh, 64, 0.125, 0)
            return a, lse
        S = self._new(N, Hh, T, TP, dtype=F32)
        ops.gemm(qkv, qkv, S, T, T, 64, a_ld=3 * W, b_ld=3 * W, b_off=W, a_role=ops.ROLE_OUT, b_role=ops.ROLE_OUT,
                 batch=N * Hh, batch_inner=Hh, a_bs=T * 3 * W, b_bs=T * 3 * W, a_bs_in=64, b_bs_in=64, ldc=TP,
                 out_bs=Hh * T * TP, out_bs_in=T * TP, alpha=0.125, block_n=64)
        P = self._new(N, Hh, T, TP)
        call("softmax_fwd", S, P, N * Hh * T, T, TP)
        a = self._new(N * T, W)
        # O[n,i,h,:] = sum_j P[n,h,i,j] V[n,j,h,:]
        ops.gemm(P, qkv, a, T, 64, T, a_ld=TP, b_mode=ops.MNMAJOR, b_ld=3 * W, b_off=2 * W, a_role=ops.ROLE_OUT,
                 b_role=ops.ROLE_OUT, batch=N * Hh, batch_inner=Hh, a_bs=Hh * T * TP, a_bs_in=T * TP, b_bs=T * 3 * W,
                 b_bs_in=64, ldc=W, out_bs=T * W, out_bs_in=64, block_n=64)
        return a, P

    def _attn_bwd(self, qkv, P, da, N, a=None):
        W, T, Hh, TP = self.W, self.T, self.Hh, self.TP
        if P is None:
            dqkv = self._new(N * T, 3 * W)
            call("mha_small_bwd", qkv, da, dqkv, N, T, Hh, 64, 0.125)
            return dqkv
        if P.dtype == F32:                    # log-sum-exp of the tiled forward
            dqkv = self._new(N * T, 3 * W)
            call("mha_flash_bwd", qkv, a, da, P, torch.empty_like(P), dqkv, N, T, Hh, 64, 0.125, 0)
            return dqkv
        kw = dict(batch=N * Hh, batch_inner=Hh, a_role=ops.ROLE_OUT, b_role=ops.ROLE_OUT, block_n=64)
        dP = self._new(N, Hh, T, TP, dtype=F32)            # dP[i,j] = sum_d dO[i,d] V[j,d]
        ops.gemm(da, qkv, dP, T, T, 64, a_ld=W, b_ld=3 * W, b_off=2 * W, a_bs=T * W, a_bs_in=64, b_bs=T * 3 * W, b_bs_in=64,
                 ldc=TP, out_bs=Hh * T * TP, out_bs_in=T * TP, **kw)
        dS = self._new(N, Hh, T, TP)
        call("softmax_bwd", P, dP, dS, N * Hh * T, T, TP, 0.125)
        del dP
        dqkv = self._new(N * T, 3 * W)
        ob = dict(ldc=3 * W, out_bs=T * 3 * W, out_bs_in=64)
        # dV[j,d] = sum_i P[i,j] dO[i,d]
        ops.gemm(P, da, dqkv, T, 64, T, a_mode=ops.MNMAJOR, a_ld=TP, b_mode=ops.MNMAJOR, b_ld=W, a_bs=Hh * T * TP,
                 a_bs_in=T * TP, b_bs=T * W, b_bs_in=64, out_off=2 * W, **ob, **kw)
        # dQ[i,d] = sum_j dS[i,j] K[j,d]
        ops.gemm(dS, qkv, dqkv, T, 64, T, a_ld=TP, b_mode=ops.MNMAJOR, b_ld=3 * W, b_off=W, a_bs=Hh * T * TP, a_bs_in=T * TP,
                 b_bs=T * 3 * W, b_bs_in=64, out_off=0, **ob, **kw)
        # dK[j,d] = sum_i dS[i,j] Q[i,d]
        ops.gemm(dS, qkv, dqkv, T, 64, T, a_mode=ops.MNMAJOR, a_ld=TP, b_mode=ops.MNMAJOR, b_ld=3 * W, b_off=0,
                 a_bs=Hh * T * TP, a_bs_in=T * TP, b_bs=T * 3 * W, b_bs_in=64, out_off=W, **ob, **kw)
        return dqkv

    def forward(self, patches):
        """patches: [N][G*G][3*patch*patch] bf16 -> (embed [N][E] fp32, saved)."""
        N = patches.shape[0]
        W, T, L, Hh, E = self.W, self.T, self.L, self.Hh, self.E
        Kp = 3 * self.patch * self.patch
        M = N * T
        pe = self._new(N * (T - 1), W)
        ops.gemm(patches, self.w["conv1"], pe, N * (T - 1), W, Kp)
        x0 = self._new(M, W)
        call("clip_assemble", pe, self.cls, self.pos, x0, N, T, W)
        del pe
        h, st_pre = self._ln(x0, "ln_pre", M)
        saved = dict(N=N, x0=x0, st_pre=st_pre, layers=[])
        for l in range(L):
            p = "transformer.resblocks.%d." % l
            n1, st1 = self._ln(h, p + "ln_1", M)
            qkv = self._new(M, 3 * W)
            ops.gemm(n1, self.w[p + "in"], qkv, M, 3 * W, W, bias=self.sd32[p + "attn.in_proj_bias"])
            del n1
            a, P = self._attn_fwd(qkv, N)
            h2 = self._new(M, W)
            ops.gemm(a, self.w[p + "out"], h2, M, W, W, bias=self.sd32[p + "attn.out_proj.bias"], res=h)
            a_keep = a if (P is not None and P.dtype == F32) else None       # the tiled backward wants O (delta = rowsum(dO * O))
            del a
            n2, st2 = self._ln(h2, p + "ln_2", M)
            u, gact = self._new(M, 4 * W), self._new(M, 4 * W)
            ops.gemm(n2, self.w[p + "fc"], gact, M, 4 * W, W, bias=self.sd32[p + "mlp.c_fc.bias"], act=self.act, pre_out=u)
            del n2
            h3 = self._new(M, W)
            ops.gemm(gact, self.w[p + "pj"], h3, M, W, 4 * W, bias=self.sd32[p + "mlp.c_proj.bias"], res=h2)
            del gact
            saved["layers"].append(dict(h=h, st1=st1, qkv=qkv, P=P, a=a_keep, h2=h2, st2=st2, u=u))
            h = h3
        xc = self._new(N, W)
        call("copy_rows", h, xc, N, W, T * W, W)                     # class-token rows
        nc, st_post = self._ln(xc, "ln_post", N)
        emb = self._new(N, E, dtype=F32)
        ops.gemm(nc, self.w["proj"], emb, N, E, W, b_mode=ops.MNMAJOR, b_ld=E)
        saved.update(xc=xc, st_post=st_post)
        return emb, saved

    def backward(self, saved, g_emb):
        """g_emb: [N][E] fp32 -> gradient w.r.t. the patches [N][G*G][3*patch*patch] bf16."""
        N = saved["N"]
        W, T, L, Hh, E = self.W, self.T, self.L, self.Hh, self.E
        Kp = 3 * self.patch * self.patch
        M = N * T
        gb = self._new(N, E)
        call("cast_f32_bf16", g_emb, gb, N * E)
        dnc = self._new(N, W)
        ops.gemm(gb, self.w["proj"], dnc, N, W, E)                    # dnc[n,w] = sum_e g[n,e] proj[w,e]
        dxc = self._ln_bwd(dnc, saved["xc"], saved["st_post"], "ln_post", N)
        dh = torch.zeros(M, W, device=self.dev, dtype=BF16)
        call("copy_rows", dxc, dh, N, W, W, T * W)
        for l in range(L - 1, -1, -1):
            p = "transformer.resblocks.%d." % l
            lv = saved["layers"][l]
            du = self._new(M, 4 * W)
            ops.linear_dgrad(dh, self.w[p + "pj"], du, M, W, 4 * W, aux=lv["u"], mul_mode=self.act)
            dn2 = self._new(M, W)
            ops.linear_dgrad(du, self.w[p + "fc"], dn2, M, 4 * W, W)
            del du
            dh2 = self._ln_bwd(dn2, lv["h2"], lv["st2"], p + "ln_2", M, add=dh)
            da = self._new(M, W)
            ops.linear_dgrad(dh2, self.w[p + "out"], da, M, W, W)
            dqkv = self._attn_bwd(lv["qkv"], lv["P"], da, N, a=lv["a"])
            dn1 = self._new(M, W)
            ops.linear_dgrad(dqkv, self.w[p + "in"], dn1, M, 3 * W, W)
            del dqkv
            dh = self._ln_bwd(dn1, lv["h"], lv["st1"], p + "ln_1", M, add=dh2)
        dx0 = self._ln_bwd(dh, saved["x0"], saved["st_pre"], "ln_pre", M)
        dpe = self._new(N * (T - 1), W)
        # rows 1..T-1 of every sequence, compacted: view dx0 as [N][T*W] and copy (T-1)*W contiguous elements per row
        call("copy_rows", dx0.view(-1)[W:], dpe, N, (T - 1) * W, T * W, (T - 1) * W)
        dpatch = self._new(N * (T - 1), Kp)
        ops.linear_dgrad(dpe, self.w["conv1"], dpatch, N * (T - 1), W, Kp)
        return dpatch.view(N, T - 1, Kp)
