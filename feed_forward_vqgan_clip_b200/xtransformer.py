"""X-transformer mapper (prompt embedding -> VQGAN latent grid), B200-native.

Drop-in for the reference's `XTransformer` (transformer.py:5-46) as built by `build_model` for `model_type: xtransformer`
(main.py:488-499: initial_proj=True, add_input=False): `proj` Linear(input_dim -> S*S*dim), then x-transformers'
ContinuousTransformerWrapper(dim_in=dim, dim_out=channels, max_seq_len=S*S+1, Decoder(dim, depth, heads)) — causal
pre-LayerNorm attention / feed-forward stack — and the (B,S,S,C)->permute output.  The package's module names are mirrored
so its state_dict keys line up (see oracle/xtransformer.py for the key list and the [recall] caveats: the arithmetic of
x-transformers==0.19.1 is not available here, parity is against the restatement).
Arithmetic in libffvc_sm100.so: tcgen05 GEMMs for every projection, per-(sample, head) Q.K^T / P.V GEMMs batched through
4-D tensor maps over the fused q|k|v activation, causal softmax forward / softmax backward, LayerNorm; forward, dgrad, wgrad.
"""
import torch
from torch import nn

from . import ops
from .ops import BF16, F32, call

DIM_HEAD = 64


class _Attn(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        inner = heads * DIM_HEAD
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.Linear(inner, dim)


class _FF(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.Sequential(nn.Sequential(nn.Linear(dim, 4 * dim), nn.GELU()), nn.Dropout(0.0), nn.Linear(4 * dim, dim))


class _PosEmb(nn.Module):
    def __init__(self, dim, n):
        super().__init__()
        self.emb = nn.Embedding(n, dim)


class XTransformer(nn.Module):
    def __init__(self, input_dim, image_size, channels, dim, depth, heads, initial_proj=True, add_input=True):
        super().__init__()
        if not initial_proj:
            raise NotImplementedError("only the initial_proj=True path (what build_model constructs, main.py:497) is built")
        self.input_dim, self.image_size, self.channels, self.dim, self.depth, self.heads = (input_dim, image_size, channels,
                                                                                             dim, depth, heads)
        T = image_size * image_size
        tr = nn.Module()
        tr.project_in = nn.Linear(dim, dim)
        tr.pos_emb = _PosEmb(dim, T + (0 if add_input else 1))
        tr.attn_layers = nn.Module()
        layers = []
        for _ in range(depth):
            layers.append(nn.ModuleList([nn.LayerNorm(dim), _Attn(dim, heads), nn.Identity()]))
            layers.append(nn.ModuleList([nn.LayerNorm(dim), _FF(dim), nn.Identity()]))
        tr.attn_layers.layers = nn.ModuleList(layers)
        tr.norm = nn.LayerNorm(dim)
        tr.project_out = nn.Linear(dim, channels)
        self.transformer = tr
        self.proj = nn.Linear(input_dim, T * dim)
        self._engine = None

    def engine(self):
        if self._engine is None or not self._engine.valid():
            self._engine = XTEngine(self)
        return self._engine

    def forward(self, x):
        return _Fn.apply(self, x, *list(self.parameters()))


class _Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        eng = module.engine()
        z, saved = eng.forward(x)
        ctx.eng, ctx.saved = eng, saved
        B, S, C = x.shape[0], eng.S, eng.C
        return z.view(B, S, S, C).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dz):
        eng = ctx.eng
        B = dz.shape[0]
        dz_tok = dz.permute(0, 2, 3, 1).contiguous().view(B * eng.T, eng.C).float()
        eng.zero_grad_arena()
        eng.backward(ctx.saved, dz_tok)
        return (None, None) + tuple(eng.grad_views)


class XTEngine:
    def __init__(self, m):
        self.m = m
        self.params = list(m.parameters())
        dev = self.params[0].device
        ops.require_cuda(dev, "the X-transformer mapper")
        self.dev = dev
        self.S, self.C, self.D, self.L, self.H, self.IN = m.image_size, m.channels, m.dim, m.depth, m.heads, m.input_dim
        self.T = self.S * self.S
        self.Wi = self.H * DIM_HEAD
        if self.D % 8 or self.IN % 8 or self.C % 8:
            raise NotImplementedError("dims must be multiples of 8")
        sizes = [p.numel() for p in self.params]
        offs, o = [], 0
        for n in sizes:
            offs.append(o)
            o += (n + 7) // 8 * 8
        self.total = o
        self.arena = torch.zeros(o, device=dev, dtype=F32)
        self.grad = torch.zeros(o, device=dev, dtype=F32)
        self.shadow = torch.empty(o, device=dev, dtype=BF16)
        self.grad_views, self._ptrs = [], []
        for p, off, n in zip(self.params, offs, sizes):
            view = self.arena[off:off + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            self.grad_views.append(self.grad[off:off + n].view(p.shape))
            self._ptrs.append(p.data_ptr())
        names = [n for n, _ in m.named_parameters()]
        self.offs = dict(zip(names, offs))
        self.numel = dict(zip(names, sizes))
        self._shadow_version = None
        self.ext_shadow_fresh = False

    def valid(self):
        return all(p.data_ptr() == q for p, q in zip(self.params, self._ptrs))

    def w(self, name):
        o = self.offs[name]
        return self.shadow[o:o + self.numel[name]]

    def wf(self, name):
        o = self.offs[name]
        return self.arena[o:o + self.numel[name]]

    def g(self, name):
        o = self.offs[name]
        return self.grad[o:o + self.numel[name]]

    def refresh_shadow(self):
        ver = tuple(p._version for p in self.params)
        if not ((self.ext_shadow_fresh and ver == getattr(self, "_adam_ver", ver)) or ver == self._shadow_version):
            call("cast_f32_bf16", self.arena, self.shadow, self.total)
        self.ext_shadow_fresh = False
        self._shadow_version = ver

    def zero_grad_arena(self):
        self.grad.zero_()

    def _new(self, *shape, dtype=BF16):
        return torch.empty(*shape, device=self.dev, dtype=dtype)

    # attention over the fused [B][T][3*Wi] q|k|v activation; per-(sample, head) GEMMs, causal
    # tiled attention (ffvc_mha_flash_*): scores / probabilities never reach HBM, the forward saves the row log-sum-exp only.
    # False selects round 1's batched tcgen05 GEMMs around a materialised T x T softmax (kept for A/B and as a cross-check).
    FLASH_ATTN = True

    def _attn_fwd(self, qkv, B):
        T, H, Wi = self.T, self.H, self.Wi
        if self.FLASH_ATTN:
            a = self._new(B * T, Wi)
            lse = self._new(B, H, T, dtype=F32)
            call("mha_flash_fwd", qkv, a, lse, B, T, H, DIM_HEAD, DIM_HEAD ** -0.5, 1)
            return a, lse
        S = self._new(B, H, T, T, dtype=F32)
        ops.gemm(qkv, qkv, S, T, T, DIM_HEAD, a_ld=3 * Wi, b_ld=3 * Wi, b_off=Wi, a_role=ops.ROLE_OUT, b_role=ops.ROLE_OUT,
                 batch=B * H, batch_inner=H, a_bs=T * 3 * Wi, b_bs=T * 3 * Wi, a_bs_in=DIM_HEAD, b_bs_in=DIM_HEAD, ldc=T,
                 out_bs=H * T * T, out_bs_in=T * T, alpha=DIM_HEAD ** -0.5)
        P = self._new(B, H, T, T)
        call("softmax_causal_fwd", S, P, B * H * T, T, T)
        del S
        a = self._new(B * T, Wi)
        ops.gemm(P, qkv, a, T, DIM_HEAD, T, a_ld=T, b_mode=ops.MNMAJOR, b_ld=3 * Wi, b_off=2 * Wi, a_role=ops.ROLE_OUT,
                 b_role=ops.ROLE_OUT, batch=B * H, batch_inner=H, a_bs=H * T * T, a_bs_in=T * T, b_bs=T * 3 * Wi,
                 b_bs_in=DIM_HEAD, ldc=Wi, out_bs=T * Wi, out_bs_in=DIM_HEAD, block_n=64)
        return a, P

    def _attn_bwd(self, qkv, P, da, B, a=None):
        T, H, Wi = self.T, self.H, self.Wi
        if P.dtype == F32:                    # the flash forward's log-sum-exp: recompute the probabilities block by block
            dqkv = self._new(B * T, 3 * Wi)
            call("mha_flash_bwd", qkv, a, da, P, torch.empty_like(P), dqkv, B, T, H, DIM_HEAD, DIM_HEAD ** -0.5, 1)
            return dqkv
        kw = dict(batch=B * H, batch_inner=H, a_role=ops.ROLE_OUT, b_role=ops.ROLE_OUT)
        dP = self._new(B, H, T, T, dtype=F32)
        ops.gemm(da, qkv, dP, T, T, DIM_HEAD, a_ld=Wi, b_ld=3 * Wi, b_off=2 * Wi, a_bs=T * Wi, a_bs_in=DIM_HEAD, b_bs=T * 3 * Wi,
                 b_bs_in=DIM_HEAD, ldc=T, out_bs=H * T * T, out_bs_in=T * T, **kw)
        dS = self._new(B, H, T, T)
        call("softmax_bwd", P, dP, dS, B * H * T, T, T, DIM_HEAD ** -0.5)     # P is 0 above the diagonal => dS is too
        del dP
        dqkv = self._new(B * T, 3 * Wi)
        ob = dict(ldc=3 * Wi, out_bs=T * 3 * Wi, out_bs_in=DIM_HEAD, block_n=64)
        ops.gemm(P, da, dqkv, T, DIM_HEAD, T, a_mode=ops.MNMAJOR, a_ld=T, b_mode=ops.MNMAJOR, b_ld=Wi, a_bs=H * T * T,
                 a_bs_in=T * T, b_bs=T * Wi, b_bs_in=DIM_HEAD, out_off=2 * Wi, **ob, **kw)                       # dV
        ops.gemm(dS, qkv, dqkv, T, DIM_HEAD, T, a_ld=T, b_mode=ops.MNMAJOR, b_ld=3 * Wi, b_off=Wi, a_bs=H * T * T,
                 a_bs_in=T * T, b_bs=T * 3 * Wi, b_bs_in=DIM_HEAD, out_off=0, **ob, **kw)                        # dQ
        ops.gemm(dS, qkv, dqkv, T, DIM_HEAD, T, a_mode=ops.MNMAJOR, a_ld=T, b_mode=ops.MNMAJOR, b_ld=3 * Wi, b_off=0,
                 a_bs=H * T * T, a_bs_in=T * T, b_bs=T * 3 * Wi, b_bs_in=DIM_HEAD, out_off=Wi, **ob, **kw)       # dK
        return dqkv

    def forward(self, x):
        """x: (B, IN) fp32 cuda -> z: (B*T, C) fp32 token-major, plus saved activations."""
        self.refresh_shadow()
        B, T, D, L, C, Wi = x.shape[0], self.T, self.D, self.L, self.C, self.Wi
        R = B * T
        xb = self._new(B, self.IN)
        call("cast_f32_bf16", x.contiguous(), xb, B * self.IN)
        h0 = self._new(R, D)
        ops.gemm(xb, self.w("proj.weight"), h0, B, T * D, self.IN, bias=self.wf("proj.bias"))
        # project_in, then + positional embedding: realised as a residual operand broadcast over the batch
        pos = self._new(R, D)
        call("broadcast_rows", self.wf("transformer.pos_emb.emb.weight"), pos, B, T * D)
        h = self._new(R, D)
        ops.gemm(h0, self.w("transformer.project_in.weight"), h, R, D, D, bias=self.wf("transformer.project_in.bias"), res=pos)
        del pos
        sv = dict(B=B, xb=xb, h0=h0, layers=[])
        for j in range(L):
            pa = "transformer.attn_layers.layers.%d." % (2 * j)
            n1, mu1, rs1 = self._new(R, D), self._new(R, dtype=F32), self._new(R, dtype=F32)
            call("layernorm_fwd", h, self.wf(pa + "0.weight"), self.wf(pa + "0.bias"), n1, mu1, rs1, R, D, 1e-5)
            qkv = self._new(R, 3 * Wi)
            for i, nm in enumerate(("to_q", "to_k", "to_v")):
                ops.gemm(n1, self.w(pa + "1.%s.weight" % nm), qkv, R, Wi, D, ldc=3 * Wi, out_off=i * Wi)
            a, P = self._attn_fwd(qkv, B)
            h2 = self._new(R, D)
            ops.gemm(a, self.w(pa + "1.to_out.weight"), h2, R, D, Wi, bias=self.wf(pa + "1.to_out.bias"), res=h)
            pf = "transformer.attn_layers.layers.%d." % (2 * j + 1)
            n2, mu2, rs2 = self._new(R, D), self._new(R, dtype=F32), self._new(R, dtype=F32)
            call("layernorm_fwd", h2, self.wf(pf + "0.weight"), self.wf(pf + "0.bias"), n2, mu2, rs2, R, D, 1e-5)
            u, gact = self._new(R, 4 * D), self._new(R, 4 * D)
            ops.gemm(n2, self.w(pf + "1.net.0.0.weight"), gact, R, 4 * D, D, bias=self.wf(pf + "1.net.0.0.bias"),
                     act=ops.ACT_GELU, pre_out=u)
            h3 = self._new(R, D)
            ops.gemm(gact, self.w(pf + "1.net.2.weight"), h3, R, D, 4 * D, bias=self.wf(pf + "1.net.2.bias"), res=h2)
            sv["layers"].append(dict(h=h, n1=n1, st1=(mu1, rs1), qkv=qkv, a=a, P=P, h2=h2, n2=n2, st2=(mu2, rs2), u=u, g=gact))
            h = h3
        nf, muf, rsf = self._new(R, D), self._new(R, dtype=F32), self._new(R, dtype=F32)
        call("layernorm_fwd", h, self.wf("transformer.norm.weight"), self.wf("transformer.norm.bias"), nf, muf, rsf, R, D, 1e-5)
        z = self._new(R, C, dtype=F32)
        ops.gemm(nf, self.w("transformer.project_out.weight"), z, R, C, D, bias=self.wf("transformer.project_out.bias"))
        sv.update(hL=h, nf=nf, stf=(muf, rsf))
        return z, sv

    # ---- data parallel: completion order of the gradients inside the flat arena (parallel.bucket_slices)
    def layer_starts(self):
        """arena offset of the first parameter of every (attention, feed-forward) layer pair, ascending"""
        return [self.offs["transformer.attn_layers.layers.%d.0.weight" % (2 * j)] for j in range(self.L)]

    def late_ranges(self):
        """`proj` is registered after the transformer (transformer.py:10-23) but its wgrad is the last GEMM of backward"""
        lo = self.offs["proj.weight"]
        hi = self.offs["proj.bias"] + (self.numel["proj.bias"] + 7) // 8 * 8
        return [(lo, hi)]

    def backward(self, sv, dz, on_layer_done=None):
        B, T, D, L, C, Wi = sv["B"], self.T, self.D, self.L, self.C, self.Wi
        R = B * T
        sp = ops.auto_splits
        dzb = self._new(R, C)
        call("cast_f32_bf16", dz.contiguous(), dzb, R * C)
        ops.linear_wgrad(dzb, sv["nf"], self.g("transformer.project_out.weight"), R, C, D, splits=sp(C, D, R))
        call("colsum", dzb, self.g("transformer.project_out.bias"), R, C)
        dnf = self._new(R, D)
        ops.linear_dgrad(dzb, self.w("transformer.project_out.weight"), dnf, R, C, D)
        dh = self._new(R, D)
        call("layernorm_bwd", dnf, sv["hL"], self.wf("transformer.norm.weight"), sv["stf"][0], sv["stf"][1], None, dh,
             self.g("transformer.norm.weight"), self.g("transformer.norm.bias"), R, D)
        for j in range(L - 1, -1, -1):
            lv = sv["layers"][j]
            pf = "transformer.attn_layers.layers.%d." % (2 * j + 1)
            ops.linear_wgrad(dh, lv["g"], self.g(pf + "1.net.2.weight"), R, D, 4 * D, splits=sp(D, 4 * D, R))
            call("colsum", dh, self.g(pf + "1.net.2.bias"), R, D)
            du = self._new(R, 4 * D)
            ops.linear_dgrad(dh, self.w(pf + "1.net.2.weight"), du, R, D, 4 * D, aux=lv["u"], mul_mode=ops.ACT_GELU)
            ops.linear_wgrad(du, lv["n2"], self.g(pf + "1.net.0.0.weight"), R, 4 * D, D, splits=sp(4 * D, D, R))
            call("colsum", du, self.g(pf + "1.net.0.0.bias"), R, 4 * D)
            dn2 = self._new(R, D)
            ops.linear_dgrad(du, self.w(pf + "1.net.0.0.weight"), dn2, R, 4 * D, D)
            del du
            dh2 = self._new(R, D)
            call("layernorm_bwd", dn2, lv["h2"], self.wf(pf + "0.weight"), lv["st2"][0], lv["st2"][1], dh, dh2,
                 self.g(pf + "0.weight"), self.g(pf + "0.bias"), R, D)
            pa = "transformer.attn_layers.layers.%d." % (2 * j)
            ops.linear_wgrad(dh2, lv["a"], self.g(pa + "1.to_out.weight"), R, D, Wi, splits=sp(D, Wi, R))
            call("colsum", dh2, self.g(pa + "1.to_out.bias"), R, D)
            da = self._new(R, Wi)
            ops.linear_dgrad(dh2, self.w(pa + "1.to_out.weight"), da, R, D, Wi)
            dqkv = self._attn_bwd(lv["qkv"], lv["P"], da, B, a=lv["a"])
            dn1 = None
            for i, nm in enumerate(("to_q", "to_k", "to_v")):
                # wgrad: dW[n,k] += sum_m dqkv[m, i*Wi + n] n1[m,k]
                ops.gemm(dqkv, lv["n1"], self.g(pa + "1.%s.weight" % nm), Wi, D, R, a_mode=ops.MNMAJOR, b_mode=ops.MNMAJOR,
                         a_ld=3 * Wi, b_ld=D, a_off=i * Wi, atomic=True, splits=sp(Wi, D, R))
                nxt = self._new(R, D)
                ops.gemm(dqkv, self.w(pa + "1.%s.weight" % nm), nxt, R, D, Wi, a_ld=3 * Wi, a_off=i * Wi, b_mode=ops.MNMAJOR,
                         b_ld=D, res=dn1)
                dn1 = nxt
            del dqkv
            dh = self._new(R, D)
            call("layernorm_bwd", dn1, lv["h"], self.wf(pa + "0.weight"), lv["st1"][0], lv["st1"][1], dh2, dh,
                 self.g(pa + "0.weight"), self.g(pa + "0.bias"), R, D)
            if on_layer_done is not None:
                on_layer_done(j)
        # h = project_in(h0) + pos
        call("colsum", dh, self.g("transformer.pos_emb.emb.weight"), B, T * D)          # sum over the batch; rows beyond T untouched
        ops.linear_wgrad(dh, sv["h0"], self.g("transformer.project_in.weight"), R, D, D, splits=sp(D, D, R))
        call("colsum", dh, self.g("transformer.project_in.bias"), R, D)
        dh0 = self._new(R, D)
        ops.linear_dgrad(dh, self.w("transformer.project_in.weight"), dh0, R, D, D)
        ops.linear_wgrad(dh0, sv["xb"], self.g("proj.weight"), B, T * D, self.IN)
        call("colsum", dh0, self.g("proj.bias"), B, T * D)
