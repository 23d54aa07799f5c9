"""SimpleVitGAN mapper (prompt embedding -> VQGAN latent grid), B200-native.

Drop-in for the reference's `vitgan.SimpleGenerator` (vitgan.py:262-305) as built by `build_model` for
`model_type: simple_vitgan` (main.py:469-478: size = vq_image_size, num_heads default 6): same constructor arguments, same
`state_dict()` keys / shapes, same construction order (seed parity), `forward(x: (B, input_dim)) -> (B, C, size, size)`.

Differences from `Generator` (vitgan_mapper.py): one token per latent position (T = size*size, 256 at 16x16), the encoder's
running state starts as `inp(noise)` viewed (B, dim, T) and permuted to (B, T, dim) plus `pos_emb1D`, and `w_out` maps
dim -> out_channels per token.  With T = 256 the attention no longer fits the small SIMT kernel of the 16-token mapper: it
runs as batched tcgen05 GEMMs (scores, softmax kernel, P.V) exactly like the X-transformer mapper.  Those GEMMs need
head-contiguous, 16-byte aligned q | k | v, while the reference interleaves the projection as '(d k h)' with a head dimension
of int(dim / heads) (170 at dim 1024 / 6 heads): the projection WEIGHTS are re-packed once per forward
(`ffvc_vitgan_pack_qkv_weight` / `_pack_out_weight`, head dimension zero-padded to a multiple of 8) and the weight gradients
un-packed (`ffvc_vitgan_unpack_*`), so activations never need a re-layout.
"""
import torch
from torch import nn

from . import ops
from .ops import BF16, F32, call
from .vitgan_mapper import _Block, _SLN


class SimpleGenerator(nn.Module):
    def __init__(self, size=8, in_channels=256, dim=384, blocks=6, num_heads=6, dim_head=None, dropout=0, out_channels=3,
                 input_dim=1024):
        super().__init__()
        if dropout != 0:
            raise NotImplementedError("dropout > 0 is not implemented on the B200 path (configs use dropout: 0)")
        self.size, self.dim, self.blocks, self.num_heads = size, dim, blocks, num_heads
        self.out_channels, self.input_dim = out_channels, input_dim
        T = size * size
        self.pos_emb1D = nn.Parameter(torch.randn(T, dim))                  # vitgan.py:283
        self.mlp = nn.Linear(input_dim, T * dim)                            # vitgan.py:285
        self.inp = nn.Linear(input_dim, T * dim)                            # vitgan.py:286
        self.Transformer_Encoder = nn.Module()
        self.Transformer_Encoder.blocks = nn.Sequential(*[_Block(dim, num_heads, dim_head) for _ in range(blocks)])
        self.w_out = nn.Sequential(nn.Linear(dim, out_channels))            # vitgan.py:290-294
        self.sln_norm = _SLN(dim)
        self._engine = None

    def engine(self):
        if self._engine is None or not self._engine.valid():
            self._engine = SimpleVitGANEngine(self)
        return self._engine

    def forward(self, noise):
        return _Fn.apply(self, noise, *list(self.parameters()))


class _Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        eng = module.engine()
        z, saved = eng.forward(x)                               # [B*T, C] fp32, token p = y*size + x
        ctx.eng, ctx.saved = eng, saved
        B, S, C = x.shape[0], eng.S, eng.C
        return z.view(B, S, S, C).permute(0, 3, 1, 2)           # vitgan.py:304

    @staticmethod
    def backward(ctx, dz):
        eng = ctx.eng
        B = dz.shape[0]
        dz_tok = dz.permute(0, 2, 3, 1).contiguous().view(B * eng.T, eng.C).float()
        eng.zero_grad_arena()
        eng.backward(ctx.saved, dz_tok)
        return (None, None) + tuple(eng.grad_views)


class SimpleVitGANEngine:
    """Flat fp32 master / grad arenas + bf16 shadow like the other mapper engines; activations bf16 [B*T, D]."""

    def __init__(self, m):
        self.m = m
        self.params = list(m.parameters())
        dev = self.params[0].device
        ops.require_cuda(dev, "the SimpleVitGAN mapper")
        self.dev = dev
        self.S, self.D, self.L, self.H = m.size, m.dim, m.blocks, m.num_heads
        self.T = self.S * self.S
        self.C, self.IN = m.out_channels, m.input_dim
        a0 = m.Transformer_Encoder.blocks[0].attn
        self.dh, self.Wd = a0.dim_head, a0.weight_dim
        self.dhp = (self.dh + 7) // 8 * 8                  # padded head dimension: 16-byte aligned head slices
        self.Wi = self.H * self.dhp                        # padded width of q, k, v and of the attention output
        if self.D % 8 or self.IN % 8 or self.C % 8 or self.T % 8:
            raise NotImplementedError("dim, input_dim, out_channels and size*size must be multiples of 8")
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
        # packed bf16 projection weights of every block (see the module docstring)
        self.wqkv = [torch.empty(3 * self.Wi, self.D, device=dev, dtype=BF16) for _ in range(self.L)]
        self.wout = [torch.empty(self.D, self.Wi, device=dev, dtype=BF16) for _ in range(self.L)]
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
        for i in range(self.L):     # 2 small re-packs per block from the fp32 masters (3.2 M + 1.1 M elements at dim 1024)
            p = "Transformer_Encoder.blocks.%d.attn." % i
            call("vitgan_pack_qkv_weight", self.wf(p + "to_qkv.weight"), self.wqkv[i], self.H, self.dh, self.dhp, self.D)
            call("vitgan_pack_out_weight", self.wf(p + "w_out.weight"), self.wout[i], self.H, self.dh, self.dhp, self.D)
        self.ext_shadow_fresh = False
        self._shadow_version = ver

    def zero_grad_arena(self):
        self.grad.zero_()

    def _new(self, *shape, dtype=BF16):
        return torch.empty(*shape, device=self.dev, dtype=dtype)

    # ---- SLN (vitgan.py:8-21), as in VitGANEngine
    def _sln_fwd(self, hl, x, p, R):
        D = self.D
        n, mu, rs = self._new(R, D), self._new(R, dtype=F32), self._new(R, dtype=F32)
        call("layernorm_fwd", hl, self.wf(p + "ln.weight"), self.wf(p + "ln.bias"), n, mu, rs, R, D, 1e-5)
        s = self._new(R, D)
        call("sln_mod_fwd", n, x, self.wf(p + "gamma"), self.wf(p + "beta"), s, R * D)
        return s, (n, mu, rs)

    def _sln_bwd(self, ds, hl, x, st, p, R, dx_acc, add):
        D = self.D
        n, mu, rs = st
        dn = self._new(R, D)
        call("sln_mod_bwd", ds, n, x, self.wf(p + "gamma"), self.wf(p + "beta"), dn, dx_acc, self.g(p + "gamma"),
             self.g(p + "beta"), R * D)
        dhl = self._new(R, D)
        call("layernorm_bwd", dn, hl, self.wf(p + "ln.weight"), mu, rs, add, dhl, self.g(p + "ln.weight"),
             self.g(p + "ln.bias"), R, D)
        return dhl

    # ---- attention over the packed [B][T][3*Wi] q | k | v activation: per-(sample, head) GEMMs, full (non-causal) softmax,
    #      score scale dim ** -0.5 (vitgan.py:65,90-93)
    def _attn_fwd(self, qkv, B):
        T, H, Wi, dhp = self.T, self.H, self.Wi, self.dhp
        S = self._new(B, H, T, T, dtype=F32)
        ops.gemm(qkv, qkv, S, T, T, dhp, a_ld=3 * Wi, b_ld=3 * Wi, b_off=Wi, a_role=ops.ROLE_OUT, b_role=ops.ROLE_OUT,
                 batch=B * H, batch_inner=H, a_bs=T * 3 * Wi, b_bs=T * 3 * Wi, a_bs_in=dhp, b_bs_in=dhp, ldc=T,
                 out_bs=H * T * T, out_bs_in=T * T, alpha=float(self.D) ** -0.5)
        P = self._new(B, H, T, T)
        call("softmax_fwd", S, P, B * H * T, T, T)
        del S
        a = self._new(B * T, Wi)
        ops.gemm(P, qkv, a, T, dhp, T, a_ld=T, b_mode=ops.MNMAJOR, b_ld=3 * Wi, b_off=2 * Wi, a_role=ops.ROLE_OUT,
                 b_role=ops.ROLE_OUT, batch=B * H, batch_inner=H, a_bs=H * T * T, a_bs_in=T * T, b_bs=T * 3 * Wi,
                 b_bs_in=dhp, ldc=Wi, out_bs=T * Wi, out_bs_in=dhp)
        return a, P

    def _attn_bwd(self, qkv, P, da, B):
        T, H, Wi, dhp = self.T, self.H, self.Wi, self.dhp
        kw = dict(batch=B * H, batch_inner=H, a_role=ops.ROLE_OUT, b_role=ops.ROLE_OUT)
        dP = self._new(B, H, T, T, dtype=F32)
        ops.gemm(da, qkv, dP, T, T, dhp, a_ld=Wi, b_ld=3 * Wi, b_off=2 * Wi, a_bs=T * Wi, a_bs_in=dhp, b_bs=T * 3 * Wi,
                 b_bs_in=dhp, ldc=T, out_bs=H * T * T, out_bs_in=T * T, **kw)
        dS = self._new(B, H, T, T)
        call("softmax_bwd", P, dP, dS, B * H * T, T, T, float(self.D) ** -0.5)
        del dP
        dqkv = self._new(B * T, 3 * Wi)
        ob = dict(ldc=3 * Wi, out_bs=T * 3 * Wi, out_bs_in=dhp)
        ops.gemm(P, da, dqkv, T, dhp, T, a_mode=ops.MNMAJOR, a_ld=T, b_mode=ops.MNMAJOR, b_ld=Wi, a_bs=H * T * T,
                 a_bs_in=T * T, b_bs=T * Wi, b_bs_in=dhp, out_off=2 * Wi, **ob, **kw)                            # dV = P^T dO
        ops.gemm(dS, qkv, dqkv, T, dhp, T, a_ld=T, b_mode=ops.MNMAJOR, b_ld=3 * Wi, b_off=Wi, a_bs=H * T * T,
                 a_bs_in=T * T, b_bs=T * 3 * Wi, b_bs_in=dhp, out_off=0, **ob, **kw)                             # dQ = dS K
        ops.gemm(dS, qkv, dqkv, T, dhp, T, a_mode=ops.MNMAJOR, a_ld=T, b_mode=ops.MNMAJOR, b_ld=3 * Wi, b_off=0,
                 a_bs=H * T * T, a_bs_in=T * T, b_bs=T * 3 * Wi, b_bs_in=dhp, out_off=Wi, **ob, **kw)            # dK = dS^T Q
        return dqkv

    # ------------------------------------------------------------------ forward
    def forward(self, noise):
        """noise: (B, IN) fp32 cuda -> z: (B*T, C) fp32 token-major (token p = y*size + x), plus saved activations."""
        self.refresh_shadow()
        B, T, D, L, C, Wi = noise.shape[0], self.T, self.D, self.L, self.C, self.Wi
        R = B * T
        nb = self._new(B, self.IN)
        call("cast_f32_bf16", noise.contiguous(), nb, B * self.IN)
        x = self._new(R, D)                                                     # modulation signal, carried unchanged
        ops.gemm(nb, self.w("mlp.weight"), x, B, T * D, self.IN, bias=self.wf("mlp.bias"))
        it = self._new(B, D * T)                                                # inp(noise), per sample [D][T] (vitgan.py:299)
        ops.gemm(nb, self.w("inp.weight"), it, B, T * D, self.IN, bias=self.wf("inp.bias"))
        e = self._new(R, D)
        call("transpose", it, e, B, D, T, 0, 0)                                 # [B][D][T] -> [B][T][D]
        del it
        pos = self._new(R, D)
        call("broadcast_rows", self.wf("pos_emb1D"), pos, B, T * D)
        hl = self._new(R, D)
        call("add_bf16", e, pos, hl, R * D)                                     # inp_emb + pos_emb1D (vitgan.py:300)
        del e, pos
        sv = dict(B=B, nb=nb, x=x, blocks=[])
        for i in range(L):
            p = "Transformer_Encoder.blocks.%d." % i
            s1, st1 = self._sln_fwd(hl, x, p + "norm1.", R)
            qkv = self._new(R, 3 * Wi)
            ops.gemm(s1, self.wqkv[i], qkv, R, 3 * Wi, D)
            a, P = self._attn_fwd(qkv, B)
            ht = self._new(R, D)
            ops.gemm(a, self.wout[i], ht, R, D, Wi, bias=self.wf(p + "attn.w_out.bias"), res=hl)
            s2, st2 = self._sln_fwd(ht, x, p + "norm2.", R)
            u, gact = self._new(R, 4 * D), self._new(R, 4 * D)
            ops.gemm(s2, self.w(p + "mlp.linear1.weight"), gact, R, 4 * D, D, bias=self.wf(p + "mlp.linear1.bias"),
                     act=ops.ACT_GELU, pre_out=u)
            hn = self._new(R, D)
            ops.gemm(gact, self.w(p + "mlp.linear2.weight"), hn, R, D, 4 * D, bias=self.wf(p + "mlp.linear2.bias"), res=ht)
            sv["blocks"].append(dict(hl=hl, st1=st1, s1=s1, qkv=qkv, a=a, P=P, ht=ht, st2=st2, s2=s2, u=u, g=gact))
            hl = hn
        sf, stf = self._sln_fwd(hl, x, "sln_norm.", R)
        z = self._new(R, C, dtype=F32)
        ops.gemm(sf, self.w("w_out.0.weight"), z, R, C, D, bias=self.wf("w_out.0.bias"))
        sv.update(hl_last=hl, stf=stf, sf=sf)
        return z, sv

    # ------------------------------------------------------------------ backward
    # ---- data parallel: completion order of the gradients inside the flat arena (parallel.bucket_slices)
    def layer_starts(self):
        """arena offset of the first parameter of every encoder block, ascending"""
        return [min(off for n, off in self.offs.items() if n.startswith("Transformer_Encoder.blocks.%d." % i)) for i in range(self.L)]

    def late_ranges(self):
        """nothing registered after the blocks finishes late: pos_emb1D and the input projections sit at the head of the arena"""
        return []

    def backward(self, sv, dz, on_layer_done=None):
        """dz: (B*T, C) fp32 token-major.  Accumulates all parameter gradients into self.grad.
        on_layer_done(i): called once the gradients of encoder block i (and of everything registered after it) are complete."""
        B, T, D, L, C, Wi = sv["B"], self.T, self.D, self.L, self.C, self.Wi
        H, dh, dhp = self.H, self.dh, self.dhp
        R = B * T
        x = sv["x"]
        sp = ops.auto_splits
        dzb = self._new(R, C)
        call("cast_f32_bf16", dz.contiguous(), dzb, R * C)
        dx_acc = torch.zeros(R, D, device=self.dev, dtype=F32)
        ops.linear_wgrad(dzb, sv["sf"], self.g("w_out.0.weight"), R, C, D, splits=sp(C, D, R))
        call("colsum", dzb, self.g("w_out.0.bias"), R, C)
        dsf = self._new(R, D)
        ops.linear_dgrad(dzb, self.w("w_out.0.weight"), dsf, R, C, D)
        dhl = self._sln_bwd(dsf, sv["hl_last"], x, sv["stf"], "sln_norm.", R, dx_acc, None)
        dwq = torch.empty(3 * Wi, D, device=self.dev, dtype=F32)               # packed-layout weight gradients (scratch)
        dwo = torch.empty(D, Wi, device=self.dev, dtype=F32)
        for i in range(L - 1, -1, -1):
            p = "Transformer_Encoder.blocks.%d." % i
            bv = sv["blocks"][i]
            # ---- MLP
            ops.linear_wgrad(dhl, bv["g"], self.g(p + "mlp.linear2.weight"), R, D, 4 * D, splits=sp(D, 4 * D, R))
            call("colsum", dhl, self.g(p + "mlp.linear2.bias"), R, D)
            du = self._new(R, 4 * D)
            ops.linear_dgrad(dhl, self.w(p + "mlp.linear2.weight"), du, R, D, 4 * D, aux=bv["u"], mul_mode=ops.ACT_GELU)
            ops.linear_wgrad(du, bv["s2"], self.g(p + "mlp.linear1.weight"), R, 4 * D, D, splits=sp(4 * D, D, R))
            call("colsum", du, self.g(p + "mlp.linear1.bias"), R, 4 * D)
            ds2 = self._new(R, D)
            ops.linear_dgrad(du, self.w(p + "mlp.linear1.weight"), ds2, R, 4 * D, D)
            del du
            dht = self._sln_bwd(ds2, bv["ht"], x, bv["st2"], p + "norm2.", R, dx_acc, dhl)
            # ---- attention output projection: packed gradient, then un-pack into the reference layout
            dwo.zero_()
            ops.linear_wgrad(dht, bv["a"], dwo, R, D, Wi, splits=sp(D, Wi, R))
            call("vitgan_unpack_out_wgrad", dwo, self.g(p + "attn.w_out.weight"), H, dh, dhp, D)
            call("colsum", dht, self.g(p + "attn.w_out.bias"), R, D)
            da = self._new(R, Wi)
            ops.linear_dgrad(dht, self.wout[i], da, R, D, Wi)
            dqkv = self._attn_bwd(bv["qkv"], bv["P"], da, B)
            dwq.zero_()
            ops.linear_wgrad(dqkv, bv["s1"], dwq, R, 3 * Wi, D, splits=sp(3 * Wi, D, R))
            call("vitgan_unpack_qkv_wgrad", dwq, self.g(p + "attn.to_qkv.weight"), H, dh, dhp, D)
            ds1 = self._new(R, D)
            ops.linear_dgrad(dqkv, self.wqkv[i], ds1, R, 3 * Wi, D)
            del dqkv
            dhl = self._sln_bwd(ds1, bv["hl"], x, bv["st1"], p + "norm1.", R, dx_acc, dht)
            if on_layer_done is not None:
                on_layer_done(i)
        # hl0 = permute(inp(noise)) + pos_emb1D: pos_emb1D receives the batch sum, inp the per-sample transposed gradient
        call("colsum", dhl, self.g("pos_emb1D"), B, T * D)
        dit = self._new(B, D * T)
        call("transpose", dhl, dit, B, T, D, 0, 0)                             # [B][T][D] -> [B][D][T]
        ops.linear_wgrad(dit, sv["nb"], self.g("inp.weight"), B, T * D, self.IN)
        call("colsum", dit, self.g("inp.bias"), B, T * D)
        # x = mlp(noise): only wgrad / bias grad (the prompt embedding needs no gradient)
        dxb = self._new(R, D)
        call("cast_f32_bf16", dx_acc, dxb, R * D)
        ops.linear_wgrad(dxb, sv["nb"], self.g("mlp.weight"), B, T * D, self.IN)
        call("colsum", dxb, self.g("mlp.bias"), B, T * D)
