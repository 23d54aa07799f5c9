"""VitGAN mapper (prompt embedding -> VQGAN latent grid), B200-native.

Drop-in for the reference's `vitgan.Generator` (vitgan.py:221-260) as built by `build_model` for `model_type: vitgan`
(main.py:459-468: initialize_size = vq_image_size // 8, num_heads default 6): same constructor arguments, same
`state_dict()` keys / shapes, same construction order (seed parity), `forward(x: (B, input_dim)) ->
(B, C, T, T)` with T = initialize_size * 8 tokens and the reference's raw `view` of the projected tokens.
The nn.Modules are parameter containers; arithmetic (forward, dgrad, wgrad) runs in libffvc_sm100.so:
tcgen05 GEMMs for every Linear, LayerNorm + SLN modulation kernels, and a small attention kernel that honours the
reference's '(d k h)' interleaved qkv split and its `dim ** -0.5` score scale (vitgan.py:65,82).
"""
import torch
from torch import nn

from . import ops
from .ops import BF16, F32, call


class _SLN(nn.Module):                       # vitgan.py:8-21
    def __init__(self, dim):
        super().__init__()
        self.ln = nn.LayerNorm(dim)
        self.gamma = nn.Parameter(torch.randn(1, 1, 1))
        self.beta = nn.Parameter(torch.randn(1, 1, 1))


class _Attention(nn.Module):                 # vitgan.py:59-67
    def __init__(self, dim, num_heads, dim_head):
        super().__init__()
        self.num_heads = num_heads
        self.dim_head = int(dim / num_heads) if dim_head is None else dim_head
        self.weight_dim = self.num_heads * self.dim_head
        self.to_qkv = nn.Linear(dim, self.weight_dim * 3, bias=False)
        self.w_out = nn.Linear(self.weight_dim, dim, bias=True)


class _MLP(nn.Module):                       # vitgan.py:24-34
    def __init__(self, dim, hid):
        super().__init__()
        self.linear1 = nn.Linear(dim, hid)
        self.linear2 = nn.Linear(hid, dim)


class _Block(nn.Module):                     # vitgan.py:120-130
    def __init__(self, dim, num_heads, dim_head):
        super().__init__()
        self.attn = _Attention(dim, num_heads, dim_head)
        self.norm1 = _SLN(dim)
        self.norm2 = _SLN(dim)
        self.mlp = _MLP(dim, dim * 4)


class Generator(nn.Module):
    def __init__(self, initialize_size=8, dim=384, blocks=6, num_heads=6, dim_head=None, dropout=0, out_channels=3,
                 input_dim=1024):
        super().__init__()
        if dropout != 0:
            raise NotImplementedError("dropout > 0 is not implemented on the B200 path (configs use dropout: 0)")
        self.initialize_size, self.dim, self.blocks, self.num_heads = initialize_size, dim, blocks, num_heads
        self.out_channels, self.input_dim = out_channels, input_dim
        T = initialize_size * 8
        self.pos_emb1D = nn.Parameter(torch.randn(T, dim))
        self.mlp = nn.Linear(input_dim, T * dim)
        self.Transformer_Encoder = nn.Module()
        self.Transformer_Encoder.blocks = nn.Sequential(*[_Block(dim, num_heads, dim_head) for _ in range(blocks)])
        self.w_out = nn.Sequential(nn.Linear(dim, T * out_channels))
        self.sln_norm = _SLN(dim)
        self._engine = None

    def engine(self):
        if self._engine is None or not self._engine.valid():
            self._engine = VitGANEngine(self)
        return self._engine

    def forward(self, noise):
        return _Fn.apply(self, noise, *list(self.parameters()))


class _Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        eng = module.engine()
        z, saved = eng.forward(x)                               # [B*T*T, C] token-major fp32
        ctx.eng, ctx.saved = eng, saved
        B, T, C = x.shape[0], eng.T, eng.C
        return z.view(B, T, T, C).permute(0, 3, 1, 2)           # == reference's x.view(B, C, T, T) values

    @staticmethod
    def backward(ctx, dz):
        eng = ctx.eng
        B = dz.shape[0]
        dz_tok = dz.permute(0, 2, 3, 1).contiguous().view(B * eng.T * eng.T, eng.C).float()
        eng.zero_grad_arena()
        eng.backward(ctx.saved, dz_tok)
        return (None, None) + tuple(eng.grad_views)


class VitGANEngine:
    """Flat fp32 master / grad arenas + bf16 shadow like MixerEngine; activations bf16 [B*T, D]."""

    PAD = 8

    def __init__(self, m):
        self.m = m
        self.params = list(m.parameters())
        dev = self.params[0].device
        ops.require_cuda(dev, "the VitGAN mapper")
        self.dev = dev
        self.T, self.D, self.L, self.H = m.initialize_size * 8, m.dim, m.blocks, m.num_heads
        self.C, self.IN = m.out_channels, m.input_dim
        self.S = self.T                                     # latent grid side, as TrainStep expects
        a0 = m.Transformer_Encoder.blocks[0].attn
        self.dh, self.Wd = a0.dim_head, a0.weight_dim
        self.Wp = (self.Wd + 7) // 8 * 8                   # padded row pitch of the attention output / w_out operand
        self.Q3 = (3 * self.Wd + 7) // 8 * 8               # padded row pitch of the qkv activations
        if self.D % 8 or self.IN % 8 or (self.T * self.C) % 8:
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
        # w_out weights are [D][Wd] with Wd = 1020 at D = 1024 / 6 heads: bf16 copies with a 16-byte-multiple row pitch
        self.wout_pad = [torch.zeros(self.D, self.Wp, device=dev, dtype=BF16) for _ in range(self.L)]
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
        fresh = (self.ext_shadow_fresh and ver == getattr(self, "_adam_ver", ver)) or ver == self._shadow_version
        if not fresh:
            call("cast_f32_bf16", self.arena, self.shadow, self.total)
        for i in range(self.L):   # the padded copies are cheap (1 M elements per block): refresh every forward
            call("cast_f32_bf16_pitched", self.wf("Transformer_Encoder.blocks.%d.attn.w_out.weight" % i), self.wout_pad[i],
                 self.D, self.Wd, self.Wp)
        self.ext_shadow_fresh = False
        self._shadow_version = ver

    def zero_grad_arena(self):
        self.grad.zero_()

    def _new(self, *shape, dtype=BF16):
        return torch.empty(*shape, device=self.dev, dtype=dtype)

    def _sln_fwd(self, hl, x, p, R):
        D = self.D
        n, mu, rs = self._new(R, D), self._new(R, dtype=F32), self._new(R, dtype=F32)
        call("layernorm_fwd", hl, self.wf(p + "ln.weight"), self.wf(p + "ln.bias"), n, mu, rs, R, D, 1e-5)
        s = self._new(R, D)
        call("sln_mod_fwd", n, x, self.wf(p + "gamma"), self.wf(p + "beta"), s, R * D)
        return s, (n, mu, rs)

    def _sln_bwd(self, ds, hl, x, st, p, R, dx_acc, add):
        """returns d(hl) (+ add); accumulates dx, dgamma, dbeta, d(ln.weight), d(ln.bias)."""
        D = self.D
        n, mu, rs = st
        dn = self._new(R, D)
        call("sln_mod_bwd", ds, n, x, self.wf(p + "gamma"), self.wf(p + "beta"), dn, dx_acc, self.g(p + "gamma"),
             self.g(p + "beta"), R * D)
        dhl = self._new(R, D)
        call("layernorm_bwd", dn, hl, self.wf(p + "ln.weight"), mu, rs, add, dhl, self.g(p + "ln.weight"),
             self.g(p + "ln.bias"), R, D)
        return dhl

    # ------------------------------------------------------------------ forward
    def forward(self, noise):
        """noise: (B, IN) fp32 cuda -> z: (B*T*T, C) fp32 token-major (token p = y*T + x), plus saved activations."""
        self.refresh_shadow()
        B, T, D, L, H, C = noise.shape[0], self.T, self.D, self.L, self.H, self.C
        Wd, Wp, Q3, dh = self.Wd, self.Wp, self.Q3, self.dh
        R = B * T
        nb = self._new(B, self.IN)
        call("cast_f32_bf16", noise.contiguous(), nb, B * self.IN)
        x = self._new(R, D)                                                     # modulation signal, carried unchanged
        ops.gemm(nb, self.w("mlp.weight"), x, B, T * D, self.IN, bias=self.wf("mlp.bias"))
        hl = self._new(R, D)
        call("broadcast_rows", self.wf("pos_emb1D"), hl, B, T * D)
        sv = dict(B=B, nb=nb, x=x, blocks=[])
        scale = float(D) ** -0.5
        for i in range(L):
            p = "Transformer_Encoder.blocks.%d." % i
            s1, st1 = self._sln_fwd(hl, x, p + "norm1.", R)
            qkv = torch.zeros(R, Q3, device=self.dev, dtype=BF16) if Q3 != 3 * Wd else self._new(R, Q3)
            ops.gemm(s1, self.w(p + "attn.to_qkv.weight"), qkv, R, 3 * Wd, D, ldc=Q3)
            a = torch.zeros(R, Wp, device=self.dev, dtype=BF16) if Wp != Wd else self._new(R, Wp)
            probs = self._new(B * H, T, T, dtype=F32)
            call("vitgan_attn_fwd", qkv, a, probs, B, T, H, dh, Q3, Wp, scale)
            ht = self._new(R, D)
            ops.gemm(a, self.wout_pad[i], ht, R, D, Wd, a_ld=Wp, b_ld=Wp, bias=self.wf(p + "attn.w_out.bias"), res=hl)
            s2, st2 = self._sln_fwd(ht, x, p + "norm2.", R)
            u, gact = self._new(R, 4 * D), self._new(R, 4 * D)
            ops.gemm(s2, self.w(p + "mlp.linear1.weight"), gact, R, 4 * D, D, bias=self.wf(p + "mlp.linear1.bias"),
                     act=ops.ACT_GELU, pre_out=u)
            hn = self._new(R, D)
            ops.gemm(gact, self.w(p + "mlp.linear2.weight"), hn, R, D, 4 * D, bias=self.wf(p + "mlp.linear2.bias"), res=ht)
            sv["blocks"].append(dict(hl=hl, st1=st1, s1=s1, qkv=qkv, a=a, probs=probs, ht=ht, st2=st2, s2=s2, u=u, g=gact))
            hl = hn
        sf, stf = self._sln_fwd(hl, x, "sln_norm.", R)
        zr = self._new(R, T * C, dtype=F32)                                     # per sample: flat [T][T*C] == [C][T][T]
        ops.gemm(sf, self.w("w_out.0.weight"), zr, R, T * C, D, bias=self.wf("w_out.0.bias"))
        z = self._new(B * T * T, C, dtype=F32)
        call("transpose", zr, z, B, C, T * T, 1, 1)                            # [B][C][T*T] -> [B][T*T][C]
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
        """dz: (B*T*T, C) fp32 token-major.  Accumulates all parameter gradients into self.grad.
        on_layer_done(i): called once the gradients of encoder block i (and of everything registered after it) are complete."""
        B, T, D, L, H, C = sv["B"], self.T, self.D, self.L, self.H, self.C
        Wd, Wp, Q3, dh = self.Wd, self.Wp, self.Q3, self.dh
        R = B * T
        x = sv["x"]
        sp = ops.auto_splits
        dzr32 = self._new(R, T * C, dtype=F32)
        call("transpose", dz.contiguous(), dzr32, B, T * T, C, 1, 1)          # back to the reference's [C][T][T] flattening
        dzr = self._new(R, T * C)
        call("cast_f32_bf16", dzr32, dzr, R * T * C)
        dx_acc = torch.zeros(R, D, device=self.dev, dtype=F32)
        ops.linear_wgrad(dzr, sv["sf"], self.g("w_out.0.weight"), R, T * C, D, splits=sp(T * C, D, R))
        call("colsum", dzr, self.g("w_out.0.bias"), R, T * C)
        dsf = self._new(R, D)
        ops.linear_dgrad(dzr, self.w("w_out.0.weight"), dsf, R, T * C, D)
        dhl = self._sln_bwd(dsf, sv["hl_last"], x, sv["stf"], "sln_norm.", R, dx_acc, None)
        scale = float(D) ** -0.5
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
            dht = self._sln_bwd(ds2, bv["ht"], x, bv["st2"], p + "norm2.", R, dx_acc, dhl)
            # ---- attention
            # dW_out[n, k] += sum_m dht[m, n] a[m, k]   (k < Wd; a has row pitch Wp)
            ops.gemm(dht, bv["a"], self.g(p + "attn.w_out.weight"), D, Wd, R, a_mode=ops.MNMAJOR, b_mode=ops.MNMAJOR, a_ld=D,
                     b_ld=Wp, atomic=True, splits=sp(D, Wd, R))
            call("colsum", dht, self.g(p + "attn.w_out.bias"), R, D)
            da = torch.zeros(R, Wp, device=self.dev, dtype=BF16) if Wp != Wd else self._new(R, Wp)
            ops.gemm(dht, self.wout_pad[i], da, R, Wd, D, b_mode=ops.MNMAJOR, b_ld=Wp, ldc=Wp)
            dqkv = torch.zeros(R, Q3, device=self.dev, dtype=BF16) if Q3 != 3 * Wd else self._new(R, Q3)
            call("vitgan_attn_bwd", bv["qkv"], bv["probs"], da, dqkv, B, T, H, dh, Q3, Wp, scale)
            ops.gemm(dqkv, bv["s1"], self.g(p + "attn.to_qkv.weight"), 3 * Wd, D, R, a_mode=ops.MNMAJOR, b_mode=ops.MNMAJOR,
                     a_ld=Q3, b_ld=D, atomic=True, splits=sp(3 * Wd, D, R))
            ds1 = self._new(R, D)
            ops.gemm(dqkv, self.w(p + "attn.to_qkv.weight"), ds1, R, D, 3 * Wd, a_ld=Q3, b_mode=ops.MNMAJOR, b_ld=D)
            dhl = self._sln_bwd(ds1, bv["hl"], x, bv["st1"], p + "norm1.", R, dx_acc, dht)
            if on_layer_done is not None:
                on_layer_done(i)
        # pos_emb1D receives the batch-summed gradient of the first block's input (it was broadcast over B)
        call("colsum", dhl, self.g("pos_emb1D"), B, T * D)
        # x = mlp(noise): only wgrad / bias grad (the prompt embedding needs no gradient)
        dxb = self._new(R, D)
        call("cast_f32_bf16", dx_acc, dxb, R * D)
        ops.linear_wgrad(dxb, sv["nb"], self.g("mlp.weight"), B, T * D, self.IN)
        call("colsum", dxb, self.g("mlp.bias"), B, T * D)
