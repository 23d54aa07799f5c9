"""MLP-Mixer mapper (prompt embedding -> VQGAN latent grid), B200-native.

Drop-in for the reference's `Mixer` (mlp_mixer_pytorch.py:70-91): same constructor signature, same
`state_dict()` keys / shapes, same construction order (so `torch.manual_seed(s)` yields the
reference's initial weights), `forward(x: (B, input_dim)) -> (B, C, S, S)` permuted view.  The nn.Modules
below are parameter containers only: all arithmetic runs in libffvc_sm100.so (tcgen05 GEMMs with fused
bias / GELU / residual epilogues, warp-shuffle LayerNorm), forward AND backward (dgrad + wgrad), through
`MixerEngine`.  fp32 master parameters live in one flat arena (one NCCL all-reduce / one fused Adam launch);
the GEMMs read a bf16 shadow of it.
"""
import torch
from torch import nn

from . import ops
from .ops import BF16, F32, call


class _PreNormResidual(nn.Module):          # parameter container mirroring mlp_mixer_pytorch.py:7-14
    def __init__(self, dim, fn):
        super().__init__()
        self.fn = fn
        self.norm = nn.LayerNorm(dim)


def _feed_forward(dim, expansion_factor, dropout, dense):   # mirrors mlp_mixer_pytorch.py:16-23
    return nn.Sequential(dense(dim, dim * expansion_factor), nn.GELU(), nn.Dropout(dropout),
                         dense(dim * expansion_factor, dim), nn.Dropout(dropout))


def _conv1d_k1(i, o):
    return nn.Conv1d(i, o, kernel_size=1)


class Mixer(nn.Module):
    def __init__(self, input_dim, image_size, channels, patch_size, dim, depth, expansion_factor=4, dropout=0.):
        super().__init__()
        if patch_size != 1:
            raise NotImplementedError("the reference only ever builds the mapper with patch_size=1 (main.py:485)")
        if dropout != 0:
            raise NotImplementedError("dropout > 0 is not implemented on the B200 path (configs use dropout: 0)")
        if expansion_factor != 4:
            raise NotImplementedError("expansion_factor != 4")
        self.input_dim, self.image_size, self.channels, self.dim, self.depth = input_dim, image_size, channels, dim, depth
        T = image_size * image_size
        # same construction order as mlp_mixer_pytorch.py:25-38,73-78 (identical RNG consumption)
        self.mixer = nn.Sequential(
            nn.Identity(),                                      # stands in for the parameter-free Rearrange
            nn.Linear(channels, dim),
            *[nn.Sequential(_PreNormResidual(dim, _feed_forward(T, 4, dropout, _conv1d_k1)),
                            _PreNormResidual(dim, _feed_forward(dim, 4, dropout, nn.Linear)))
              for _ in range(depth)],
            nn.LayerNorm(dim),
        )
        self.proj = nn.Linear(input_dim, T * channels)
        self.final_proj = nn.Linear(dim, channels)
        self._engine = None

    # ------------------------------------------------------------------ engine plumbing
    def engine(self):
        if self._engine is None or not self._engine.valid():
            self._engine = MixerEngine(self)
        return self._engine

    def forward(self, x):
        return _MixerFn.apply(self, x, *list(self.parameters()))


class _MixerFn(torch.autograd.Function):
    """autograd bridge: parameters are listed as inputs so autograd sees the dependency; their gradients are
    written straight into the flat gradient arena by the engine (returned as views)."""

    @staticmethod
    def forward(ctx, module, x, *params):
        eng = module.engine()
        z, saved = eng.forward(x)
        ctx.eng, ctx.saved, ctx.n = eng, saved, len(params)
        B, S, C = x.shape[0], module.image_size, module.channels
        return z.view(B, S, S, C).permute(0, 3, 1, 2)          # like mlp_mixer_pytorch.py:89-90

    @staticmethod
    def backward(ctx, dz):
        eng = ctx.eng
        B = dz.shape[0]
        dz_tok = dz.permute(0, 2, 3, 1).contiguous().view(B * eng.T, eng.C).float()
        eng.zero_grad_arena()
        eng.backward(ctx.saved, dz_tok)
        return (None, None) + tuple(eng.grad_views)


class MixerEngine:
    """Explicit forward / backward of the mixer on flat arenas.  Layout: activations bf16 token-major
    [B*T, D]; the token-mixing GEMMs contract over the token axis by reading the activations MN-major."""

    def __init__(self, m):
        self.m = m
        self.params = list(m.parameters())
        dev = self.params[0].device
        ops.require_cuda(dev, "Mixer")
        self.dev = dev
        self.S, self.C, self.D, self.L, self.IN = m.image_size, m.channels, m.dim, m.depth, m.input_dim
        self.T = self.S * self.S
        if self.IN % 8 or self.C % 8 or self.D % 8 or self.T % 8:
            raise NotImplementedError("dims must be multiples of 8 (TMA 16-byte stride rule)")
        # ---- flat fp32 master arena, parameters re-pointed at views of it
        sizes = [p.numel() for p in self.params]
        offs, o = [], 0
        for n in sizes:
            offs.append(o)
            o += (n + 7) // 8 * 8          # keep every tensor 32-byte aligned inside the arenas
        self.total = o
        self.arena = torch.zeros(o, device=dev, dtype=F32)
        self.grad = torch.zeros(o, device=dev, dtype=F32)
        self.shadow = torch.empty(o, device=dev, dtype=BF16)
        self.grad_views = []
        self._ptrs = []
        for p, off, n in zip(self.params, offs, sizes):
            view = self.arena[off:off + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            self.grad_views.append(self.grad[off:off + n].view(p.shape))
            self._ptrs.append(p.data_ptr())
        self.offs = dict(zip([n for n, _ in m.named_parameters()], offs))
        self.shapes = dict((n, p.shape) for n, p in m.named_parameters())
        self._shadow_version = None
        self.ext_shadow_fresh = False       # set by FusedAdam when it refreshed the shadow itself
        from . import _lib
        # scratch of the fused LayerNorm backward (per-CTA partial sums of dgamma / dbeta / bias gradients)
        self.ln_ws = torch.empty(int(_lib.load().ffvc_layernorm_bwd_ws_bytes(self.D, self.T)) // 4, device=dev, dtype=F32)

    def valid(self):
        return all(p.data_ptr() == q for p, q in zip(self.params, self._ptrs))

    # ---- views into the arenas by state_dict name
    def w(self, name):
        off = self.offs[name]
        n = 1
        for s in self.shapes[name]:
            n *= s
        return self.shadow[off:off + n]

    def wf(self, name):
        off = self.offs[name]
        n = 1
        for s in self.shapes[name]:
            n *= s
        return self.arena[off:off + n]

    def g(self, name):
        off = self.offs[name]
        n = 1
        for s in self.shapes[name]:
            n *= s
        return self.grad[off:off + n]

    def refresh_shadow(self):
        ver = tuple(p._version for p in self.params)
        if (self.ext_shadow_fresh and ver == getattr(self, "_adam_ver", ver)) or ver == self._shadow_version:
            self.ext_shadow_fresh = False
            self._shadow_version = ver
            return
        call("cast_f32_bf16", self.arena, self.shadow, self.total)
        self._shadow_version = ver

    def zero_grad_arena(self):
        self.grad.zero_()

    # ------------------------------------------------------------------ forward
    def forward(self, x):
        """x: (B, IN) fp32 cuda -> z: (B*T, C) fp32 token-major, plus the saved-activation record."""
        self.refresh_shadow()
        B, T, C, D, L, IN = x.shape[0], self.T, self.C, self.D, self.L, self.IN
        dev = self.dev
        R = B * T

        def new(*shape, dtype=BF16):
            return torch.empty(*shape, device=dev, dtype=dtype)

        sv = {"B": B}
        xb = new(B, IN)
        call("cast_f32_bf16", x.contiguous(), xb, B * IN)
        sv["xb"] = xb
        P = new(B, C * T)
        ops.gemm(xb, self.w("proj.weight"), P, B, C * T, IN, bias=self.wf("proj.bias"))
        tok = new(R, C)
        call("transpose", P, tok, B, C, T, 0, 0)                       # [B][C][T] -> [B][T][C]
        sv["tok"] = tok
        H = new(R, D)
        ops.gemm(tok, self.w("mixer.1.weight"), H, R, D, C, bias=self.wf("mixer.1.bias"))
        layers = []
        for i in range(2, L + 2):
            p = "mixer.%d." % i
            lv = {"Ha": H}
            n1, mu1, rs1 = new(R, D), new(R, dtype=F32), new(R, dtype=F32)
            call("layernorm_fwd", H, self.wf(p + "0.norm.weight"), self.wf(p + "0.norm.bias"), n1, mu1, rs1, R, D, 1e-5)
            U1, G1 = new(B, 4 * T, D), new(B, 4 * T, D)
            # token mixing 1: U1[b,j,d] = sum_t Wt1[j,t] n1[b,t,d] + b[j]; G1 = gelu(U1)
            ops.gemm(self.w(p + "0.fn.0.weight"), n1, G1, 4 * T, D, T, b_mode=ops.MNMAJOR, b_ld=D, b_role=ops.ROLE_OUT,
                     b_bs=T * D, batch=B, out_bs=4 * T * D, bias=self.wf(p + "0.fn.0.bias"), bias_mode=2,
                     act=ops.ACT_GELU, pre_out=U1)
            Hb = new(R, D)
            ops.gemm(self.w(p + "0.fn.3.weight"), G1, Hb, T, D, 4 * T, b_mode=ops.MNMAJOR, b_ld=D, b_role=ops.ROLE_OUT,
                     b_bs=4 * T * D, batch=B, out_bs=T * D, bias=self.wf(p + "0.fn.3.bias"), bias_mode=2, res=H)
            n2, mu2, rs2 = new(R, D), new(R, dtype=F32), new(R, dtype=F32)
            call("layernorm_fwd", Hb, self.wf(p + "1.norm.weight"), self.wf(p + "1.norm.bias"), n2, mu2, rs2, R, D, 1e-5)
            U2, G2 = new(R, 4 * D), new(R, 4 * D)
            ops.gemm(n2, self.w(p + "1.fn.0.weight"), G2, R, 4 * D, D, bias=self.wf(p + "1.fn.0.bias"),
                     act=ops.ACT_GELU, pre_out=U2)
            Hn = new(R, D)
            ops.gemm(G2, self.w(p + "1.fn.3.weight"), Hn, R, D, 4 * D, bias=self.wf(p + "1.fn.3.bias"), res=Hb)
            lv.update(n1=n1, mu1=mu1, rs1=rs1, U1=U1, G1=G1, Hb=Hb, n2=n2, mu2=mu2, rs2=rs2, U2=U2, G2=G2)
            layers.append(lv)
            H = Hn
        q = "mixer.%d." % (L + 2)
        nf, muf, rsf = new(R, D), new(R, dtype=F32), new(R, dtype=F32)
        call("layernorm_fwd", H, self.wf(q + "weight"), self.wf(q + "bias"), nf, muf, rsf, R, D, 1e-5)
        z = new(R, C, dtype=F32)
        ops.gemm(nf, self.w("final_proj.weight"), z, R, C, D, bias=self.wf("final_proj.bias"))
        sv.update(layers=layers, HL=H, nf=nf, muf=muf, rsf=rsf)
        return z, sv

    # ------------------------------------------------------------------ backward (dgrad + wgrad, into self.grad)
    def layer_starts(self):
        """arena offset of the first parameter of every mixer layer (mixer.2 .. mixer.L+1), ascending"""
        return [min(off for n, off in self.offs.items() if n.startswith("mixer.%d." % i)) for i in range(2, self.L + 2)]

    def late_ranges(self):
        """arena ranges whose gradient is complete only when backward has ended although they are registered after the layers:
        the input projection `proj` (its wgrad / bias sum are the last launches of backward) — see parallel.bucket_slices"""
        out = []
        for n in ("proj.weight", "proj.bias"):
            lo = self.offs[n]
            hi = lo + (self.shapes[n].numel() + 7) // 8 * 8          # slots are padded to 8 elements in the arenas
            if out and out[-1][1] == lo:
                out[-1] = (out[-1][0], hi)
            else:
                out.append((lo, hi))
        return out

    def backward(self, sv, dz, on_layer_done=None):
        """dz: (B*T, C) fp32.  Accumulates every parameter gradient into the flat fp32 arena `self.grad`.
        on_layer_done(k): called right after the gradients of mixer layer k (0-based, last layer first) — and of everything
        registered after it — are complete: the data-parallel step uses it to start bucketed all-reduces during backward."""
        B, T, C, D, L, IN = sv["B"], self.T, self.C, self.D, self.L, self.IN
        dev = self.dev
        R = B * T

        def new(*shape, dtype=BF16):
            return torch.empty(*shape, device=dev, dtype=dtype)

        sp = ops.auto_splits
        dzb = new(R, C)
        call("cast_f32_bf16", dz.contiguous(), dzb, R * C)
        # final_proj
        ops.linear_wgrad(dzb, sv["nf"], self.g("final_proj.weight"), R, C, D, splits=sp(C, D, R))
        # bias-gradient sums do not feed the next GEMM: they run on a side stream next to it (ops.fork / ops.join)
        ops.fork(lambda: call("colsum", dzb, self.g("final_proj.bias"), R, C))
        dnf = new(R, D)
        ops.linear_dgrad(dzb, self.w("final_proj.weight"), dnf, R, C, D)
        q = "mixer.%d." % (L + 2)
        dH = new(R, D)
        # every LayerNorm backward also emits the bias gradient that is a plain sum of its output dx (one fused pass):
        # column sums -> bias of the Linear whose output gradient dx is; per-token row sums -> bias of the token-mixing Conv1d
        call("layernorm_bwd_sums", dnf, sv["HL"], self.wf(q + "weight"), sv["muf"], sv["rsf"], None, dH,
             self.g(q + "weight"), self.g(q + "bias"), self.g("mixer.%d.1.fn.3.bias" % (L + 1)), None, 0, self.ln_ws, R, D)
        for i in range(L + 1, 1, -1):
            p = "mixer.%d." % i
            lv = sv["layers"][i - 2]
            # ---- channel mixing
            ops.linear_wgrad(dH, lv["G2"], self.g(p + "1.fn.3.weight"), R, D, 4 * D, splits=sp(D, 4 * D, R))
            # (bias gradient of 1.fn.3 = colsum(dH): emitted by the LayerNorm backward that produced dH)
            dU2 = new(R, 4 * D)
            ops.linear_dgrad(dH, self.w(p + "1.fn.3.weight"), dU2, R, D, 4 * D, aux=lv["U2"], mul_mode=ops.ACT_GELU)
            ops.fork(lambda t=dU2, q=p: call("colsum", t, self.g(q + "1.fn.0.bias"), R, 4 * D))
            ops.linear_wgrad(dU2, lv["n2"], self.g(p + "1.fn.0.weight"), R, 4 * D, D, splits=sp(4 * D, D, R))
            dn2 = new(R, D)
            ops.linear_dgrad(dU2, self.w(p + "1.fn.0.weight"), dn2, R, 4 * D, D)
            ops.join()                      # before dU2 (read by the forked column sum) goes back to the allocator
            del dU2
            dHb = new(R, D)
            call("layernorm_bwd_sums", dn2, lv["Hb"], self.wf(p + "1.norm.weight"), lv["mu2"], lv["rs2"], dH, dHb,
                 self.g(p + "1.norm.weight"), self.g(p + "1.norm.bias"), None, self.g(p + "0.fn.3.bias"), T, self.ln_ws, R, D)
            # ---- token mixing:  Hb[b] = Wt2 . G1[b] + bt2 + Ha[b]
            seg_splits = sp(T, 4 * T, B * D)
            ops.gemm(dHb, lv["G1"], self.g(p + "0.fn.3.weight"), T, 4 * T, D, a_role=ops.ROLE_SEG, a_bs=T * D,
                     b_role=ops.ROLE_SEG, b_bs=4 * T * D, k_segs=B, splits=seg_splits, atomic=True)
            dU1 = new(B, 4 * T, D)
            ops.gemm(self.w(p + "0.fn.3.weight"), dHb, dU1, 4 * T, D, T, a_mode=ops.MNMAJOR, a_ld=4 * T,
                     b_mode=ops.MNMAJOR, b_ld=D, b_role=ops.ROLE_OUT, b_bs=T * D, batch=B, out_bs=4 * T * D,
                     aux=lv["U1"], mul_mode=ops.ACT_GELU)
            ops.fork(lambda t=dU1, q=p: call("rowsum", t, self.g(q + "0.fn.0.bias"), B, 4 * T, D))
            ops.gemm(dU1, lv["n1"], self.g(p + "0.fn.0.weight"), 4 * T, T, D, a_role=ops.ROLE_SEG, a_bs=4 * T * D,
                     b_role=ops.ROLE_SEG, b_bs=T * D, k_segs=B, splits=sp(4 * T, T, B * D), atomic=True)
            dn1 = new(R, D)
            ops.gemm(self.w(p + "0.fn.0.weight"), dU1, dn1, T, D, 4 * T, a_mode=ops.MNMAJOR, a_ld=T,
                     b_mode=ops.MNMAJOR, b_ld=D, b_role=ops.ROLE_OUT, b_bs=4 * T * D, batch=B, out_bs=T * D)
            ops.join()
            del dU1
            dHa = new(R, D)
            nxt = "mixer.%d.1.fn.3.bias" % (i - 1) if i > 2 else "mixer.1.bias"     # the Linear whose output gradient dHa is
            call("layernorm_bwd_sums", dn1, lv["Ha"], self.wf(p + "0.norm.weight"), lv["mu1"], lv["rs1"], dHb, dHa,
                 self.g(p + "0.norm.weight"), self.g(p + "0.norm.bias"), self.g(nxt), None, 0, self.ln_ws, R, D)
            dH = dHa
            if on_layer_done is not None:
                on_layer_done(i - 2)
        # mixer.1 (Linear C -> D)
        ops.linear_wgrad(dH, sv["tok"], self.g("mixer.1.weight"), R, D, C, splits=sp(D, C, R))
        dtok = new(R, C)
        ops.linear_dgrad(dH, self.w("mixer.1.weight"), dtok, R, D, C)
        dP = new(B, C * T)
        call("transpose", dtok, dP, B, T, C, 0, 0)                    # [B][T][C] -> [B][C][T]
        ops.fork(lambda: call("colsum", dP, self.g("proj.bias"), B, C * T))
        ops.linear_wgrad(dP, sv["xb"], self.g("proj.weight"), B, C * T, IN)
        ops.join()
