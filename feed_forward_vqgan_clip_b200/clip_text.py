"""CLIP text encoder (frozen), forward only, B200-native — `perceptor.encode_text(tokens)` of main.py:733,1035.

Only used when the dataset holds integer token ids (main.py:733 `inp.dtype == torch.long`); the benchmark path feeds
pre-computed embeddings.  Parameter containers carry OpenAI-CLIP's key names (token_embedding.weight,
positional_embedding, transformer.resblocks.N.*, ln_final.*, text_projection) so real checkpoints load.
Follows the in-tree twin cloob.TextTransformer (cloob.py:258-323): token + positional embedding, causal pre-LN
transformer (QuickGELU MLP), ln_final, features at the EOT token (arg-max token id), @ text_projection.
Arithmetic: tcgen05 GEMMs (projections, per-head Q.K^T and P.V through 4-D batched tensor maps), causal softmax,
LayerNorm — all in libffvc_sm100.so.
"""
import torch
from torch import nn

from . import ops
from .clip_vit import _Block
from .ops import BF16, F32, call

TEXT_B32 = dict(embed_dim=512, context_length=77, vocab_size=49408, transformer_width=512, transformer_heads=8,
                transformer_layers=12)


class TextTransformer(nn.Module):
    def __init__(self, embed_dim=512, context_length=77, vocab_size=49408, transformer_width=512, transformer_heads=8,
                 transformer_layers=12, act="quick_gelu"):
        super().__init__()
        if transformer_width // transformer_heads != 64:
            raise NotImplementedError("head_dim must be 64")
        self.cfg = dict(embed_dim=embed_dim, context_length=context_length, vocab_size=vocab_size,
                        transformer_width=transformer_width, transformer_heads=transformer_heads,
                        transformer_layers=transformer_layers)
        self.act = act
        self.transformer = nn.Module()
        self.transformer.resblocks = nn.Sequential(*[_Block(transformer_width) for _ in range(transformer_layers)])
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width).normal_(std=0.01))
        self.ln_final = nn.LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim).normal_(std=transformer_width ** -0.5))
        self._engine = None

    def engine(self):
        if self._engine is None or self._engine.ptr != self.text_projection.data_ptr():
            self._engine = TextEngine(self)
        return self._engine

    @torch.no_grad()
    def forward(self, text):
        return self.engine().forward(text)


class TextEngine:
    TP = 80          # 77 context positions padded to a 16-byte-multiple row pitch

    def __init__(self, m):
        cfg = m.cfg
        self.dev = m.text_projection.device
        ops.require_cuda(self.dev, "the CLIP text encoder")
        self.ptr = m.text_projection.data_ptr()
        self.W, self.L, self.Hh, self.E, self.T = (cfg["transformer_width"], cfg["transformer_layers"],
                                                   cfg["transformer_heads"], cfg["embed_dim"], cfg["context_length"])
        if self.T > self.TP:
            raise NotImplementedError("context_length > 80")
        self.act = ops.ACT_QUICKGELU if m.act == "quick_gelu" else ops.ACT_GELU
        sd = {k: v.detach() for k, v in m.state_dict().items()}
        self.sd32 = {k: v.float().contiguous() for k, v in sd.items() if v.dim() == 1}
        self.emb = sd["token_embedding.weight"].float().contiguous()
        self.pos = sd["positional_embedding"].float().contiguous()
        self.w = {"proj": sd["text_projection"].contiguous().to(BF16)}
        for l in range(self.L):
            p = "transformer.resblocks.%d." % l
            self.w[p + "in"] = sd[p + "attn.in_proj_weight"].contiguous().to(BF16)
            self.w[p + "out"] = sd[p + "attn.out_proj.weight"].contiguous().to(BF16)
            self.w[p + "fc"] = sd[p + "mlp.c_fc.weight"].contiguous().to(BF16)
            self.w[p + "pj"] = sd[p + "mlp.c_proj.weight"].contiguous().to(BF16)

    def _new(self, *shape, dtype=BF16):
        return torch.empty(*shape, device=self.dev, dtype=dtype)

    def _ln(self, x, name, rows):
        y = self._new(rows, self.W)
        call("layernorm_fwd", x, self.sd32[name + ".weight"], self.sd32[name + ".bias"], y, None, None, rows, self.W, 1e-5)
        return y

    def forward(self, text):
        """text: (B, T) int64 token ids on the GPU -> (B, embed_dim) fp32."""
        B, T = text.shape
        assert T == self.T and text.dtype == torch.long
        W, L, Hh, E, TP = self.W, self.L, self.Hh, self.E, self.TP
        M = B * T
        text = text.contiguous()
        h = self._new(M, W)
        call("embed_tokens", text, self.emb, self.pos, h, M, T, W)
        for l in range(L):
            p = "transformer.resblocks.%d." % l
            n1 = self._ln(h, p + "ln_1", M)
            qkv = self._new(M, 3 * W)
            ops.gemm(n1, self.w[p + "in"], qkv, M, 3 * W, W, bias=self.sd32[p + "attn.in_proj_bias"])
            S = self._new(B, Hh, T, TP, dtype=F32)
            ops.gemm(qkv, qkv, S, T, T, 64, a_ld=3 * W, b_ld=3 * W, b_off=W, a_role=ops.ROLE_OUT, b_role=ops.ROLE_OUT,
                     batch=B * Hh, batch_inner=Hh, a_bs=T * 3 * W, b_bs=T * 3 * W, a_bs_in=64, b_bs_in=64, ldc=TP,
                     out_bs=Hh * T * TP, out_bs_in=T * TP, alpha=0.125, block_n=128)
            P = self._new(B, Hh, T, TP)
            call("softmax_causal_fwd", S, P, B * Hh * T, T, TP)
            a = self._new(M, W)
            ops.gemm(P, qkv, a, T, 64, T, a_ld=TP, b_mode=ops.MNMAJOR, b_ld=3 * W, b_off=2 * W, a_role=ops.ROLE_OUT,
                     b_role=ops.ROLE_OUT, batch=B * Hh, batch_inner=Hh, a_bs=Hh * T * TP, a_bs_in=T * TP, b_bs=T * 3 * W,
                     b_bs_in=64, ldc=W, out_bs=T * W, out_bs_in=64, block_n=64)
            h2 = self._new(M, W)
            ops.gemm(a, self.w[p + "out"], h2, M, W, W, bias=self.sd32[p + "attn.out_proj.bias"], res=h)
            n2 = self._ln(h2, p + "ln_2", M)
            g = self._new(M, 4 * W)
            ops.gemm(n2, self.w[p + "fc"], g, M, 4 * W, W, bias=self.sd32[p + "mlp.c_fc.bias"], act=self.act)
            h = self._new(M, W)
            ops.gemm(g, self.w[p + "pj"], h, M, W, 4 * W, bias=self.sd32[p + "mlp.c_proj.bias"], res=h2)
        eot = text.argmax(dim=-1)                                   # position of the EOT token (plumbing on B integers)
        xe = self._new(B, W)
        call("gather_rows", h, eot, xe, B, T, W)
        ne = self._ln(xe, "ln_final", B)                           # LayerNorm is per row: normalising only the EOT rows is exact
        out = self._new(B, E, dtype=F32)
        ops.gemm(ne, self.w["proj"], out, B, E, W, b_mode=ops.MNMAJOR, b_ld=E)
        return out
