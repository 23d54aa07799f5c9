"""ORACLE (test infrastructure, not product): CPU fp32 restatement of the CLIP text encoder, following the in-tree twin
cloob.TextTransformer.forward (cloob.py:312-323) + Transformer / ResidualAttentionBlock with the causal mask of
cloob.py:304-310.  Pinned by tests/golden/clip_text.pt (outputs of the real cloob.TextTransformer)."""
import torch
import torch.nn.functional as F


def encode_text(sd, text, heads, act="quick_gelu"):
    x = sd["token_embedding.weight"][text] + sd["positional_embedding"]
    B, T, W = x.shape
    dh = W // heads
    mask = torch.full((T, T), float("-inf")).triu_(1)
    l = 0
    while "transformer.resblocks.%d.ln_1.weight" % l in sd:
        p = "transformer.resblocks.%d." % l
        n = F.layer_norm(x, (W,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"])
        q, k, v = F.linear(n, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]).split(W, dim=-1)
        q, k, v = (t.reshape(B, T, heads, dh).transpose(1, 2) for t in (q, k, v))
        a = torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5 + mask, dim=-1)
        o = (a @ v).transpose(1, 2).reshape(B, T, W)
        x = x + F.linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
        n = F.layer_norm(x, (W,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"])
        u = F.linear(n, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])
        u = u * torch.sigmoid(1.702 * u) if act == "quick_gelu" else F.gelu(u)
        x = x + F.linear(u, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
        l += 1
    x = F.layer_norm(x, (W,), sd["ln_final.weight"], sd["ln_final.bias"])
    return x[torch.arange(B, device=x.device), text.argmax(dim=-1)] @ sd["text_projection"]
