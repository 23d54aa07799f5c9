"""ORACLE (test infrastructure, not product): CPU fp32 restatement of the reference VitGAN mapper.

Follows /root/reference/vitgan.py:
  - Generator.forward                      vitgan.py:254-260   (T = initialize_size*8 tokens, raw view to (B,C,T,T))
  - GTransformerEncoder / GEncoderBlock    vitgan.py:120-164   (hl starts as pos_emb1D with no batch dim; x is carried unchanged)
  - SLN: gamma * w * LN(hl) + beta * w     vitgan.py:8-21      (scalar gamma / beta)
  - Attention                              vitgan.py:44-97     (no-bias to_qkv, split 'b t (d k h) -> k b h t d',
                                                                scale = dim ** -0.5 (FULL model dim), w_out with bias)
  - MLP (Linear, exact GELU, Linear)       vitgan.py:24-41
  - SimpleGenerator.forward                vitgan.py:262-305   (T = size*size tokens, hl0 = permuted inp(noise) + pos_emb1D)
Pinned by tests/test_oracle_golden.py against outputs + parameter gradients of the real reference module
(tests/golden/vitgan.pt and simple_vitgan.pt, produced by tests/golden/make_golden.py importing /root/reference/vitgan.py)."""
import torch
import torch.nn.functional as F


def _sln(sd, p, hl, w):
    n = F.layer_norm(hl, (hl.shape[-1],), sd[p + "ln.weight"], sd[p + "ln.bias"])
    return sd[p + "gamma"] * w * n + sd[p + "beta"] * w


def vitgan_blocks(sd):
    n = 0
    while "Transformer_Encoder.blocks.%d.attn.to_qkv.weight" % n in sd:
        n += 1
    return n


def _encoder(sd, hl, x, num_heads):
    """GTransformerEncoder (vitgan.py:138-164): hl is transformed, x (the modulation signal) is carried unchanged."""
    B, T, D = x.shape
    for i in range(vitgan_blocks(sd)):
        p = "Transformer_Encoder.blocks.%d." % i
        s = _sln(sd, p + "norm1.", hl, x)
        qkv = F.linear(s, sd[p + "attn.to_qkv.weight"])
        Wd = qkv.shape[-1] // 3
        dh = Wd // num_heads
        qkv = qkv.view(B, T, dh, 3, num_heads).permute(3, 0, 4, 1, 2)          # 'b t (d k h) -> k b h t d'
        q, k, v = qkv[0], qkv[1], qkv[2]
        a = torch.softmax(torch.einsum("bhid,bhjd->bhij", q, k) * D ** -0.5, dim=-1)
        r = torch.einsum("bhij,bhjd->bhid", a, v).permute(0, 2, 1, 3).reshape(B, T, Wd)   # 'b h t d -> b t (h d)'
        hl_temp = F.linear(r, sd[p + "attn.w_out.weight"], sd[p + "attn.w_out.bias"]) + hl
        s2 = _sln(sd, p + "norm2.", hl_temp, x)
        u = F.gelu(F.linear(s2, sd[p + "mlp.linear1.weight"], sd[p + "mlp.linear1.bias"]))
        hl = F.linear(u, sd[p + "mlp.linear2.weight"], sd[p + "mlp.linear2.bias"]) + hl_temp
    return hl


def vitgan_forward(sd, noise, out_channels, num_heads):
    T, D = sd["pos_emb1D"].shape
    B = noise.shape[0]
    x = F.linear(noise, sd["mlp.weight"], sd["mlp.bias"]).view(B, T, D)
    hl = _encoder(sd, sd["pos_emb1D"], x, num_heads)
    y = _sln(sd, "sln_norm.", hl, x)
    y = F.linear(y, sd["w_out.0.weight"], sd["w_out.0.bias"])
    return y.reshape(B, out_channels, T, T)


def simple_vitgan_forward(sd, noise, out_channels, num_heads):
    """SimpleGenerator.forward (vitgan.py:296-305; build_model model_type 'simple_vitgan', main.py:469-478): T = size*size
    tokens; the encoder's hl starts as inp(noise) viewed (B, dim, T), permuted to (B, T, dim), plus pos_emb1D; w_out maps
    dim -> out_channels per token and the result is permuted to (B, C, size, size)."""
    T, D = sd["pos_emb1D"].shape
    B = noise.shape[0]
    S = int(round(T ** 0.5))
    inp = F.linear(noise, sd["inp.weight"], sd["inp.bias"])
    x = F.linear(noise, sd["mlp.weight"], sd["mlp.bias"]).view(B, T, D)
    hl = inp.view(B, D, T).permute(0, 2, 1) + sd["pos_emb1D"]
    hl = _encoder(sd, hl, x, num_heads)
    y = _sln(sd, "sln_norm.", hl, x)
    y = F.linear(y, sd["w_out.0.weight"], sd["w_out.0.bias"])
    return y.view(B, S, S, out_channels).permute(0, 3, 1, 2)
