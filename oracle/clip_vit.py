"""ORACLE (test infrastructure, not product): CPU fp32 restatement of the CLIP ViT image encoder.

The arithmetic of `perceptor.encode_image` (call site main.py:799) lives in `clip-anytorch==2.2.0`
(requirements.txt:3; absent).  Its architecture has an in-tree, line-for-line twin that IS importable:
/root/reference/cloob.py:170-255 (LayerNorm, QuickGELU, ResidualAttentionBlock, Transformer,
VisualTransformer).  This file follows that twin and is PINNED against it: tests/golden/make_golden.py runs
cloob.VisualTransformer on seeded inputs and tests/test_oracle_golden.py compares outputs and input-gradients.

  conv1 patch embed (no bias)           cloob.py:224,237-239
  class token + positional embedding    cloob.py:240-243
  ln_pre / resblocks / ln_post / proj   cloob.py:244-253
  ResidualAttentionBlock                cloob.py:184-205   (nn.MultiheadAttention, heads = width/64)
  QuickGELU x*sigmoid(1.702x)           cloob.py:179-181   (OpenCLIP ViT-B-32 uses exact GELU: act="gelu")
"""
import torch
import torch.nn.functional as F

VIT_B32 = dict(input_resolution=224, patch_size=32, width=768, layers=12, heads=12, output_dim=512)


def _ln(x, sd, p):
    return F.layer_norm(x.float(), (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps=1e-5)


def _act(x, act):
    return x * torch.sigmoid(1.702 * x) if act == "quick_gelu" else F.gelu(x)


def encode_image(sd, x, cfg=VIT_B32, act="quick_gelu"):
    """x: (N, 3, R, R) normalised image -> (N, output_dim)."""
    W, Hh = cfg["width"], cfg["heads"]
    dh = W // Hh
    h = F.conv2d(x, sd["conv1.weight"], stride=cfg["patch_size"])            # (N, W, g, g)
    N = h.shape[0]
    h = h.reshape(N, W, -1).permute(0, 2, 1)                                 # (N, g*g, W)
    cls = sd["class_embedding"].expand(N, 1, W)
    h = torch.cat([cls, h], dim=1) + sd["positional_embedding"]
    h = _ln(h, sd, "ln_pre")
    T = h.shape[1]
    for l in range(cfg["layers"]):
        p = "transformer.resblocks.%d." % l
        n = _ln(h, sd, p + "ln_1")
        qkv = F.linear(n, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
        q, k, v = qkv.split(W, dim=-1)
        q = q.reshape(N, T, Hh, dh).transpose(1, 2)
        k = k.reshape(N, T, Hh, dh).transpose(1, 2)
        v = v.reshape(N, T, Hh, dh).transpose(1, 2)
        a = torch.softmax((q @ k.transpose(-1, -2)) * dh ** -0.5, dim=-1)
        o = (a @ v).transpose(1, 2).reshape(N, T, W)
        h = h + F.linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
        n = _ln(h, sd, p + "ln_2")
        u = _act(F.linear(n, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]), act)
        h = h + F.linear(u, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
    h = _ln(h[:, 0, :], sd, "ln_post")
    return h @ sd["proj"]


def init_clip_state_dict(cfg=VIT_B32, seed=0):
    """Deterministic random init with the CLIP visual state_dict key names/shapes."""
    g = torch.Generator().manual_seed(seed)
    W, P = cfg["width"], cfg["patch_size"]
    T = (cfg["input_resolution"] // P) ** 2 + 1
    scale = W ** -0.5
    sd = {}

    def rn(*s, std=1.0):
        return torch.randn(*s, generator=g) * std

    def ln(name):
        sd[name + ".weight"] = 1.0 + 0.1 * rn(W)
        sd[name + ".bias"] = 0.1 * rn(W)

    sd["conv1.weight"] = rn(W, 3, P, P, std=(3 * P * P) ** -0.5)
    sd["class_embedding"] = rn(W, std=scale)
    sd["positional_embedding"] = rn(T, W, std=scale)
    sd["proj"] = rn(W, cfg["output_dim"], std=scale)
    ln("ln_pre")
    ln("ln_post")
    for l in range(cfg["layers"]):
        p = "transformer.resblocks.%d." % l
        sd[p + "attn.in_proj_weight"] = rn(3 * W, W, std=scale)
        sd[p + "attn.in_proj_bias"] = rn(3 * W, std=0.02)
        sd[p + "attn.out_proj.weight"] = rn(W, W, std=scale)
        sd[p + "attn.out_proj.bias"] = rn(W, std=0.02)
        ln(p + "ln_1")
        ln(p + "ln_2")
        sd[p + "mlp.c_fc.weight"] = rn(4 * W, W, std=scale)
        sd[p + "mlp.c_fc.bias"] = rn(4 * W, std=0.02)
        sd[p + "mlp.c_proj.weight"] = rn(W, 4 * W, std=(4 * W) ** -0.5)
        sd[p + "mlp.c_proj.bias"] = rn(W, std=0.02)
    return sd
