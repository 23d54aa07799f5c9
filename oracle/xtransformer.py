"""ORACLE (test infrastructure, not product): CPU fp32 restatement of the reference X-transformer mapper.

`XTransformer.forward` itself is in-tree (transformer.py:28-46) but its arithmetic lives in `x-transformers==0.19.1`
(requirements.txt:20), which is absent from /root/reference and from this image: PARITY UNPINNED by the package itself.  The stack (positions,
causal pre-LN blocks, final norm) is cross-checked against an independent implementation of the same architecture,
transformers' GPT2Model with mapped random weights (output + input gradient, tests/test_oracle_golden.py).  Restated from the
package's published architecture: ContinuousTransformerWrapper(project_in Linear, learned absolute
positional embedding, Decoder = causal pre-LayerNorm AttentionLayers alternating Attention (to_q/to_k/to_v without bias,
dim_head 64, scale 64**-0.5, causal mask, to_out with bias) and FeedForward (Linear, exact GELU, Linear, mult 4), final
LayerNorm, project_out Linear).  Uncertain detail, chosen and documented: the absolute positional embedding is added
unscaled (later package versions scale it by dim**-0.5).
Keys follow the package's module names so a real checkpoint's state_dict would load:
  proj.*, transformer.project_in.*, transformer.pos_emb.emb.weight, transformer.attn_layers.layers.{i}.0.{weight,bias} (norm),
  ...layers.{2j}.1.{to_q,to_k,to_v}.weight, ...to_out.{weight,bias}, ...layers.{2j+1}.1.net.0.0.*, ...net.2.*,
  transformer.norm.*, transformer.project_out.*
build_model passes initial_proj=True, add_input=False (main.py:497-498) -> the `proj` path of transformer.py:30-32.
"""
import torch
import torch.nn.functional as F


def xt_depth(sd):
    n = 0
    while "transformer.attn_layers.layers.%d.0.weight" % (2 * n) in sd:
        n += 1
    return n


def xtransformer_forward(sd, x, image_size, channels, heads):
    B = x.shape[0]
    T = image_size * image_size
    dim = sd["transformer.project_in.weight"].shape[0]
    h = F.linear(x, sd["proj.weight"], sd["proj.bias"]).view(B, T, dim)                    # transformer.py:30-32
    h = F.linear(h, sd["transformer.project_in.weight"], sd["transformer.project_in.bias"])
    h = h + sd["transformer.pos_emb.emb.weight"][:T][None]
    dh = 64
    mask = torch.ones(T, T, dtype=torch.bool, device=x.device).triu_(1)
    for j in range(xt_depth(sd)):
        p = "transformer.attn_layers.layers.%d." % (2 * j)
        n = F.layer_norm(h, (dim,), sd[p + "0.weight"], sd[p + "0.bias"])
        q = F.linear(n, sd[p + "1.to_q.weight"]).view(B, T, heads, dh).transpose(1, 2)
        k = F.linear(n, sd[p + "1.to_k.weight"]).view(B, T, heads, dh).transpose(1, 2)
        v = F.linear(n, sd[p + "1.to_v.weight"]).view(B, T, heads, dh).transpose(1, 2)
        dots = (q @ k.transpose(-1, -2)) * dh ** -0.5
        dots = dots.masked_fill(mask, -torch.finfo(dots.dtype).max)
        o = (torch.softmax(dots, dim=-1) @ v).transpose(1, 2).reshape(B, T, heads * dh)
        h = F.linear(o, sd[p + "1.to_out.weight"], sd[p + "1.to_out.bias"]) + h
        p = "transformer.attn_layers.layers.%d." % (2 * j + 1)
        n = F.layer_norm(h, (dim,), sd[p + "0.weight"], sd[p + "0.bias"])
        u = F.gelu(F.linear(n, sd[p + "1.net.0.0.weight"], sd[p + "1.net.0.0.bias"]))
        h = F.linear(u, sd[p + "1.net.2.weight"], sd[p + "1.net.2.bias"]) + h
    h = F.layer_norm(h, (dim,), sd["transformer.norm.weight"], sd["transformer.norm.bias"])
    h = F.linear(h, sd["transformer.project_out.weight"], sd["transformer.project_out.bias"])
    return h.view(B, image_size, image_size, channels).permute(0, 3, 1, 2)                 # transformer.py:44-46
