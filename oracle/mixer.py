"""ORACLE (test infrastructure, not product): CPU fp32 restatement of the reference MLP-Mixer mapper.

Follows /root/reference/mlp_mixer_pytorch.py:
  - Mixer.forward                    mlp_mixer_pytorch.py:82-91
  - MLPMixer (Rearrange, Linear, blocks, final LayerNorm)   :25-38
  - PreNormResidual                  :7-14
  - FeedForward (token-mix uses Conv1d k=1 over the token axis, channel-mix uses Linear)  :16-23
Parameters are taken from a state_dict with the reference's key names, so the same
weights drive the reference module, this oracle and the CUDA path.

Pinned: tests/test_oracle_golden.py checks this file against outputs AND parameter gradients of the real
reference module (tests/golden/make_golden.py imports /root/reference/mlp_mixer_pytorch.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import torch
import torch.nn.functional as F


def mixer_depth(sd):
    d = 0
    while "mixer.%d.0.norm.weight" % (d + 2) in sd:
        d += 1
    return d


def mixer_forward(sd, x, image_size, channels):
    """x: (B, input_dim) fp32 -> z: (B, channels, S, S) (permuted view, like the reference)."""
    S, C = image_size, channels
    T = S * S
    B = x.shape[0]
    h = F.linear(x, sd["proj.weight"], sd["proj.bias"])            # :85
    h = h.view(B, C, T).transpose(1, 2)                            # :86 + Rearrange 'b c h w -> b (h w) c' (:31, p=1)
    h = F.linear(h, sd["mixer.1.weight"], sd["mixer.1.bias"])      # :32
    depth = mixer_depth(sd)
    for i in range(2, depth + 2):
        p = "mixer.%d." % i
        # token mixing: PreNormResidual(FeedForward(num_patches, dense=Conv1d k=1))  (:34)
        n = F.layer_norm(h, (h.shape[-1],), sd[p + "0.norm.weight"], sd[p + "0.norm.bias"])
        u = torch.einsum("jt,btd->bjd", sd[p + "0.fn.0.weight"][:, :, 0], n) + sd[p + "0.fn.0.bias"][None, :, None]
        u = F.gelu(u)
        v = torch.einsum("tj,bjd->btd", sd[p + "0.fn.3.weight"][:, :, 0], u) + sd[p + "0.fn.3.bias"][None, :, None]
        h = v + h
        # channel mixing: PreNormResidual(FeedForward(dim))  (:35)
        n = F.layer_norm(h, (h.shape[-1],), sd[p + "1.norm.weight"], sd[p + "1.norm.bias"])
        u = F.gelu(F.linear(n, sd[p + "1.fn.0.weight"], sd[p + "1.fn.0.bias"]))
        h = F.linear(u, sd[p + "1.fn.3.weight"], sd[p + "1.fn.3.bias"]) + h
    q = "mixer.%d." % (depth + 2)
    h = F.layer_norm(h, (h.shape[-1],), sd[q + "weight"], sd[q + "bias"])   # :37
    h = F.linear(h, sd["final_proj.weight"], sd["final_proj.bias"])          # :88
    return h.view(B, S, S, C).permute(0, 3, 1, 2)                            # :89-90


def init_mixer_state_dict(input_dim, image_size, channels, dim, depth, seed=0, dtype=torch.float32):
    """Random-init parameters with the reference's shapes and PyTorch-default-like scales (deterministic;
    does not depend on the reference being importable)."""
    g = torch.Generator().manual_seed(seed)
    T = image_size * image_size

    def lin(out_f, in_f, conv=False):
        bound = 1.0 / (in_f ** 0.5)
        w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
        b = (torch.rand(out_f, generator=g) * 2 - 1) * bound
        return (w[:, :, None] if conv else w).to(dtype), b.to(dtype)

    sd = {}
    sd["proj.weight"], sd["proj.bias"] = lin(T * channels, input_dim)
    sd["mixer.1.weight"], sd["mixer.1.bias"] = lin(dim, channels)
    for i in range(2, depth + 2):
        p = "mixer.%d." % i
        sd[p + "0.norm.weight"] = torch.ones(dim, dtype=dtype) + 0.1 * torch.randn(dim, generator=g)
        sd[p + "0.norm.bias"] = 0.1 * torch.randn(dim, generator=g)
        sd[p + "0.fn.0.weight"], sd[p + "0.fn.0.bias"] = lin(4 * T, T, conv=True)
        sd[p + "0.fn.3.weight"], sd[p + "0.fn.3.bias"] = lin(T, 4 * T, conv=True)
        sd[p + "1.norm.weight"] = torch.ones(dim, dtype=dtype) + 0.1 * torch.randn(dim, generator=g)
        sd[p + "1.norm.bias"] = 0.1 * torch.randn(dim, generator=g)
        sd[p + "1.fn.0.weight"], sd[p + "1.fn.0.bias"] = lin(4 * dim, dim)
        sd[p + "1.fn.3.weight"], sd[p + "1.fn.3.bias"] = lin(dim, 4 * dim)
    q = "mixer.%d." % (depth + 2)
    sd[q + "weight"] = torch.ones(dim, dtype=dtype) + 0.1 * torch.randn(dim, generator=g)
    sd[q + "bias"] = 0.1 * torch.randn(dim, generator=g)
    sd["final_proj.weight"], sd["final_proj.bias"] = lin(channels, dim)
    return sd
