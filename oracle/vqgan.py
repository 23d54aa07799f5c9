"""ORACLE (test infrastructure, not product): CPU fp32 restatement of the VQGAN pieces on the hot path.

Third-party boundary: the decoder arithmetic lives in `taming-transformers-rom1504==0.0.6` (requirements.txt:2), which
is neither under /root/reference nor installed here, so it is NOT pinned by taming itself.  It is cross-checked against an
independent implementation of the same decoder that IS importable here — transformers' JanusVQVAEDecoder, random weights
mapped onto taming's key names: output and input gradient agree to 1e-5 (tests/test_oracle_golden.py).
This file restates the published architecture (taming/modules/diffusionmodules/model.py:
Decoder, ResnetBlock, AttnBlock, Upsample, Normalize, nonlinearity; taming/models/vqgan.py: VQModel.decode)
for the `vqgan_imagenet_f16_16384` config, with the package's state_dict key names.
What IS pinned against the reference's own code (tests/golden/make_golden.py imports /root/reference/main.py):
`vector_quantize`, `ReplaceGrad`, `ClampWithGrad`, `synth` glue (main.py:105-143) — see `synth` below and
tests/test_oracle_golden.py.

Reference call sites: main.py:87-89 (construction), main.py:141-142 (quantize.embedding.weight, decode).
"""
import torch
import torch.nn.functional as F

F16_16384 = dict(ch=128, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, attn_resolutions=(16,), resolution=256,
                 z_channels=256, out_ch=3, embed_dim=256, n_embed=16384)


def _gn(x, sd, p):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)   # taming Normalize


def _swish(x):
    return x * torch.sigmoid(x)                                                 # taming nonlinearity


def _conv(x, sd, p, pad):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=pad)


def _resblock(x, sd, p):
    h = _conv(_swish(_gn(x, sd, p + ".norm1")), sd, p + ".conv1", 1)
    h = _conv(_swish(_gn(h, sd, p + ".norm2")), sd, p + ".conv2", 1)          # dropout(0) elided
    if (p + ".nin_shortcut.weight") in sd:
        x = _conv(x, sd, p + ".nin_shortcut", 0)
    return x + h


def _attn(x, sd, p):
    h = _gn(x, sd, p + ".norm")
    q, k, v = _conv(h, sd, p + ".q", 0), _conv(h, sd, p + ".k", 0), _conv(h, sd, p + ".v", 0)
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w = torch.bmm(q, k) * (int(c) ** (-0.5))
    w = F.softmax(w, dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(h, sd, p + ".proj_out", 0)


def decoder_layout(cfg):
    """[(i_level, [(cin, cout)]*nblocks, has_attn, has_upsample)] from top (lowest res) to bottom."""
    ch, ch_mult, nrb = cfg["ch"], cfg["ch_mult"], cfg["num_res_blocks"]
    nres = len(ch_mult)
    block_in = ch * ch_mult[-1]
    curr = cfg["resolution"] // 2 ** (nres - 1)
    levels = []
    for i_level in reversed(range(nres)):
        block_out = ch * ch_mult[i_level]
        blocks = []
        for _ in range(nrb + 1):
            blocks.append((block_in, block_out))
            block_in = block_out
        levels.append((i_level, blocks, curr in cfg["attn_resolutions"], i_level != 0))
        if i_level != 0:
            curr *= 2
    return levels


def decode(sd, z_q, cfg=F16_16384):
    """VQModel.decode: z_q (B, embed_dim, S, S) -> image (B, out_ch, 16S, 16S) in ~[-1, 1]."""
    h = _conv(z_q, sd, "post_quant_conv", 0)
    h = _conv(h, sd, "decoder.conv_in", 1)
    h = _resblock(h, sd, "decoder.mid.block_1")
    h = _attn(h, sd, "decoder.mid.attn_1")
    h = _resblock(h, sd, "decoder.mid.block_2")
    for i_level, blocks, has_attn, has_up in decoder_layout(cfg):
        for j in range(len(blocks)):
            h = _resblock(h, sd, "decoder.up.%d.block.%d" % (i_level, j))
            if has_attn:
                h = _attn(h, sd, "decoder.up.%d.attn.%d" % (i_level, j))
        if has_up:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(h, sd, "decoder.up.%d.upsample.conv" % i_level, 1)
    h = _swish(_gn(h, sd, "decoder.norm_out"))
    return _conv(h, sd, "decoder.conv_out", 1)


# ---- glue, restating main.py:105-143 (pinned against the reference functions by the golden test)
class _ReplaceGrad(torch.autograd.Function):      # main.py:105-113
    @staticmethod
    def forward(ctx, x_forward, x_backward):
        ctx.shape = x_backward.shape
        return x_forward

    @staticmethod
    def backward(ctx, g):
        return None, g.sum_to_size(ctx.shape)


class _ClampWithGrad(torch.autograd.Function):    # main.py:118-129
    @staticmethod
    def forward(ctx, x, lo, hi):
        ctx.lo, ctx.hi = lo, hi
        ctx.save_for_backward(x)
        return x.clamp(lo, hi)

    @staticmethod
    def backward(ctx, g):
        x, = ctx.saved_tensors
        return g * (g * (x - x.clamp(ctx.lo, ctx.hi)) >= 0), None, None


clamp_with_grad = _ClampWithGrad.apply


def vector_quantize(x, codebook, force_idx=None):                 # main.py:134-138
    """force_idx (tests only): use these code indices instead of the argmin — lets a parity test compare gradients
    downstream of the quantiser without the few near-tie flips bf16 mapper noise produces."""
    d = x.pow(2).sum(dim=-1, keepdim=True) + codebook.pow(2).sum(dim=1) - 2 * x @ codebook.T
    idx = d.argmin(-1) if force_idx is None else force_idx.view(d.shape[:-1]).long()
    x_q = F.one_hot(idx, codebook.shape[0]).to(d.dtype) @ codebook
    return _ReplaceGrad.apply(x_q, x), idx


def synth(sd, z, cfg=F16_16384, return_indices=False, force_idx=None):   # main.py:140-143
    z_q, idx = vector_quantize(z.movedim(1, 3), sd["quantize.embedding.weight"], force_idx)
    x = clamp_with_grad(decode(sd, z_q.movedim(3, 1), cfg).add(1).div(2), 0, 1)
    return (x, idx) if return_indices else x


def init_vqgan_state_dict(cfg=F16_16384, seed=0, codebook_std=1.0):
    """Deterministic random init with taming's key names/shapes.  Codebook ~ N(0, codebook_std) rather than
    taming's U(+-1/n_embed) so that nearest-code distances are not near-ties (SURVEY §7 hard parts)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, cout, cin, k):
        fan = cin * k * k
        bound = 1.0 / fan ** 0.5
        sd[name + ".weight"] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    def norm(name, c):
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g)

    def res(name, cin, cout):
        norm(name + ".norm1", cin)
        conv(name + ".conv1", cout, cin, 3)
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".nin_shortcut", cout, cin, 1)

    def attn(name, c):
        norm(name + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            conv(name + "." + n, c, c, 1)

    sd["quantize.embedding.weight"] = codebook_std * torch.randn(cfg["n_embed"], cfg["embed_dim"], generator=g)
    conv("post_quant_conv", cfg["z_channels"], cfg["embed_dim"], 1)
    block_in = cfg["ch"] * cfg["ch_mult"][-1]
    conv("decoder.conv_in", block_in, cfg["z_channels"], 3)
    res("decoder.mid.block_1", block_in, block_in)
    attn("decoder.mid.attn_1", block_in)
    res("decoder.mid.block_2", block_in, block_in)
    last = block_in
    for i_level, blocks, has_attn, has_up in decoder_layout(cfg):
        for j, (cin, cout) in enumerate(blocks):
            res("decoder.up.%d.block.%d" % (i_level, j), cin, cout)
            if has_attn:
                attn("decoder.up.%d.attn.%d" % (i_level, j), cout)
            last = cout
        if has_up:
            conv("decoder.up.%d.upsample.conv" % i_level, last, last, 3)
    norm("decoder.norm_out", last)
    conv("decoder.conv_out", cfg["out_ch"], last, 3)
    return sd
