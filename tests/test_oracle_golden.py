"""CPU tests: pin the oracle (oracle/*.py) against golden vectors produced by the REFERENCE's own code
(tests/golden/make_golden.py, run against /root/reference).  No GPU, no reference needed at test time."""
import os

import pytest
import torch

import oracle.clip_vit as oclip
import oracle.loss as oloss
import oracle.mixer as omix
import oracle.vqgan as ovq

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_mixer_oracle_matches_reference_forward_and_grads():
    gold = torch.load(os.path.join(G, "mixer.pt"))
    for name in ("tiny", "s8"):
        g = gold[name]
        sd = {k: v.clone().requires_grad_(True) for k, v in g["state_dict"].items()}
        y = omix.mixer_forward(sd, g["x"], g["cfg"]["image_size"], g["cfg"]["channels"])
        assert y.shape == g["y"].shape
        assert torch.allclose(y, g["y"], rtol=1e-5, atol=1e-6), (y - g["y"]).abs().max()
        (y * g["w"]).sum().backward()
        for k, ref in g["grads"].items():
            assert torch.allclose(sd[k].grad, ref, rtol=1e-4, atol=1e-5), k
    # known answers (SURVEY §8c [probe])
    assert gold["count_8x128"] == 38948480
    sd = omix.init_mixer_state_dict(512, 16, 256, 128, 8)
    assert sum(v.numel() for v in sd.values()) == 38948480


def test_vitgan_oracle_matches_reference_forward_and_grads():
    import oracle.vitgan as ovit
    g = torch.load(os.path.join(G, "vitgan.pt"))
    sd = {k: v.clone().requires_grad_(True) for k, v in g["state_dict"].items()}
    y = ovit.vitgan_forward(sd, g["x"], g["cfg"]["out_channels"], g["cfg"]["num_heads"])
    assert torch.allclose(y, g["y"], rtol=1e-5, atol=1e-6)
    (y * g["w"]).sum().backward()
    for k, ref in g["grads"].items():
        assert torch.allclose(sd[k].grad, ref, rtol=1e-4, atol=1e-5), k


def test_simple_vitgan_oracle_matches_reference_forward_and_grads():
    """model_type 'simple_vitgan' (main.py:469-478): oracle/vitgan.simple_vitgan_forward against the real vitgan.SimpleGenerator"""
    import oracle.vitgan as ovit
    g = torch.load(os.path.join(G, "simple_vitgan.pt"))
    sd = {k: v.clone().requires_grad_(True) for k, v in g["state_dict"].items()}
    y = ovit.simple_vitgan_forward(sd, g["x"], g["cfg"]["out_channels"], g["cfg"]["num_heads"])
    assert y.shape == g["y"].shape and torch.allclose(y, g["y"], rtol=1e-5, atol=1e-6)
    (y * g["w"]).sum().backward()
    for k, ref in g["grads"].items():
        assert torch.allclose(sd[k].grad, ref, rtol=1e-4, atol=1e-5), k


def test_clip_oracle_matches_reference_twin():
    g = torch.load(os.path.join(G, "clip_vit.pt"))
    x = g["x"].clone().requires_grad_(True)
    y = oclip.encode_image(g["state_dict"], x, g["cfg"], act="quick_gelu")
    assert torch.allclose(y, g["y"], rtol=1e-4, atol=1e-5), (y - g["y"]).abs().max()
    (y * g["w"]).sum().backward()
    assert torch.allclose(x.grad, g["dx"], rtol=1e-3, atol=1e-5), (x.grad - g["dx"]).abs().max()
    assert g["count_vitb32"] == 87849216
    assert sum(v.numel() for v in oclip.init_clip_state_dict().values()) == 87849216


def test_clip_text_oracle_matches_reference_twin():
    import oracle.clip_text as otext
    g = torch.load(os.path.join(G, "clip_text.pt"))
    y = otext.encode_text(g["state_dict"], g["text"], g["cfg"]["transformer_heads"])
    assert torch.allclose(y, g["y"], rtol=1e-4, atol=1e-5), (y - g["y"]).abs().max()


def test_glue_oracle_matches_reference_main_py():
    g = torch.load(os.path.join(G, "glue.pt"))
    c = g["clamp"]
    x = c["x"].clone().requires_grad_(True)
    ovq.clamp_with_grad(x, 0, 1).backward(c["g"])
    assert torch.equal(x.grad, c["gx"])
    assert c["gx"].tolist() == [0.0, 1.0, -1.0, -0.0, 1.0, -1.0]
    v = g["vq"]
    z = v["z"].clone().requires_grad_(True)
    zq, idx = ovq.vector_quantize(z, v["cb"])
    assert torch.equal(idx, v["idx"]) and torch.allclose(zq, v["zq"])
    (zq * v["w"]).sum().backward()
    assert torch.allclose(z.grad, v["dz"])                   # straight-through: dz == w
    s = g["synth"]
    # synth glue around a stand-in linear decode (same stand-in as the golden script)
    z2 = s["z"].clone().requires_grad_(True)
    zq2, _ = ovq.vector_quantize(z2.movedim(1, 3), v["cb"])
    dec = torch.einsum("oc,bchw->bohw", s["lin"], zq2.movedim(3, 1)) * 0.7
    xr = ovq.clamp_with_grad(dec.add(1).div(2), 0, 1)
    assert torch.allclose(xr, s["xr"], atol=1e-6)
    (xr * s["w"]).sum().backward()
    assert torch.allclose(z2.grad, s["dz"], atol=1e-6)
    t = g["tv"]
    assert torch.allclose(oloss.tv_loss(t["img"]), t["tv"])
    l = g["loss"]
    e = l["embed"].clone().requires_grad_(True)
    d = oloss.spherical_dist_loss(e, l["feats"], l["cutn"])
    assert torch.allclose(d, l["dists"], atol=1e-6)
    d.backward()
    assert torch.allclose(e.grad, l["dembed"], atol=1e-6)


def test_vqgan_decoder_oracle_shapes_and_param_count():
    # parity UNPINNED at the taming boundary (package absent): structural known answers only
    sd = ovq.init_vqgan_state_dict()
    dec = sum(v.numel() for k, v in sd.items() if k.startswith("decoder.") or k.startswith("post_quant"))
    assert 42.0e6 < dec < 43.0e6, dec                       # "~42.5 M params" 
    assert sd["quantize.embedding.weight"].shape == (16384, 256)
    small = dict(ch=32, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(4,), resolution=8, z_channels=32, out_ch=3,
                 embed_dim=32, n_embed=64)
    sds = ovq.init_vqgan_state_dict(small, seed=3)
    z = torch.randn(2, 32, 4, 4, requires_grad=True)
    x = ovq.synth(sds, z, small)
    assert x.shape == (2, 3, 8, 8) and float(x.min()) >= 0 and float(x.max()) <= 1
    x.sum().backward()
    assert z.grad.shape == z.shape


def test_cutout_oracle_runs_and_is_differentiable():
    import oracle.cutouts as oc
    from feed_forward_vqgan_clip_b200.cutouts import sample_params
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 3, 64, 64, generator=g, requires_grad=True)
    prm = sample_params(6, 32, g)
    y = oc.make_cutouts(x, 3, prm, 32)
    assert y.shape == (6, 3, 32, 32)
    y.sum().backward()
    assert torch.isfinite(x.grad).all()
    # identity parameters reduce to pooled + noise, normalised
    ident = dict(affine_inv=torch.eye(3).repeat(6, 1, 1), persp_inv=torch.eye(3).repeat(6, 1, 1), sat=torch.ones(6),
                 hue=torch.zeros(6), erase=[0, 0, 0, 0], noise=torch.zeros(6, 3, 32, 32))
    y0 = oc.make_cutouts(x.detach(), 3, ident, 32, normalize=False)
    pooled = (torch.nn.functional.adaptive_avg_pool2d(x.detach(), 32) + torch.nn.functional.adaptive_max_pool2d(x.detach(), 32)) / 2
    assert torch.allclose(y0, pooled.repeat(3, 1, 1, 1), atol=2e-6)


def test_lpips_vgg16_taps_match_torchvision_vgg16_features():
    """oracle/lpips.py restates taming's `vgg16` slices (relu1_2, relu2_2, relu3_3, relu4_3, relu5_3 of torchvision's VGG16
    `features`).  taming is absent, torchvision is not: pin the STRUCTURE (layer order, pooling positions, tap
    points, key numbering) against torchvision.models.vgg16 with the same random weights."""
    torchvision = pytest.importorskip("torchvision")
    import oracle.lpips as ol
    torch.manual_seed(11)
    feats = torchvision.models.vgg16(weights=None).features.eval()
    sd = {}
    for s, idxs in ol.VGG16_SLICES:
        for i in idxs:
            assert isinstance(feats[i], torch.nn.Conv2d) and tuple(feats[i].weight.shape[:2]) == ol.VGG16_CH[i][::-1]
            sd["slice%d.%d.weight" % (s, i)] = feats[i].weight.detach().clone()
            sd["slice%d.%d.bias" % (s, i)] = feats[i].bias.detach().clone()
    x = torch.randn(2, 3, 64, 64)
    mine = ol.vgg_taps(sd, x)
    ends = [4, 9, 16, 23, 30]                      # taming lpips.vgg16: slices end after relu1_2 ... relu5_3
    h, start, ref = x, 0, []
    with torch.no_grad():
        for e in ends:
            for i in range(start, e):
                h = feats[i](h)
            ref.append(h)
            start = e
    assert len(mine) == 5
    for a, b in zip(mine, ref):
        assert a.shape == b.shape and torch.allclose(a, b, atol=1e-5, rtol=1e-5)


def test_clip_vit_restatement_matches_transformers_clip_vision_model():
    """secondary cross-check: an independent implementation of the CLIP ViT — transformers'
    CLIPVisionModelWithProjection, random init — gives the oracle's embedding when its weights are mapped to the CLIP
    `visual.*` key names the oracle (and the reference's cloob twin) use."""
    transformers = pytest.importorskip("transformers")
    import oracle.clip_vit as oc
    cfg = dict(input_resolution=64, patch_size=32, width=64, layers=2, heads=2, output_dim=32)
    hf_cfg = transformers.CLIPVisionConfig(hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                                           image_size=64, patch_size=32, projection_dim=32, hidden_act="quick_gelu",
                                           layer_norm_eps=1e-5, attention_dropout=0.0)
    torch.manual_seed(12)
    m = transformers.CLIPVisionModelWithProjection(hf_cfg).eval()
    with torch.no_grad():
        for p in m.parameters():                   # HF initialises LayerNorms to (1, 0) and biases to 0: make every term matter
            p.add_(0.05 * torch.randn_like(p))
    hf = {k: v.detach() for k, v in m.state_dict().items()}
    v = "vision_model."
    sd = {"conv1.weight": hf[v + "embeddings.patch_embedding.weight"], "class_embedding": hf[v + "embeddings.class_embedding"],
          "positional_embedding": hf[v + "embeddings.position_embedding.weight"], "proj": hf["visual_projection.weight"].t(),
          "ln_pre.weight": hf[v + "pre_layrnorm.weight"], "ln_pre.bias": hf[v + "pre_layrnorm.bias"],
          "ln_post.weight": hf[v + "post_layernorm.weight"], "ln_post.bias": hf[v + "post_layernorm.bias"]}
    for l in range(2):
        a, b = v + "encoder.layers.%d." % l, "transformer.resblocks.%d." % l
        sd[b + "attn.in_proj_weight"] = torch.cat([hf[a + "self_attn.%s_proj.weight" % n] for n in "qkv"])
        sd[b + "attn.in_proj_bias"] = torch.cat([hf[a + "self_attn.%s_proj.bias" % n] for n in "qkv"])
        sd[b + "attn.out_proj.weight"], sd[b + "attn.out_proj.bias"] = hf[a + "self_attn.out_proj.weight"], hf[a + "self_attn.out_proj.bias"]
        for mine, theirs in (("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"), ("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")):
            sd[b + mine + ".weight"], sd[b + mine + ".bias"] = hf[a + theirs + ".weight"], hf[a + theirs + ".bias"]
    x = torch.randn(3, 3, 64, 64)
    with torch.no_grad():
        ref = m(pixel_values=x).image_embeds
        out = oc.encode_image(sd, x, cfg, act="quick_gelu")
    assert out.shape == ref.shape == (3, 32)
    assert torch.allclose(out, ref, atol=2e-5, rtol=1e-4), (out - ref).abs().max()


def test_vqgan_decoder_restatement_matches_transformers_janus_vqvae_decoder():
    """taming-transformers (the reference's VQGAN, main.py:29,84-103,142) is absent from the reference tree and the image, so
    oracle/vqgan.py restates its Decoder from the published architecture.  transformers ships an independent
    implementation of the same decoder family (JanusVQVAEDecoder: GroupNorm(32, 1e-6) + swish ResnetBlocks with nin_shortcut,
    single-head AttnBlock at the lowest resolution, nearest-2x + conv Upsample, mid block_1 / attn_1 / block_2, norm_out, conv_out).
    With its random weights mapped onto taming's key names (only the level index differs: HF appends levels in processing
    order) the restatement reproduces its output and input gradient — every key and shape corresponds one to one."""
    pytest.importorskip("transformers")
    from transformers.models.janus import modeling_janus as mj
    from transformers.models.janus.configuration_janus import JanusVQVAEConfig
    hf_cfg = JanusVQVAEConfig(embed_dim=16, num_embeddings=64, double_latent=False, latent_channels=16, in_channels=3, out_channels=3,
                              base_channels=32, channel_multiplier=[1, 2, 2], num_res_blocks=2, dropout=0.0)
    torch.manual_seed(21)
    dec = mj.JanusVQVAEDecoder(hf_cfg).eval()
    with torch.no_grad():
        for p in dec.parameters():                 # GroupNorm starts at (1, 0): make every term matter
            p.add_(0.05 * torch.randn_like(p))
    cfg = dict(ch=32, ch_mult=(1, 2, 2), num_res_blocks=2, attn_resolutions=(8,), resolution=32, z_channels=16, out_ch=3,
               embed_dim=16, n_embed=64)
    ref_sd = ovq.init_vqgan_state_dict(cfg, seed=0)
    levels = len(cfg["ch_mult"])
    sd = {}
    for k, v in dec.state_dict().items():
        parts = k.split(".")
        if parts[0] == "up":
            parts[1] = str(levels - 1 - int(parts[1]))
        sd["decoder." + ".".join(parts)] = v.detach().clone()
    sd["post_quant_conv.weight"] = torch.eye(16).view(16, 16, 1, 1)            # identity: compare the Decoder alone
    sd["post_quant_conv.bias"] = torch.zeros(16)
    sd["quantize.embedding.weight"] = ref_sd["quantize.embedding.weight"]
    assert set(sd) == set(ref_sd)
    for k in ref_sd:
        assert sd[k].shape == ref_sd[k].shape, k
    z = torch.randn(2, 16, 8, 8)
    w = torch.randn(2, 3, 32, 32)
    za, zb = z.clone().requires_grad_(True), z.clone().requires_grad_(True)
    a = dec(za)
    b = ovq.decode(sd, zb, cfg)
    assert a.shape == b.shape == (2, 3, 32, 32)
    assert torch.allclose(a, b, atol=1e-5, rtol=1e-5), (a - b).abs().max()
    (a * w).sum().backward()
    (b * w).sum().backward()
    assert torch.allclose(za.grad, zb.grad, atol=1e-4, rtol=1e-4), (za.grad - zb.grad).abs().max()


def test_cutout_warp_and_hue_restatements_match_torchvision():
    """kornia (the reference's augmentation library, main.py:18,170-200) is absent; torchvision is not.  Two conventions of
    oracle/cutouts.py are checked against torchvision's independent implementations of the same operations:
      * the inverse affine map built by cutouts.sample_params (rotate about the centre, then translate; pixel centres at integer
        coordinates; bilinear taps) against transforms.v2.functional.affine — interior pixels (the padding modes differ),
      * the HSV round trip + hue shift against adjust_hue (hue factor f <-> 2 pi f radians)."""
    pytest.importorskip("torchvision")
    import math
    import torchvision.transforms.v2.functional as TF
    from torchvision.transforms import InterpolationMode
    import oracle.cutouts as oc
    torch.manual_seed(31)
    P = 64
    img = torch.rand(2, 3, P, P)
    for f in (0.07, -0.04, 0.1):
        a = oc.color_jitter(img, torch.ones(2), torch.full((2,), f * 2 * math.pi))
        assert torch.allclose(a, TF.adjust_hue(img, f), atol=1e-5), f
    c = (P - 1) / 2.0
    for ang_deg, tx, ty in ((10.0, 3.0, -2.0), (-14.0, -5.0, 4.0), (0.0, 6.0, 6.0)):
        ang = ang_deg * math.pi / 180
        ca, sa = math.cos(ang), math.sin(ang)
        inv = torch.zeros(2, 3, 3)                      # same expressions as cutouts.sample_params ("Af")
        inv[:, 0, 0], inv[:, 0, 1], inv[:, 0, 2] = ca, sa, c - ca * (c + tx) - sa * (c + ty)
        inv[:, 1, 0], inv[:, 1, 1], inv[:, 1, 2] = -sa, ca, c + sa * (c + tx) - ca * (c + ty)
        inv[:, 2, 2] = 1.0
        a = oc.warp(img, inv, "zeros")
        b = TF.affine(img, angle=ang_deg, translate=[tx, ty], scale=1.0, shear=[0.0, 0.0], interpolation=InterpolationMode.BILINEAR,
                      fill=0.0)
        m = slice(14, P - 14)
        assert torch.allclose(a[..., m, m], b[..., m, m], atol=2e-5), (ang_deg, (a - b)[..., m, m].abs().max())


def test_xtransformer_stack_restatement_matches_transformers_gpt2():
    """x-transformers (transformer.py:3,11-20) is absent from the reference tree and the image, so oracle/xtransformer.py
    restates ContinuousTransformerWrapper + Decoder from the published architecture.  transformers' GPT2Model
    is an independent implementation of the same stack — learned absolute positions added unscaled, causal pre-LayerNorm
    blocks (attention scaled by head_dim**-0.5, projections q|k|v -> out, residual; LayerNorm, Linear-GELU-Linear x4, residual),
    final LayerNorm.  It only exists with heads * head_dim == dim, so the check runs at dim 128 = 2 heads x 64 (x-transformers'
    fixed dim_head); with its weights mapped onto the package's key names (Conv1D stores [in, out]; the qkv bias, which
    x-transformers does not have, is zeroed) the restatement reproduces GPT-2's hidden states and input gradient."""
    transformers = pytest.importorskip("transformers")
    import oracle.xtransformer as ox
    S, C, dim, heads, depth, in_dim = 3, 16, 128, 2, 2, 24
    T = S * S
    hf_cfg = transformers.GPT2Config(vocab_size=8, n_positions=T + 1, n_embd=dim, n_layer=depth, n_head=heads, n_inner=4 * dim,
                                     activation_function="gelu", resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0,
                                     layer_norm_epsilon=1e-5, scale_attn_weights=True, scale_attn_by_inverse_layer_idx=False,
                                     reorder_and_upcast_attn=False)
    torch.manual_seed(41)
    m = transformers.GPT2Model(hf_cfg).eval()
    with torch.no_grad():
        for p in m.parameters():                   # LayerNorms start at (1, 0), biases at 0: make every term matter
            p.add_(0.05 * torch.randn_like(p))
        for l in range(depth):
            m.h[l].attn.c_attn.bias.zero_()
    hf = {k: v.detach() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(42)
    sd = {"proj.weight": torch.randn(T * dim, in_dim, generator=g) * 0.2, "proj.bias": torch.randn(T * dim, generator=g) * 0.1,
          "transformer.project_in.weight": torch.randn(dim, dim, generator=g) * 0.1,
          "transformer.project_in.bias": torch.randn(dim, generator=g) * 0.1,
          "transformer.pos_emb.emb.weight": hf["wpe.weight"],
          "transformer.norm.weight": hf["ln_f.weight"], "transformer.norm.bias": hf["ln_f.bias"],
          "transformer.project_out.weight": torch.randn(C, dim, generator=g) * 0.1,
          "transformer.project_out.bias": torch.randn(C, generator=g) * 0.1}
    for l in range(depth):
        a, f, h = "transformer.attn_layers.layers.%d." % (2 * l), "transformer.attn_layers.layers.%d." % (2 * l + 1), "h.%d." % l
        sd[a + "0.weight"], sd[a + "0.bias"] = hf[h + "ln_1.weight"], hf[h + "ln_1.bias"]
        wq, wk, wv = hf[h + "attn.c_attn.weight"].split(dim, dim=1)                   # Conv1D: [in, 3 * out]
        sd[a + "1.to_q.weight"], sd[a + "1.to_k.weight"], sd[a + "1.to_v.weight"] = wq.t(), wk.t(), wv.t()
        sd[a + "1.to_out.weight"], sd[a + "1.to_out.bias"] = hf[h + "attn.c_proj.weight"].t(), hf[h + "attn.c_proj.bias"]
        sd[f + "0.weight"], sd[f + "0.bias"] = hf[h + "ln_2.weight"], hf[h + "ln_2.bias"]
        sd[f + "1.net.0.0.weight"], sd[f + "1.net.0.0.bias"] = hf[h + "mlp.c_fc.weight"].t(), hf[h + "mlp.c_fc.bias"]
        sd[f + "1.net.2.weight"], sd[f + "1.net.2.bias"] = hf[h + "mlp.c_proj.weight"].t(), hf[h + "mlp.c_proj.bias"]
    assert ox.xt_depth(sd) == depth
    x = torch.randn(3, in_dim, generator=g)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    F = torch.nn.functional
    # the wrapper around the stack is transformer.py:30-32,44-46 + two Linears; the stack itself runs in GPT-2
    e = F.linear(F.linear(xa, sd["proj.weight"], sd["proj.bias"]).view(3, T, dim), sd["transformer.project_in.weight"],
                 sd["transformer.project_in.bias"])
    hs = m(inputs_embeds=e).last_hidden_state
    ref = F.linear(hs, sd["transformer.project_out.weight"], sd["transformer.project_out.bias"]).view(3, S, S, C).permute(0, 3, 1, 2)
    out = ox.xtransformer_forward(sd, xb, S, C, heads)
    assert out.shape == ref.shape == (3, C, S, S)
    assert torch.allclose(out, ref, atol=2e-5, rtol=1e-4), (out - ref).abs().max()
    w = torch.randn(3, C, S, S, generator=g)
    (ref * w).sum().backward()
    (out * w).sum().backward()
    assert torch.allclose(xa.grad, xb.grad, atol=1e-4, rtol=1e-4), (xa.grad - xb.grad).abs().max()
    # causality: the latent at token t must not depend on projected tokens after t (Decoder = causal=True, transformer.py:15)
    sd2 = dict(sd)
    pw = sd["proj.weight"].clone().view(T, dim, in_dim)
    pw[T - 1] += 1.0                                     # perturb only the last token's projection
    sd2["proj.weight"] = pw.view(T * dim, in_dim)
    with torch.no_grad():
        o2 = ox.xtransformer_forward(sd2, x, S, C, heads)
    a1, a2 = out.detach().permute(0, 2, 3, 1).reshape(3, T, C), o2.permute(0, 2, 3, 1).reshape(3, T, C)
    assert torch.equal(a1[:, :T - 1], a2[:, :T - 1]) and not torch.allclose(a1[:, T - 1], a2[:, T - 1])


def test_cutout_perspective_restatement_matches_torchvision():
    """RandomPerspective's arithmetic (main.py:177-178; kornia absent): the inverse homography cutouts.sample_params builds in
    closed form (quadrilateral of perturbed corners -> the cut_size square, pixel centres at integer coordinates) and
    oracle/cutouts.warp's bilinear sampling with zero padding, against torchvision's independent perspective() — which solves the
    8x8 system for the same four point pairs and samples with pixel centres at half-integers (hence the +0.5 on both point sets)."""
    pytest.importorskip("torchvision")
    import torchvision.transforms.v2.functional as TF
    from torchvision.transforms import InterpolationMode
    import oracle.cutouts as oc
    from feed_forward_vqgan_clip_b200.cutouts import _persp_coeffs, _quad_to_square
    torch.manual_seed(51)
    P = 48
    q = float(P - 1)
    img = torch.rand(1, 3, P, P)
    g = torch.Generator().manual_seed(52)
    r = torch.rand(5, 8, generator=g, dtype=torch.float64) * (0.7 * q / 2)            # distortion_scale 0.7, as in sample_params
    dst = torch.stack([torch.stack([r[:, 0], r[:, 1]], -1), torch.stack([q - r[:, 2], r[:, 3]], -1),
                       torch.stack([q - r[:, 4], q - r[:, 5]], -1), torch.stack([r[:, 6], q - r[:, 7]], -1)], dim=1)
    src = torch.tensor([[0, 0], [q, 0], [q, q], [0, q]], dtype=torch.float64)
    inv = _quad_to_square(dst, q)
    assert torch.allclose(inv, _persp_coeffs(dst, src.expand(5, 4, 2)), atol=1e-9)    # closed form == the general solve
    for i in range(5):
        a = oc.warp(img, inv[i:i + 1].float(), "zeros")
        b = TF.perspective(img, startpoints=(src + 0.5).tolist(), endpoints=(dst[i] + 0.5).tolist(),
                           interpolation=InterpolationMode.BILINEAR, fill=None)   # plain zero padding (an explicit fill
        assert torch.allclose(a, b, atol=5e-5), (i, (a - b).abs().max())              # attenuates edge pixels twice)
    # RandomErasing(value=0) rectangle convention ([x0, y0, x1, y1), one rectangle for the whole batch) against TF.erase
    x = torch.rand(2, 3, P, P)
    prm = dict(affine_inv=torch.eye(3).repeat(4, 1, 1), persp_inv=torch.eye(3).repeat(4, 1, 1), sat=torch.ones(4), hue=torch.zeros(4),
               erase=[5, 9, 30, 21], noise=torch.zeros(4, 3, P, P))
    y = oc.make_cutouts(x, 2, prm, P, normalize=False)
    assert torch.allclose(y, TF.erase(x.repeat(2, 1, 1, 1), i=9, j=5, h=12, w=25, v=torch.zeros(1)), atol=1e-5)   # the identity jitter still runs the HSV round trip
