"""Generates the golden vectors under tests/golden/ by running the REFERENCE's own code (imported from
/root/reference, read-only) on seeded inputs.  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Importable reference pieces: mlp_mixer_pytorch.Mixer, cloob.VisualTransformer (the in-tree CLIP ViT
twin) and main.py's pure-torch glue (ReplaceGrad / ClampWithGrad / vector_quantize / synth / tv_loss + the loss
expression), the latter after stubbing main.py's absent third-party imports with MagicMock.
"""
import os
import sys
from unittest.mock import MagicMock

import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)


def golden_mixer():
    from mlp_mixer_pytorch import Mixer
    torch.manual_seed(0)
    cfgs = {"tiny": dict(input_dim=32, image_size=4, channels=16, patch_size=1, dim=32, depth=2),
            "s8": dict(input_dim=24, image_size=8, channels=8, patch_size=1, dim=32, depth=1)}
    out = {}
    for name, cfg in cfgs.items():
        net = Mixer(**cfg)
        x = torch.randn(3, cfg["input_dim"])
        w = torch.randn(3, cfg["channels"], cfg["image_size"], cfg["image_size"])
        y = net(x)
        (y * w).sum().backward()
        out[name] = dict(cfg=cfg, state_dict={k: v.detach().clone() for k, v in net.state_dict().items()}, x=x, w=w,
                         y=y.detach().contiguous(), grads={k: p.grad.clone() for k, p in net.named_parameters()},
                         n_params=sum(p.numel() for p in net.parameters()))
    # parameter-count known answers
    out["count_8x128"] = sum(p.numel() for p in Mixer(512, 16, 256, 1, 128, 8).parameters())
    torch.save(out, os.path.join(OUT, "mixer.pt"))


def golden_vitgan():
    from vitgan import Generator
    torch.manual_seed(4)
    cfg = dict(initialize_size=1, dim=24, blocks=2, num_heads=3, out_channels=4, input_dim=16)      # T = 8, dim_head = 8
    net = Generator(**cfg)
    x = torch.randn(3, 16)
    y = net(x)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    big = Generator(initialize_size=2, dim=1024, blocks=32, num_heads=6, out_channels=256, input_dim=512)
    torch.save(dict(cfg=cfg, state_dict={k: v.detach().clone() for k, v in net.state_dict().items()}, x=x, w=w, y=y.detach(),
                    grads={k: p.grad.clone() for k, p in net.named_parameters()},
                    count_32x1024=sum(p.numel() for p in big.parameters()),
                    keys_32x1024=[(k, tuple(v.shape)) for k, v in big.state_dict().items() if ".blocks." not in k or ".blocks.0." in k]),
               os.path.join(OUT, "vitgan.pt"))


def golden_simple_vitgan():
    from vitgan import SimpleGenerator
    torch.manual_seed(6)
    cfg = dict(size=3, dim=24, blocks=2, num_heads=3, out_channels=4, input_dim=16)      # T = 9, dim_head = 8
    net = SimpleGenerator(**cfg)
    x = torch.randn(3, 16)
    y = net(x)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    big = SimpleGenerator(size=16, dim=256, blocks=1, num_heads=6, out_channels=256, input_dim=512)
    torch.save(dict(cfg=cfg, state_dict={k: v.detach().clone() for k, v in net.state_dict().items()}, x=x, w=w, y=y.detach(),
                    grads={k: p.grad.clone() for k, p in net.named_parameters()},
                    keys_16x256=[(k, tuple(v.shape)) for k, v in big.state_dict().items()]),
               os.path.join(OUT, "simple_vitgan.pt"))


def golden_clip():
    from cloob import VisualTransformer
    torch.manual_seed(1)
    cfg = dict(input_resolution=64, patch_size=32, width=64, layers=2, heads=1, output_dim=32)
    net = VisualTransformer(**cfg)
    for p in net.parameters():          # default init leaves LayerNorm at identity; perturb everything a little
        p.data.add_(0.02 * torch.randn_like(p))
    x = torch.randn(3, 3, 64, 64, requires_grad=True)
    w = torch.randn(3, 32)
    y = net(x)
    (y * w).sum().backward()
    torch.save(dict(cfg=cfg, state_dict={k: v.detach().clone() for k, v in net.state_dict().items()}, x=x.detach(), w=w,
                    y=y.detach(), dx=x.grad.clone(),
                    count_vitb32=sum(p.numel() for p in VisualTransformer(224, 32, 768, 12, 12, 512).parameters())),
               os.path.join(OUT, "clip_vit.pt"))


def golden_clip_text():
    from cloob import TextTransformer
    torch.manual_seed(5)
    cfg = dict(embed_dim=32, context_length=77, vocab_size=300, transformer_width=128, transformer_heads=2, transformer_layers=2)
    net = TextTransformer(**cfg)
    for p in net.parameters():
        p.data.add_(0.02 * torch.randn_like(p))
    text = torch.zeros(4, 77, dtype=torch.long)
    for b in range(4):
        n = 5 + 7 * b
        text[b, :n] = torch.randint(1, 298, (n,))
        text[b, n] = 299                                   # EOT = highest id
    with torch.no_grad():
        y = net(text)
    torch.save(dict(cfg=cfg, state_dict={k: v.detach().clone() for k, v in net.state_dict().items()}, text=text, y=y),
               os.path.join(OUT, "clip_text.pt"))


def golden_glue():
    for name in ["clize", "omegaconf", "kornia", "kornia.augmentation", "taming", "taming.models",
                 "taming.models.cond_transformer", "taming.models.vqgan", "taming.modules", "taming.modules.losses",
                 "taming.modules.losses.lpips", "clip", "clip.simple_tokenizer", "transformer",
                 "torch.utils.tensorboard"]:
        sys.modules.setdefault(name, MagicMock())
    os.environ["USE_HOROVOD"] = "false"
    import main as ref
    torch.manual_seed(2)
    out = {}
    # ClampWithGrad truth table + random
    x = torch.tensor([-1.0, 2.0, -1.0, 2.0, 0.5, 0.5], requires_grad=True)
    g = torch.tensor([1.0, 1.0, -1.0, -1.0, 1.0, -1.0])
    ref.clamp_with_grad(x, 0, 1).backward(g)
    out["clamp"] = dict(x=x.detach(), g=g, gx=x.grad.clone())
    # vector_quantize: values, indices, straight-through gradient
    cb = torch.randn(50, 8)
    z = torch.randn(2, 4, 4, 8, requires_grad=True)
    zq = ref.vector_quantize(z, cb)
    wq = torch.randn_like(zq)
    (zq * wq).sum().backward()
    d = z.detach().pow(2).sum(-1, keepdim=True) + cb.pow(2).sum(1) - 2 * z.detach() @ cb.T
    out["vq"] = dict(cb=cb, z=z.detach(), zq=zq.detach(), idx=d.argmin(-1), w=wq, dz=z.grad.clone())
    # synth with a stand-in decode (so that the glue around decode is what is pinned)
    class Stub:
        pass
    model = Stub()
    model.quantize = Stub()
    model.quantize.embedding = Stub()
    model.quantize.embedding.weight = cb
    lin = torch.randn(3, 8)
    model.decode = lambda zq_: torch.einsum("oc,bchw->bohw", lin, zq_) * 0.7
    z2 = torch.randn(2, 8, 4, 4, requires_grad=True)
    xr = ref.synth(model, z2)
    wi = torch.randn_like(xr)
    (xr * wi).sum().backward()
    out["synth"] = dict(lin=lin, z=z2.detach(), xr=xr.detach(), w=wi, dz=z2.grad.clone())
    # tv_loss and the spherical loss expression (main.py:801-811)
    img = torch.rand(2, 3, 9, 7)
    out["tv"] = dict(img=img, tv=ref.tv_loss(img))
    cutn, B, D = 4, 3, 16
    embed = torch.randn(cutn * B, D, requires_grad=True)
    feats = torch.randn(B, D) * 0.45
    F = torch.nn.functional
    H = F.normalize(feats.repeat(cutn, 1).view(cutn, 1, B, D), dim=-1).view(-1, D)
    e = F.normalize(embed, dim=1)
    dists = 1.0 * ((H.sub(e).norm(dim=-1).div(2).arcsin().pow(2).mul(2)).mean())
    dists.backward()
    out["loss"] = dict(embed=embed.detach(), feats=feats, cutn=cutn, dists=dists.detach(), dembed=embed.grad.clone())
    # (appended last so that the draws above keep their values) vector_quantize at a size the CUDA search accepts: 64-dim codes
    cb64 = torch.randn(52, 64)
    z64 = (torch.randn(2, 4, 4, 64) * 1.2).requires_grad_(True)
    zq64 = ref.vector_quantize(z64, cb64)
    w64 = torch.randn_like(zq64)
    (zq64 * w64).sum().backward()
    d64 = z64.detach().pow(2).sum(-1, keepdim=True) + cb64.pow(2).sum(1) - 2 * z64.detach() @ cb64.T
    out["vq64"] = dict(cb=cb64, z=z64.detach(), zq=zq64.detach(), idx=d64.argmin(-1), w=w64, dz=z64.grad.clone())
    torch.save(out, os.path.join(OUT, "glue.pt"))


if __name__ == "__main__":
    if len(sys.argv) > 1:                # regenerate only the named fixtures, e.g. `make_golden.py simple_vitgan`
        for name in sys.argv[1:]:
            globals()["golden_" + name]()
        sys.exit(0)
    golden_mixer()
    golden_vitgan()
    golden_simple_vitgan()
    golden_clip()
    golden_clip_text()
    golden_glue()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
