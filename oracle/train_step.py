"""ORACLE (test infrastructure, not product): the reference's train-step body (main.py:729-837) restated on the
CPU in fp32 from the oracle pieces — mapper (oracle/mixer.py) -> clamp_with_grad -> synth (oracle/vqgan.py) ->
MakeCutouts + normalise (oracle/cutouts.py) -> encode_image (oracle/clip_vit.py) -> spherical loss (oracle/loss.py)
-> backward -> Adam on the mapper parameters only (main.py:591).
Used by tests/ (parity of the CUDA step) and by bench.py's cpu_baseline / --impl reference legs (timing of the
reference's CPU path).  Never imported by the product package."""
import torch

from . import clip_vit as oclip
from . import cutouts as ocut
from . import loss as oloss
from . import mixer as omix
from . import vqgan as ovq


class OracleTrainer:
    def __init__(self, sd_mixer, sd_vq, sd_clip, image_size, channels, vq_cfg=ovq.F16_16384, clip_cfg=oclip.VIT_B32,
                 cutn=8, cut_size=224, lr=1e-3, act="quick_gelu", l2_coef=0.0, tv_coef=0.0, mapper="mixer", num_heads=6,
                 repeat=1, diversity_coef=0.0, sd_vgg=None, diversity_mode="between_same_prompts", input_loss_coef=0.0,
                 normalize_input=False):
        self.params = {k: v.clone().requires_grad_(True) for k, v in sd_mixer.items()}
        self.sd_vq, self.sd_clip = sd_vq, sd_clip
        self.S, self.C = image_size, channels
        self.vq_cfg, self.clip_cfg, self.cutn, self.cut_size, self.act = vq_cfg, clip_cfg, cutn, cut_size, act
        self.opt = torch.optim.Adam(list(self.params.values()), lr=lr)       # main.py:591
        cb = sd_vq["quantize.embedding.weight"]
        self.z_lo, self.z_hi = float(cb.min()), float(cb.max())              # main.py:645-646,763
        self.last_indices = None
        self.l2_coef, self.tv_coef, self.mapper, self.num_heads = l2_coef, tv_coef, mapper, num_heads
        self.repeat, self.diversity_coef, self.sd_vgg, self.diversity_mode = repeat, diversity_coef, sd_vgg, diversity_mode
        self.input_loss_coef, self.normalize_input = input_loss_coef, normalize_input

    def step(self, inp_feats, out_feats, prm, force_idx=None):
        if self.normalize_input:                                                             # main.py:734-735
            inp_feats = torch.nn.functional.normalize(inp_feats, dim=1)
        if self.repeat > 1:                                                                  # main.py:739-740
            inp_feats, out_feats = inp_feats.repeat(self.repeat, 1), out_feats.repeat(self.repeat, 1)
        src_feats = inp_feats
        if prm.get("mapper_noise") is not None:                                              # main.py:741-750 (explicit noise rows)
            inp_feats = torch.cat((inp_feats, prm["mapper_noise"].to(inp_feats.device)), dim=1)
        if self.mapper == "vitgan":
            from . import vitgan as ovit
            z = ovit.vitgan_forward(self.params, inp_feats, self.C, self.num_heads).contiguous()
        elif self.mapper == "simple_vitgan":
            from . import vitgan as ovit
            z = ovit.simple_vitgan_forward(self.params, inp_feats, self.C, self.num_heads).contiguous()
        elif self.mapper == "xtransformer":
            from . import xtransformer as oxt
            z = oxt.xtransformer_forward(self.params, inp_feats, self.S, self.C, self.num_heads).contiguous()
        else:
            z = omix.mixer_forward(self.params, inp_feats, self.S, self.C).contiguous()      # main.py:754-757
        z.retain_grad()
        self.last_z = z
        l2 = (z ** 2).mean() if self.l2_coef > 0 else 0.0                                    # main.py:758-762
        z = ovq.clamp_with_grad(z, self.z_lo, self.z_hi)                                     # main.py:763
        xr, idx = ovq.synth(self.sd_vq, z, self.vq_cfg, return_indices=True, force_idx=force_idx)   # main.py:767
        self.last_indices = idx
        x = ocut.make_cutouts(xr, self.cutn, prm, self.cut_size, normalize=True)             # main.py:796-797
        embed = oclip.encode_image(self.sd_clip, x, self.clip_cfg, act=self.act).float()     # main.py:799
        tv = oloss.tv_loss(xr) if self.tv_coef > 0 else 0.0                                  # main.py:769-773
        dists = oloss.spherical_dist_loss(embed, out_feats, self.cutn)                       # main.py:801-811
        if self.input_loss_coef:                                                             # main.py:812-824
            dists = dists + oloss.spherical_dist_loss(embed, src_feats, self.cutn, self.input_loss_coef)
        self.last_terms = tuple(float(t.detach()) if torch.is_tensor(t) else float(t) for t in (dists, l2, tv))
        loss = dists + self.l2_coef * l2 + self.tv_coef * tv                                 # main.py:831
        if self.diversity_coef and self.sd_vgg is not None and (self.repeat > 1 or self.diversity_mode == "all"):
            from . import lpips as olpips                                                    # main.py:776-791,831
            mean = torch.tensor(ocut.CLIP_MEAN, device=xr.device).view(1, 3, 1, 1)
            std = torch.tensor(ocut.CLIP_STD, device=xr.device).view(1, 3, 1, 1)
            div = olpips.diversity(self.sd_vgg, xr, self.repeat, xr.shape[0] // self.repeat, mean, std, self.diversity_mode)
            loss = loss - self.diversity_coef * div
        self.opt.zero_grad()                                                                 # main.py:825
        loss.backward()                                                                      # main.py:832
        self.grads = {k: p.grad.clone() for k, p in self.params.items()}
        self.opt.step()                                                                      # main.py:835
        return loss.item()
