#!/usr/bin/env python
"""GPU diagnostic: where does the whole-step gradient of the tiny VitGAN / Mixer configs deviate from the oracle?
Prints index agreement, per-parameter cosine with free and with forced VQ indices, and the cosine of dz (the gradient
entering the mapper).  python tests/diag_e2e.py > gpurun_out/diag_e2e.log"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import oracle.clip_vit as oclip  # noqa: E402
import oracle.vqgan as ovq  # noqa: E402
from oracle.train_step import OracleTrainer  # noqa: E402
from feed_forward_vqgan_clip_b200.clip_vit import CLIP  # noqa: E402
from feed_forward_vqgan_clip_b200.cutouts import sample_params  # noqa: E402
from feed_forward_vqgan_clip_b200.mixer import Mixer  # noqa: E402
from feed_forward_vqgan_clip_b200.train_step import TrainStep  # noqa: E402
from feed_forward_vqgan_clip_b200.vitgan_mapper import Generator as VitGAN  # noqa: E402
from feed_forward_vqgan_clip_b200.vqgan import VQModel  # noqa: E402

DEV = "cuda:0"
SMALL_VQ = dict(ch=64, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(16,), resolution=32, z_channels=64, out_ch=3,
                embed_dim=64, n_embed=512)
SMALL_CLIP = dict(input_resolution=224, patch_size=32, width=128, layers=2, heads=2, output_dim=64)


def r16(sd):
    return {k: (v.to(torch.bfloat16).float() if v.dim() >= 2 else v.clone()) for k, v in sd.items()}


def cos(a, b):
    return torch.nn.functional.cosine_similarity(a.detach().float().cpu().flatten(), b.detach().float().cpu().flatten(), dim=0).item()


def run(kind, l2, tv):
    if kind == "vitgan":
        torch.manual_seed(13)
        net = VitGAN(initialize_size=2, dim=128, blocks=2, num_heads=6, out_channels=64, input_dim=64)
        with torch.no_grad():
            net.w_out[0].weight.mul_(4.0)
    else:
        torch.manual_seed(7)
        net = Mixer(input_dim=64, image_size=16, channels=64, patch_size=1, dim=128, depth=2)
        with torch.no_grad():
            net.final_proj.weight.mul_(6.0)
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() >= 2 and p.numel() > 1:
                p.copy_(p.to(torch.bfloat16).float())
    sd_m = {k: v.clone() for k, v in net.state_dict().items()}
    sd_v = r16(ovq.init_vqgan_state_dict(SMALL_VQ, seed=8))
    vq = VQModel(SMALL_VQ)
    vq.load_state_dict(sd_v)
    vq = vq.to(DEV).eval().requires_grad_(False)
    sd_c = r16(oclip.init_clip_state_dict(SMALL_CLIP, seed=9))
    clip = CLIP(SMALL_CLIP)
    clip.visual.load_state_dict(sd_c)
    clip = clip.to(DEV).eval().requires_grad_(False)
    net = net.to(DEV)
    g = torch.Generator().manual_seed(14)
    x = (torch.randn(2, 64, generator=g) * 0.45).to(torch.bfloat16).float()
    prm = sample_params(8, 224, g)
    ts = TrainStep(net, vq, clip, cutn=4, lr=1e-3, l2_coef=l2, tv_coef=tv)
    eng = net.engine()
    cap = {}
    real_bwd = eng.backward

    def bwd(sv, dz):
        cap["dz"] = dz.detach().clone()
        return real_bwd(sv, dz)

    eng.backward = bwd
    loss = ts.step(x.to(DEV), None, prm)
    torch.cuda.synchronize()
    grads = {n: gv.detach().clone().cpu() for (n, p), gv in zip(net.named_parameters(), eng.grad_views)}
    for forced in (False, True):
        otr = OracleTrainer(sd_m, sd_v, sd_c, 16, 64, SMALL_VQ, SMALL_CLIP, cutn=4, lr=1e-3, l2_coef=l2, tv_coef=tv,
                            mapper=kind, num_heads=6)
        otr.step(x, x, prm, force_idx=ts.last_indices.cpu().long() if forced else None)
        agree = (ts.last_indices.cpu().long().view(-1) == otr.last_indices.view(-1)).float().mean().item()
        sims = {n: round(cos(grads[n], otr.grads[n]), 4) for n in grads if grads[n].numel() >= 4096}
        zg = otr.last_z.grad                                   # (B, C, S, S)
        dz_ref = zg.permute(0, 2, 3, 1).reshape(-1, zg.shape[1])
        print("%s l2=%g tv=%g forced=%s: loss %.5f vs %.5f, idx agree %.4f, cos(dz) %.4f, min/max param cos %.4f / %.4f"
              % (kind, l2, tv, forced, loss.item(), otr.last_terms[0], agree, cos(cap["dz"], dz_ref), min(sims.values()), max(sims.values())))
        sys.stdout.flush()


if __name__ == "__main__":
    for kind in ("mixer", "vitgan"):
        for l2, tv in ((0.0, 0.0), (0.1, 0.0), (0.0, 0.5), (0.1, 0.5)):
            run(kind, l2, tv)
