"""Parity of the CUDA train step with the fp32 oracle at BASELINE config #2's REAL architecture (Mixer 32 x 1024, the full VQGAN
f16/16384 decoder, CLIP ViT-B/32, 256 x 256, 8 cutouts), with the oracle (plain torch, fp32, TF32 off) running on the GPU so that a
batch of prompts finishes in seconds.  Test infrastructure: imports oracle/.

What is compared (report(); asserted by tests/test_zz_d_full_size_gpu.py, written to profiles/ by `python tests/fullsize_parity.py`):

  forward   z (mapper), code indices, image (decoder on the same codes), embeddings (cutouts + CLIP on the CUDA image), loss
  backward  stage by stage, each stage's vector-Jacobian product at the SAME inputs and the SAME cotangent:
              d(image)  : oracle cutouts -> CLIP -> loss differentiated at the CUDA path's image
              d(z_q)    : oracle decoder vjp at the same codes with the CUDA path's d(image)
              d(params) : oracle mapper vjp with the CUDA path's d(z)
            and end to end: every parameter gradient against the oracle's own backward (codes forced to the CUDA path's).

Why stage-wise: the reference's gradient is DISCONTINUOUS in the image — (avg + max) pooling routes half of every pooled
pixel's gradient to the arg-max of its window (main.py:218), the HSV jitter switches branch at every sector boundary, clamps
switch at 0 / 1 — so two exact evaluations whose images differ by bf16 rounding (0.5 - 1 % after ~60 layers) have visibly
different gradients although each is the exact gradient of its own forward.  `sensitivity` measures that on the ORACLE ALONE: the
fp32 oracle against itself with its image perturbed by noise of the size of the CUDA path's image error."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DEV = "cuda"
MIXER = dict(input_dim=512, image_size=16, channels=256, patch_size=1, dim=1024, depth=32)
CUTN, CUT = 8, 224


def _cos_rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-300)), float((a - b).norm() / (b.norm() + 1e-300))


def _maxrel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-300))


def report(nb=8, seed=3):
    import oracle.clip_vit as oclip
    import oracle.cutouts as ocut
    import oracle.loss as oloss
    import oracle.mixer as omix
    import oracle.vqgan as ovq
    from feed_forward_vqgan_clip_b200.clip_vit import CLIP
    from feed_forward_vqgan_clip_b200.cutouts import sample_params
    from feed_forward_vqgan_clip_b200.mixer import Mixer
    from feed_forward_vqgan_clip_b200.train_step import TrainStep
    from feed_forward_vqgan_clip_b200.vqgan import VQModel

    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        def r16(sd):
            return {k: (v.to(torch.bfloat16).float() if v.dim() >= 2 else v.clone()) for k, v in sd.items()}
        sd_m = r16(omix.init_mixer_state_dict(512, 16, 256, 1024, 32, seed=0))
        sd_v, sd_c = r16(ovq.init_vqgan_state_dict(seed=1)), r16(oclip.init_clip_state_dict(seed=2))
        net = Mixer(**MIXER)
        net.load_state_dict(sd_m)
        vq = VQModel()
        vq.load_state_dict(sd_v)
        clip = CLIP()
        clip.visual.load_state_dict(sd_c)
        net, vq, clip = net.to(DEV), vq.to(DEV).eval().requires_grad_(False), clip.to(DEV).eval().requires_grad_(False)
        g = torch.Generator().manual_seed(seed)
        x = (torch.randn(nb, 512, generator=g) * 0.45).to(torch.bfloat16).float().to(DEV)
        prm = sample_params(CUTN * nb, CUT, g)
        prm_d = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in prm.items()}
        ts = TrainStep(net, vq, clip, cutn=CUTN, lr=1e-3)
        ts.debug = {}
        loss = float(ts.step(x, None, prm).item())
        dbg, ts.debug = ts.debug, None
        idx = ts.last_indices.long().view(-1)
        eng = net.engine()
        S, C, H = 16, 256, 256
        out = {"prompts": nb, "loss_cuda": loss}

        # ---------------------------------------------------------------- oracle on the GPU, fp32
        sd_m = {k: v.to(DEV) for k, v in sd_m.items()}
        sd_v = {k: v.to(DEV) for k, v in sd_v.items()}
        sd_c = {k: v.to(DEV) for k, v in sd_c.items()}
        cb = sd_v["quantize.embedding.weight"]
        z_lo, z_hi = float(cb.min()), float(cb.max())
        params = {k: v.clone().requires_grad_(True) for k, v in sd_m.items()}

        # forward: mapper
        z_ref = omix.mixer_forward(params, x, S, C)                                   # (B, C, S, S)
        z_tok_ref = z_ref.permute(0, 2, 3, 1).reshape(nb * S * S, C)
        out["z"] = dict(zip(("cos", "rel"), _cos_rel(dbg["z"], z_tok_ref.detach())), maxrel=_maxrel(dbg["z"], z_tok_ref.detach()))
        zt = z_tok_ref.detach().clamp(z_lo, z_hi)
        own = torch.cat([((zt[i:i + 512] ** 2).sum(1, keepdim=True) + (cb ** 2).sum(1)[None] - 2 * zt[i:i + 512] @ cb.t()).argmin(1)
                         for i in range(0, zt.shape[0], 512)])
        out["idx_agreement"] = float((own == idx).float().mean())

        # end to end: the oracle's own backward on the CUDA path's codes
        zc = ovq.clamp_with_grad(z_ref, z_lo, z_hi)
        xr, _ = ovq.synth(sd_v, zc, return_indices=True, force_idx=idx.view(nb, S, S))
        xr.retain_grad()
        emb = oclip.encode_image(sd_c, ocut.make_cutouts(xr, CUTN, prm_d, CUT, normalize=True)).float()
        loss_ref = oloss.spherical_dist_loss(emb, x, CUTN)
        loss_ref.backward()
        out["loss_oracle"] = float(loss_ref)
        img_ref = xr.detach().permute(0, 2, 3, 1)                                    # NHWC like the engine's image
        out["img"] = dict(zip(("cos", "rel"), _cos_rel(dbg["img"], img_ref)), maxrel=_maxrel(dbg["img"], img_ref))
        e2e, worst, worst_name = {}, 1.0, None
        flat_a, flat_b = [], []
        for (n, p), gv in zip(net.named_parameters(), eng.grad_views):
            a, b = gv.detach().float(), params[n].grad
            flat_a.append(a.flatten())
            flat_b.append(b.flatten())
            if p.numel() >= 65536:
                c = _cos_rel(a, b)[0]
                e2e[n] = c
                if c < worst:
                    worst, worst_name = c, n
        out["e2e_grad"] = dict(zip(("cos", "rel"), _cos_rel(torch.cat(flat_a), torch.cat(flat_b))), worst_cos=worst, worst_param=worst_name)
        out["e2e_dimg"] = dict(zip(("cos", "rel"), _cos_rel(dbg["dimg"], xr.grad.permute(0, 2, 3, 1))))
        grads_e2e = {k: p.grad.clone() for k, p in params.items()}

        # ---------------------------------------------------------------- stage-wise vector-Jacobian products
        # (1) cutouts -> CLIP -> loss at the CUDA path's image
        img_c = dbg["img"].float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
        emb1 = oclip.encode_image(sd_c, ocut.make_cutouts(img_c, CUTN, prm_d, CUT, normalize=True)).float()
        l1 = oloss.spherical_dist_loss(emb1, x, CUTN)
        l1.backward()
        out["loss_at_cuda_image"] = float(l1)
        out["emb"] = dict(zip(("cos", "rel"), _cos_rel(dbg["emb"], emb1.detach())), maxrel=_maxrel(dbg["emb"], emb1.detach()))
        dimg_ref = img_c.grad.permute(0, 2, 3, 1)
        out["stage_dimg"] = dict(zip(("cos", "rel"), _cos_rel(dbg["dimg"], dimg_ref)))
        # the oracle against ITSELF: the same stage at an image perturbed by noise of the size of the CUDA path's image error
        img_err = out["img"]["rel"]
        gn = torch.Generator(device=DEV).manual_seed(11)
        noise = torch.randn(img_ref.shape, device=DEV, generator=gn)
        pert = (img_ref + noise * (img_err * img_ref.norm() / noise.norm())).clamp(0, 1)
        outs = []
        for im in (img_ref, pert):
            t = im.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
            e = oclip.encode_image(sd_c, ocut.make_cutouts(t, CUTN, prm_d, CUT, normalize=True)).float()
            oloss.spherical_dist_loss(e, x, CUTN).backward()
            outs.append(t.grad.clone())
        out["sensitivity"] = dict(zip(("cos", "rel"), _cos_rel(outs[1], outs[0])), image_perturbation_rel=img_err,
                                  what="fp32 oracle d(image) at its own image vs at that image + noise of the CUDA image error's size")
        del outs, emb1, l1

        # (2) decoder vjp at the same codes, cotangent = the CUDA path's d(image)
        zq_in = cb[idx].view(nb, S, S, C).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
        xr2 = ovq.clamp_with_grad(ovq.decode(sd_v, zq_in).add(1).div(2), 0, 1)
        xr2.backward(dbg["dimg"].permute(0, 3, 1, 2).contiguous())
        dzq_ref = zq_in.grad.permute(0, 2, 3, 1).reshape(nb * S * S, C)
        out["stage_dzq"] = dict(zip(("cos", "rel"), _cos_rel(dbg["dzq"], dzq_ref)))
        del xr2

        # (3) mapper vjp, cotangent = the CUDA path's d(z)
        for p in params.values():
            p.grad = None
        z3 = omix.mixer_forward(params, x, S, C)
        z3.backward(dbg["dz"].view(nb, S, S, C).permute(0, 3, 1, 2).contiguous())
        worst3, worst3_name, fa, fb = 1.0, None, [], []
        for (n, p), gv in zip(net.named_parameters(), eng.grad_views):
            a, b = gv.detach().float(), params[n].grad
            fa.append(a.flatten())
            fb.append(b.flatten())
            if p.numel() >= 65536:
                c = _cos_rel(a, b)[0]
                if c < worst3:
                    worst3, worst3_name = c, n
        out["stage_dparams"] = dict(zip(("cos", "rel"), _cos_rel(torch.cat(fa), torch.cat(fb))), worst_cos=worst3, worst_param=worst3_name)
        out["e2e_per_param_min5"] = sorted(e2e.items(), key=lambda kv: kv[1])[:5]
        del grads_e2e
        return out
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32


def to_markdown(r):
    L = ["# Full-size parity, config #2 real architecture (%d prompts, oracle fp32 on the GPU, TF32 off)" % r["prompts"], "",
         "loss: CUDA %.6f, oracle %.6f (rel %.2e); oracle loss at the CUDA image %.6f" % (
             r["loss_cuda"], r["loss_oracle"], abs(r["loss_cuda"] - r["loss_oracle"]) / abs(r["loss_oracle"]), r["loss_at_cuda_image"]),
         "code indices equal to the oracle's own arg-min: %.4f" % r["idx_agreement"], "",
         "| quantity | cos | rel (L2) | max abs err / max abs ref |", "|---|---|---|---|"]
    for k in ("z", "img", "emb"):
        L.append("| %s (forward) | %.6f | %.4f | %.4f |" % (k, r[k]["cos"], r[k]["rel"], r[k]["maxrel"]))
    for k, name in (("stage_dimg", "d(image): cutouts + CLIP + loss vjp at the CUDA image"),
                    ("stage_dzq", "d(z_q): decoder vjp, CUDA d(image) as cotangent"),
                    ("stage_dparams", "d(params): mapper vjp, CUDA d(z) as cotangent"),
                    ("e2e_dimg", "d(image) end to end (each path at its own image)"),
                    ("e2e_grad", "d(params) end to end"),
                    ("sensitivity", "ORACLE vs ORACLE: d(image) under an image perturbation of rel %.4f" % r["sensitivity"]["image_perturbation_rel"])):
        L.append("| %s | %.6f | %.4f | |" % (name, r[k]["cos"], r[k]["rel"]))
    L += ["", "worst parameter (>= 65536 elements): stage-wise %.5f (%s), end to end %.5f (%s)" % (
        r["stage_dparams"]["worst_cos"], r["stage_dparams"]["worst_param"], r["e2e_grad"]["worst_cos"], r["e2e_grad"]["worst_param"])]
    return "\n".join(L) + "\n"


if __name__ == "__main__":
    r = report(int(os.environ.get("PARITY_B", 8)))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/r02_parity_fullsize.json", "w") as f:
        json.dump(r, f, indent=1)
    with open("gpurun_out/r02_parity_fullsize.md", "w") as f:
        f.write(to_markdown(r))
    print(to_markdown(r))
