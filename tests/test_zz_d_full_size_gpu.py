"""Full-size checks of the hot path on the GPU (BASELINE.json configs[1]: MLP-Mixer 32x1024 -> VQGAN f16/16384 -> 8 cutouts
-> CLIP ViT-B/32, 256x256, 64 prompts) through size-independent properties — the CPU oracle needs minutes per prompt at
this size, so parity at full size is argued from properties of the path instead (the small-size tests compare with the
oracle directly):

  * batch sharding: the step is a mean over independent prompts, so the global-batch gradient equals the average of the
    shard gradients and the VQ indices (integer work) are the same whichever way the batch is cut — this is also the
    data-parallel contract of SURVEY §8e (main.py:627: Horovod averages the ranks' gradients);
  * the quantised latents are rows of the codebook, and no other code is closer (fp32 check on a sample of rows);
  * the synthesised image lies in [0, 1] (clamp_with_grad, main.py:142), the loss in the range of 2*asin(d/2)^2.

The file sorts last on purpose: it allocates ~45 GB.
"""
import math

import pytest
import torch

from feed_forward_vqgan_clip_b200 import parallel
from feed_forward_vqgan_clip_b200.clip_vit import CLIP
from feed_forward_vqgan_clip_b200.cutouts import sample_params
from feed_forward_vqgan_clip_b200.mixer import Mixer
from feed_forward_vqgan_clip_b200.train_step import TrainStep
from feed_forward_vqgan_clip_b200.vqgan import VQModel

pytestmark = pytest.mark.gpu
DEV = "cuda"
MIXER = dict(input_dim=512, image_size=16, channels=256, patch_size=1, dim=1024, depth=32)
CUTN, B, CUT = 8, 64, 224


def _restore(eng, keep):
    """put the mapper's master weights back (the step ends with Adam) and force a fresh bf16 shadow"""
    eng.arena.copy_(keep)
    eng.ext_shadow_fresh = False
    eng._shadow_version = None


def test_config2_full_size_step_is_invariant_to_batch_sharding():
    torch.manual_seed(0)
    net = Mixer(**MIXER)
    vq = VQModel()
    with torch.no_grad():
        vq.quantize.embedding.weight.normal_(0, 1)
    clip = CLIP()
    net, vq, clip = net.to(DEV), vq.to(DEV).eval().requires_grad_(False), clip.to(DEV).eval().requires_grad_(False)
    ts = TrainStep(net, vq, clip, cutn=CUTN, lr=1e-3)
    eng = ts.mix
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(B, 512, generator=g) * 0.45).float().to(DEV)
    prm = sample_params(CUTN * B, CUT, g, with_noise=False)
    gd = torch.Generator(device=DEV).manual_seed(2)
    N = CUTN * B
    prm["facs"] = torch.rand(N, device=DEV, generator=gd) * 0.1                      # main.py:223-225
    prm["noise_raw"] = torch.randn(N, 3, CUT, CUT, device=DEV, generator=gd)
    for k in ("affine_inv", "persp_inv", "sat", "hue"):
        prm[k] = prm[k].to(DEV)

    keep = eng.arena.clone()
    loss_full = float(ts.step(x, None, prm).item())
    g_full = eng.grad.clone()
    idx_full = ts.last_indices.clone().view(B, -1)

    # ---- properties of the full-size step itself
    assert math.isfinite(loss_full) and 0.0 <= loss_full <= 2.0 * (math.pi / 2) ** 2 + 1e-3, loss_full
    assert torch.isfinite(g_full).all() and float(g_full.abs().max()) > 0
    cb = vq.quantize.embedding.weight.detach().float()
    assert int(idx_full.min()) >= 0 and int(idx_full.max()) < cb.shape[0]
    assert idx_full.unique().numel() > 64, "the VQ search collapsed onto a handful of codes"

    # ---- the same 64 prompts in 4 shards of 16
    world = 4
    acc = torch.zeros_like(g_full)
    losses, idx = [], []
    for r in range(world):
        _restore(eng, keep)
        lo, hi = parallel.shard_range(B, r, world)
        ps = parallel.shard_cutout_params(prm, CUTN, B, lo, hi)
        losses.append(float(ts.step(x[lo:hi].contiguous(), None, ps).item()))
        acc += eng.grad
        idx.append(ts.last_indices.clone().view(hi - lo, -1))
    acc /= world
    torch.cuda.synchronize()
    # Every forward kernel is reproducible and batch-invariant (a GEMM output element is the same K-ordered sum whatever the tile
    # configuration; GroupNorm statistics are fixed-order reductions whose partial sums depend on the sample's shape only — round 2),
    # so each prompt's z, code indices, image and embedding are BIT-IDENTICAL however the batch is cut (asserted below on the
    # indices and the loss; profiles/r02_diag_fullsize_after_fix.md for the tensors).  Round 1 failed here with cosine 0.9867:
    # float atomics in the GroupNorm statistics made the image differ by 0.85 % from run to run, and the reference's discontinuous
    # gradient (max-pool arg-max routing, HSV sectors) turned that into a 16 % gradient difference — of the step against ITSELF.
    # What remains is backward-only and not a defect: a shard's loss is the mean over a 4x smaller batch, so its gradients are 4x
    # larger before the average; the cutout backward accumulates its bilinear scatter in 2^-40 fixed point (reproducible, see
    # test_config2_full_size_step_is_reproducible), whose grid is not scale-invariant — d(image) agrees to 1e-12 instead of bit for
    # bit, one bf16 rounding flips here and there, and the decoder's / mapper's activation-gradient chain grows that to the bf16
    # noise floor (measured: d(z_q) cosine 0.99991, gradient cosine 0.99977, rel 0.021).
    assert torch.equal(torch.cat(idx), idx_full)
    mean_loss = sum(losses) / world
    assert abs(mean_loss - loss_full) <= 1e-5 * abs(loss_full), (mean_loss, loss_full)
    a, b = acc.double(), g_full.double()
    cosine = float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-300))
    rel = float((a - b).norm() / (b.norm() + 1e-300))
    assert cosine >= 0.9995 and rel <= 3e-2, (cosine, rel)


def test_config2_full_size_step_is_reproducible():
    """The same 64-prompt step twice: every activation and every activation gradient (image, embeddings, d(image), d(z_q), d(z)) is
    BIT-IDENTICAL — no float atomics anywhere on the activation path (GroupNorm statistics: fixed-order reductions; cutout scatter:
    fixed-point accumulators) — and the flat parameter gradient agrees to the order of fp32 additions in the wgrad split-K
    (relative difference <= 1e-5).  Round 1's step differed from itself by 16 % here."""
    torch.manual_seed(0)
    net = Mixer(**MIXER)
    vq = VQModel()
    with torch.no_grad():
        vq.quantize.embedding.weight.normal_(0, 1)
    clip = CLIP()
    net, vq, clip = net.to(DEV), vq.to(DEV).eval().requires_grad_(False), clip.to(DEV).eval().requires_grad_(False)
    ts = TrainStep(net, vq, clip, cutn=CUTN, lr=1e-3)
    eng = ts.mix
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(B, 512, generator=g) * 0.45).float().to(DEV)
    prm = sample_params(CUTN * B, CUT, g, with_noise=False)
    gd = torch.Generator(device=DEV).manual_seed(2)
    prm["facs"] = torch.rand(CUTN * B, device=DEV, generator=gd) * 0.1
    prm["noise_raw"] = torch.randn(CUTN * B, 3, CUT, CUT, device=DEV, generator=gd)
    keep = eng.arena.clone()
    runs = []
    for _ in range(2):
        _restore(eng, keep)
        ts.debug = {}
        loss = float(ts.step(x, None, prm).item())
        runs.append((loss, ts.debug, eng.grad.clone(), ts.last_indices.clone()))
        ts.debug = None
    (l0, d0, g0, i0), (l1, d1, g1, i1) = runs
    assert torch.equal(i0, i1)
    assert abs(l0 - l1) <= 1e-6 * abs(l0)        # the reported scalar is summed over the cutouts with float atomics (not on the gradient path)
    for k in ("z", "img", "patches", "emb", "demb", "dimg", "dzq", "dz"):
        assert torch.equal(d0[k], d1[k]), k
    rel = float((g0.double() - g1.double()).norm() / g0.double().norm())
    assert rel <= 1e-5, rel


def test_full_size_vq_rows_are_nearest_codebook_rows():
    """ffvc_vq_nearest_tc at the full 16384 latents x 16384 codes: z_q rows are codebook rows, and on a sample of rows no
    other code is closer in fp32 than the one chosen (up to the fp32 rounding of the distance expression, main.py:135-137)."""
    torch.manual_seed(3)
    vq = VQModel()
    with torch.no_grad():
        vq.quantize.embedding.weight.normal_(0, 1)
    vq = vq.to(DEV).eval().requires_grad_(False)
    dec = vq.engine()
    cb = vq.quantize.embedding.weight.detach().float()
    lo, hi = float(cb.min()), float(cb.max())
    z = torch.randn(B * 256, 256, device=DEV) * 1.5
    zq, idx, zc = dec.quantize(z, lo, hi)
    idx = idx.view(-1).long()
    assert torch.equal(zq.view(-1, 256), cb[idx].to(torch.bfloat16))      # the decoder reads bf16 rows of the codebook
    assert torch.equal(zc, z.clamp(lo, hi))
    rows = torch.arange(0, z.shape[0], 16, device=DEV)
    zs = z[rows].clamp(lo, hi).double()
    d = (zs * zs).sum(1, keepdim=True) + (cb.double() ** 2).sum(1)[None] - 2.0 * zs @ cb.double().t()
    best = d.min(dim=1).values
    chosen = d.gather(1, idx[rows, None])[:, 0]
    assert float((chosen - best).max()) <= 2e-3, float((chosen - best).max())     # a different pick must be a numerical tie
    assert (d.argmin(dim=1) == idx[rows]).float().mean().item() >= 0.999


def test_config2_full_architecture_step_vs_oracle():
    """Parity at BASELINE config #2's real architecture (not a scaled-down stand-in): Mixer 32 x 1024, the full VQGAN f16/16384
    decoder, CLIP ViT-B/32, 256 x 256, 8 cutouts, 8 prompts — the CUDA step against the fp32 oracle (plain torch on the GPU, TF32
    off) on the same bf16-rounded weights, embeddings and augmentation parameters (tests/fullsize_parity.py).

    The stated tolerance at this depth (bf16 compute, 32 + ~60 + 12 layers, against fp32):
      forward   L2-relative error <= 2e-2 for z, the image and the embeddings; loss within 1e-3; >= 95 % of the code indices
      backward  every stage's vector-Jacobian product, evaluated at the same inputs with the same cotangent: cosine >= 0.99
                (measured 0.9999 cutouts + CLIP + loss, 0.9990 decoder, 0.9998 mapper — profiles/r02_parity_fullsize.md)
      end to end  each path differentiates at ITS OWN image, and the reference's gradient is discontinuous in the image (max-pool
                arg-max, HSV sectors, clamps; main.py:218,171-172,142): the fp32 oracle against ITSELF, with its image perturbed by
                noise of the size of the CUDA path's image error (0.8 %), keeps only cosine ~0.93 of d(image).  The end-to-end
                gradient must be at least as close to the oracle as the oracle is to its own perturbed evaluation, and >= 0.97."""
    from fullsize_parity import report
    r = report(nb=8)
    assert abs(r["loss_cuda"] - r["loss_oracle"]) <= 1e-3 * abs(r["loss_oracle"]), r
    assert r["idx_agreement"] >= 0.95, r["idx_agreement"]
    for k in ("z", "img", "emb"):
        assert r[k]["rel"] <= 2e-2 and r[k]["cos"] >= 0.9995, (k, r[k])
    for k in ("stage_dimg", "stage_dzq", "stage_dparams"):
        assert r[k]["cos"] >= 0.99, (k, r[k])
    assert r["stage_dparams"]["worst_cos"] >= 0.99, r["stage_dparams"]
    assert r["e2e_grad"]["cos"] >= max(0.97, r["sensitivity"]["cos"]), (r["e2e_grad"], r["sensitivity"])
    assert r["e2e_grad"]["worst_cos"] >= 0.97, r["e2e_grad"]
