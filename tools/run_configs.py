#!/usr/bin/env python
"""Full-size sanity run of the BASELINE.json configurations #2..#5 on ONE GPU (reduced per-GPU batch where memory needs it):
one warm-up + `--steps` timed eager train steps each; reports ms/step, prompts/s, loss and peak memory.  Not a bench line —
it checks that every mapper / image size / loss-term combination runs at its real dimensions.
    python tools/run_configs.py [--steps 2] [--only 3]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from feed_forward_vqgan_clip_b200 import api  # noqa: E402
from feed_forward_vqgan_clip_b200.lpips import LpipsVGG16  # noqa: E402
from feed_forward_vqgan_clip_b200.train_step import TrainStep  # noqa: E402

CONFIGS = {
    2: dict(name="#2 mlp_mixer 32x1024, ViT-B/32, 256x256", cfg=dict(model_type="mlp_mixer", dim=1024, depth=32, vq_image_size=16), B=64),
    3: dict(name="#3 vitgan 32x1024 heads 6, ViT-B/32, 256x256", cfg=dict(model_type="vitgan", dim=1024, depth=32, vq_image_size=16), B=64),
    4: dict(name="#4 xtransformer 256x16 heads 6, ViT-B/32, 512x512", cfg=dict(model_type="xtransformer", dim=256, depth=16, vq_image_size=32), B=16),
    5: dict(name="#5 mlp_mixer 32x1024, open_clip ViT-B-32 (exact GELU), 512x512, diversity + TV, repeat 2",
            cfg=dict(model_type="mlp_mixer", dim=1024, depth=32, vq_image_size=32, clip_model="open_clip:ViT-B-32"), B=8,
            kw=dict(tv_coef=0.1, diversity_coef=0.1, repeat=2), lpips=True),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--only", type=int, default=0)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    for k, c in CONFIGS.items():
        if args.only and k != args.only:
            continue
        torch.manual_seed(0)
        t0 = time.time()
        try:
            net = api.build_model(c["cfg"]).to(dev)
            vq = api.load_vqgan_model()
            with torch.no_grad():
                vq.quantize.embedding.weight.normal_(0, 1)
            vq = vq.to(dev).eval().requires_grad_(False)
            clip = api.load_clip_model(c["cfg"].get("clip_model", "ViT-B/32")).to(dev).eval().requires_grad_(False)
            kw = dict(c.get("kw", {}))
            if c.get("lpips"):
                kw["lpips_net"] = LpipsVGG16().to(dev).eval().requires_grad_(False)
            ts = TrainStep(net, vq, clip, cutn=8, lr=1e-3, **kw)
            B = c["B"]
            x = (torch.randn(B, 512, generator=torch.Generator().manual_seed(1)) * 0.45).to(dev)
            torch.cuda.reset_peak_memory_stats()
            loss = ts.step(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                loss = ts.step(x)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            prompts = B * kw.get("repeat", 1)
            out = dict(config=c["name"], per_gpu_batch=B, images_per_step=prompts, ms_per_step=round(ms, 2),
                       images_per_s=round(prompts / ms * 1e3, 1), loss=round(float(loss), 5), aux_loss=[round(v, 5) for v in ts.aux_loss.tolist()],
                       params_M=round(sum(p.numel() for p in net.parameters()) / 1e6, 1),
                       peak_mem_GB=round(torch.cuda.max_memory_allocated() / 2 ** 30, 1), setup_s=round(time.time() - t0, 1), ok=True)
            assert loss == loss, "loss is NaN"
        except Exception as e:  # noqa: BLE001
            import traceback
            out = dict(config=c["name"], ok=False, error=repr(e)[:300], tb=traceback.format_exc()[-800:])
        print(json.dumps(out))
        sys.stdout.flush()
        del out
        for name in ("ts", "net", "vq", "clip"):
            locals().pop(name, None)
        import gc
        gc.collect()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
