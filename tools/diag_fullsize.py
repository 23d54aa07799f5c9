"""Diagnosis of the config #2 full-size sharding-invariance gap: the 64-prompt step against 4 shards of 16
with controls — (0) the same 64-prompt step twice (run-to-run noise of the fp32 atomics), (a) the shards decode the code indices
of the full run (no VQ flips), (b) split-K pinned to 1 (no batch-size dependent accumulation split), (c) both — and, for each,
where along the path the two computations part: z, code indices, image, cutouts, embeddings, d(embed), d(image), d(z_q), and the
flat gradient arena per parameter group.  Writes gpurun_out/diag_fullsize.json (+ .md).  GPU only; not part of the test suite."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from feed_forward_vqgan_clip_b200 import ops, parallel                      # noqa: E402
from feed_forward_vqgan_clip_b200.clip_vit import CLIP                      # noqa: E402
from feed_forward_vqgan_clip_b200.cutouts import sample_params              # noqa: E402
from feed_forward_vqgan_clip_b200.mixer import Mixer                        # noqa: E402
from feed_forward_vqgan_clip_b200.train_step import TrainStep               # noqa: E402
from feed_forward_vqgan_clip_b200.vqgan import VQModel                      # noqa: E402

DEV = "cuda"
MIXER = dict(input_dim=512, image_size=16, channels=256, patch_size=1, dim=1024, depth=32)
CUTN, B, CUT, WORLD = 8, int(os.environ.get("DIAG_B", 64)), 224, 4


def cos_rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    c = float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-300))
    r = float((a - b).norm() / (b.norm() + 1e-300))
    return round(c, 6), round(r, 6)


def main():
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
    prm["facs"] = torch.rand(N, device=DEV, generator=gd) * 0.1
    prm["noise_raw"] = torch.randn(N, 3, CUT, CUT, device=DEV, generator=gd)
    for k in ("affine_inv", "persp_inv", "sat", "hue"):
        prm[k] = prm[k].to(DEV)
    keep = eng.arena.clone()
    names = [n for n, _ in net.named_parameters()]

    def restore():
        eng.arena.copy_(keep)
        eng.ext_shadow_fresh = False
        eng._shadow_version = None

    def groups():
        """arena slices per parameter group: proj, mixer.1, every 8 layers' token / channel mixing weights, tail"""
        out = {}
        for n in names:
            off = eng.offs[n]
            num = eng.shapes[n].numel()
            if n.startswith("mixer.") and n.split(".")[1].isdigit() and 2 <= int(n.split(".")[1]) <= MIXER["depth"] + 1:
                li = int(n.split(".")[1]) - 2
                kind = "tok" if ".0.fn." in n else ("chan" if ".1.fn." in n else "norm")
                if n.endswith("bias"):
                    kind += "_bias" if kind != "norm" else ""
                key = "L%02d-%02d.%s" % (li // 8 * 8, li // 8 * 8 + 7, kind)
            else:
                key = n
            out.setdefault(key, []).append((off, off + num))
        return out

    def run_full():
        restore()
        ts.debug = {}
        loss = float(ts.step(x, None, prm).item())
        d = ts.debug
        ts.debug = None
        return loss, eng.grad.clone(), ts.last_indices.clone().view(B, -1), d

    def run_shards(force=None):
        acc = torch.zeros_like(eng.grad)
        losses, idx, dbg = [], [], []
        for r in range(WORLD):
            restore()
            lo, hi = parallel.shard_range(B, r, WORLD)
            ps = parallel.shard_cutout_params(prm, CUTN, B, lo, hi)
            ts.force_idx = None if force is None else force[lo:hi].contiguous()
            ts.debug = {}
            losses.append(float(ts.step(x[lo:hi].contiguous(), None, ps).item()))
            ts.force_idx = None
            dbg.append(ts.debug)
            ts.debug = None
            acc += eng.grad
            idx.append(ts.last_indices.clone().view(hi - lo, -1))
        acc /= WORLD
        # re-assemble the intermediates in the global batch's order
        d = {}
        for k in ("z", "img", "dimg", "dzq"):                    # prompt-major
            sc = (1.0 / WORLD) if k.startswith("d") else 1.0
            d[k] = torch.cat([db[k].float().reshape(B // WORLD, -1) for db in dbg]) * sc
        for k in ("patches", "emb", "demb"):                     # cutout-major: row c * B + j
            sc = (1.0 / WORLD) if k.startswith("d") else 1.0
            parts = [db[k].float().reshape(CUTN, B // WORLD, -1) for db in dbg]
            d[k] = torch.cat(parts, dim=1).reshape(N, -1) * sc
        return sum(losses) / WORLD, acc, torch.cat(idx), d

    def compare(tag, full, other, report):
        lf, gf, idf, df = full
        lo_, go, ido, do = other
        rec = {"loss_full": lf, "loss_other": lo_, "idx_same": float((idf == ido).float().mean()),
               "flipped_codes": int((idf != ido).sum()), "prompts_with_flips": int((idf != ido).any(dim=1).sum()),
               "grad": cos_rel(go, gf)}
        for k in ("z", "img", "patches", "emb", "demb", "dimg", "dzq"):
            a, b = do[k].reshape(df[k].numel()), df[k].float().reshape(-1)
            rec[k] = cos_rel(a, b) + (bool(torch.equal(a, b)),)
        # per-prompt: is the d(z_q) gap concentrated on prompts with flipped codes?
        a, b = do["dzq"].reshape(B, -1).double(), df["dzq"].float().reshape(B, -1).double()
        per = ((a - b).norm(dim=1) / (b.norm(dim=1) + 1e-300))
        flips = (idf != ido).any(dim=1)
        rec["dzq_rel_prompts_with_flips"] = float(per[flips].mean()) if flips.any() else None
        rec["dzq_rel_prompts_without_flips"] = float(per[~flips].mean()) if (~flips).any() else None
        rec["groups"] = {k: cos_rel(torch.cat([go[l:h] for l, h in v]), torch.cat([gf[l:h] for l, h in v])) for k, v in groups().items()}
        report[tag] = rec
        print(tag, json.dumps({k: v for k, v in rec.items() if k != "groups"}), flush=True)

    report = {"B": B, "world": WORLD}
    full = run_full()
    compare("0_full_vs_full_again", full, run_full(), report)
    compare("A_shards", full, run_shards(), report)
    compare("a_shards_forced_idx", full, run_shards(force=full[2]), report)
    keep_sp = ops.auto_splits
    ops.auto_splits = lambda *a, **k: 1
    os.environ["FFVC_GEMM_TUNED"] = "0"
    full1 = run_full()
    compare("b0_full_splits1_vs_full", full, full1, report)
    compare("b_shards_splits1", full1, run_shards(), report)
    compare("c_shards_splits1_forced_idx", full1, run_shards(force=full1[2]), report)
    ops.auto_splits = keep_sp
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/diag_fullsize.json", "w") as f:
        json.dump(report, f, indent=1)
    with open("gpurun_out/diag_fullsize.md", "w") as f:
        f.write("| control | idx same | flipped codes (prompts) | loss full / other | grad cos | grad rel | z | img | emb | demb | dimg | dzq |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for tag, r in report.items():
            if not isinstance(r, dict):
                continue
            f.write("| %s | %.5f | %d (%d) | %.6f / %.6f | %.5f | %.4f | %s |\n" % (
                tag, r["idx_same"], r["flipped_codes"], r["prompts_with_flips"], r["loss_full"], r["loss_other"], r["grad"][0],
                r["grad"][1], " | ".join("%.5f / %.4f%s" % (r[k][0], r[k][1], " =" if r[k][2] else "") for k in
                                         ("z", "img", "emb", "demb", "dimg", "dzq"))))
        f.write("\nPer parameter group (cos / rel):\n\n")
        tags = [t for t, r in report.items() if isinstance(r, dict)]
        f.write("| group | " + " | ".join(tags) + " |\n|---|" + "---|" * len(tags) + "\n")
        for gk in report[tags[0]]["groups"]:
            f.write("| %s | " % gk + " | ".join("%.5f / %.4f" % tuple(report[t]["groups"][gk]) for t in tags) + " |\n")


if __name__ == "__main__":
    main()
