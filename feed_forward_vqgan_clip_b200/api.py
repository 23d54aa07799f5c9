"""The reference's Python call surface for the train-step path, backed by the B200 engines.

    build_model(config)                      main.py:448-502   (model_type mlp_mixer, vitgan, simple_vitgan, xtransformer)
    load_vqgan_model(config_path, ckpt)      main.py:84-103
    load_model(path)                         main.py:1273-1290 (checkpoint.th dictionaries and legacy pickled mappers)
    load_clip_model(name, path)              main.py:1308-1333 (OpenAI ViT-B/32 and OpenCLIP ViT-B-32 architectures)
    MakeCutouts / synth / clamp_with_grad / vector_quantize      main.py:105-229
    LPIPS / normalize_tensor                 main.py:30-31,532-537,776-787 (diversity term; `LPIPS().net(x)` -> the five VGG16 taps)
    train_step(net, vq, perceptor, ...)      the body of main.py:729-837 as one fused object (TrainStep)
    CheckpointWriter(train_step, folder)     main.py:904-911 (checkpoint.th / checkpoint_ema.th / opt.th), asynchronous, optionally sharded

`config` may be any mapping / attribute object with the keys of configs/example.yaml (an OmegaConf DictConfig or a
plain dict both work); unknown keys are ignored, optional keys defaulted exactly like main.py:450-457,466,496-498.
"""
import os

import torch

from .checkpoint import CheckpointWriter, load_sharded  # noqa: F401
from .clip_vit import CLIP, VIT_B32
from .cutouts import MakeCutouts, sample_params  # noqa: F401
from .lpips import LPIPS, normalize_tensor  # noqa: F401
from .mixer import Mixer
from .simple_vitgan_mapper import SimpleGenerator as SimpleVitGAN
from .vitgan_mapper import Generator as VitGAN
from .xtransformer import XTransformer
from .train_step import FusedAdam, TrainStep  # noqa: F401
from .vqgan import F16_16384, VQModel, clamp_with_grad, synth, vector_quantize  # noqa: F401

CLIP_SIZE = {"ViT-B/32": 224, "ViT-B-32": 224, "ViT-B-32-quickgelu": 224}       # main.py:53-66 (ViT-B/32 family)
CLIP_DIM = {"ViT-B/32": 512, "ViT-B-32": 512, "ViT-B-32-quickgelu": 512}        # main.py:67-80


def _get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    if hasattr(cfg, "get"):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def load_config(path):
    """YAML -> dict (stands in for OmegaConf.load, main.py:506; omegaconf is not a dependency here)."""
    import yaml
    with open(path) as f:
        return yaml.safe_load(f)


def build_model(config, vq_channels=256):
    clip_model = _get(config, "clip_model", "ViT-B/32")
    clip_dim = _get(config, "clip_dim", CLIP_DIM.get(clip_model.split(":")[-1] if ":" in clip_model else clip_model, 512))
    vq_image_size = _get(config, "vq_image_size", 16)
    noise_dim = _get(config, "noise_dim", 0) or 0
    model_type = _get(config, "model_type", "mlp_mixer")
    if model_type == "mlp_mixer":
        return Mixer(input_dim=clip_dim + noise_dim, image_size=vq_image_size, channels=vq_channels, patch_size=1,
                     dim=_get(config, "dim"), depth=_get(config, "depth"), dropout=_get(config, "dropout", 0))
    if model_type == "vitgan":                                                   # main.py:459-468
        return VitGAN(initialize_size=vq_image_size // 8, dropout=_get(config, "dropout", 0), out_channels=vq_channels,
                      input_dim=clip_dim + noise_dim, dim=_get(config, "dim"), num_heads=_get(config, "num_heads", 6),
                      blocks=_get(config, "depth"))
    if model_type == "simple_vitgan":                                            # main.py:469-478
        return SimpleVitGAN(size=vq_image_size, dropout=_get(config, "dropout", 0), out_channels=vq_channels,
                            input_dim=clip_dim + noise_dim, dim=_get(config, "dim"), num_heads=_get(config, "num_heads", 6),
                            blocks=_get(config, "depth"))
    if model_type == "xtransformer":                                             # main.py:488-499
        return XTransformer(input_dim=clip_dim + noise_dim, image_size=vq_image_size, channels=vq_channels, dim=_get(config, "dim"),
                            depth=_get(config, "depth"), heads=_get(config, "num_heads", 6),
                            initial_proj=_get(config, "initial_proj", True), add_input=_get(config, "add_input", False))
    raise ValueError("model_type should be 'vitgan' or  'mlp_mixer' or 'xtransformer'")      # main.py:501


def load_model(model_path, vq_channels=256):
    """main.py:1273-1290 (what `test`, `evaluate` and predict.py call): a `checkpoint.th` dictionary {"config", "state_dict", ...}
    — the form train() writes (main.py:904) and the published models ship in — or, for backward compatibility, a pickled module
    instance carrying `.config` (unpickling that needs the class it was saved from to be importable).  Returns this package's mapper
    with those weights and `net.config` set."""
    ckpt = torch.load(model_path, map_location="cpu", weights_only=False)
    if isinstance(ckpt, dict):
        config, sd = ckpt["config"], ckpt["state_dict"]
    else:
        config, sd = ckpt.config, ckpt.state_dict()
    net = build_model(config, vq_channels=vq_channels)
    net.load_state_dict(sd)
    net.config = config
    for k in ("step", "epoch"):
        if isinstance(ckpt, dict) and k in ckpt:
            setattr(net, k, ckpt[k])
    return net


def load_vqgan_model(config_path=None, checkpoint_path=None):
    """VQModel branch of main.py:84-103 (the only branch BASELINE's configs use).  The architecture of
    vqgan_imagenet_f16_16384.yaml is built in; the checkpoint is optional (random init without it)."""
    model = VQModel(F16_16384)
    model.eval().requires_grad_(False)
    if checkpoint_path and os.path.exists(checkpoint_path):
        model.init_from_ckpt(checkpoint_path)
    return model


def _read_state_dict(path):
    """weights from a plain state_dict file, a {"state_dict": ...} checkpoint, a pickled module, or a TorchScript archive (the
    form OpenAI publishes ViT-B/32 in; clip.load(..., jit=False) rebuilds the model from its state_dict the same way)"""
    try:
        obj = torch.load(path, map_location="cpu", weights_only=False)
    except Exception:
        obj = torch.jit.load(path, map_location="cpu")
    if hasattr(obj, "state_dict") and callable(obj.state_dict):
        return obj.state_dict()
    return obj.get("state_dict", obj)


def load_clip_model(model_type="ViT-B/32", path=None):
    if model_type.startswith("open_clip:"):
        arch = model_type.split(":")[1]
        act = "quick_gelu" if "quickgelu" in arch else "gelu"       # OpenCLIP ViT-B-32 = exact GELU
        if not arch.startswith("ViT-B-32"):
            raise NotImplementedError(arch)
    elif model_type == "ViT-B/32":
        act = "quick_gelu"
    else:
        raise NotImplementedError("perceptor %r (only the ViT-B/32 family is on the benchmark path)" % model_type)
    from .clip_text import TEXT_B32
    model = CLIP(VIT_B32, act=act, text_cfg=TEXT_B32)
    if path and os.path.exists(path):
        sd = _read_state_dict(path)
        vis = {k[len("visual."):]: v.float() for k, v in sd.items() if k.startswith("visual.")}
        model.visual.load_state_dict(vis)
        txt = {k: v.float() for k, v in sd.items() if k in model.text.state_dict()}
        model.text.load_state_dict(txt, strict=False)
        if "logit_scale" in sd:                                       # used by the evaluation block (main.py:700,1192)
            with torch.no_grad():
                model.logit_scale.copy_(sd["logit_scale"].float())
    return model.eval().requires_grad_(False)


@torch.no_grad()
def generate(net, vq, inp_feats, return_indices=False):
    """Inference path of `test` / `predict` (main.py:1056-1059, predict.py:113-117): mapper -> clamp -> VQ -> decode,
    forward only, every kernel from libffvc_sm100.so.  inp_feats: (B, clip_dim [+ noise_dim]) fp32 CUDA tensor.
    Returns the generated images (B, 3, 16*S, 16*S) in [0, 1] (what the reference hands to `make_grid`)."""
    mix, dec = net.engine(), vq.engine()
    x = inp_feats.contiguous().float()
    B = x.shape[0]
    z, _ = mix.forward(x)                                    # [B*S*S, C] fp32 token-major
    lo, hi = float(dec.codebook.min()), float(dec.codebook.max())      # main.py:1057-1058: clamp to the codebook range
    zq, idx, _ = dec.quantize(z, lo, hi)
    img, _ = dec.forward(zq.view(B, mix.S, mix.S, mix.C))    # [B, H, W, 3] fp32, already (x+1)/2 clamped to [0,1]
    img = img.permute(0, 3, 1, 2)
    return (img, idx.view(B, mix.S, mix.S)) if return_indices else img


def train_step(net, vq, perceptor, config=None, **kw):
    """Build the fused step object for `net` (Mixer), `vq` (VQModel) and `perceptor` (CLIP), all on the same GPU."""
    cfg = config or {}
    return TrainStep(net, vq, perceptor, cutn=_get(cfg, "cutn", 8), lr=_get(cfg, "lr", 1e-3),
                     target_loss_coef=_get(cfg, "target_loss_coef", 1.0), l2_coef=_get(cfg, "l2_coef", 0.0) or 0.0,
                     tv_coef=_get(cfg, "tv_coef", 0.0) or 0.0, diversity_coef=_get(cfg, "diversity_coef", 0.0) or 0.0,
                     repeat=_get(cfg, "repeat", 1) or 1, clip_grad_norm=_get(cfg, "clip_grad_norm", None),
                     scheduler=_get(cfg, "scheduler", None), use_ema=bool(_get(cfg, "use_ema", False)),
                     total_steps=kw.pop("total_steps", None) or _get(cfg, "max_steps", 0) or 0,      # main.py:704-705: T_max = config.max_steps
                     ema_decay=_get(cfg, "ema_decay", 0.995),
                     diversity_mode=_get(cfg, "diversity_mode", "between_same_prompts"),
                     input_loss=bool(_get(cfg, "input_loss", False)), input_loss_coef=_get(cfg, "input_loss_coef", 1),   # main.py:690-691
                     normalize_input=bool(_get(cfg, "normalize_input", False)),                                      # main.py:696
                     noise_dim=_get(cfg, "noise_dim", 0) or 0, nb_noise=_get(cfg, "nb_noise", None), **kw)           # main.py:457,649
    # (`tv_exponent`, main.py:699, is read by the reference and never used: tv_loss takes no exponent, main.py:423-428)
