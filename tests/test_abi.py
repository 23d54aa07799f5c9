"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/ffvc.h declares, the
ctypes struct mirrors the C struct, and the product fails loudly without CUDA (no fallback)."""
import ctypes as C
import os

import pytest
import torch

from feed_forward_vqgan_clip_b200 import _lib


def test_library_exports_every_declared_symbol():
    decls = _lib.header_declarations()
    assert len(decls) >= 40
    lib = _lib.load()
    for name in decls:
        assert hasattr(lib, name), name
    assert lib.ffvc_arch() == 100
    assert lib.ffvc_sizeof(b"ffvc_gemm_params") == C.sizeof(_lib.GemmParams)


def test_no_cpu_fallback():
    from feed_forward_vqgan_clip_b200 import ops
    from feed_forward_vqgan_clip_b200.mixer import Mixer
    with pytest.raises(RuntimeError):
        ops.call("cast_f32_bf16", torch.zeros(4), torch.zeros(4, dtype=torch.bfloat16), 4)
    net = Mixer(32, 4, 16, 1, 32, 1)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 32))


def test_mixer_state_dict_contract_and_seeded_init_match_reference_golden():
    """Same keys / shapes as the reference module and — because construction order is mirrored — the same
    initial weights under the same seed (tests/golden/mixer.pt was produced by the reference with manual_seed(0))."""
    from feed_forward_vqgan_clip_b200.mixer import Mixer
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "mixer.pt"))
    torch.manual_seed(0)
    net = Mixer(**gold["tiny"]["cfg"])
    sd = net.state_dict()
    ref = gold["tiny"]["state_dict"]
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert sd[k].shape == ref[k].shape, k
        assert torch.equal(sd[k], ref[k]), k
    big = Mixer(512, 16, 256, 1, 128, 8)
    assert sum(p.numel() for p in big.parameters()) == 38948480


def test_vqgan_and_clip_state_dict_keys_match_oracle_layout():
    import oracle.clip_vit as oclip
    import oracle.vqgan as ovq
    from feed_forward_vqgan_clip_b200.clip_vit import VisualTransformer
    from feed_forward_vqgan_clip_b200.vqgan import VQModel
    small = dict(ch=64, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(16,), resolution=32, z_channels=64, out_ch=3,
                 embed_dim=64, n_embed=128)
    vq = VQModel(small)
    ref = ovq.init_vqgan_state_dict(small)
    assert set(vq.state_dict().keys()) == set(ref.keys())
    for k, v in vq.state_dict().items():
        assert v.shape == ref[k].shape, k
    cfg = dict(input_resolution=64, patch_size=32, width=64, layers=2, heads=1, output_dim=32)
    vis = VisualTransformer(**cfg)
    refc = oclip.init_clip_state_dict(cfg)
    assert set(vis.state_dict().keys()) == set(refc.keys())
    for k, v in vis.state_dict().items():
        assert v.shape == refc[k].shape, k


def test_vitgan_state_dict_contract_and_seeded_init_match_reference_golden():
    from feed_forward_vqgan_clip_b200.vitgan_mapper import Generator
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "vitgan.pt"))
    torch.manual_seed(4)
    net = Generator(**gold["cfg"])
    sd, ref = net.state_dict(), gold["state_dict"]
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert sd[k].shape == ref[k].shape and torch.equal(sd[k], ref[k]), k
    # full-size known answers: 415,078,530 parameters, to_qkv (3060, 1024), w_out (1024, 1020)
    shapes = dict(gold["keys_32x1024"])
    assert gold["count_32x1024"] == 415078530
    assert shapes["Transformer_Encoder.blocks.0.attn.to_qkv.weight"] == (3060, 1024)
    assert shapes["Transformer_Encoder.blocks.0.attn.w_out.weight"] == (1024, 1020)


def test_simple_vitgan_state_dict_contract_and_seeded_init_match_reference_golden():
    """model_type simple_vitgan (main.py:469-478): same keys, shapes, construction order (=> same seeded init) as the reference's
    vitgan.SimpleGenerator; build_model routes to it with the reference's argument mapping."""
    from feed_forward_vqgan_clip_b200 import api
    from feed_forward_vqgan_clip_b200.simple_vitgan_mapper import SimpleGenerator
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "simple_vitgan.pt"))
    torch.manual_seed(6)
    net = SimpleGenerator(**gold["cfg"])
    sd, ref = net.state_dict(), gold["state_dict"]
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert sd[k].shape == ref[k].shape and torch.equal(sd[k], ref[k]), k
    big = api.build_model(dict(model_type="simple_vitgan", clip_model="ViT-B/32", vq_image_size=16, noise_dim=0, dim=256, depth=1,
                               dropout=0))
    assert [(k, tuple(v.shape)) for k, v in big.state_dict().items()] == [(k, tuple(s)) for k, s in gold["keys_16x256"]]
    with pytest.raises(ValueError):
        api.build_model(dict(model_type="no_such_mapper", dim=8, depth=1))


def test_option_switches_and_env_override():
    lib = _lib.load()
    assert lib.ffvc_get_option(b"no_such_option") == -1
    for name in (b"ln_fwd_v2", b"ln_bwd_v2", b"pool_v2", b"gn_ring", b"halo_epi16"):
        old = lib.ffvc_get_option(name)
        assert old in (0, 1, 2)
        assert lib.ffvc_set_option(name, 0) == old and lib.ffvc_get_option(name) == 0
        lib.ffvc_set_option(name, old)


def test_fused_adam_state_dict_is_torch_adam_compatible():
    """opt.th (main.py:593-596,911) round trip: torch.optim.Adam loads FusedAdam's dict and vice versa (host logic only)."""
    from types import SimpleNamespace
    from feed_forward_vqgan_clip_b200.train_step import FusedAdam
    a, b = torch.nn.Parameter(torch.randn(2, 4)), torch.nn.Parameter(torch.randn(8))
    arena = torch.zeros(16)
    arena[:8], arena[8:] = a.data.view(-1), b.data
    a.data, b.data = arena[:8].view(2, 4), arena[8:]
    eng = SimpleNamespace(arena=arena, grad=torch.zeros(16), shadow=None, total=16, dev=torch.device("cpu"), params=[a, b])
    opt = FusedAdam(eng, lr=3e-4)
    assert opt.state_dict()["state"] == {}                        # like a fresh torch optimizer
    opt.m.normal_()
    opt.v.uniform_()
    opt.hyper[8] = 3
    sd = opt.state_dict()
    assert sd["state"][0]["exp_avg"].shape == (2, 4) and sd["state"][1]["exp_avg_sq"].shape == (8,)
    t = torch.optim.Adam([torch.nn.Parameter(torch.zeros(2, 4)), torch.nn.Parameter(torch.zeros(8))])
    t.load_state_dict(sd)
    for q in t.param_groups[0]["params"]:
        q.grad = torch.ones_like(q)
    t.step()
    assert int(t.state_dict()["state"][0]["step"]) == 4 and t.param_groups[0]["lr"] == pytest.approx(3e-4)
    opt2 = FusedAdam(eng)
    opt2.load_state_dict(t.state_dict())
    assert float(opt2.hyper[8]) == 4.0 and float(opt2.hyper[0]) == pytest.approx(3e-4)
    assert torch.equal(opt2.m[:8].view(2, 4), t.state_dict()["state"][0]["exp_avg"])


def test_perspective_inverse_map_closed_form_matches_linear_solve():
    """cutouts._quad_to_square (closed form used by sample_params) against the general 8x8 solve it replaced"""
    from feed_forward_vqgan_clip_b200 import cutouts as C
    torch.manual_seed(3)
    n, q = 1000, 223.0
    r = torch.rand(n, 8, dtype=torch.float64) * (0.7 * q / 2)
    dst = torch.stack([torch.stack([r[:, 0], r[:, 1]], -1), torch.stack([q - r[:, 2], r[:, 3]], -1),
                       torch.stack([q - r[:, 4], q - r[:, 5]], -1), torch.stack([r[:, 6], q - r[:, 7]], -1)], dim=1)
    src = torch.tensor([[0, 0], [q, 0], [q, q], [0, q]], dtype=torch.float64).expand(n, 4, 2)
    a, b = C._persp_coeffs(dst, src), C._quad_to_square(dst, q)
    assert ((a - b).abs() / (a.abs() + 1e-9)).max().item() < 1e-8
    pts = torch.cat([dst, torch.ones(n, 4, 1, dtype=torch.float64)], -1) @ b.transpose(1, 2)
    assert (pts[..., :2] / pts[..., 2:] - src).abs().max().item() < 1e-9
    eye = C._quad_to_square(src[:3].clone(), q)                       # undistorted corners -> identity
    assert torch.allclose(eye, torch.eye(3, dtype=torch.float64).expand(3, 3, 3), atol=1e-12)


def test_replay_host_logic_draws_next_parameters_after_the_launch():
    """TrainStep.replay on stand-in objects (no GPU): static buffers are filled from the host tensors, the graph is launched
    once, the parameters of the next call are drawn after the launch and consumed by it, explicit parameters bypass that"""
    from types import SimpleNamespace
    from feed_forward_vqgan_clip_b200.cutouts import sample_params
    from feed_forward_vqgan_clip_b200.train_step import TrainStep
    B, cutn = 2, 3
    N = B * cutn
    st = dict(inp=torch.zeros(B, 8), out=torch.zeros(B, 8), affine_inv=torch.zeros(N, 3, 3), persp_inv=torch.zeros(N, 3, 3),
              sat=torch.zeros(N), hue=torch.zeros(N), erase=torch.zeros(4, dtype=torch.int32))
    events = []
    fake = SimpleNamespace(static=st, graph=SimpleNamespace(replay=lambda: events.append("launch")), loss=torch.zeros(1), cutn=cutn,
                           cut_size=224, gen=torch.Generator().manual_seed(1), _next_prm=None, repeat=1, noise_bank=None)
    fake.load_static = lambda *a: TrainStep.load_static(fake, *a)

    def new_params(b):
        events.append("sample")
        return TrainStep.new_params(fake, b)

    fake.new_params = new_params
    x = torch.randn(B, 8)
    TrainStep.replay(fake, x)
    assert events == ["sample", "launch", "sample"] and torch.equal(st["inp"], x) and torch.equal(st["out"], x)
    ahead = fake._next_prm
    assert ahead is not None
    TrainStep.replay(fake, x)
    assert events == ["sample", "launch", "sample", "launch", "sample"]
    assert torch.equal(st["affine_inv"], ahead["affine_inv"]) and torch.equal(st["sat"], ahead["sat"])
    explicit = sample_params(N, 224, torch.Generator().manual_seed(9), with_noise=False)
    kept = fake._next_prm
    TrainStep.replay(fake, x, None, explicit)
    assert events[-1] == "launch" and fake._next_prm is kept and torch.equal(st["persp_inv"], explicit["persp_inv"])
    # repeat > 1 (main.py:739-740): a batch of B / repeat prompts is repeated into the static buffers like step() does
    fake.repeat = 2
    half = torch.randn(B // 2, 8)
    TrainStep.replay(fake, half, None, explicit)
    assert torch.equal(st["inp"], half.repeat(2, 1)) and torch.equal(st["out"], half.repeat(2, 1))


def test_load_clip_and_vqgan_checkpoints_in_the_published_formats(tmp_path):
    """api.load_clip_model / load_vqgan_model read what the reference reads (main.py:84-103,1308-1333): CLIP weights as a plain
    fp16 state_dict (visual.* + text tower keys) or a TorchScript archive, the VQGAN as taming's {"state_dict": ...} checkpoint
    with encoder / loss keys the decoder-only model ignores.  Host logic only (no engine is built)."""
    from feed_forward_vqgan_clip_b200 import api
    from feed_forward_vqgan_clip_b200.clip_vit import CLIP, VIT_B32
    from feed_forward_vqgan_clip_b200.clip_text import TEXT_B32
    src = CLIP(VIT_B32, text_cfg=TEXT_B32)
    sd = {"visual." + k: v.half() for k, v in src.visual.state_dict().items()}
    sd.update({k: v.half() for k, v in src.text.state_dict().items()})
    sd["logit_scale"] = torch.tensor(3.25)
    torch.save(sd, tmp_path / "clip_sd.pt")
    m = api.load_clip_model("ViT-B/32", str(tmp_path / "clip_sd.pt"))
    k = "transformer.resblocks.3.mlp.c_fc.weight"
    assert m.visual.state_dict()[k].dtype == torch.float32
    assert torch.equal(m.visual.state_dict()[k], src.visual.state_dict()[k].half().float())
    assert torch.equal(m.text.state_dict()["text_projection"], src.text.state_dict()["text_projection"].half().float())
    assert not any(p.requires_grad for p in m.parameters()) and abs(float(m.logit_scale) - 3.25) < 1e-6

    class Holder(torch.nn.Module):                       # a TorchScript archive only has to expose state_dict()
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.arange(4.0))

        def forward(self, x):
            return x + self.w
    torch.jit.script(Holder()).save(str(tmp_path / "jit.pt"))
    assert torch.equal(api._read_state_dict(str(tmp_path / "jit.pt"))["w"], torch.arange(4.0))

    vq = api.load_vqgan_model()
    ck = {"state_dict": {k: v + 1.0 for k, v in vq.state_dict().items()}}
    ck["state_dict"]["encoder.conv_in.weight"] = torch.zeros(3)          # taming's encoder / loss keys are not ours
    ck["state_dict"]["loss.discriminator.main.0.weight"] = torch.zeros(3)
    torch.save(ck, tmp_path / "vqgan.ckpt")
    vq2 = api.load_vqgan_model(None, str(tmp_path / "vqgan.ckpt"))
    for k2, v in vq.state_dict().items():
        assert torch.equal(vq2.state_dict()[k2], v + 1.0), k2
    assert vq2.quantize.embedding.weight.shape == (16384, 256) and not vq2.training


def test_load_model_reads_checkpoint_dictionaries_and_legacy_pickles(tmp_path):
    """api.load_model (main.py:1273-1290): the {"config", "state_dict"} dictionary train() writes, and — where the reference tree is
    present — a pickled instance of the REFERENCE's own Mixer (the legacy `model.th` form): both come back as this package's mapper
    with the same weights."""
    import sys
    from feed_forward_vqgan_clip_b200 import api
    cfg = dict(model_type="mlp_mixer", clip_model="ViT-B/32", clip_dim=32, vq_image_size=4, noise_dim=0, dim=32, depth=2, dropout=0)
    torch.manual_seed(5)
    src = api.build_model(cfg, vq_channels=16)
    torch.save({"state_dict": src.state_dict(), "config": cfg, "step": 7, "epoch": 1}, tmp_path / "checkpoint.th")
    net = api.load_model(str(tmp_path / "checkpoint.th"), vq_channels=16)
    assert type(net) is type(src) and net.config == cfg and net.step == 7 and net.epoch == 1
    for (k, a), b in zip(net.state_dict().items(), src.state_dict().values()):
        assert torch.equal(a, b), k
    if os.path.exists("/root/reference/mlp_mixer_pytorch.py"):
        sys.path.insert(0, "/root/reference")
        try:
            from mlp_mixer_pytorch import Mixer as RefMixer
            torch.manual_seed(6)
            legacy = RefMixer(input_dim=32, image_size=4, channels=16, patch_size=1, dim=32, depth=2)
            legacy.config = dict(cfg)                        # main.py:590 attaches the config to the module before pickling it
            torch.save(legacy, tmp_path / "model.th")
            net2 = api.load_model(str(tmp_path / "model.th"), vq_channels=16)
            assert type(net2) is type(src)
            for (k, a), b in zip(net2.state_dict().items(), legacy.state_dict().values()):
                assert torch.equal(a, b), k
        finally:
            sys.path.remove("/root/reference")


def test_every_compute_entry_point_has_a_cpu_statement_of_its_contract():
    """tests/abi_model.py states the contract of every compute entry point of include/ffvc.h (the control / query functions —
    options, launch counter, workspace sizes, error string — need none): a new kernel cannot be added to the ABI without its
    host-checkable statement.  The single-kernel GroupNorm forms must agree with the two-pass model they alias."""
    import abi_model
    control = {"ffvc_arch", "ffvc_last_error", "ffvc_launch_count", "ffvc_reset_launch_count", "ffvc_set_option", "ffvc_get_option",
               "ffvc_sizeof", "ffvc_gemm_set_stream_k", "ffvc_gemm_max_quads", "ffvc_gemm_set_tma_store", "ffvc_groupnorm_set_pipeline",
               "ffvc_groupnorm_ws_bytes", "ffvc_groupnorm_ws_doubles", "ffvc_layernorm_bwd_ws_bytes"}
    declared = set(_lib.header_declarations())
    modelled = {"ffvc_" + n[2:] for n in dir(abi_model) if n.startswith("k_")} | {"ffvc_gemm"}
    assert declared - modelled == control and not (modelled - declared)
    g = torch.Generator().manual_seed(0)
    N, HW, C = 2, 64, 64
    x = torch.randn(N * HW, C, generator=g).to(torch.bfloat16)
    dy = torch.randn(N * HW, C, generator=g).to(torch.bfloat16)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    m1, r1, m2, r2 = (torch.empty(N * 32) for _ in range(4))
    y1, y2, d1, d2 = (torch.empty(N * HW, C, dtype=torch.bfloat16) for _ in range(4))
    ws = torch.empty(4096, dtype=torch.float64)
    abi_model.k_groupnorm_stats(x, ws, m1, r1, N, HW, C, 32, 1e-6)
    abi_model.k_groupnorm_apply(x, m1, r1, gamma, beta, y1, N, HW, C, 32, 1)
    abi_model.k_groupnorm_fused_fwd(x, gamma, beta, y2, m2, r2, ws, N, HW, C, 32, 1, 1e-6)
    assert torch.equal(y1, y2) and torch.equal(m1, m2) and torch.equal(r1, r2)
    abi_model.k_groupnorm_bwd(dy, x, m1, r1, gamma, beta, ws, None, d1, N, HW, C, 32, 1)
    abi_model.k_groupnorm_fused_bwd(dy, x, m1, r1, gamma, beta, ws, None, d2, N, HW, C, 32, 1)
    assert torch.equal(d1, d2)
    img = torch.randn(1, 8, 8, 3, generator=g)
    w = torch.randn(8, 27, generator=g)
    out = torch.empty(64, 8, dtype=torch.bfloat16)
    abi_model.k_conv3x3_cin3(img, w, out, 1, 8, 8, 8)
    ref = torch.nn.functional.conv2d(img.permute(0, 3, 1, 2), w.view(8, 3, 3, 3).permute(0, 3, 1, 2), padding=1)
    assert torch.allclose(out.float().view(1, 8, 8, 8), ref.permute(0, 2, 3, 1), atol=2e-2, rtol=1e-2)      # bf16 output


def test_abi_model_satisfies_the_expectations_of_the_kernel_tests():
    """tests/abi_model.py is what the CPU host-logic tests stand on, so it must itself be held to the kernels' contracts: the
    per-kernel GPU tests (tests/test_ops_gpu.py — written against torch / the oracle and green on the B200) are run here on the
    CPU with the `abi_model_mode` plugin routing every ffvc_* launch to the model.  The whole GPU suite runs the same way with
    `PYTHONPATH=tests python -m pytest tests -m gpu -p abi_model_mode` (177 of the 206 GPU-verified tests apply and pass)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=os.path.join(root, "tests") + os.pathsep + root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_ops_gpu.py"), "-m", "gpu", "-p", "abi_model_mode",
                        "-q", "-x", "-p", "no:cacheprovider"], capture_output=True, text=True, env=env, cwd=root, timeout=900)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:]
    assert r.returncode == 0 and " passed" in tail and "failed" not in tail, r.stdout[-3000:] + r.stderr[-1000:]
    assert int(tail.split(" passed")[0].split()[-1]) >= 90, tail


def test_documented_ctypes_stub_matches_the_header():
    """INTEGRATION.md's ctypes mirror of ffvc_gemm_params (generated by tools/gen_ctypes_stub.py) is executed as written: its size
    equals ffvc_sizeof("ffvc_gemm_params") and its fields are the package's own mirror's, in order (round 1 shipped a stale stub)."""
    import ctypes
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.search(r"<!-- BEGIN GENERATED STRUCT -->\s*```python\n(.*?)```", doc, flags=re.S).group(1)
    ns = {}
    exec(block, ns)
    stub = ns["ffvc_gemm_params"]
    assert ctypes.sizeof(stub) == _lib.load().ffvc_sizeof(b"ffvc_gemm_params") == ctypes.sizeof(_lib.GemmParams)
    assert [f[0] for f in stub._fields_] == [f[0] for f in _lib.GemmParams._fields_]
    for (n1, t1), (n2, t2) in zip(stub._fields_, _lib.GemmParams._fields_):
        assert ctypes.sizeof(t1) == ctypes.sizeof(t2), (n1, t1, t2)
    # regenerating from the header changes nothing
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_ctypes_stub", os.path.join(root, "tools", "gen_ctypes_stub.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    assert gen.stub().strip() in block
