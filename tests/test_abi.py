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
