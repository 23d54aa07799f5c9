"""checkpoint.CheckpointWriter's device path (D2D snapshot on the training stream, pinned D2H on a side stream, background
write) on the GPU: the files hold the values at the time of save() even though the arenas keep changing afterwards."""
import os
from types import SimpleNamespace

import pytest
import torch

from feed_forward_vqgan_clip_b200 import checkpoint as ck
from feed_forward_vqgan_clip_b200.mixer import Mixer
from feed_forward_vqgan_clip_b200.train_step import FusedAdam

pytestmark = pytest.mark.gpu


def test_async_checkpoint_snapshot_is_consistent(tmp_path):
    torch.manual_seed(0)
    cfg = dict(input_dim=64, image_size=16, channels=64, patch_size=1, dim=128, depth=2)
    net = Mixer(**cfg).to("cuda")
    eng = net.engine()
    opt = FusedAdam(eng, lr=1e-3)
    opt.enable_ema(0.99)
    opt.m.normal_()
    opt.v.uniform_()
    ts = SimpleNamespace(mix=eng, opt=opt)
    before = {k: v.detach().clone().cpu() for k, v in net.state_dict().items()}
    m_before = opt.m.clone().cpu()
    w = ck.CheckpointWriter(ts, str(tmp_path), config={"model_type": "mlp_mixer"})
    w.save(step=5, epoch=1)
    eng.arena.add_(1.0)                                  # "the next step": must not leak into the files
    opt.m.zero_()
    w.wait()
    c = torch.load(tmp_path / "checkpoint.th", weights_only=False)
    assert c["step"] == 5 and list(c["state_dict"].keys()) == list(before.keys())
    for k, v in before.items():
        assert torch.equal(c["state_dict"][k], v), k
    assert os.path.exists(tmp_path / "checkpoint_ema.th")
    o = torch.load(tmp_path / "opt.th", weights_only=False)
    assert o["state"] == {}                              # no optimizer step taken yet: like a fresh torch.optim.Adam
    fresh = Mixer(**cfg)
    fresh.load_state_dict(c["state_dict"])               # main.py:581
    # after one tick the moments appear under the parameter indices, at the offsets of the arena
    opt.m.copy_(m_before.cuda())
    opt.hyper[8:9].fill_(1.0)
    w.save(step=6, epoch=1, blocking=True)
    o = torch.load(tmp_path / "opt.th", weights_only=False)
    lay = ck.param_layout(eng)
    name, off, n, shape = lay[3]
    assert torch.equal(o["state"][3]["exp_avg"], m_before[off:off + n].view(shape))
    torch.optim.Adam(fresh.parameters()).load_state_dict(o)   # main.py:593-596
