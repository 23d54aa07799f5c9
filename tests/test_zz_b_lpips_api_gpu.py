"""`LPIPS().net(x)` / `normalize_tensor` — the diversity term's call surface in the reference's loop (main.py:30-31,532-537,778-787)
— on the GPU against the oracle's VGG16 tap network: the five taps and the gradient w.r.t. the image.  (The fused form inside
TrainStep is covered by test_ops_gpu.py::test_lpips_diversity_engine_vs_oracle; this entry point was added after the round's
last GPU session and re-uses the same engine passes.)"""
import pytest
import torch

import oracle.lpips as ol
from feed_forward_vqgan_clip_b200 import api

pytestmark = pytest.mark.gpu


def test_lpips_net_taps_and_input_gradient_vs_oracle():
    sd = {k: (v.to(torch.bfloat16).float() if v.dim() == 4 else v) for k, v in ol.init_vgg_state_dict(seed=3).items()}
    model = api.LPIPS()
    model.net.load_state_dict(sd)
    model = model.to("cuda")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 256, 256, generator=g)     # relu5_3 works on 16 x 16: the smallest map the implicit-GEMM conv tiles
    xa, xb = x.clone().cuda().requires_grad_(True), x.clone().requires_grad_(True)
    mine, ref = model.net(xa), ol.vgg_taps(sd, xb)
    assert len(mine) == 5
    la = lb = 0
    for i, (a, b) in enumerate(zip(mine, ref)):
        assert a.shape == b.shape, i
        err = (a.detach().float().cpu() - b.detach()).abs().max().item()
        assert err <= 3e-2 * b.detach().abs().max().item(), (i, err)
        w = torch.randn(b.shape, generator=g)
        la, lb = la + (api.normalize_tensor(a) * w.cuda()).sum(), lb + (ol.normalize_tensor(b) * w).sum()
    la.backward()
    lb.backward()
    ga, gb = xa.grad.float().cpu().flatten(), xb.grad.flatten()
    assert float(torch.dot(ga, gb) / (ga.norm() * gb.norm())) > 0.98
