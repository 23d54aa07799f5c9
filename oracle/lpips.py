"""ORACLE (test infrastructure, not product): CPU fp32 restatement of the LPIPS-VGG16 diversity term of the train step,
main.py:776-791 (modes 'between_same_prompts' and 'all') + `normalize_tensor` and the `vgg16` feature slices of
taming.modules.losses.lpips (absent package; torchvision VGG16 `features` split at relu1_2, relu2_2, relu3_3, relu4_3,
relu5_3 — SURVEY App. A.5).  The slice structure is PINNED against torchvision.models.vgg16().features with shared random weights
(tests/test_oracle_golden.py); taming's normalize_tensor and the pretrained weights are absent (restated / random).  The loop over taps and the pairwise-difference expression are the reference's own lines."""
import torch
import torch.nn.functional as F

# (slice, [torchvision features index of each conv]) ; a max-pool precedes every slice but the first
VGG16_SLICES = [(1, [0, 2]), (2, [5, 7]), (3, [10, 12, 14]), (4, [17, 19, 21]), (5, [24, 26, 28])]
VGG16_CH = {0: (3, 64), 2: (64, 64), 5: (64, 128), 7: (128, 128), 10: (128, 256), 12: (256, 256), 14: (256, 256),
            17: (256, 512), 19: (512, 512), 21: (512, 512), 24: (512, 512), 26: (512, 512), 28: (512, 512)}


def init_vgg_state_dict(seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for s, idxs in VGG16_SLICES:
        for i in idxs:
            cin, cout = VGG16_CH[i]
            sd["slice%d.%d.weight" % (s, i)] = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
            sd["slice%d.%d.bias" % (s, i)] = 0.05 * torch.randn(cout, generator=g)
    return sd


def vgg_taps(sd, x):
    outs = []
    h = x
    for s, idxs in VGG16_SLICES:
        if s > 1:
            h = F.max_pool2d(h, 2, 2)
        for i in idxs:
            h = F.relu(F.conv2d(h, sd["slice%d.%d.weight" % (s, i)], sd["slice%d.%d.bias" % (s, i)], padding=1))
        outs.append(h)
    return outs


def normalize_tensor(x, eps=1e-10):                       # taming lpips.normalize_tensor
    return x / (torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True)) + eps)


def diversity(sd, xr, repeat, bs, mean, std, mode="between_same_prompts"):             # main.py:776-789
    div = 0
    for feats in vgg_taps(sd, (xr - mean) / std):
        feats = normalize_tensor(feats)
        _, cc, hh, ww = feats.shape
        if mode == "between_same_prompts":
            div = div + ((feats.view(repeat, 1, bs, cc, hh, ww) - feats.view(1, repeat, bs, cc, hh, ww)) ** 2).sum(dim=3).mean()
        elif mode == "all":                                                                    # main.py:783-787
            nb = len(feats)
            div = div + ((feats.view(nb, 1, cc, hh, ww) - feats.view(1, nb, cc, hh, ww)) ** 2).sum(dim=2).mean()
        else:
            raise ValueError("diversity_mode should be 'between_same_prompts' lr 'all'")
    return div
