"""ORACLE (test infrastructure, not product): the loss expression of the train step, restating
main.py:801-811 (spherical distance between the normalised target embedding, repeated cutn times, and the
normalised image embedding), main.py:423-428 (tv_loss) and main.py:758-762 (z L2).  Plain torch; pinned by
tests/test_oracle_golden.py against values computed with the reference's own lines."""
import torch
import torch.nn.functional as F


def spherical_dist_loss(embed, out_feats, cutn, coef=1.0):
    H = F.normalize(out_feats.repeat(cutn, 1), dim=-1)
    e = F.normalize(embed, dim=1)
    return coef * (H.sub(e).norm(dim=-1).div(2).arcsin().pow(2).mul(2)).mean()


def tv_loss(y):
    return 0.5 * (torch.abs(y[:, :, 1:, :] - y[:, :, :-1, :]).mean() + torch.abs(y[:, :, :, 1:] - y[:, :, :, :-1]).mean())
