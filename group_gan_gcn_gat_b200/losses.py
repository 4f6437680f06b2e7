"""Losses and displacement metrics of the reference (sgan/losses.py:5-119) -- host-side keep, O(batch) elementwise.

`label_rng` makes the label-smoothing draws explicit: the reference pulls them from Python's global `random`
(sgan/losses.py:32,45-46); under data parallelism every rank must draw the same value, so the training helpers
pass one identically-seeded `random.Random` per rank.
"""
import random as _random

import torch


def bce_loss(input, target):
    """Stable binary cross entropy with logits, mean over the minibatch (sgan/losses.py:5-21)."""
    return (input.clamp(min=0) - input * target + torch.log1p(torch.exp(-input.abs()))).mean()


def _draw(lo, hi, label_rng):
    return (label_rng or _random).uniform(lo, hi)


def gan_g_loss(scores_fake, label_rng=None):
    return bce_loss(scores_fake, torch.ones_like(scores_fake) * _draw(0.7, 1.2, label_rng))


def gan_d_loss(scores_real, scores_fake, label_rng=None):
    y_real = torch.ones_like(scores_real) * _draw(0.7, 1.2, label_rng)
    y_fake = torch.zeros_like(scores_fake) * _draw(0, 0.3, label_rng)      # = 0, the draw keeps the RNG stream aligned
    return bce_loss(scores_real, y_real) + bce_loss(scores_fake, y_fake)


def l2_loss(pred_traj, pred_traj_gt, loss_mask, random=0, mode='average'):
    """Masked squared error; mode in sum | average | raw (per pedestrian) -- sgan/losses.py:52-71."""
    err = loss_mask.unsqueeze(2) * (pred_traj_gt.permute(1, 0, 2) - pred_traj.permute(1, 0, 2)) ** 2
    if mode == 'sum':
        return err.sum()
    if mode == 'average':
        return err.sum() / loss_mask.numel()
    if mode == 'raw':
        return err.sum(dim=2).sum(dim=1)
    raise ValueError('unknown mode %r' % (mode,))


def displacement_error(pred_traj, pred_traj_gt, consider_ped=None, mode='sum'):
    """Sum over time of the Euclidean error per pedestrian (ADE numerator) -- sgan/losses.py:74-95."""
    d = torch.sqrt(((pred_traj_gt.permute(1, 0, 2) - pred_traj.permute(1, 0, 2)) ** 2).sum(dim=2)).sum(dim=1)
    if consider_ped is not None:
        d = d * consider_ped
    return d if mode == 'raw' else d.sum()


def final_displacement_error(pred_pos, pred_pos_gt, consider_ped=None, mode='sum'):
    """Euclidean error at the last step (FDE numerator) -- sgan/losses.py:98-119."""
    d = torch.sqrt(((pred_pos_gt - pred_pos) ** 2).sum(dim=1))
    if consider_ped is not None:
        d = d * consider_ped
    return d if mode == 'raw' else d.sum()
