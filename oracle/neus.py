"""NeuS SDF->alpha and the variance network -- CPU oracle of models/neus.py:15-43,117-139,90-101.

PINNED against the reference (golden vectors produced by calling NeuSModel.get_alpha /
VarianceNetwork on stub-imported reference code, tests/golden/make_golden.py).
"""
import torch
import torch.nn.functional as F


def inv_s_from_variance(variance):
    """VarianceNetwork.inv_s (neus.py:27-32, modulate=False) then the clip of neus.py:118."""
    return torch.exp(variance * 10.0).clip(1e-6, 1e6)


def get_alpha(sdf, normal, dirs, dists, inv_s, cos_anneal_ratio):
    """models/neus.py:117-139."""
    true_cos = (dirs * normal).sum(-1, keepdim=True)
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio) + F.relu(-true_cos) * cos_anneal_ratio)
    est_next = sdf[..., None] + iter_cos * dists.reshape(-1, 1) * 0.5
    est_prev = sdf[..., None] - iter_cos * dists.reshape(-1, 1) * 0.5
    prev_cdf = torch.sigmoid(est_prev * inv_s)
    next_cdf = torch.sigmoid(est_next * inv_s)
    p = prev_cdf - next_cdf
    c = prev_cdf
    return ((p + 1e-5) / (c + 1e-5)).view(-1).clip(0.0, 1.0)


def occ_alpha(sdf, inv_s, render_step_size):
    """occ_eval_fn of models/neus.py:90-101 (fronto-parallel step)."""
    est_next = sdf[..., None] - render_step_size * 0.5
    est_prev = sdf[..., None] + render_step_size * 0.5
    prev_cdf = torch.sigmoid(est_prev * inv_s)
    next_cdf = torch.sigmoid(est_next * inv_s)
    return ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).view(-1, 1).clip(0.0, 1.0)
