"""Activations of models/utils.py:53-97 (trunc_exp, get_activation table) -- CPU oracle.

PINNED against the reference (tests/golden/make_golden.py imports models.utils.get_activation).
"""
import torch
import torch.nn.functional as F


class _TruncExp(torch.autograd.Function):
    # models/utils.py:53-68: forward exp(x); backward g*exp(clamp(x, max=15))
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        x = ctx.saved_tensors[0]
        return g * torch.exp(torch.clamp(x, max=15))


trunc_exp = _TruncExp.apply


def get_activation(name):
    """models/utils.py:71-97."""
    if name is None:
        return lambda x: x
    name = name.lower()
    if name == 'none':
        return lambda x: x
    if name.startswith('scale'):
        s = float(name[5:])
        return lambda x: x.clamp(0., s) / s
    if name.startswith('clamp'):
        c = float(name[5:])
        return lambda x: x.clamp(0., c)
    if name.startswith('mul'):
        m = float(name[3:])
        return lambda x: x * m
    if name == 'lin2srgb':
        return lambda x: torch.where(x > 0.0031308, torch.pow(torch.clamp(x, min=0.0031308), 1.0 / 2.4) * 1.055 - 0.055,
                                     12.92 * x).clamp(0., 1.)
    if name == 'trunc_exp':
        return trunc_exp
    if name.startswith('+') or name.startswith('-'):
        return lambda x: x + float(name)
    if name == 'sigmoid':
        return torch.sigmoid
    if name == 'tanh':
        return torch.tanh
    return getattr(F, name)
