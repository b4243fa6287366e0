"""Loss blocks of the reference's training steps -- CPU oracle (torch).

Restates systems/nerf.py:97 (masked smooth-L1) and systems/neus.py:98-121 (rgb mse / l1 on the valid rays, eikonal, mask and opaque
binary cross entropy on the clamped opacity, sparsity) together with systems/criterions.py:155-159 (``binary_cross_entropy``) and the
dynamic ray-count rule of systems/nerf.py:93-95 / systems/neus.py:93-95.  PINNED: tests/test_reference_dropin.py executes the
reference's own ``training_step`` on the CPU and compares (tests/helpers/reference_system.py)."""
import torch
import torch.nn.functional as F


def binary_cross_entropy(inp, target):
    """systems/criterions.py:155-159"""
    return -(target * torch.log(inp) + (1 - target) * torch.log(1 - inp)).mean()


def nerf_loss(out, rgb, lambda_rgb=1.0):
    """systems/nerf.py:97-99"""
    v = out['rays_valid'][..., 0]
    return F.smooth_l1_loss(out['comp_rgb'][v], rgb[v]) * lambda_rgb


def neus_loss_terms(out, rgb, fg_mask, sparsity_scale=1.0):
    """systems/neus.py:98-121 -> dict of the six unweighted terms"""
    v = out['rays_valid_full'][..., 0]
    opacity = torch.clamp(out['opacity'].squeeze(-1), 1.e-3, 1. - 1.e-3)
    return {
        'rgb_mse': F.mse_loss(out['comp_rgb_full'][v], rgb[v]),
        'rgb_l1': F.l1_loss(out['comp_rgb_full'][v], rgb[v]),
        'eikonal': ((torch.linalg.norm(out['sdf_grad_samples'], ord=2, dim=-1) - 1.) ** 2).mean(),
        'mask': binary_cross_entropy(opacity, fg_mask.float()),
        'opaque': binary_cross_entropy(opacity, opacity),
        'sparsity': torch.exp(-sparsity_scale * out['sdf_samples'].abs()).mean(),
    }


def neus_loss(out, rgb, fg_mask, lambdas, has_mask=True):
    """weighted sum as in systems/neus.py:98-121 (lambdas: dict name -> weight for rgb_mse, rgb_l1, eikonal, mask, opaque, sparsity)"""
    t = neus_loss_terms(out, rgb, fg_mask, lambdas.get('sparsity_scale', 1.0))
    total = 0.
    for name, value in t.items():
        w = lambdas.get(name, 0.0)
        if name == 'mask' and not has_mask:
            w = 0.0
        total = total + value * w
    return total, t


def next_train_num_rays(train_num_rays, train_num_samples, num_samples, max_train_num_rays):
    """systems/nerf.py:93-95"""
    target = int(train_num_rays * (train_num_samples / num_samples))
    return min(int(train_num_rays * 0.9 + target * 0.1), max_train_num_rays)
