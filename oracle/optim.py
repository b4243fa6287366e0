"""AdamW update rule -- CPU oracle (TEST INFRASTRUCTURE, see oracle/__init__.py).

The reference's optimizer IS ``torch.optim.AdamW`` (systems/utils.py:314-325: ``getattr(torch.optim, config.name)(params, **config.args)``;
configs/nerf-blender.yaml:74-79: lr 1e-2, betas (0.9, 0.99), eps 1e-15, weight_decay = torch's default 1e-2), so this restatement is
PINNED against torch itself in tests/test_oracle_kat.py::test_adamw_oracle_matches_torch.  Follows torch's single-tensor path:
decoupled decay, exp_avg.lerp_(grad, 1 - beta1), exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2), bias corrections and
step size in Python doubles, denom = sqrt(v) / sqrt(bc2) + eps.
"""
import math

import numpy as np


def adamw_step(p, g, m, v, step, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, weight_decay=1e-2, inv_grad_scale=1.0):
    """one update in fp32 (numpy arrays are modified in place); ``step`` is the 1-based number of this update."""
    f = np.float32
    b1, b2 = betas
    g = (g * f(inv_grad_scale)).astype(np.float32)
    p *= f(1.0 - lr * weight_decay)
    m += (g - m) * f(1.0 - b1)
    v *= f(b2)
    v += f(1.0 - b2) * g * g
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    denom = np.sqrt(v) / f(math.sqrt(bc2)) + f(eps)
    p -= f(lr / bc1) * (m / denom)
    return p, m, v
