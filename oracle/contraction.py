"""Scene contraction -- CPU oracle of models/geometry.py:17-29 + models/utils.py:108-113.

PINNED against the reference (golden vectors from models.geometry.contract_to_unisphere).
"""
import torch

AABB = 0
UN_BOUNDED_SPHERE = 2  # nerfacc.ContractionType enum values: AABB=0, UN_BOUNDED_TANH=1, UN_BOUNDED_SPHERE=2


def scale_anything(dat, inp_scale, tgt_scale):
    dat = (dat - inp_scale[0]) / (inp_scale[1] - inp_scale[0])
    return dat * (tgt_scale[1] - tgt_scale[0]) + tgt_scale[0]


def contract_to_unisphere(x, radius, contraction_type):
    if contraction_type == AABB:
        return scale_anything(x, (-radius, radius), (0, 1))
    if contraction_type == UN_BOUNDED_SPHERE:
        x = scale_anything(x, (-radius, radius), (0, 1))
        x = x * 2 - 1
        mag = x.norm(dim=-1, keepdim=True)
        scale = torch.where(mag > 1, (2 - 1 / mag) / mag, torch.ones_like(mag))
        x = x * scale
        return x / 4 + 0.5
    raise NotImplementedError


def contract_inv(u, radius, contraction_type):
    """unit cube -> world (nerfacc ``contract_inv``; used by OccupancyGrid._update)."""
    if contraction_type == AABB:
        return scale_anything(u, (0, 1), (-radius, radius))
    if contraction_type == UN_BOUNDED_SPHERE:
        x = (u - 0.5) * 4
        mag = x.norm(dim=-1, keepdim=True)
        scale = torch.where(mag > 1, 1 / (2 - mag) / mag, torch.ones_like(mag))
        x = x * scale
        return scale_anything((x + 1) / 2, (0, 1), (-radius, radius))
    raise NotImplementedError
