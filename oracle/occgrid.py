"""Occupancy grid state + update rule (nerfacc 0.3.3 ``OccupancyGrid``) -- CPU oracle.

[3P, parity unpinned] restated from SURVEY.md Appendix A.3; reference call sites
models/nerf.py:36-55, models/neus.py:63-74,108-111.  The cell selection and the in-cell jitter are
INPUTS here (the product draws them with torch's RNG) so product and oracle see identical points.
"""
import numpy as np
import torch

from .contraction import contract_inv, UN_BOUNDED_SPHERE


def cell_coords(cells, R):
    """flat cell index (ix*R*R + iy*R + iz) -> integer coords [n,3]."""
    ix = cells // (R * R)
    iy = (cells // R) % R
    iz = cells % R
    return torch.stack([ix, iy, iz], dim=-1)


def update(occs, cells, jitter, occ_eval_fn, radius, contraction, R, ema_decay=0.95, occ_thre=0.01):
    """occs float32 [R^3]; cells int64 [n]; jitter float32 [n,3] in [0,1).
    Returns (new occs, binary bool [R,R,R])."""
    x = (cell_coords(cells, R).float() + jitter) / R
    if contraction == UN_BOUNDED_SPHERE:
        keep = (x - 0.5).norm(dim=-1) < 0.5
        cells, x = cells[keep], x[keep]
    xw = contract_inv(x, radius, contraction)
    occ = occ_eval_fn(xw).reshape(-1).float()
    occs = occs.clone()
    occs[cells] = torch.maximum(occs[cells] * ema_decay, occ)
    thre = min(float(occs.mean()), occ_thre)
    binary = (occs > thre).view(R, R, R)
    return occs, binary


def pack_bits(binary):
    """bool [R,R,R] -> uint32 words, bit (idx & 31) of word (idx >> 5), idx = ix*R*R + iy*R + iz."""
    flat = np.asarray(binary).reshape(-1).astype(np.uint8)
    pad = (-len(flat)) % 32
    flat = np.concatenate([flat, np.zeros(pad, np.uint8)])
    return np.packbits(flat.reshape(-1, 32), axis=1, bitorder='little').view(np.uint32).reshape(-1)
