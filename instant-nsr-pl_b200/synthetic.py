"""Seeded synthetic inputs of NeRF-Synthetic-lego shape: no dataset files exist offline.

* cameras: 100 poses on a sphere of radius 4.0311 looking at the origin (seed 42), W=H=800,
  focal = 400/tan(0.5*0.6911112) (datasets/blender.py:48), pixel centres +0.5, ray directions as
  models/ray_utils.py:9-43 (restated), normalised as systems/nerf.py:66.
* scene: radius 1.5, 128^3 occupancy = centred ball r=0.75 U a baseplate-like slab.
* parameters: table U(-1e-4,1e-4) (tcnn init) plus a smooth bump written into one dense level, and a
  density-MLP row wired to it, so that density = trunc_exp(out0 - 1) is O(100) at the ball centre and
  falls to exp(-1) outside: the visibility filter and early termination are exercised.
All numpy/torch-CPU and deterministic; bench.py, tests and the CPU baseline share it.
"""
import math

import numpy as np
import torch

RADIUS = 1.5
CAM_DIST = 4.0311
IMG_W = IMG_H = 800
FOCAL = 0.5 * IMG_W / math.tan(0.5 * 0.6911112)


def cameras(n=100, seed=42, dist=CAM_DIST):
    """c2w [n,3,4] (OpenGL convention: camera looks down -z, +y up), on the upper hemisphere."""
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(n, 3))
    v[:, 2] = np.abs(v[:, 2]) * 0.6 + 0.05
    pos = v / np.linalg.norm(v, axis=1, keepdims=True) * dist
    fwd = -pos / np.linalg.norm(pos, axis=1, keepdims=True)
    up = np.array([0., 0., 1.])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right, axis=1, keepdims=True)
    true_up = np.cross(right, fwd)
    c2w = np.zeros((n, 3, 4), np.float32)
    c2w[:, :, 0], c2w[:, :, 1], c2w[:, :, 2], c2w[:, :, 3] = right, true_up, -fwd, pos
    return c2w


def sample_rays(n_rays, seed=0, c2w=None, W=IMG_W, H=IMG_H, focal=FOCAL):
    """uniform (image, x, y) draws (systems/nerf.py:38-47) -> rays [n,6] float32 (o | normalised d)."""
    c2w = cameras() if c2w is None else c2w
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, len(c2w), n_rays)
    x = rng.integers(0, W, n_rays).astype(np.float32) + 0.5
    y = rng.integers(0, H, n_rays).astype(np.float32) + 0.5
    dirs = np.stack([(x - W / 2) / focal, -(y - H / 2) / focal, -np.ones_like(x)], -1).astype(np.float32)
    R = c2w[idx][:, :, :3]
    d = (dirs[:, None, :] * R).sum(-1)
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    o = c2w[idx][:, :, 3]
    return np.concatenate([o, d], axis=1).astype(np.float32)


def occupancy(R=128, radius=RADIUS, ball=0.75):
    g = (np.arange(R) + 0.5) / R * 2 * radius - radius
    X, Y, Z = np.meshgrid(g, g, g, indexing='ij')
    binary = (X ** 2 + Y ** 2 + Z ** 2 < ball ** 2)
    binary |= (np.abs(Z + 0.85) < 0.06) & (np.abs(X) < 1.1) & (np.abs(Y) < 1.1)
    return binary


def render_step_size(radius=RADIUS, num_samples_per_ray=1024):
    return 1.732 * 2 * radius / num_samples_per_ray  # models/nerf.py:31


def shape_density(params, grid_spec, mlp_n_params, peak_logit=10.0, ball=0.75, radius=RADIUS, level=4, hidden_unit=0):
    """In-place edit of a NetworkWithInputEncoding flat parameter vector (MLP first, then grid):
    feature 0 of dense level `level` := max(1 - |p|/ball, slab bump) at every vertex; density-MLP hidden
    unit `hidden_unit` := that feature; output row 0 := peak_logit * hidden unit.  Density is then
    exp(peak_logit * bump - 1): O(1e3+) deep inside the ball / slab, exp(-1) outside."""
    assert bool(grid_spec.dense[level])
    r, off, scale = int(grid_spec.res[level]), int(grid_spec.offset[level]), float(grid_spec.scale[level])
    g = (torch.arange(r, dtype=torch.float64) - 0.5) / scale          # vertex -> x01
    X, Y, Z = torch.meshgrid(g, g, g, indexing='ij')                   # index x + y*r + z*r*r
    p = torch.stack([X, Y, Z], -1) * 2 * radius - radius
    bump = 1.0 - p.norm(dim=-1) / ball
    slab = (1.0 - (p[..., 2] + 0.85).abs() / 0.10) * ((p[..., 0].abs() < 1.15) & (p[..., 1].abs() < 1.15))
    bump = torch.maximum(bump, slab).permute(2, 1, 0).reshape(-1)      # flat order: idx = x + y*r + z*r^2
    table = params[mlp_n_params:].view(-1, 2)
    table[off:off + r ** 3, 0] = bump.float()
    in_pad = grid_spec.n_output_dims
    W1 = params[:64 * in_pad].view(64, in_pad)
    W2 = params[64 * in_pad:64 * in_pad + 16 * 64].view(16, 64)
    W1[hidden_unit].zero_()
    W1[hidden_unit, 2 * level] = 1.0
    W2[0].zero_()
    W2[0, hidden_unit] = peak_logit
    return params
