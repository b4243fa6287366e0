"""Isosurface extraction: the drop-in for ``MarchingCubeHelper`` / ``BaseImplicitGeometry.isosurface_`` of the reference
(models/geometry.py:32-112).  The reference evaluates the level field on the GPU in chunks, parks every chunk on the CPU and runs
PyMCubes there; here the level grid stays in HBM and the mesh is extracted by four streaming kernels (csrc/mcubes.cu) -- count,
scan, vertices, faces -- with one host read (the two totals) to size the outputs.  There is no CPU path."""
import ctypes as C

import torch
import torch.nn as nn

from .lib import lib, ptr, stream, check_cuda, contig

_BLOCK = 256  # grid points per CTA in csrc/mcubes.cu (size of the block-offset workspace)


def marching_cubes(level, threshold=0.0, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0), negate=True):
    """level: CUDA fp32 [nx,ny,nz].  Surface {value = threshold}, value = -level when ``negate`` (what the reference hands to mcubes,
    geometry.py:62: the inside is where -level > threshold), triangles wound with outward normals.
    -> (verts f32 [V,3] spanning the box lo..hi, faces int64 [F,3]), both on the device."""
    check_cuda(level, what='marching_cubes')
    if level.dim() != 3:
        raise ValueError(f'marching_cubes: level must be [nx,ny,nz], got {tuple(level.shape)}')
    f = contig(level.detach(), torch.float32)
    nx, ny, nz = f.shape
    dev = f.device
    nb = (f.numel() + _BLOCK - 1) // _BLOCK
    offsets = torch.empty(2 * nb, dtype=torch.int32, device=dev)
    totals = torch.empty(2, dtype=torch.int64, device=dev)
    lib.call('nsr_mc_count', ptr(f), nx, ny, nz, float(threshold), int(bool(negate)), ptr(offsets), ptr(totals), stream())
    n_verts, n_faces = (int(v) for v in totals.tolist())  # the one host sync: output sizes
    verts = torch.empty(n_verts, 3, device=dev)
    faces = torch.empty(n_faces, 3, dtype=torch.int64, device=dev)
    if n_verts == 0:
        return verts, faces
    vid_map = torch.empty(f.numel(), dtype=torch.int32, device=dev)
    lo_h, hi_h = (C.c_float * 3)(*[float(v) for v in lo]), (C.c_float * 3)(*[float(v) for v in hi])
    lib.call('nsr_mc_emit', ptr(f), nx, ny, nz, float(threshold), int(bool(negate)), ptr(offsets), lo_h, hi_h, ptr(vid_map), ptr(verts), n_verts,
             ptr(faces), n_faces, stream())
    return verts, faces


class MarchingCubeHelper(nn.Module):
    """models/geometry.py:32-70 with the same surface: ``grid_vertices()`` (points of the unit cube in 'ij' order) and
    ``forward(level, threshold) -> {'v_pos' in [0,1]^3, 't_pos_idx'}`` returned on the CPU as the reference does."""

    def __init__(self, resolution, use_torch=True):
        super().__init__()
        self.resolution = int(resolution)
        self.points_range = (0, 1)
        self.verts = None

    def grid_vertices(self):
        if self.verts is None:
            r = self.resolution
            x = y = z = torch.linspace(*self.points_range, r)
            x, y, z = torch.meshgrid(x, y, z, indexing='ij')
            self.verts = torch.stack([x.reshape(-1), y.reshape(-1), z.reshape(-1)], dim=-1)
        return self.verts

    def forward(self, level, threshold=0.):
        r = self.resolution
        v, f = marching_cubes(level.float().view(r, r, r), threshold)
        return {'v_pos': v.cpu(), 't_pos_idx': f.cpu()}


@torch.no_grad()
def level_grid(forward_level, resolution, vmin, vmax, chunk, device):
    """the level field on the resolution^3 lattice spanning [vmin, vmax] (geometry.py:86-97), evaluated in ``chunk``-point slices; the
    lattice points are generated on the device (same values as torch.linspace(0, 1, R) scaled into the box) and the result stays there."""
    r = int(resolution)
    lin = torch.linspace(0, 1, r, device=device)
    axes = [lin * (float(vmax[a]) - float(vmin[a])) + float(vmin[a]) for a in range(3)]
    out = torch.empty(r * r * r, device=device)
    for s in range(0, r * r * r, int(chunk)):
        idx = torch.arange(s, min(s + int(chunk), r * r * r), device=device)
        pts = torch.stack([axes[0][idx // (r * r)], axes[1][(idx // r) % r], axes[2][idx % r]], dim=-1)
        out[s:s + idx.numel()] = forward_level(pts).reshape(-1).float()
    return out.view(r, r, r)


@torch.no_grad()
def isosurface(forward_level, radius, resolution, threshold, chunk, device):
    """BaseImplicitGeometry.isosurface (geometry.py:106-112): coarse pass over [-radius, radius]^3, then a second pass over the coarse
    mesh's bounding box enlarged by 10 % -> {'v_pos' [V,3] world coordinates, 't_pos_idx' [F,3]} on the CPU."""
    def one_pass(vmin, vmax):
        level = level_grid(forward_level, resolution, vmin, vmax, chunk, device)
        return marching_cubes(level, threshold, vmin, vmax)

    r = float(radius)
    v, f = one_pass((-r, -r, -r), (r, r, r))
    if v.shape[0] > 0:
        lo, hi = v.amin(dim=0), v.amax(dim=0)
        lo_, hi_ = (lo - (hi - lo) * 0.1).clamp(-r, r), (hi + (hi - lo) * 0.1).clamp(-r, r)
        v, f = one_pass(lo_.tolist(), hi_.tolist())
    return {'v_pos': v.cpu(), 't_pos_idx': f.cpu()}
