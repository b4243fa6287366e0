"""Multiresolution hash grid (tiny-cuda-nn ``HashGrid`` = Grid/Hash/Linear, fp16 params) -- CPU oracle.

[3P, parity unpinned] restated from tiny-cuda-nn's published behaviour (SURVEY.md §8a
"Hash-grid geometry", Appendix A.5).  Reference call sites: models/network_utils.py:47,90,209;
configs/nerf-blender.yaml:43-49, configs/neus-blender.yaml:47-54.

Everything is plain differentiable torch, so first- and second-order derivatives w.r.t. the
input position and the table come from autograd (that is what pins the CUDA
bwd / bwd_input / bwd_bwd kernels).
"""
import math
import numpy as np
import torch

PRIME_Y = 2654435761
PRIME_Z = 805459861


def level_table(cfg, n_input_dims=3):
    """Per-level geometry.  cfg: the reference's encoding-config dict.

    scale_l = exp2(l*log2(pls))*base - 1   (computed in fp32, shared verbatim with the kernels,
    which take this table instead of recomputing it); res_l = ceil(scale_l)+1;
    size_l = min(roundup8(res_l^3), 2^log2_hashmap_size); dense iff res_l^3 <= size_l.
    """
    assert n_input_dims == 3
    L = int(cfg['n_levels'])
    F = int(cfg.get('n_features_per_level', 2))
    T = 1 << int(cfg.get('log2_hashmap_size', 19))
    base = np.float32(cfg.get('base_resolution', 16))
    log2_pls = np.log2(np.float32(cfg.get('per_level_scale', 2.0))).astype(np.float32)
    scale = np.zeros(L, np.float32)
    res = np.zeros(L, np.int64)
    size = np.zeros(L, np.int64)
    dense = np.zeros(L, bool)
    for l in range(L):
        s = np.float32(np.exp2(np.float32(np.float32(l) * log2_pls))) * base - np.float32(1.0)
        scale[l] = np.float32(s)
        r = int(math.ceil(float(scale[l]))) + 1
        res[l] = r
        n = r ** 3
        n8 = (n + 7) // 8 * 8
        size[l] = min(n8, T)
        dense[l] = n <= size[l]
    offset = np.zeros(L + 1, np.int64)
    offset[1:] = np.cumsum(size)
    return dict(n_levels=L, n_features=F, scale=scale, res=res, size=size, offset=offset, dense=dense,
                n_params=int(offset[-1]) * F, n_output_dims=L * F)


def fma_f32(a, b, c):
    """float32 fused multiply-add emulated through float64 (exact product, one extra rounding)."""
    return (a.double() * b.double() + c.double()).float()


def corner_index(ix, iy, iz, res, size, dense):
    """ix,iy,iz int64 tensors (uint32 semantics).  Dense: x + y*res + z*res^2; hashed:
    x ^ y*2654435761 ^ z*805459861 (mod 2^32); both then mod size."""
    m = 0xFFFFFFFF
    if dense:
        idx = (ix + iy * res + iz * res * res) & m
    else:
        idx = ((ix & m) ^ ((iy * PRIME_Y) & m) ^ ((iz * PRIME_Z) & m)) & m
    return idx % size


def hashgrid_fwd(x, table, lt, compute_dtype=torch.float64, one_gather=False):
    """x [N,3] in [0,1]; table [n_entries, F] (any float dtype; values as stored, e.g. already
    rounded to fp16).  Returns [N, L*F] in compute_dtype (level-major, feature-minor).
    one_gather=True evaluates the same expression with ONE indexing op over all L*8 corners instead of L*8 separate ones: torch's
    autograd materialises a table-sized dense gradient per indexing op (128 x 50 MB zero-fills per backward for the 12.6 M-entry
    table), which says nothing about the algorithm -- the timed CPU baseline (bench.py) uses this form; results agree to rounding."""
    if one_gather:
        return _hashgrid_fwd_one_gather(x, table, lt, compute_dtype)
    N = x.shape[0]
    L, F = lt['n_levels'], lt['n_features']
    xs = x.to(compute_dtype)
    tab = table.to(compute_dtype)
    outs = []
    for l in range(L):
        scale = float(lt['scale'][l])
        # cell decision uses the fp32 fma the kernel uses; the fractional part is differentiable
        pos32 = fma_f32(x.detach().float(), torch.tensor(scale, dtype=torch.float32), torch.tensor(0.5))
        cell = torch.floor(pos32)
        frac = xs * scale + 0.5 - cell.to(compute_dtype)
        ci = cell.to(torch.int64)
        res, size, dense, off = int(lt['res'][l]), int(lt['size'][l]), bool(lt['dense'][l]), int(lt['offset'][l])
        acc = torch.zeros(N, F, dtype=compute_dtype)
        for c in range(8):
            bx, by, bz = c & 1, (c >> 1) & 1, (c >> 2) & 1
            w = (frac[:, 0] if bx else 1 - frac[:, 0]) * (frac[:, 1] if by else 1 - frac[:, 1]) * \
                (frac[:, 2] if bz else 1 - frac[:, 2])
            idx = corner_index(ci[:, 0] + bx, ci[:, 1] + by, ci[:, 2] + bz, res, size, dense) + off
            acc = acc + w[:, None] * tab[idx]
        outs.append(acc)
    return torch.cat(outs, dim=-1)


def _hashgrid_fwd_one_gather(x, table, lt, compute_dtype):
    N, L, F = x.shape[0], lt['n_levels'], lt['n_features']
    xs = x.to(compute_dtype)
    idx_all, w_all = [], []
    for l in range(L):
        scale = float(lt['scale'][l])
        pos32 = fma_f32(x.detach().float(), torch.tensor(scale, dtype=torch.float32), torch.tensor(0.5))
        cell = torch.floor(pos32)
        frac = xs * scale + 0.5 - cell.to(compute_dtype)
        ci = cell.to(torch.int64)
        res, size, dense, off = int(lt['res'][l]), int(lt['size'][l]), bool(lt['dense'][l]), int(lt['offset'][l])
        for c in range(8):
            bx, by, bz = c & 1, (c >> 1) & 1, (c >> 2) & 1
            w_all.append((frac[:, 0] if bx else 1 - frac[:, 0]) * (frac[:, 1] if by else 1 - frac[:, 1]) *
                         (frac[:, 2] if bz else 1 - frac[:, 2]))
            idx_all.append(corner_index(ci[:, 0] + bx, ci[:, 1] + by, ci[:, 2] + bz, res, size, dense) + off)
    idx = torch.stack(idx_all, dim=1)                       # [N, L*8]
    w = torch.stack(w_all, dim=1).view(N, L, 8, 1)
    vals = table[idx].to(compute_dtype).view(N, L, 8, F)    # the one gather (its backward is one scatter-add into the table)
    return (w * vals).sum(dim=2).reshape(N, L * F)


def init_table(lt, seed=1337, dtype=torch.float32):
    """tcnn init: U(-1e-4, 1e-4)."""
    g = torch.Generator().manual_seed(seed)
    n = int(lt['offset'][-1])
    return (torch.rand(n, lt['n_features'], generator=g, dtype=dtype) * 2 - 1) * 1e-4
