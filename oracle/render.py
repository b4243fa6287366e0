"""Volume-rendering helpers (nerfacc 0.3.3 ``render_visibility``, ``render_weight_from_density``,
``render_weight_from_alpha``, ``accumulate_along_rays``) -- CPU oracle.

[3P, parity unpinned] restated from SURVEY.md Appendix A.1 item 7 / A.2; reference call sites
models/nerf.py:105-109, models/neus.py:181-184,237-243.  Differentiable torch (float64 by default)
so gradients come from autograd.  Samples of one ray are contiguous and rays ascend (marcher
output), as nerfacc's scan-by-key path requires.
"""
import torch


def _segments(ray_indices, n_rays):
    counts = torch.bincount(ray_indices.long(), minlength=n_rays)
    starts = torch.cumsum(counts, 0) - counts
    return starts, counts


def _to_dense(v, ray_indices, n_rays, fill):
    """[K] packed -> ([n_rays, Lmax] padded, position index) for per-ray scans."""
    starts, counts = _segments(ray_indices, n_rays)
    K = v.shape[0]
    pos = torch.arange(K) - starts[ray_indices.long()]
    Lmax = int(counts.max().item()) if K > 0 else 0
    dense = torch.full((n_rays, max(Lmax, 1)), fill, dtype=v.dtype)
    dense[ray_indices.long(), pos] = v
    return dense, pos


def transmittance_from_alpha(alphas, ray_indices, n_rays):
    """T_i = prod_{j<i} (1 - alpha_j) within each ray (exclusive)."""
    a = alphas.reshape(-1)
    if a.numel() == 0:
        return a
    dense, pos = _to_dense(1.0 - a, ray_indices, n_rays, 1.0)
    cp = torch.cumprod(dense, dim=1)
    excl = torch.cat([torch.ones(n_rays, 1, dtype=a.dtype), cp[:, :-1]], dim=1)
    return excl[ray_indices.long(), pos]


def transmittance_from_density(sigmas, deltas, ray_indices, n_rays):
    """T_i = exp(-sum_{j<i} sigma_j delta_j) within each ray (exclusive)."""
    sd = (sigmas.reshape(-1) * deltas.reshape(-1))
    if sd.numel() == 0:
        return sd
    dense, pos = _to_dense(sd, ray_indices, n_rays, 0.0)
    cs = torch.cumsum(dense, dim=1)
    excl = torch.cat([torch.zeros(n_rays, 1, dtype=sd.dtype), cs[:, :-1]], dim=1)
    return torch.exp(-excl[ray_indices.long(), pos])


def render_visibility(alphas, ray_indices, n_rays, early_stop_eps=1e-4, alpha_thre=0.0):
    """keep mask of nerfacc's ray_marching pre-pass: T_i >= early_stop_eps (and alpha >= thre)."""
    T = transmittance_from_alpha(alphas, ray_indices, n_rays)
    keep = T >= early_stop_eps
    if alpha_thre > 0:
        keep = keep & (alphas.reshape(-1) >= alpha_thre)
    return keep, T


def render_weight_from_density(t_starts, t_ends, sigmas, ray_indices, n_rays):
    d = (t_ends - t_starts).reshape(-1)
    T = transmittance_from_density(sigmas.reshape(-1), d, ray_indices, n_rays)
    return (T * (1.0 - torch.exp(-sigmas.reshape(-1) * d))).reshape(-1, 1)


def render_weight_from_alpha(alphas, ray_indices, n_rays):
    T = transmittance_from_alpha(alphas, ray_indices, n_rays)
    return (T * alphas.reshape(-1)).reshape(-1, 1)


def accumulate_along_rays(weights, ray_indices, values, n_rays):
    src = weights if values is None else weights * values
    out = torch.zeros(n_rays, src.shape[-1], dtype=src.dtype)
    return out.index_add(0, ray_indices.long(), src)
