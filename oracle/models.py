"""Module-level CPU oracle: the per-ray render of models/nerf.py:61-127 and models/neus.py:141-287,
assembled from the oracle pieces (hashgrid, sh, mlp, march, render, neus).  Pure torch/numpy, fp32
field arithmetic with optional fp16 emulation of what the kernels store; differentiable w.r.t. every
parameter tensor that has requires_grad.  Test infrastructure / CPU baseline only."""
import numpy as np
import torch
import torch.nn.functional as F

from . import hashgrid, sh, mlp, march, render, neus, contraction
from .activations import trunc_exp


def _t(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


class NerfParams:
    """Parameters of the nerf-blender fields: NetworkWithInputEncoding flat vector (MLP first, then
    grid; geometry.py:117-120) and the colour network flat vector (texture.py:12-21)."""

    def __init__(self, grid_cfg, density_flat, color_flat, feature_dim=16, density_hidden=1, color_hidden=2, density_bias=-1.0):
        self.lt = hashgrid.level_table(grid_cfg)
        self.density_flat, self.color_flat = density_flat, color_flat
        self.feature_dim, self.density_hidden, self.color_hidden, self.density_bias = feature_dim, density_hidden, color_hidden, density_bias
        self.n_mlp = mlp.ffmlp_layout(self.lt['n_output_dims'], feature_dim, 64, density_hidden)[1]


def nerf_field(P, positions, dirs, radius, emulate_fp16=True, density_only=False, ctype=contraction.AABB):
    """VolumeDensity.forward + VolumeRadiance.forward (geometry.py:122-130, texture.py:23-30)."""
    q = mlp.round_half if emulate_fp16 else (lambda t: t)   # values only: see oracle.mlp._RoundHalf
    x01 = contraction.contract_to_unisphere(positions, radius, ctype)
    table = P.density_flat[P.n_mlp:].view(-1, 2)
    table = q(table) if emulate_fp16 else table
    enc = q(hashgrid.hashgrid_fwd(x01.detach(), table, P.lt, compute_dtype=torch.float32, one_gather=getattr(P, 'one_gather', False)))
    out = q(mlp.ffmlp_fwd(enc, P.density_flat[:P.n_mlp], P.lt['n_output_dims'], P.feature_dim, 64, P.density_hidden, 'ReLU', 'None',
                          emulate_fp16=emulate_fp16))
    density = trunc_exp(out[:, 0] + P.density_bias)
    if density_only:
        return density, None
    shv = q(sh.sh4((dirs + 1.) / 2.))
    rgb_raw = mlp.ffmlp_fwd(torch.cat([out, shv], dim=-1), P.color_flat, P.feature_dim + 16, 3, 64, P.color_hidden, 'ReLU', 'None',
                            emulate_fp16=emulate_fp16)
    return density, torch.sigmoid(rgb_raw)


def nerf_render(P, rays, binary, radius, step, bg_color, jitter=None, emulate_fp16=True, training=True, early_stop_eps=1e-4):
    """NeRFModel.forward_ (models/nerf.py:61-127), AABB / cone_angle 0 configuration."""
    rays = np.asarray(rays, np.float32)
    o, d = rays[:, :3], rays[:, 3:6]
    n_rays = len(rays)
    aabb = np.array([-radius] * 3 + [radius] * 3, np.float32)
    t0, t1 = march.ray_interval(o, d, aabb, None, None, step, jitter)
    ri, ts, te, _ = march.march_lattice(o, d, aabb, binary, step, t0, t1)
    ri_t, ts_t, te_t = torch.from_numpy(ri).long(), _t(ts)[:, None], _t(te)[:, None]
    ot, dt = _t(o), _t(d)
    with torch.no_grad():  # sigma_fn pre-pass + render_visibility
        pos = ot[ri_t] + dt[ri_t] * (ts_t + te_t) / 2.
        sig, _ = nerf_field(P, pos, None, radius, emulate_fp16, density_only=True)
        alphas = 1.0 - torch.exp(-sig[:, None] * (te_t - ts_t))
        keep, T_pre = render.render_visibility(alphas.view(-1), ri_t, n_rays, early_stop_eps, 0.0)
    n_marched = len(ri)
    ri_t, ts_t, te_t = ri_t[keep], ts_t[keep], te_t[keep]
    mid = (ts_t + te_t) / 2.
    pos = ot[ri_t] + dt[ri_t] * mid
    density, rgb = nerf_field(P, pos, dt[ri_t], radius, emulate_fp16)
    w = render.render_weight_from_density(ts_t, te_t, density[:, None], ri_t, n_rays)
    opacity = render.accumulate_along_rays(w, ri_t, None, n_rays)
    depth = render.accumulate_along_rays(w, ri_t, mid, n_rays)
    comp = render.accumulate_along_rays(w, ri_t, rgb, n_rays) + bg_color * (1.0 - opacity)
    out = {'comp_rgb': comp, 'opacity': opacity, 'depth': depth, 'rays_valid': opacity > 0,
           'num_samples': torch.tensor([len(ts_t)], dtype=torch.int32), 'num_marched': n_marched, 'keep_mask': keep, 'trans_pre': T_pre,
           'density': density, 'rgb': rgb}
    if training:
        out.update({'weights': w.view(-1), 'points': mid.view(-1), 'intervals': (te_t - ts_t).view(-1), 'ray_indices': ri_t.view(-1)})
    return out


def nerf_unbounded_render(P, rays, binary, radius, step, cone_angle, near_plane, far_plane, bg_color, emulate_fp16=True, early_stop_eps=1e-4):
    """NeRFModel.forward_ with learned_background (models/nerf.py:21-27,61-127; configs/nerf-colmap.yaml): no scene box, near / far planes,
    blind cone stepping through the 256^3 grid under the UN_BOUNDED_SPHERE contraction, fields evaluated on contracted positions."""
    rays = np.asarray(rays, np.float32)
    o, d = rays[:, :3], rays[:, 3:6]
    n_rays = len(rays)
    roi = np.array([-radius] * 3 + [radius] * 3, np.float32)
    t0, t1 = march.ray_interval(o, d, None, near_plane, far_plane, step, None)
    ri, ts, te, _ = march.march_sequential(o, d, roi, binary, step, cone_angle, t0, t1, march.UN_BOUNDED_SPHERE)
    ri_t, ts_t, te_t = torch.from_numpy(ri).long(), _t(ts)[:, None], _t(te)[:, None]
    ot, dt = _t(o), _t(d)
    S = contraction.UN_BOUNDED_SPHERE
    with torch.no_grad():
        sig, _ = nerf_field(P, ot[ri_t] + dt[ri_t] * (ts_t + te_t) / 2., None, radius, emulate_fp16, density_only=True, ctype=S)
        keep, _ = render.render_visibility((1.0 - torch.exp(-sig[:, None] * (te_t - ts_t))).view(-1), ri_t, n_rays, early_stop_eps, 0.0)
    n_marched = len(ri)
    ri_t, ts_t, te_t = ri_t[keep], ts_t[keep], te_t[keep]
    mid = (ts_t + te_t) / 2.
    density, rgb = nerf_field(P, ot[ri_t] + dt[ri_t] * mid, dt[ri_t], radius, emulate_fp16, ctype=S)
    w = render.render_weight_from_density(ts_t, te_t, density[:, None], ri_t, n_rays)
    opacity = render.accumulate_along_rays(w, ri_t, None, n_rays)
    depth = render.accumulate_along_rays(w, ri_t, mid, n_rays)
    comp = render.accumulate_along_rays(w, ri_t, rgb, n_rays) + bg_color * (1.0 - opacity)
    return {'comp_rgb': comp, 'opacity': opacity, 'depth': depth, 'rays_valid': opacity > 0,
            'num_samples': torch.tensor([len(ts_t)], dtype=torch.int32), 'num_marched': n_marched,
            'weights': w.view(-1), 'points': mid.view(-1), 'intervals': (te_t - ts_t).view(-1), 'ray_indices': ri_t.view(-1)}


def smooth_l1_masked(comp_rgb, target, valid):
    """systems/nerf.py:97."""
    v = valid.view(-1)
    return F.smooth_l1_loss(comp_rgb[v], target[v])


# --------------------------------------------------------------------------------------------------
# CPU baseline field of config 1: VanillaFrequency + VanillaMLP (the only field arithmetic the reference
# itself implements in torch; network_utils.py:14-37,95-139)
# --------------------------------------------------------------------------------------------------
class VanillaNerfFields(torch.nn.Module):
    def __init__(self, n_freq_xyz=10, n_freq_dir=4, feature_dim=16, seed=0):
        super().__init__()
        torch.manual_seed(seed)
        self.enc_xyz = mlp.VanillaFrequency(3, {'n_frequencies': n_freq_xyz})
        self.enc_dir = mlp.VanillaFrequency(3, {'n_frequencies': n_freq_dir})
        self.geo = mlp.VanillaMLP(self.enc_xyz.n_output_dims, feature_dim, {'n_neurons': 64, 'n_hidden_layers': 1, 'output_activation': 'none'})
        self.tex = mlp.VanillaMLP(feature_dim + self.enc_dir.n_output_dims, 3, {'n_neurons': 64, 'n_hidden_layers': 2, 'output_activation': 'none'})

    def field(self, positions, dirs, radius, density_only=False):
        out = self.geo(self.enc_xyz(contraction.contract_to_unisphere(positions, radius, contraction.AABB)))
        density = trunc_exp(out[:, 0] - 1.0)
        if density_only:
            return density, None
        rgb = torch.sigmoid(self.tex(torch.cat([out, self.enc_dir((dirs + 1.) / 2.)], dim=-1)))
        return density, rgb


def vanilla_nerf_render(fields, rays, binary, radius, step, bg_color, jitter=None, early_stop_eps=1e-4):
    rays = np.asarray(rays, np.float32)
    o, d = rays[:, :3], rays[:, 3:6]
    n_rays = len(rays)
    aabb = np.array([-radius] * 3 + [radius] * 3, np.float32)
    t0, t1 = march.ray_interval(o, d, aabb, None, None, step, jitter)
    ri, ts, te, _ = march.march_lattice(o, d, aabb, binary, step, t0, t1)
    ri_t, ts_t, te_t = torch.from_numpy(ri).long(), _t(ts)[:, None], _t(te)[:, None]
    ot, dt = _t(o), _t(d)
    with torch.no_grad():
        sig, _ = fields.field(ot[ri_t] + dt[ri_t] * (ts_t + te_t) / 2., None, radius, density_only=True)
        keep, _ = render.render_visibility((1.0 - torch.exp(-sig[:, None] * (te_t - ts_t))).view(-1), ri_t, n_rays, early_stop_eps, 0.0)
    n_marched = len(ri)
    ri_t, ts_t, te_t = ri_t[keep], ts_t[keep], te_t[keep]
    mid = (ts_t + te_t) / 2.
    density, rgb = fields.field(ot[ri_t] + dt[ri_t] * mid, dt[ri_t], radius)
    w = render.render_weight_from_density(ts_t, te_t, density[:, None], ri_t, n_rays)
    opacity = render.accumulate_along_rays(w, ri_t, None, n_rays)
    comp = render.accumulate_along_rays(w, ri_t, rgb, n_rays) + bg_color * (1.0 - opacity)
    return {'comp_rgb': comp, 'opacity': opacity, 'rays_valid': opacity > 0, 'num_samples': len(ts_t), 'num_marched': n_marched}


# --------------------------------------------------------------------------------------------------
# NeuS
# --------------------------------------------------------------------------------------------------
class NeusParams:
    """neus-blender fields: hash table (VolumeSDF.encoding), VanillaMLP SDF network (weight-norm,
    sphere-init), colour network flat vector (FullyFused 32->64->64->3), variance scalar."""

    def __init__(self, grid_cfg, table_flat, sdf_mlp, color_flat, variance, feature_dim=13, color_hidden=2):
        self.lt = hashgrid.level_table(grid_cfg)
        self.table_flat, self.sdf_mlp, self.color_flat, self.variance = table_flat, sdf_mlp, color_flat, variance
        self.feature_dim, self.color_hidden = feature_dim, color_hidden


def neus_render(P, rays, binary, radius, step, bg_color, cos_anneal_ratio, jitter=None, emulate_fp16=True):
    """NeuSModel.forward_ without learned background (models/neus.py:205-287)."""
    q = mlp.round_half if emulate_fp16 else (lambda t: t)   # values only: see oracle.mlp._RoundHalf
    rays = np.asarray(rays, np.float32)
    o, d = rays[:, :3], rays[:, 3:6]
    n_rays = len(rays)
    aabb = np.array([-radius] * 3 + [radius] * 3, np.float32)
    t0, t1 = march.ray_interval(o, d, aabb, None, None, step, jitter)
    ri, ts, te, _ = march.march_lattice(o, d, aabb, binary, step, t0, t1)
    ri_t, ts_t, te_t = torch.from_numpy(ri).long(), _t(ts)[:, None], _t(te)[:, None]
    ot, dt = _t(o), _t(d)
    mid = (ts_t + te_t) / 2.
    t_dirs = dt[ri_t]
    pos = (ot[ri_t] + t_dirs * mid).requires_grad_(True)
    dists = te_t - ts_t
    x01 = contraction.contract_to_unisphere(pos, radius, contraction.AABB)
    table = P.table_flat.view(-1, 2)
    table = q(table) if emulate_fp16 else table
    enc = hashgrid.hashgrid_fwd(x01, table, P.lt, compute_dtype=torch.float32)
    out = P.sdf_mlp(torch.cat([x01 * 2. - 1., enc], dim=-1)).float()
    sdf, feature = out[:, 0], out
    grad, = torch.autograd.grad(sdf, pos, torch.ones_like(sdf), create_graph=True)
    normal = F.normalize(grad, p=2, dim=-1)
    inv_s = neus.inv_s_from_variance(P.variance)
    alpha = neus.get_alpha(sdf, normal, t_dirs, dists, inv_s, cos_anneal_ratio)[:, None]
    shv = q(sh.sh4((t_dirs + 1.) / 2.))
    if getattr(P, 'color_mlp', None) is not None:   # neus-dtu.yaml:58-70: the colour network is the reference's VanillaMLP (fp32, biases)
        rgb = torch.sigmoid(P.color_mlp(torch.cat([feature, shv, normal], dim=-1)))
    else:
        rgb = torch.sigmoid(mlp.ffmlp_fwd(torch.cat([feature, shv, normal], dim=-1), P.color_flat, P.feature_dim + 16 + 3, 3, 64, P.color_hidden,
                                          'ReLU', 'None', emulate_fp16=emulate_fp16))
    w = render.render_weight_from_alpha(alpha, ri_t, n_rays)
    opacity = render.accumulate_along_rays(w, ri_t, None, n_rays)
    depth = render.accumulate_along_rays(w, ri_t, mid, n_rays)
    comp = render.accumulate_along_rays(w, ri_t, rgb, n_rays)
    comp_normal = F.normalize(render.accumulate_along_rays(w, ri_t, normal, n_rays), p=2, dim=-1)
    return {'comp_rgb': comp, 'comp_normal': comp_normal, 'opacity': opacity, 'depth': depth, 'rays_valid': opacity > 0,
            'num_samples': torch.tensor([len(ts_t)], dtype=torch.int32), 'sdf_samples': sdf, 'sdf_grad_samples': grad,
            'weights': w.view(-1), 'points': mid.view(-1), 'intervals': dists.view(-1), 'ray_indices': ri_t,
            'comp_rgb_full': comp + bg_color * (1.0 - opacity), 'inv_s': torch.exp(P.variance * 10.0), 'alpha': alpha.view(-1), 'rgb': rgb}


# --------------------------------------------------------------------------------------------------
# NeuS with learned background (config C4: neus-dtu.yaml)
# --------------------------------------------------------------------------------------------------
class NeusBgParams:
    """background fields of neus-dtu.yaml:72-105: hash table + VanillaMLP density network (32 -> 64 -> 8) and VanillaMLP colour network
    ([feature 8 | SH4 16] -> 64 -> 64 -> 3), both passed as modules (the reference's own arithmetic, pinned by the golden vectors)."""

    def __init__(self, grid_cfg, table_flat, density_mlp, color_mlp, density_bias=-1.0):
        self.lt = hashgrid.level_table(grid_cfg)
        self.table_flat, self.density_mlp, self.color_mlp, self.density_bias = table_flat, density_mlp, color_mlp, density_bias


def neus_bg_field(P, positions, dirs, radius, emulate_fp16=True, density_only=False):
    """VolumeDensity (UN_BOUNDED_SPHERE contraction) + VolumeRadiance with VanillaMLPs (geometry.py:122-130, texture.py:23-30)"""
    q = mlp.round_half if emulate_fp16 else (lambda t: t)   # values only: see oracle.mlp._RoundHalf
    x01 = contraction.contract_to_unisphere(positions, radius, contraction.UN_BOUNDED_SPHERE)
    table = P.table_flat.view(-1, 2)
    table = q(table) if emulate_fp16 else table
    enc = q(hashgrid.hashgrid_fwd(x01.detach(), table, P.lt, compute_dtype=torch.float32, one_gather=True))
    out = P.density_mlp(enc).float()
    density = trunc_exp(out[:, 0] + P.density_bias)
    if density_only:
        return density, None
    shv = q(sh.sh4((dirs + 1.) / 2.))
    return density, torch.sigmoid(P.color_mlp(torch.cat([out, shv], dim=-1)))


def neus_bg_render(P, rays, binary_bg, radius, step, cone_angle, near_plane, far_plane, bg_color, emulate_fp16=True, early_stop_eps=1e-4):
    """NeuSModel.forward_bg_ (models/neus.py:141-203): start where the ray leaves the foreground box (or at near_plane when it misses
    it), blind cone stepping through the contracted 256^3 grid, sigma_fn visibility pre-pass, NeRF-style compositing."""
    rays = np.asarray(rays, np.float32)
    o, d = rays[:, :3], rays[:, 3:6]
    n_rays = len(rays)
    aabb = np.array([-radius] * 3 + [radius] * 3, np.float32)
    _, t_box = march.ray_aabb_intersect(o, d, aabb)
    near = np.where(t_box > 1e9, np.float32(near_plane), t_box).astype(np.float32)   # neus.py:157
    t0, t1 = march.ray_interval(o, d, None, near, far_plane, step, None)
    ri, ts, te, _ = march.march_sequential(o, d, aabb, binary_bg, step, cone_angle, t0, t1, march.UN_BOUNDED_SPHERE)
    ri_t, ts_t, te_t = torch.from_numpy(ri).long(), _t(ts)[:, None], _t(te)[:, None]
    ot, dt = _t(o), _t(d)
    with torch.no_grad():
        sig, _ = neus_bg_field(P, ot[ri_t] + dt[ri_t] * (ts_t + te_t) / 2., None, radius, emulate_fp16, density_only=True)
        keep, _ = render.render_visibility((1.0 - torch.exp(-sig[:, None] * (te_t - ts_t))).view(-1), ri_t, n_rays, early_stop_eps, 0.0)
    n_marched = len(ri)
    ri_t, ts_t, te_t = ri_t[keep], ts_t[keep], te_t[keep]
    mid = (ts_t + te_t) / 2.
    density, rgb = neus_bg_field(P, ot[ri_t] + dt[ri_t] * mid, dt[ri_t], radius, emulate_fp16)
    w = render.render_weight_from_density(ts_t, te_t, density[:, None], ri_t, n_rays)
    opacity = render.accumulate_along_rays(w, ri_t, None, n_rays)
    depth = render.accumulate_along_rays(w, ri_t, mid, n_rays)
    comp = render.accumulate_along_rays(w, ri_t, rgb, n_rays) + bg_color * (1.0 - opacity)
    return {'comp_rgb': comp, 'opacity': opacity, 'depth': depth, 'rays_valid': opacity > 0,
            'num_samples': torch.tensor([len(ts_t)], dtype=torch.int32), 'num_marched': n_marched,
            'weights': w.view(-1), 'points': mid.view(-1), 'intervals': (te_t - ts_t).view(-1), 'ray_indices': ri_t.view(-1)}


def neus_dtu_render(P, Pbg, rays, binary, binary_bg, radius, step, bg_step, bg_cone_angle, bg_near, bg_far, bg_color, cos_anneal_ratio,
                    emulate_fp16=True):
    """NeuSModel.forward_ with learned_background (models/neus.py:205-287): foreground NeuS pass, background pass, composition
    ``comp_rgb_full = comp_rgb + comp_rgb_bg * (1 - opacity)`` and the *_bg / *_full dictionary keys."""
    fg = neus_render(P, rays, binary, radius, step, bg_color, cos_anneal_ratio, jitter=None, emulate_fp16=emulate_fp16)
    bg = neus_bg_render(Pbg, rays, binary_bg, radius, bg_step, bg_cone_angle, bg_near, bg_far, bg_color, emulate_fp16=emulate_fp16)
    out = {k: v for k, v in fg.items() if k not in ('comp_rgb_full', 'alpha', 'rgb', 'inv_s')}
    out.update({k + '_bg': v for k, v in bg.items() if k != 'num_marched'})
    out['comp_rgb_full'] = fg['comp_rgb'] + bg['comp_rgb'] * (1.0 - fg['opacity'])
    out['num_samples_full'] = fg['num_samples'] + bg['num_samples']
    out['rays_valid_full'] = fg['rays_valid'] | bg['rays_valid']
    return out

