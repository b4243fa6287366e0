"""CPU stand-ins for the two third-party CUDA packages the reference imports (``tinycudann``, ``nerfacc`` 0.3.3), built from the oracle's
PER-OP restatements (oracle/hashgrid.py, mlp.py, sh.py, march.py, render.py).  Test infrastructure only: they let the UNMODIFIED reference
``models/nerf.py`` / ``models/neus.py`` execute ``forward_`` on the CPU, so that the oracle's ORCHESTRATION (oracle/models.py:
``nerf_render`` / ``neus_render``, restated from those files) can be pinned against the reference's own code running here.  Everything is
fp32 without fp16 rounding (the comparison is about glue, not about half precision)."""
import enum
import types

import numpy as np
import torch
import torch.nn as nn

from oracle import hashgrid, mlp as omlp, sh as osh, march, render


# ---- tinycudann ---------------------------------------------------------------------------------------------------------------
class Encoding(nn.Module):
    def __init__(self, n_input_dims, encoding_config, dtype=None):
        super().__init__()
        self.n_input_dims, self.cfg = n_input_dims, dict(encoding_config)
        if self.cfg['otype'] == 'HashGrid':
            self.lt = hashgrid.level_table(self.cfg)
            self.n_output_dims = self.lt['n_output_dims']
            self.params = nn.Parameter(hashgrid.init_table(self.lt).flatten() * 200.0)   # rougher than tcnn's 1e-4 so the table matters
        elif self.cfg['otype'] == 'SphericalHarmonics':
            assert self.cfg['degree'] == 4
            self.n_output_dims = 16
            self.params = nn.Parameter(torch.zeros(0))
        else:
            raise NotImplementedError(self.cfg['otype'])

    def forward(self, x):
        if self.cfg['otype'] == 'HashGrid':
            return hashgrid.hashgrid_fwd(x.float(), self.params.view(-1, 2), self.lt, compute_dtype=torch.float32, one_gather=True)
        return osh.sh4(x.float())


class Network(nn.Module):
    def __init__(self, n_input_dims, n_output_dims, network_config):
        super().__init__()
        self.n_input_dims, self.n_output_dims, self.cfg = n_input_dims, n_output_dims, dict(network_config)
        self.params = nn.Parameter(omlp.ffmlp_init(n_input_dims, n_output_dims, 64, self.cfg['n_hidden_layers'], seed=11))

    def forward(self, x):
        return omlp.ffmlp_fwd(x.float(), self.params, self.n_input_dims, self.n_output_dims, 64, self.cfg['n_hidden_layers'],
                              self.cfg.get('activation', 'ReLU'), self.cfg.get('output_activation', 'None'), emulate_fp16=False)


class NetworkWithInputEncoding(nn.Module):
    def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config):
        super().__init__()
        self.n_input_dims, self.n_output_dims, self.ncfg = n_input_dims, n_output_dims, dict(network_config)
        self.lt = hashgrid.level_table(dict(encoding_config))
        self.n_mlp = omlp.ffmlp_layout(self.lt['n_output_dims'], n_output_dims, 64, self.ncfg['n_hidden_layers'])[1]
        table = hashgrid.init_table(self.lt).flatten() * 200.0
        self.params = nn.Parameter(torch.cat([omlp.ffmlp_init(self.lt['n_output_dims'], n_output_dims, 64, self.ncfg['n_hidden_layers'], seed=12), table]))

    def forward(self, x):
        enc = hashgrid.hashgrid_fwd(x.float().detach(), self.params[self.n_mlp:].view(-1, 2), self.lt, compute_dtype=torch.float32, one_gather=True)
        return omlp.ffmlp_fwd(enc, self.params[:self.n_mlp], self.lt['n_output_dims'], self.n_output_dims, 64, self.ncfg['n_hidden_layers'],
                              self.ncfg.get('activation', 'ReLU'), self.ncfg.get('output_activation', 'None'), emulate_fp16=False)


def tinycudann_module():
    m = types.ModuleType('tinycudann')
    m.Encoding, m.Network, m.NetworkWithInputEncoding, m.free_temporary_memory = Encoding, Network, NetworkWithInputEncoding, lambda: None
    return m


# ---- nerfacc 0.3.3 ------------------------------------------------------------------------------------------------------------
class ContractionType(enum.Enum):
    AABB = 0
    UN_BOUNDED_TANH = 1
    UN_BOUNDED_SPHERE = 2


class OccupancyGrid(nn.Module):
    def __init__(self, roi_aabb, resolution=128, contraction_type=ContractionType.AABB):
        super().__init__()
        self.register_buffer('_roi_aabb', torch.as_tensor(roi_aabb, dtype=torch.float32))
        self.contraction_type = contraction_type
        self.register_buffer('resolution', torch.tensor([resolution] * 3, dtype=torch.int32))     # nerfacc 0.3.3's checkpointed state
        self.register_buffer('occs', torch.zeros(resolution ** 3))
        self.register_buffer('_binary', torch.zeros([resolution] * 3, dtype=torch.bool))

    @property
    def binary(self):
        return self._binary

    def every_n_step(self, step, occ_eval_fn, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16):
        # the grid itself is set by the test; keep what the reference handed over so its occupancy functions can be checked
        self.last_call = dict(step=step, occ_eval_fn=occ_eval_fn, occ_thre=occ_thre)


@torch.no_grad()
def ray_marching(rays_o, rays_d, t_min=None, t_max=None, scene_aabb=None, grid=None, sigma_fn=None, alpha_fn=None, early_stop_eps=1e-4,
                 alpha_thre=0.0, near_plane=None, far_plane=None, render_step_size=1e-3, stratified=False, cone_angle=0.0, jitter=None):
    """nerfacc 0.3.3 ray_marching semantics from the oracle's marchers: the step lattice for AABB grids with
    cone_angle 0, blind cone stepping through the grid's own (contracted) region otherwise; then the sigma_fn visibility filter"""
    assert alpha_fn is None and grid is not None
    if stratified and jitter is None:   # nerfacc draws one U[0,1) offset per ray
        jitter = torch.rand(rays_o.shape[0])
    o, d = rays_o.numpy().astype(np.float32), rays_d.numpy().astype(np.float32)
    step = np.float32(render_step_size)
    near = near_plane.numpy().astype(np.float32) if torch.is_tensor(near_plane) else near_plane
    box = None if scene_aabb is None else scene_aabb.numpy().astype(np.float32)
    jit = jitter.cpu().numpy().astype(np.float32) if (stratified and jitter is not None) else None
    t0, t1 = march.ray_interval(o, d, box, near, far_plane, step, jit)
    roi = grid._roi_aabb.numpy().astype(np.float32)
    if grid.contraction_type.value == ContractionType.AABB.value and cone_angle == 0.0:   # (by value: the product has its own enum class)
        ri, ts, te, _ = march.march_lattice(o, d, roi, grid.binary.numpy(), step, t0, t1)
    else:
        ri, ts, te, _ = march.march_sequential(o, d, roi, grid.binary.numpy(), step, cone_angle, t0, t1, grid.contraction_type.value)
    ri_t, ts_t, te_t = torch.from_numpy(ri).long(), torch.from_numpy(ts)[:, None], torch.from_numpy(te)[:, None]
    if sigma_fn is not None:
        sig = sigma_fn(ts_t, te_t, ri_t)
        alphas = 1.0 - torch.exp(-sig * (te_t - ts_t))
        keep, _ = render.render_visibility(alphas.view(-1), ri_t, len(o), early_stop_eps, alpha_thre)
        ri_t, ts_t, te_t = ri_t[keep], ts_t[keep], te_t[keep]
    return ri_t.int(), ts_t, te_t


def render_weight_from_density(t_starts, t_ends, sigmas, *, packed_info=None, ray_indices=None, n_rays=None):
    return render.render_weight_from_density(t_starts, t_ends, sigmas, ray_indices, n_rays)


def render_weight_from_alpha(alphas, *, packed_info=None, ray_indices=None, n_rays=None):
    return render.render_weight_from_alpha(alphas, ray_indices, n_rays)


def accumulate_along_rays(weights, ray_indices, values=None, n_rays=None):
    return render.accumulate_along_rays(weights, ray_indices, values, n_rays)


def nerfacc_modules():
    m = types.ModuleType('nerfacc')
    m.ContractionType, m.OccupancyGrid, m.ray_marching = ContractionType, OccupancyGrid, ray_marching
    m.render_weight_from_density, m.render_weight_from_alpha, m.accumulate_along_rays = render_weight_from_density, render_weight_from_alpha, \
        accumulate_along_rays
    inter = types.ModuleType('nerfacc.intersection')
    inter.ray_aabb_intersect = lambda o, d, aabb: tuple(torch.from_numpy(a) for a in march.ray_aabb_intersect(o.numpy(), d.numpy(), aabb.numpy()))
    m.intersection = inter
    return m, inter
