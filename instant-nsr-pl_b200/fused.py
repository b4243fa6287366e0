"""Fused NeRF render path: host-side orchestration of the fused kernels behind ``NeRFModel.forward_``
(models/nerf.py:61-127 of the reference).  Per step:

    march (count / scan / write)  ->  density pre-pass  ->  visibility scan  ->  prefix compaction
    ->  render forward (hash + both MLPs + compositing, one launch)
    <-  ray backward (compositing)  <-  field backward (MLPs + hash scatter, one launch)

All arithmetic is in libnsr_b200.so; this file only allocates outputs, passes pointers and wires autograd.
"""
import ctypes

import torch

from . import ops
from .lib import lib, ptr, stream, check_cuda, contig, NerfT
from .nerfacc import ContractionType


class _NerfRender(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dparams, cparams, fused, rays, jitter, keep_aux):
        st = fused.trace(rays, jitter)
        n_rays, k = rays.shape[0], st['k']
        dev = rays.device
        acc_rgb = torch.zeros(n_rays, 3, device=dev)
        opacity = torch.zeros(n_rays, 1, device=dev)
        depth = torch.zeros(n_rays, 1, device=dev)
        sig = torch.empty(k, device=dev)
        rgbs = torch.empty(k, 3, device=dev)
        weights = torch.empty(k, device=dev)
        need_grad = dparams.requires_grad or cparams.requires_grad
        enc = torch.empty(k, 32, dtype=torch.float16, device=dev) if need_grad else None
        dh, ch = fused.dparams_half(), fused.cparams_half()
        lib.call('nsr_nerf_render_fwd', fused.ref(), ptr(rays), ptr(st['ri']), ptr(st['ts']), ptr(st['te']), ptr(st['trans']), ptr(dh),
                 ptr(ch), ptr(enc), ptr(sig), ptr(rgbs), ptr(weights), ptr(acc_rgb), ptr(opacity), ptr(depth), k, stream())
        ctx.fused = fused
        ctx.n_rays, ctx.k = n_rays, k
        ctx.save_for_backward(rays, st['ri'], st['ts'], st['te'], st['trans'], st['offsets_k'], enc, sig, rgbs, weights, dh, ch)
        ctx.mark_non_differentiable(st['ri'], st['ts'], st['te'])
        fused.last_stats = {'n_marched': st['m'], 'n_kept': k}
        return acc_rgb, opacity, depth, weights, st['ri'], st['ts'], st['te']

    @staticmethod
    def backward(ctx, g_rgb, g_op, g_depth, g_w, *_):
        fused = ctx.fused
        rays, ri, ts, te, trans, offsets_k, enc, sig, rgbs, weights, dh, ch = ctx.saved_tensors
        dev = rays.device
        k, n_rays = ctx.k, ctx.n_rays
        gd = torch.zeros(fused.n_dparams, device=dev)
        gc = torch.zeros(fused.n_cparams, device=dev)
        if k > 0:
            d_sraw = torch.empty(k, device=dev)
            d_rgb = torch.empty(k, 3, device=dev)
            amax = torch.zeros(1, device=dev)
            lib.call('nsr_nerf_ray_bwd', ptr(offsets_k), ptr(ts), ptr(te), ptr(trans), ptr(weights), ptr(sig), ptr(rgbs),
                     ptr(contig(g_rgb, torch.float32)), ptr(contig(g_op, torch.float32)), ptr(contig(g_depth, torch.float32)),
                     ptr(contig(g_w, torch.float32)), ptr(d_sraw), ptr(d_rgb), ptr(amax), n_rays, stream())
            lib.call('nsr_nerf_field_bwd', fused.ref(), ptr(rays), ptr(ri), ptr(ts), ptr(te), ptr(enc), ptr(dh), ptr(ch), ptr(d_sraw),
                     ptr(d_rgb), ptr(gd), ptr(gc), float(fused.loss_scale), ptr(amax), k, stream())
        return gd, gc, None, None, None, None


class NerfFused:
    """Fused executor attached to a NeRFModel whose config has the nerf-blender shape."""

    def __init__(self, model):
        self.model = model
        geo, tex = model.geometry, model.texture
        self.net = geo.encoding_with_network           # tcnn.NetworkWithInputEncoding
        self.cnet = tex.network                        # tcnn.Network
        self.grid = self.net.grid
        self.n_dparams, self.n_cparams = self.net.params.numel(), self.cnet.params.numel()
        s = NerfT()
        s.grid = self.grid.struct
        s.radius = float(model.config.radius)
        s.density_bias = float(geo.config.density_bias)
        s.feature_dim, s.density_hidden, s.color_hidden = 16, 1, 2
        self.struct = s
        self.loss_scale = 0.0  # <= 0: chosen on the device from the incoming gradient magnitude
        self.early_stop_eps, self.alpha_thre = 1e-4, 0.0
        self.last_stats = {}

    @staticmethod
    def try_build(model):
        """Return a NerfFused if the model is exactly the shape the fused kernels implement, else None."""
        from . import tcnn
        from .models.fields import VolumeDensity, VolumeRadiance
        cfg = model.config
        geo, tex = model.geometry, model.texture
        try:
            ok = (not cfg.learned_background and cfg.grid_prune and isinstance(geo, VolumeDensity) and isinstance(tex, VolumeRadiance)
                  and isinstance(geo.encoding_with_network, tcnn.NetworkWithInputEncoding)
                  and geo.encoding_with_network.grid.n_levels == 16 and geo.encoding_with_network.mlp.n_hidden == 1
                  and geo.n_output_dims == 16 and geo.config.get('density_activation') == 'trunc_exp'
                  and 'feature_activation' not in geo.config
                  and isinstance(tex.network, tcnn.Network) and tex.network.mlp.n_hidden == 2 and tex.network.mlp.n_in == 32
                  and isinstance(tex.encoding.encoding, tcnn.Encoding) and tex.encoding.encoding.otype == 'SphericalHarmonics'
                  and not tex.encoding.include_xyz and tex.config.input_feature_dim == 16)
            net_act = str(tex.network.network_config.get('output_activation', 'None')).lower()
            col_act = str(tex.config.get('color_activation', 'none')).lower()
            ok = ok and sorted([net_act, col_act]) == ['none', 'sigmoid']
        except AttributeError:
            ok = False
        return NerfFused(model) if ok else None

    def ref(self):
        return ctypes.byref(self.struct)

    def dparams_half(self):
        return self.net._params_half()

    def cparams_half(self):
        return self.cnet._params_half()

    @torch.no_grad()
    def density(self, positions):
        """density at world positions (occ_eval_fn, models/nerf.py:49-52)."""
        check_cuda(positions, what='NerfFused.density')
        p = contig(positions.reshape(-1, 3), torch.float32)
        out = torch.empty(p.shape[0], device=p.device)
        lib.call('nsr_nerf_density', self.ref(), ptr(p), ptr(self.dparams_half()), ptr(out), p.shape[0], stream())
        return out.reshape(positions.shape[:-1])

    @torch.no_grad()
    def trace(self, rays, jitter=None):
        """march + sigma_fn visibility pre-pass + compaction: the `with torch.no_grad(): ray_marching(...)`
        block of models/nerf.py:82-93.  Returns the kept samples with their exclusive transmittance."""
        m = self.model
        dev = rays.device
        n = rays.shape[0]
        rays_o, rays_d = rays[:, 0:3].contiguous(), rays[:, 3:6].contiguous()
        t_min, t_max = ops.ray_aabb_intersect(rays_o, rays_d, m.scene_aabb)
        if m.randomized:
            u = torch.rand(n, device=dev) if jitter is None else jitter.to(dev, torch.float32)
            t_min = t_min + u * m.render_step_size
        grid = m.occupancy_grid
        ms = ops.march_struct(grid.roi_aabb.tolist(), grid._res, ContractionType.AABB.value, m.render_step_size, 0.0)
        ri_m, ts_m, te_m, off_m = ops.march(ms, rays_o, rays_d, t_min.contiguous(), t_max.contiguous(), grid.bits())
        mcount = ri_m.shape[0]
        offsets_k = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        if mcount == 0:
            e = torch.empty(0, device=dev)
            return {'ri': ri_m, 'ts': e, 'te': e, 'trans': e, 'offsets_k': offsets_k, 'k': 0, 'm': 0}
        alphas = torch.empty(mcount, device=dev)
        lib.call('nsr_nerf_prepass', self.ref(), ptr(rays), ptr(ri_m), ptr(ts_m), ptr(te_m), ptr(self.dparams_half()), ptr(alphas), mcount,
                 stream())
        keep = torch.empty(mcount, dtype=torch.uint8, device=dev)
        trans = torch.empty(mcount, device=dev)
        kept = torch.empty(n, dtype=torch.int32, device=dev)
        lib.call('nsr_visibility', ptr(alphas), ptr(off_m), ptr(keep), ptr(trans), ptr(kept), self.early_stop_eps, self.alpha_thre, n, stream())
        lib.call('nsr_scan_counts', ptr(kept), ptr(offsets_k), n, stream())
        k = int(offsets_k[n].item())  # exact-size outputs are part of the reference's return contract
        ri = torch.empty(k, dtype=torch.int32, device=dev)
        ts, te, tr = torch.empty(k, device=dev), torch.empty(k, device=dev), torch.empty(k, device=dev)
        if k > 0:
            lib.call('nsr_compact_prefix', ptr(off_m), ptr(offsets_k), ptr(ri_m), ptr(ts_m), ptr(te_m), ptr(trans), ptr(ri), ptr(ts), ptr(te),
                     ptr(tr), n, stream())
        return {'ri': ri, 'ts': ts, 'te': te, 'trans': tr, 'offsets_k': offsets_k, 'k': k, 'm': mcount}

    def render(self, rays, jitter=None):
        """NeRFModel.forward_ (models/nerf.py:61-127) -> the reference's output dict."""
        m = self.model
        check_cuda(rays, what='NeRFModel')
        rays = contig(rays, torch.float32)
        acc_rgb, opacity, depth, weights, ri, ts, te = _NerfRender.apply(self.net.params, self.cnet.params, self, rays, jitter, m.training)
        comp_rgb = acc_rgb + m.background_color * (1.0 - opacity)
        out = {'comp_rgb': comp_rgb, 'opacity': opacity, 'depth': depth, 'rays_valid': opacity > 0,
               'num_samples': torch.as_tensor([ts.shape[0]], dtype=torch.int32, device=rays.device)}
        if m.training:
            out.update({'weights': weights.view(-1), 'points': ((ts + te) / 2.).view(-1), 'intervals': (te - ts).view(-1),
                        'ray_indices': ri.long().view(-1)})
        return out
