"""Fused NeRF render path: host-side orchestration of the fused kernels behind ``NeRFModel.forward_``
(models/nerf.py:61-127 of the reference).  Per step:

    march (count / scan / write)  ->  density pre-pass  ->  visibility scan  ->  prefix compaction
    ->  render forward (hash + both MLPs + compositing, one launch)
    <-  ray backward (compositing)  <-  field backward (MLPs + hash scatter, one launch)

Sample counts stay on the device (capacity-sized buffers + device-side counters), so the pipeline contains no
host synchronisation and can be captured in a CUDA graph (``nsr_b200.graph.GraphedStep``).  The reference's
exact-size output contract (``ray_indices/weights/...`` of length K) costs one device->host read at the end of
the forward.  All arithmetic is in libnsr_b200.so; this file only allocates, passes pointers and wires autograd.
"""
import ctypes
import math

import torch

from . import ops
from .lib import lib, ptr, stream, check_cuda, contig, NerfT
from .nerfacc import ContractionType


class _NerfRender(torch.autograd.Function):
    """(dparams, cparams) -> per-ray sums + per-sample weights; everything else rides along non-differentiably."""

    @staticmethod
    def forward(ctx, dparams, cparams, fused, rays, jitter):
        st = fused.trace(rays, jitter)
        n_rays, cap = rays.shape[0], st['cap']
        dev = rays.device
        acc_rgb = torch.zeros(n_rays, 3, device=dev)
        opacity = torch.zeros(n_rays, 1, device=dev)
        depth = torch.zeros(n_rays, 1, device=dev)
        sig = torch.empty(cap, device=dev)
        rgbs = torch.empty(cap, 3, device=dev)
        weights = torch.empty(cap, device=dev)
        need_grad = dparams.requires_grad or cparams.requires_grad
        enc = torch.empty(cap, 32, dtype=torch.float16, device=dev) if need_grad else None
        dh, ch = fused.dparams_half(), fused.cparams_half()
        k_dev = st['offsets_k'][n_rays:]
        lib.call('nsr_nerf_render_fwd', fused.ref(), ptr(rays), ptr(st['ri']), ptr(st['ts']), ptr(st['te']), ptr(st['trans']), ptr(dh),
                 ptr(ch), ptr(enc), ptr(sig), ptr(rgbs), ptr(weights), ptr(acc_rgb), ptr(opacity), ptr(depth), cap, ptr(k_dev), stream())
        ctx.fused, ctx.n_rays, ctx.cap = fused, n_rays, cap
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(rays, st['ri'], st['ts'], st['te'], st['trans'], st['offsets_k'], enc, sig, rgbs, weights, dh, ch)
        counts = torch.cat([st['offsets_m'][n_rays:], k_dev])  # [M, K] on the device
        ctx.mark_non_differentiable(st['ri'], st['ts'], st['te'], counts)
        return acc_rgb, opacity, depth, weights, st['ri'], st['ts'], st['te'], counts

    @staticmethod
    def backward(ctx, g_rgb, g_op, g_depth, g_w, *_):
        fused = ctx.fused
        rays, ri, ts, te, trans, offsets_k, enc, sig, rgbs, weights, dh, ch = ctx.saved_tensors
        dev = rays.device
        n_rays, cap = ctx.n_rays, ctx.cap
        gd = torch.zeros(fused.n_dparams, device=dev)
        gc = torch.zeros(fused.n_cparams, device=dev)
        if cap > 0 and enc is not None:
            d_sraw = torch.empty(cap, device=dev)
            d_rgb = torch.empty(cap, 3, device=dev)
            amax = torch.zeros(1, device=dev)
            f32 = lambda t: None if t is None else contig(t, torch.float32)
            lib.call('nsr_nerf_ray_bwd', ptr(offsets_k), ptr(ts), ptr(te), ptr(trans), ptr(weights), ptr(sig), ptr(rgbs), ptr(f32(g_rgb)),
                     ptr(f32(g_op)), ptr(f32(g_depth)), ptr(f32(g_w)), ptr(d_sraw), ptr(d_rgb), ptr(amax), n_rays, stream())
            lib.call('nsr_nerf_field_bwd', fused.ref(), ptr(rays), ptr(ri), ptr(ts), ptr(te), ptr(enc), ptr(dh), ptr(ch), ptr(d_sraw),
                     ptr(d_rgb), ptr(gd), ptr(gc), float(fused.loss_scale), ptr(amax), cap, ptr(offsets_k[n_rays:]), None, None, stream())
        return gd, gc, None, None, None


class _NerfRenderRays(torch.autograd.Function):
    """Default fused path.  Forward: mask march (+ longest-rays-first order) -> ONE persistent per-ray kernel (gather, both
    MLPs, compositing with early ray termination) -> index pack.  Backward: per-ray compositing backward + the
    load-balanced sample-tile field backward, reading the per-ray ("loose") buffers through the packed->loose index.
    Loose layout: ray r's kept samples at offsets_m[r] + j, j < kept[r].
    fused.bwd_kernel = 'rays' swaps in the single per-ray backward kernel (nsr_nerf_rays_bwd; measured slower)."""

    @staticmethod
    def forward(ctx, dparams, cparams, fused, rays, jitter):
        m = fused.model
        dev = rays.device
        n = rays.shape[0]
        cap = n * fused.cap_per_ray
        mref = ctypes.byref(fused.march)
        u = None
        if m.randomized:
            u = torch.rand(n, device=dev) if jitter is None else contig(jitter.to(dev), torch.float32)
        grid = m.occupancy_grid
        bits, coarse = grid.bits(), grid.coarse_bits()
        i32 = lambda k: torch.empty(k, dtype=torch.int32, device=dev)
        f32 = lambda *k: torch.empty(*k, dtype=torch.float32, device=dev)
        i64 = lambda k: torch.empty(k, dtype=torch.int64, device=dev)
        words = (fused.cap_per_ray + 31) // 32
        need_grad = (dparams.requires_grad or cparams.requires_grad) and fused._want_grad   # (grad mode is always off inside Function.forward)
        # The backward accumulates into zeroed flat gradient buffers (50 MB for the table).  Zero them NOW on a side stream: the fill runs
        # beside the marcher / forward kernels instead of in front of the backward's first kernel; the backward joins the side stream.
        ctx.grad_bufs = None
        if need_grad and fused.prezero_grads:
            cur = torch.cuda.current_stream()
            side = fused.side_stream(dev)
            if fused.direct_grads is not None:
                gd0, gc0 = fused.direct_grads
            else:
                gd0, gc0 = torch.empty(fused.n_dparams, device=dev), torch.empty(fused.n_cparams, device=dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                gd0.zero_()
                gc0.zero_()
            gd0.record_stream(side)
            gc0.record_stream(side)
            ctx.grad_bufs = (gd0, gc0, side)
        masks, t_min, counts = i32(n * words), f32(n), i32(n)
        # one fill: the forward's ray ticket, the backward's gradient amax, the marcher's row allocator and its 8 queue-group counters
        nb = (n + 255) // 256
        zz = torch.zeros(12 + nb, dtype=torch.int32, device=dev)
        tick, amax0, m_total, bin_counts = zz[0:1], zz[1:2].view(torch.float32), zz[2:4].view(torch.int64), zz[4:12]
        kept_blocks = zz[12:] if fused.fuse_kept_scan else None   # per-256-ray sums of the kept counts (the pack kernel's prefix sum)
        if fused.march_alloc:
            # the marcher reserves every ray's rows and its place in the longest-first queue itself (atomics): no scan kernel behind it
            offsets_m, order = i64(n), i32(8 * n)
            lib.call('nsr_march_rays_alloc', mref, ptr(rays), ptr(u), ptr(bits), ptr(coarse), ptr(masks), words, ptr(t_min), ptr(counts), ptr(offsets_m),
                     ptr(m_total), ptr(bin_counts), ptr(order), n, stream())
        else:
            offsets_m, order = i64(n + 1), i32(n)
            lib.call('nsr_march_rays_mask', mref, ptr(rays), ptr(u), ptr(bits), ptr(coarse), ptr(masks), words, ptr(t_min), ptr(counts), n, stream())
            lib.call('nsr_scan_counts_order', ptr(counts), ptr(offsets_m), ptr(order), n, stream())
            m_total = offsets_m[n:]
        enc = torch.empty(cap, 32, dtype=torch.float16, device=dev) if need_grad else None
        sig, rgbs, weights, trans, kidx = f32(cap), f32(cap, 3), f32(cap), f32(cap), i32(cap)
        acc_rgb, opacity, depth, kept = f32(n, 3), f32(n, 1), f32(n, 1), i32(n)
        offsets_k = i64(n + 1)
        dh, ch = fused.dparams_half(), fused.cparams_half()
        step = float(m.render_step_size)
        lib.call('nsr_nerf_rays_fwd', fused.ref(), ptr(rays), ptr(masks), words, ptr(t_min), ptr(offsets_m), ptr(order), step,
                 float(fused.early_stop_eps), ptr(dh), ptr(ch), ptr(enc), ptr(sig), ptr(rgbs), ptr(weights), ptr(trans), ptr(kidx),
                 ptr(acc_rgb), ptr(opacity), ptr(depth), ptr(kept), ptr(tick), n, ptr(counts), ptr(bin_counts) if fused.march_alloc else None, ptr(kept_blocks), stream())
        if not fused.fuse_kept_scan:
            lib.call('nsr_scan_counts', ptr(kept), ptr(offsets_k), n, stream())
        # packed view of the kept samples: the reference's per-sample outputs + the row index of the tile backward
        # plus (training, tile backward) the backward's inputs in packed row order: encodings, unit-cube position + view direction
        ri, ts, te, pos = i32(cap), f32(cap), f32(cap), i64(cap)
        packed_bwd = need_grad and fused.bwd_kernel in ('tiles', 'tiles_split', 'tc') and fused.packed_bwd_inputs
        tiled = 1 if fused.bwd_kernel == 'tc' else 0   # tcgen05 backward: encodings in canonical 128-row UMMA tiles (one TMA bulk copy per tile)
        # (+pad rows: the tile backwards prefetch whole tiles -- 64 rows with cp.async, 128 rows with cp.async.bulk -- the last one may reach past K)
        pad = 256 - cap % 128 if tiled else 64
        enc_k = torch.empty(cap + pad, 32, dtype=torch.float16, device=dev) if packed_bwd else None
        xyzdir = f32(cap + pad, 6) if packed_bwd else None
        if fused.fuse_kept_scan:   # the packed offsets are computed inside the pack kernel (one launch and a one-CTA scan less)
            lib.call('nsr_pack_kept_scan', ptr(offsets_m), ptr(kept), ptr(offsets_k), ptr(t_min), step, ptr(kidx), ptr(weights), ptr(ri), ptr(ts),
                     ptr(te), None, ptr(pos), fused.ref(), ptr(rays), ptr(enc), ptr(enc_k), ptr(xyzdir), tiled, n, ptr(kept_blocks), stream())
        else:
            lib.call('nsr_pack_kept', ptr(offsets_m), ptr(offsets_k), ptr(t_min), step, ptr(kidx), ptr(weights), ptr(ri), ptr(ts), ptr(te), None,
                     ptr(pos), fused.ref(), ptr(rays), ptr(enc), ptr(enc_k), ptr(xyzdir), tiled, n, stream())
        ctx.fused, ctx.n_rays, ctx.cap = fused, n, cap
        ctx.set_materialize_grads(False)
        if packed_bwd:
            enc = None   # the loose copy is not needed any more
        ctx.save_for_backward(rays, t_min, offsets_m, offsets_k, kept, enc, sig, rgbs, weights, trans, kidx, ri, ts, te, pos, dh, ch,
                              enc_k, xyzdir, amax0)
        ctx.mark_non_differentiable(ri, ts, te, pos, offsets_m, offsets_k, m_total, counts)
        return acc_rgb, opacity, depth, weights, ri, ts, te, pos, offsets_m, offsets_k, m_total, counts

    @staticmethod
    def backward(ctx, g_rgb, g_op, g_depth, g_w, *_):
        fused = ctx.fused
        rays, t_min, offsets_m, offsets_k, kept, enc, sig, rgbs, weights, trans, kidx, ri, ts, te, pos, dh, ch, enc_k, xyzdir, amax0 = ctx.saved_tensors
        dev = rays.device
        n, cap = ctx.n_rays, ctx.cap
        step = float(fused.model.render_step_size)
        direct = fused.direct_grads
        pre, ctx.grad_bufs = getattr(ctx, 'grad_bufs', None), None   # (a retained-graph second backward allocates fresh buffers below)
        if pre is not None and (direct is None or pre[0] is direct[0]):
            gd, gc, side = pre                     # zeroed on the side stream while the forward ran
            torch.cuda.current_stream().wait_stream(side)
        elif direct is not None:   # accumulate straight into caller-owned buffers (the symmetric exchange buffer): autograd is bypassed
            gd, gc = direct
            gd.zero_()
            gc.zero_()
        else:
            gd = torch.zeros(fused.n_dparams, device=dev)
            gc = torch.zeros(fused.n_cparams, device=dev)
        if (enc is not None or enc_k is not None) and n > 0:
            f32 = lambda t: None if t is None else contig(t, torch.float32)
            amax = amax0   # zeroed together with the forward's ticket (a retained-graph second backward only makes the scale smaller)
            if fused.bwd_kernel == 'rays':
                tick = torch.zeros(1, dtype=torch.int32, device=dev)
                lib.call('nsr_nerf_rays_bwd', fused.ref(), ptr(rays), ptr(t_min), ptr(offsets_m), ptr(kept), step, ptr(enc), ptr(sig), ptr(rgbs),
                         ptr(weights), ptr(trans), ptr(kidx), ptr(dh), ptr(ch), ptr(f32(g_rgb)), ptr(f32(g_op)), ptr(f32(g_depth)),
                         ptr(f32(g_w)), ptr(gd), ptr(gc), float(fused.loss_scale), ptr(amax), float(fused.t_bound), ptr(tick), n, stream())
            else:
                packed = enc_k is not None
                pad = 256 - cap % 128 if fused.bwd_kernel == 'tc' else 64
                d_sraw = torch.empty(cap + pad, device=dev)
                d_rgb = torch.empty(cap + pad, 3, device=dev)   # gradients + encodings in packed row order: no index chains in front of the tile math
                lib.call('nsr_nerf_ray_bwd_loose', ptr(offsets_m), ptr(kept), ptr(t_min), step, ptr(kidx), ptr(trans), ptr(weights), ptr(sig),
                         ptr(rgbs), ptr(f32(g_rgb)), ptr(f32(g_op)), ptr(f32(g_depth)), ptr(f32(g_w)), ptr(d_sraw), ptr(d_rgb), ptr(amax),
                         ptr(offsets_k) if packed else None, n, stream())
                if packed and fused.bwd_kernel == 'tc':
                    # Blackwell-native backward: tcgen05 GEMM chain + TMA-staged tiles + scatter warps in one kernel (csrc/nerf_bwd_tc.cu)
                    if fused._tc_status is None or fused._tc_status.device != dev:
                        fused._tc_status = torch.zeros(1, dtype=torch.int32, device=dev)
                    lib.call('nsr_nerf_field_bwd_tc', fused.ref(), ptr(enc_k), ptr(dh), ptr(ch), ptr(d_sraw), ptr(d_rgb), ptr(gd), ptr(gc),
                             float(fused.loss_scale), ptr(amax), cap, ptr(offsets_k[n:]), ptr(xyzdir), ptr(fused._tc_status), stream())
                elif packed and fused.bwd_kernel == 'tiles_split':
                    # network half + table half as two launches: the REDs come from a kernel with 64 light warps per SM (csrc/nerf_fused_bwd.cu)
                    denc = torch.empty(cap, 32, dtype=torch.float16, device=dev)
                    lib.call('nsr_nerf_field_bwd_net', fused.ref(), ptr(enc_k), ptr(dh), ptr(ch), ptr(d_sraw), ptr(d_rgb), ptr(gd), ptr(gc),
                             float(fused.loss_scale), ptr(amax), cap, ptr(offsets_k[n:]), ptr(xyzdir), ptr(denc), stream())
                    # table half; data-parallel runs scatter level groups in separate launches and hand each finished group to the gradient
                    # exchange (parallel.P2PGradSync.bind_pipelined), whose kernel runs beside the next group's scatter (which leaves it one CTA slot per SM)
                    groups = fused.level_groups or ((0, 16),)
                    for gi, (l0, l1) in enumerate(groups):
                        lib.call('nsr_nerf_table_scatter', ctypes.byref(fused.struct.grid), ptr(xyzdir), 6, ptr(denc), float(fused.loss_scale), ptr(amax),
                                 ptr(gd[fused.net.mlp.n_params:]), cap, ptr(offsets_k[n:]), l0, l1, 4 if (fused.exchange_hook is not None and gi > 0) else 0,
                                 stream())
                        if fused.exchange_hook is not None:
                            fused.exchange_hook(gi)
                else:
                    lib.call('nsr_nerf_field_bwd', fused.ref(), ptr(rays), ptr(ri), ptr(ts), ptr(te), ptr(enc_k if packed else enc), ptr(dh), ptr(ch),
                             ptr(d_sraw), ptr(d_rgb), ptr(gd), ptr(gc), float(fused.loss_scale), ptr(amax), cap, ptr(offsets_k[n:]),
                             None if packed else ptr(pos), ptr(xyzdir), stream())
        if direct is not None:
            return None, None, None, None, None
        return gd, gc, None, None, None


class NerfFused:
    """Fused executor attached to a NeRFModel whose config has the nerf-blender shape."""

    def __init__(self, model):
        self.model = model
        geo, tex = model.geometry, model.texture
        self.net = geo.encoding_with_network           # tcnn.NetworkWithInputEncoding
        self.cnet = tex.network                        # tcnn.Network
        self.grid = self.net.grid
        self.n_dparams, self.n_cparams = self.net.params.numel(), self.cnet.params.numel()
        r = float(model.config.radius)
        s = NerfT()
        s.grid = self.grid.struct
        s.radius = r
        s.density_bias = float(geo.config.density_bias)
        s.feature_dim, s.density_hidden, s.color_hidden = 16, 1, 2
        self.struct = s
        self.march = ops.march_struct([-r, -r, -r, r, r, r], model.occupancy_grid_res, ContractionType.AABB.value, model.render_step_size, 0.0)
        # a ray crosses at most the box diagonal: upper bound on marched samples per ray (capacity of the static buffers)
        self.cap_per_ray = int(math.ceil(2.0 * math.sqrt(3.0) * r / model.render_step_size)) + 2
        self.loss_scale = 0.0  # <= 0: chosen on the device from the incoming gradient magnitude
        self.early_stop_eps, self.alpha_thre = 1e-4, 0.0
        self.last_stats = {}
        self._ticket = None
        self.mode = 'per_ray'   # 'per_ray' (persistent per-ray forward kernel) | 'two_pass' (pre-pass / compaction / sample-tile kernels)
        self.lean_static_outputs = False   # static=True: skip the per-ray outputs the fused loss op produces itself (comp_rgb, rays_valid)
        self.packed_bwd_inputs = True   # tile backward reads its inputs in packed row order (written by nsr_pack_kept)
        from .config import experimental
        self.fuse_kept_scan = experimental('pack_scan')   # nsr_pack_kept_scan instead of nsr_scan_counts + nsr_pack_kept (not yet timed)
        # 'tiles_split' (default: network half + high-occupancy table scatter, 189 us) | 'tiles' (one sample-tile backward kernel, REDs from the
        # MMA warps, 202 us) | 'tc' (tcgen05 / TMA warp-specialised kernel, 208 us: csrc/nerf_bwd_tc.cu) | 'rays' (single per-ray backward kernel)
        import os
        self.bwd_kernel = os.environ.get('NSR_BWD_KERNEL', 'tiles_split')
        self._tc_status = None
        # (gd, gc) flat fp32 buffers the backward zeroes and accumulates into INSTEAD of handing gradients to autograd (per-ray path only;
        # set by parallel.P2PGradSync.bind_direct: the buffers are views of the peer-mapped exchange buffer and become .grad after the exchange)
        # the marcher allocates every ray's rows and queue slot itself (nsr_march_rays_alloc) instead of a one-CTA scan kernel behind it
        self.march_alloc = os.environ.get('NSR_MARCH_ALLOC', '1') == '1'
        self.direct_grads = None
        # zero the backward's gradient buffers beside the forward on a side stream instead of in front of the backward.  Measured on B200:
        # 0.411 vs 0.401 ms/step -- the fill kernel takes SMs from the persistent forward kernel at its launch -- so it stays opt-in.
        self.prezero_grads = os.environ.get('NSR_PREZERO_GRADS', '0') == '1'
        self._side = None
        self._want_grad = True
        self.level_groups = None    # ((l0, l1), ...): the split backward's table scatter as one launch per level group (top levels first)
        self.exchange_hook = None   # callable(group index): called behind each group's scatter launch (the gradient exchange of that group)
        self.t_bound = 16.0     # bound on the ray parameter t for the loss-scale estimate (depth gradient term)

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
            # the kernels hard-code ReLU hidden layers and a linear density output: anything else runs on the composed path
            dm, cm = geo.encoding_with_network.mlp.struct, tex.network.mlp.struct
            ok = ok and dm.activation == 1 and dm.out_activation == 0 and cm.activation == 1
            net_act = str(tex.network.network_config.get('output_activation', 'None')).lower()
            col_act = str(tex.config.get('color_activation', 'none')).lower()
            ok = ok and sorted([net_act, col_act]) == ['none', 'sigmoid']
        except AttributeError:
            ok = False
        return NerfFused(model) if ok else None

    def ref(self):
        return ctypes.byref(self.struct)

    def side_stream(self, dev):
        if self._side is None or self._side.device != dev:
            self._side = torch.cuda.Stream(device=dev)
        return self._side

    def ticket(self, dev):
        if self._ticket is None or self._ticket.device != dev:
            self._ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        return self._ticket

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
        """march + sigma_fn visibility pre-pass + compaction: the ``with torch.no_grad(): ray_marching(...)`` block of
        models/nerf.py:82-93, without a host sync.  Buffers have capacity n_rays * cap_per_ray; the true counts
        are offsets_m[n_rays] (marched) and offsets_k[n_rays] (kept) on the device."""
        m = self.model
        dev = rays.device
        n = rays.shape[0]
        cap = n * self.cap_per_ray
        mref = ctypes.byref(self.march)
        u = None
        if m.randomized:
            u = torch.rand(n, device=dev) if jitter is None else contig(jitter.to(dev), torch.float32)
        grid = m.occupancy_grid
        bits, coarse = grid.bits(), grid.coarse_bits()
        i32 = lambda k: torch.empty(k, dtype=torch.int32, device=dev)
        f32 = lambda k: torch.empty(k, dtype=torch.float32, device=dev)
        words = (self.cap_per_ray + 31) // 32
        masks, t_min = i32(n * words), f32(n)
        counts, offsets_m = i32(n), torch.empty(n + 1, dtype=torch.int64, device=dev)
        lib.call('nsr_march_rays_mask', mref, ptr(rays), ptr(u), ptr(bits), ptr(coarse), ptr(masks), words, ptr(t_min), ptr(counts), n, stream())
        lib.call('nsr_scan_counts', ptr(counts), ptr(offsets_m), n, stream())
        ri_m, ts_m, te_m = i32(cap), f32(cap), f32(cap)
        lib.call('nsr_march_rays_expand', mref, ptr(masks), words, ptr(t_min), ptr(offsets_m), ptr(ri_m), ptr(ts_m), ptr(te_m), n, stream())
        alphas = f32(cap)
        lib.call('nsr_nerf_prepass', self.ref(), ptr(rays), ptr(ri_m), ptr(ts_m), ptr(te_m), ptr(self.dparams_half()), ptr(alphas), cap,
                 ptr(offsets_m[n:]), stream())
        keep = torch.empty(cap, dtype=torch.uint8, device=dev)
        trans, kept = f32(cap), i32(n)
        lib.call('nsr_visibility', ptr(alphas), ptr(offsets_m), ptr(keep), ptr(trans), ptr(kept), self.early_stop_eps, self.alpha_thre, n, stream())
        offsets_k = torch.empty(n + 1, dtype=torch.int64, device=dev)
        lib.call('nsr_scan_counts', ptr(kept), ptr(offsets_k), n, stream())
        ri, ts, te, tr = i32(cap), f32(cap), f32(cap), f32(cap)
        lib.call('nsr_compact_prefix', ptr(offsets_m), ptr(offsets_k), ptr(ri_m), ptr(ts_m), ptr(te_m), ptr(trans), ptr(ri), ptr(ts), ptr(te),
                 ptr(tr), n, stream())
        return {'ri': ri, 'ts': ts, 'te': te, 'trans': tr, 'offsets_m': offsets_m, 'offsets_k': offsets_k, 'cap': cap}

    def render(self, rays, jitter=None, static=False):
        """NeRFModel.forward_ (models/nerf.py:61-127) -> the reference's output dict.

        static=False: exact-size per-sample outputs (one device->host read of the counts).
        static=True : no host sync -- per-sample tensors keep their capacity length (entries past
        ``num_samples`` are undefined); this is the form CUDA-graph capture uses."""
        m = self.model
        check_cuda(rays, what='NeRFModel')
        rays = contig(rays, torch.float32)
        if self.mode == 'two_pass':
            return self._render_two_pass(rays, jitter, static)
        self._want_grad = torch.is_grad_enabled()
        acc_rgb, opacity, depth, weights, ri, ts, te, pos, offsets_m, offsets_k, m_total, counts_m = _NerfRenderRays.apply(
            self.net.params, self.cnet.params, self, rays, jitter)
        n = rays.shape[0]
        counts = (m_total, offsets_k[n:])   # (marched, kept) on the device
        if static and self.lean_static_outputs:
            # graph capture with the fused loss: comp_rgb / rays_valid come out of nsr_nerf_loss_fwd, nothing else reads them
            out = {'opacity': opacity, 'depth': depth, 'num_samples_dev': counts[1]}
        else:
            comp_rgb = acc_rgb + m.background_color * (1.0 - opacity)
            out = {'comp_rgb': comp_rgb, 'opacity': opacity, 'depth': depth, 'rays_valid': opacity > 0,
                   'num_samples': counts[1].to(torch.int32)}
        if static:
            self.last_stats = {'counts_dev': counts}
            out['acc_rgb'] = acc_rgb  # pre-blend per-ray colour sum (input of nsr_b200.losses.nerf_rgb_loss)
            if m.training:
                # capacity-length buffers, first num_samples entries valid: packed t_starts / t_ends / ray_indices; `weights` is in the
                # loose layout (ray r's kept samples at offsets_loose[r] + j); packed row j lives at loose position loose_pos[j]
                # offsets_loose[r]: first row of ray r in the loose buffers (NOT monotonic in r when the marcher allocates, fused.march_alloc);
                # counts_loose[r]: its marched samples
                out.update({'weights': weights, 't_starts': ts, 't_ends': te, 'ray_indices': ri, 'loose_pos': pos,
                            'offsets_loose': offsets_m, 'counts_loose': counts_m, 'offsets_packed': offsets_k})
            return out
        n_marched, k = torch.cat(counts).tolist()
        self.last_stats = {'n_marched': n_marched, 'n_kept': k}
        if m.training:
            ts_, te_ = ts[:k], te[:k]
            wk = weights.index_select(0, pos[:k])  # packed view; keeps `weights` differentiable (distortion-loss consumers)
            out.update({'weights': wk.view(-1), 'points': ((ts_ + te_) / 2.).view(-1), 'intervals': (te_ - ts_).view(-1),
                        'ray_indices': ri[:k].long().view(-1)})
        return out

    def _render_two_pass(self, rays, jitter, static):
        m = self.model
        acc_rgb, opacity, depth, weights, ri, ts, te, counts = _NerfRender.apply(self.net.params, self.cnet.params, self, rays, jitter)
        comp_rgb = acc_rgb + m.background_color * (1.0 - opacity)
        out = {'comp_rgb': comp_rgb, 'opacity': opacity, 'depth': depth, 'rays_valid': opacity > 0,
               'num_samples': counts[1:].to(torch.int32)}
        if static:
            self.last_stats = {'counts_dev': counts}
            k = None
        else:
            n_marched, k = counts.tolist()
            self.last_stats = {'n_marched': n_marched, 'n_kept': k}
        if m.training and static:
            # capacity-length raw buffers (no per-sample torch ops over the capacity): entries past num_samples are undefined
            out.update({'weights': weights, 't_starts': ts, 't_ends': te, 'ray_indices': ri})
        elif m.training:
            w, ts_, te_, ri_ = weights[:k], ts[:k], te[:k], ri[:k]
            out.update({'weights': w.view(-1), 'points': ((ts_ + te_) / 2.).view(-1), 'intervals': (te_ - ts_).view(-1),
                        'ray_indices': ri_.long().view(-1)})
        return out
