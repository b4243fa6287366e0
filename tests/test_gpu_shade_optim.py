"""GPU parity of the NeuS shading kernels (alpha + normal, compositing, fused VolumeRadiance) and the fused AdamW, through the C ABI.

Checkers: oracle.neus.get_alpha / oracle.render (CPU, fp64 autograd for the gradients), the composed per-op path for the radiance
MLP (same fp16 tensor-core arithmetic => tight tolerance) and torch.optim.AdamW on the CPU for the optimizer (it IS the reference's
optimizer, systems/utils.py:314-325).  Tolerances: alpha/normal/composited sums 1e-5 abs (fp32 vs fp64), their gradients 1e-4 of the
largest entry; radiance rgb 2e-3 vs the composed fp16 path, gradients cosine >= 0.999; AdamW parameters 2e-6 relative after 6 steps."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import neus as oneus, render as orender, optim as ooptim

D = torch.device('cuda:0')


def cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def _segments(n_rays, seed, max_len=70):
    rng = np.random.default_rng(seed)
    counts = rng.integers(0, max_len, n_rays)
    counts[rng.random(n_rays) < 0.15] = 0            # empty rays
    counts[0] = 0
    counts[-1] = 97                                  # > 3 chunks of 32
    offsets = np.zeros(n_rays + 1, np.int64)
    offsets[1:] = np.cumsum(counts)
    ray_indices = np.repeat(np.arange(n_rays), counts)
    return torch.from_numpy(offsets), torch.from_numpy(ray_indices)


@pytest.mark.parametrize('anneal', [0.0, 0.25, 1.0])
def test_neus_alpha_matches_oracle_forward_and_backward(anneal):
    from nsr_b200 import ops
    g = torch.Generator().manual_seed(3)
    k = 5000
    sdf = (torch.rand(k, generator=g) - 0.5) * 0.2
    grad = torch.randn(k, 3, generator=g) * torch.rand(k, 1, generator=g) * 2
    grad[:5] = 0.0                                   # the normalisation clamp
    dirs = F.normalize(torch.randn(k, 3, generator=g), dim=-1)
    dists = torch.rand(k, generator=g) * 0.01 + 1e-3
    var = torch.tensor(0.3)
    ga, gn = torch.randn(k, generator=g), torch.randn(k, 3, generator=g)

    # oracle in fp64
    s64, g64, v64 = sdf.double().requires_grad_(), grad.double().requires_grad_(), var.double().requires_grad_()
    inv_s = oneus.inv_s_from_variance(v64)
    n64 = F.normalize(g64, p=2, dim=-1)
    a64 = oneus.get_alpha(s64, n64, dirs.double(), dists.double(), inv_s, anneal)
    ((a64 * ga.double()).sum() + (n64 * gn.double()).sum()).backward()

    sd, gd = sdf.to(D).requires_grad_(), grad.to(D).requires_grad_()
    vd = var.to(D).requires_grad_()
    inv_s_d = torch.exp(vd * 10.0).clip(1e-6, 1e6).reshape(1)
    alpha, normal = ops.neus_alpha(sd, gd, inv_s_d, dirs.to(D), dists.to(D), anneal)
    ((alpha * ga.to(D)).sum() + (normal * gn.to(D)).sum()).backward()
    assert float((alpha.cpu().double() - a64.detach()).abs().max()) < 2e-5
    assert float((normal.cpu().double()[5:] - n64.detach()[5:]).abs().max()) < 1e-5
    for got, want, name in ((sd.grad, s64.grad, 'sdf'), (gd.grad[5:], g64.grad[5:], 'sdf_grad'), (vd.grad, v64.grad, 'variance')):
        err = float((got.cpu().double() - want).abs().max()) / float(want.abs().max())
        assert err < 2e-4, (name, err)


def test_neus_composite_matches_oracle_forward_and_backward():
    from nsr_b200 import ops
    n_rays = 257
    offsets, ray_indices = _segments(n_rays, 1)
    k = int(offsets[-1])
    g = torch.Generator().manual_seed(9)
    alpha = torch.rand(k, generator=g) * 0.3
    alpha[::11] = 0.0
    alpha[5::97] = 1.0 - 1e-7
    rgb, normal = torch.rand(k, 3, generator=g), F.normalize(torch.randn(k, 3, generator=g), dim=-1)
    ts = torch.rand(k, generator=g) * 3
    te = ts + 0.005
    gw, gop, gd, grgb, gnrm = (torch.randn(k, generator=g), torch.randn(n_rays, 1, generator=g), torch.randn(n_rays, 1, generator=g),
                               torch.randn(n_rays, 3, generator=g), torch.randn(n_rays, 3, generator=g))

    a64, c64, n64 = alpha.double().requires_grad_(), rgb.double().requires_grad_(), normal.double().requires_grad_()
    w = orender.render_weight_from_alpha(a64[:, None], ray_indices, n_rays)
    mid = ((ts + te) / 2).double()[:, None]
    op = orender.accumulate_along_rays(w, ray_indices, None, n_rays)
    dp = orender.accumulate_along_rays(w, ray_indices, mid, n_rays)
    cr = orender.accumulate_along_rays(w, ray_indices, c64, n_rays)
    cn = orender.accumulate_along_rays(w, ray_indices, n64, n_rays)
    ((w.view(-1) * gw.double()).sum() + (op * gop.double()).sum() + (dp * gd.double()).sum() + (cr * grgb.double()).sum()
     + (cn * gnrm.double()).sum()).backward()

    ad, cd, nd = alpha.to(D).requires_grad_(), rgb.to(D).requires_grad_(), normal.to(D).requires_grad_()
    W, OP, DP, CR, CN = ops.neus_composite(ad, cd, nd, ts.to(D), te.to(D), offsets.to(D))
    ((W * gw.to(D)).sum() + (OP * gop.to(D)).sum() + (DP * gd.to(D)).sum() + (CR * grgb.to(D)).sum() + (CN * gnrm.to(D)).sum()).backward()
    for got, want, name in ((W, w.view(-1), 'weights'), (OP, op, 'opacity'), (DP, dp, 'depth'), (CR, cr, 'rgb'), (CN, cn, 'normal')):
        assert float((got.cpu().double() - want.detach()).abs().max()) < 2e-5, name
    # d alpha divides by (1 - alpha): exclude the saturated samples from the relative check (kernel clamps the divisor at 1e-10)
    ok = alpha < 0.999
    for got, want, name in ((ad.grad[ok.to(D)], a64.grad[ok], 'alpha'), (cd.grad, c64.grad, 'rgb'), (nd.grad, n64.grad, 'normal')):
        err = float((got.cpu().double() - want).abs().max()) / float(want.abs().max())
        assert err < 2e-4, (name, err)


@pytest.mark.parametrize('n_feat,n_extra,oact,color_act', [(13, 3, 'none', 'sigmoid'), (16, 0, 'Sigmoid', None), (16, 0, 'none', None)])
def test_fused_radiance_matches_composed_path(n_feat, n_extra, oact, color_act):
    """same module, config key fused=False -> per-op path (SH kernel, cat, fp16 MLP kernel, casts)"""
    from nsr_b200 import models
    cfg = dict(name='volume-radiance', input_feature_dim=n_feat + n_extra, dir_encoding_config=dict(otype='SphericalHarmonics', degree=4),
               mlp_network_config=dict(otype='FullyFusedMLP', activation='ReLU', output_activation=oact, n_neurons=64, n_hidden_layers=2))
    if color_act:
        cfg['color_activation'] = color_act
    tex = models.make('volume-radiance', dict(cfg)).to(D)
    tex_ref = models.make('volume-radiance', dict(cfg, fused=False)).to(D)
    tex_ref.load_state_dict(tex.state_dict())
    g = torch.Generator().manual_seed(11)
    k = 3001                                         # ragged last tile
    feat = torch.randn(k, n_feat, generator=g).to(D)
    dirs = F.normalize(torch.randn(k, 3, generator=g), dim=-1).to(D)
    extra = [F.normalize(torch.randn(k, 3, generator=g), dim=-1).to(D)] if n_extra else []
    go = (torch.randn(k, 3, generator=g) * 1e-4).to(D)   # realistic magnitude: exercises the automatic dgrad scale
    outs = []
    for m in (tex, tex_ref):
        f = feat.clone().requires_grad_()
        e = [x.clone().requires_grad_() for x in extra]
        rgb = m(f, dirs, *e)
        (rgb * go).sum().backward()
        outs.append((rgb.detach(), f.grad, e[0].grad if e else None, m.network.params.grad.clone()))
    assert tex._rspec is not None and tex_ref._rspec is None
    (rgb, df, de, dp), (rgb_r, df_r, de_r, dp_r) = outs
    assert rgb.dtype == torch.float32 and float((rgb - rgb_r).abs().max()) < 2e-3
    assert cos(df, df_r) > 0.999 and float((df - df_r).abs().max()) < 3e-2 * float(df_r.abs().max())
    if de is not None:
        assert cos(de, de_r) > 0.999
    assert cos(dp, dp_r) > 0.999 and float((dp - dp_r).abs().max()) < 3e-2 * float(dp_r.abs().max())
    assert tex(feat[:0], dirs[:0], *[x[:0] for x in extra]).shape == (0, 3)


def test_fused_adamw_matches_torch_adamw_and_refreshes_fp16_copy():
    from nsr_b200 import tcnn
    from nsr_b200.optim import FusedAdamW
    enc = tcnn.Encoding(3, dict(otype='HashGrid', n_levels=4, n_features_per_level=2, log2_hashmap_size=12, base_resolution=4,
                                per_level_scale=1.5)).to(D)
    n = enc.params.numel()
    extra = torch.nn.Parameter(torch.randn(1027, device=D))   # not a multiple of 4: scalar tail
    holder = torch.nn.Module()
    holder.enc, holder.extra = enc, extra
    opt = FusedAdamW.for_model(holder, lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    ref_p = [torch.nn.Parameter(enc.params.detach().cpu().clone()), torch.nn.Parameter(extra.detach().cpu().clone())]
    ref = torch.optim.AdamW(ref_p, lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    half0 = enc._params_half()
    g = torch.Generator().manual_seed(2)
    for step in range(6):
        for p, r in zip((enc.params, extra), ref_p):
            gr = torch.randn(p.numel(), generator=g) * 10.0 ** (-step)
            gr[::5] = 0
            p.grad, r.grad = gr.to(D), gr.clone()
        opt.step()
        ref.step()
    for p, r in zip((enc.params, extra), ref_p):
        np.testing.assert_allclose(p.detach().cpu().numpy(), r.detach().numpy(), rtol=2e-6, atol=2e-7 * float(r.detach().abs().max()))
        np.testing.assert_allclose(opt.state[p]['exp_avg_sq'].cpu().numpy(), ref.state[r]['exp_avg_sq'].numpy(), rtol=2e-6, atol=1e-12)
    # the fp16 copy the kernels read was refreshed in place by the same kernel, and the module's cache key follows the new version
    half = enc._params_half()
    assert half.data_ptr() == half0.data_ptr()
    assert torch.equal(half, enc.params.detach().half())
    # oracle restatement agrees too (one more step from the current state)
    p_np = enc.params.detach().cpu().numpy().copy()
    m_np, v_np = opt.state[enc.params]['exp_avg'].cpu().numpy().copy(), opt.state[enc.params]['exp_avg_sq'].cpu().numpy().copy()
    gr = torch.randn(n, generator=g)
    enc.params.grad = gr.to(D)
    extra.grad = None
    opt.step()
    ooptim.adamw_step(p_np, gr.numpy().copy(), m_np, v_np, 7)
    np.testing.assert_allclose(enc.params.detach().cpu().numpy(), p_np, rtol=2e-6, atol=2e-7 * float(np.abs(p_np).max()))


def test_fused_adamw_skip_on_found_inf_and_capturable_mode():
    from nsr_b200.optim import FusedAdamW
    from nsr_b200.lib import lib, ptr, stream
    p = torch.nn.Parameter(torch.randn(4096, device=D))
    p0 = p.detach().clone()
    opt = FusedAdamW([p], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    p.grad = torch.randn(4096, device=D)
    p.grad[17] = float('inf')
    found = torch.zeros(1, device=D)
    lib.call('nsr_grad_nonfinite', ptr(p.grad), ptr(found), p.numel(), stream())
    assert float(found) == 1.0
    opt.found_inf = found
    opt.step()
    assert torch.equal(p.detach(), p0)               # step skipped, parameters untouched
    # capturable: lr and step on the device, same numbers as the host-side mode
    q = torch.nn.Parameter(p0.clone())
    r = torch.nn.Parameter(p0.clone())
    oc = FusedAdamW([q], lr=1e-2, betas=(0.9, 0.99), eps=1e-15, capturable=True)
    oh = FusedAdamW([r], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    g = torch.Generator().manual_seed(4)
    for _ in range(4):
        gr = torch.randn(4096, generator=g).to(D)
        q.grad, r.grad = gr, gr.clone()
        oc.step()
        oh.step()
    np.testing.assert_allclose(q.detach().cpu().numpy(), r.detach().cpu().numpy(), rtol=5e-6, atol=1e-8)


@pytest.mark.parametrize('ctype_name,R', [('AABB', 32), ('UN_BOUNDED_SPHERE', 24)])
def test_occupancy_refresh_kernels_match_oracle(ctype_name, R):
    """csrc/occgrid.cu (points -> EMA-max update -> threshold + packed bitfield + coarse field) vs oracle/occgrid.py on the same
    cells and jitter: occupancy values within 1e-5 relative (fma contraction of the point transform), binary equal away from
    threshold ties, bit packing exact."""
    from oracle import occgrid as oocc, contraction as ocon
    from nsr_b200 import nerfacc
    ctype = getattr(nerfacc.ContractionType, ctype_name)
    radius = 1.5
    aabb = torch.tensor([-radius] * 3 + [radius] * 3)
    grid = nerfacc.OccupancyGrid(aabb, R, ctype).to(D)
    grid.train()
    C = R ** 3
    g = torch.Generator().manual_seed(7)
    occs0 = torch.rand(C, generator=g) * 0.02
    grid.occs.copy_(occs0)
    fn = lambda x: (torch.exp(-4.0 * x.norm(dim=-1, keepdim=True)) * 0.05).float()
    # sparse update with unique cells (exact comparison) ...
    cells = torch.randperm(C, generator=g)[:C // 4]
    jitter = torch.rand(cells.shape[0], 3, generator=g)
    grid._update_cuda(512, fn, 0.01, 0.95, 256, cells=cells.to(D), jitter=jitter.to(D))
    want_occs, want_bin = oocc.update(occs0.clone(), cells, jitter, fn, radius, ctype.value, R, ema_decay=0.95, occ_thre=0.01)
    got = grid.occs.cpu()
    assert float((got - want_occs).abs().max()) <= 1e-5 * float(want_occs.abs().max())
    thr = min(float(want_occs.mean()), 0.01)
    tie = (want_occs - thr).abs() < 1e-4 * thr
    assert bool(((grid.binary.cpu().view(-1) == want_bin.view(-1)) | tie).all()) and int(tie.sum()) < 10
    # ... the packed bitfield and the coarse "any bit in 4^3" field come out of the same kernel
    bits = grid.bits().cpu().numpy().view(np.uint32)
    assert np.array_equal(bits, oocc.pack_bits(grid.binary.cpu().numpy()))
    if R % 4 == 0:
        b = grid.binary.cpu().view(R // 4, 4, R // 4, 4, R // 4, 4).any(dim=5).any(dim=3).any(dim=1)
        assert np.array_equal(grid.coarse_bits().cpu().numpy().view(np.uint32), oocc.pack_bits(b.numpy()))
    # duplicates: the maximum of the duplicated samples wins, deterministically (AABB: every sample is valid)
    if ctype_name == 'AABB':
        grid.occs.copy_(occs0)
        dup = torch.cat([cells[:100], cells[:100]])
        j2 = torch.rand(200, 3, generator=g)
        grid._update_cuda(512, fn, 0.01, 0.95, 256, cells=dup.to(D), jitter=j2.to(D))
        a, _ = oocc.update(occs0.clone(), dup[:100], j2[:100], fn, radius, ctype.value, R)
        b, _ = oocc.update(occs0.clone(), dup[100:], j2[100:], fn, radius, ctype.value, R)
        assert float((grid.occs.cpu() - torch.maximum(a, b)).abs().max()) <= 1e-5 * float(a.abs().max())
    # dense (warm-up) update through the public entry point: every cell once
    grid.occs.zero_()
    torch.manual_seed(3)
    grid.every_n_step(step=0, occ_eval_fn=fn, occ_thre=0.01)
    inside = grid.occs > 0
    assert inside.any() and (ctype_name == 'AABB') == bool(inside.all())   # sphere: cells outside the unit ball are never touched


def test_fused_neus_losses_match_reference_formulas():
    """systems/neus.py:98-121 restated with torch ops (boolean-mask means, criterions.binary_cross_entropy) vs nsr_neus_loss_fwd/bwd"""
    from nsr_b200.losses import neus_losses
    g = torch.Generator().manual_seed(21)
    n, k = 1500, 20000
    comp = torch.rand(n, 3, generator=g)
    valid = torch.rand(n, 1, generator=g) > 0.3
    opacity = torch.rand(n, 1, generator=g)
    opacity[:10] = 0.0
    opacity[10:20] = 1.0                                  # outside the clamp: zero gradient
    rgb, mask = torch.rand(n, 3, generator=g), (torch.rand(n, generator=g) > 0.5).float()
    sg, s = torch.randn(k, 3, generator=g) * 1.3, torch.randn(k, generator=g) * 0.2
    lam = dict(lambda_rgb_mse=10.0, lambda_rgb_l1=0.7, lambda_eikonal=0.1, lambda_mask=0.1, lambda_opaque=0.05, lambda_sparsity=0.02,
               sparsity_scale=3.0)

    def bce(i, t):
        return -(t * torch.log(i) + (1 - t) * torch.log(1 - i)).mean()

    c64, o64, sg64, s64 = comp.double().requires_grad_(), opacity.double().requires_grad_(), sg.double().requires_grad_(), s.double().requires_grad_()
    v = valid[:, 0]
    parts = [F.mse_loss(c64[v], rgb.double()[v]), F.l1_loss(c64[v], rgb.double()[v]), ((torch.linalg.norm(sg64, ord=2, dim=-1) - 1.) ** 2).mean()]
    oc = torch.clamp(o64.squeeze(-1), 1e-3, 1 - 1e-3)
    parts += [bce(oc, mask.double()), bce(oc, oc), torch.exp(-3.0 * s64.abs()).mean()]
    total = (10.0 * parts[0] + 0.7 * parts[1] + 0.1 * parts[2] + 0.1 * parts[3] + 0.05 * parts[4] + 0.02 * parts[5])
    (total * 1.7).backward()

    cd, od, sgd, sd = comp.to(D).requires_grad_(), opacity.to(D).requires_grad_(), sg.to(D).requires_grad_(), s.to(D).requires_grad_()
    out = {'comp_rgb_full': cd, 'rays_valid_full': valid.to(D), 'opacity': od, 'sdf_grad_samples': sgd, 'sdf_samples': sd}
    tot, pp = neus_losses(out, rgb.to(D), mask.to(D), **lam)
    (tot * 1.7).backward()
    assert abs(float(tot) - float(total)) < 2e-5 * abs(float(total))
    for a, b in zip(pp.tolist(), parts):
        assert abs(a - float(b)) < 2e-5 * abs(float(b)) + 1e-7
    for got, want, name in ((cd.grad, c64.grad, 'comp_rgb'), (od.grad, o64.grad, 'opacity'), (sgd.grad, sg64.grad, 'sdf_grad'), (sd.grad, s64.grad, 'sdf')):
        err = float((got.cpu().double() - want).abs().max()) / float(want.abs().max())
        assert err < 1e-4, (name, err)
    # no mask in the dataset -> no mask term, no NaNs; no samples at all -> finite
    tot2, _ = neus_losses(out, rgb.to(D), None, **lam)
    assert abs(float(tot2) - float(total - 0.1 * parts[3])) < 2e-5 * abs(float(total))
