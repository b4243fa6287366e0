"""Per-stage CUDA-event timings of the fused NeRF step at the bench workload (development aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nsr_b200 import models, configs, synthetic, ops
from nsr_b200.lib import lib, ptr, stream

D = torch.device('cuda:0')


def build(n_rays=8192, fused=True):
    cfg = configs.nerf_blender(); cfg['fused'] = fused
    torch.manual_seed(0)
    m = models.make('nerf', cfg).to(D)
    net = m.geometry.encoding_with_network
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        p = net.params.detach().cpu().clone()
        p[net.mlp.n_params:] = (torch.rand(net.grid.n_params, generator=g) * 2 - 1) * 0.1
        synthetic.shape_density(p, net.grid, net.mlp.n_params)
        net.params.copy_(p.to(D))
    m.occupancy_grid.set_binary(torch.from_numpy(synthetic.occupancy()))
    m.background_color = torch.tensor([0.3, 0.6, 0.9], device=D)
    m.train()
    rays = torch.from_numpy(synthetic.sample_rays(n_rays, seed=0)).to(D)
    return m, rays


def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    m, rays = build(n)
    f = m._fused
    target = torch.rand(n, 3, device=D)
    res = {}
    out0 = m.forward_(rays)
    res['M'], res['K'] = f.last_stats['n_marched'], f.last_stats['n_kept']
    res['trace_ms'] = timeit(lambda: f.trace(rays))

    def step():
        out = m.forward_(rays)
        loss = torch.nn.functional.smooth_l1_loss(out['comp_rgb'][out['rays_valid'][..., 0]], target[out['rays_valid'][..., 0]])
        for p in m.parameters(): p.grad = None
        loss.backward()
    res['step_ms'] = timeit(step)

    def fwd_only():
        with torch.no_grad(): m.forward_(rays)
    res['fwd_ms'] = timeit(fwd_only)
    # composed (per-op) path for comparison
    mc, _ = build(n, fused=False)

    def step_c():
        out = mc.forward_(rays)
        loss = torch.nn.functional.smooth_l1_loss(out['comp_rgb'][out['rays_valid'][..., 0]], target[out['rays_valid'][..., 0]])
        for p in mc.parameters(): p.grad = None
        loss.backward()
    res['composed_step_ms'] = timeit(step_c, iters=5, warm=2)
    res['rays_per_s_fused'] = n / res['step_ms'] * 1e3
    print(json.dumps(res))
