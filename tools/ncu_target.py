"""Short driver for ncu captures: a few fused training steps of the bench workload (C2, 8192 rays)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device('cuda:0')
model = bench.build_model(dev)
from nsr_b200 import synthetic
rays = torch.from_numpy(synthetic.sample_rays(bench.N_RAYS, seed=0)).to(dev)
target = torch.rand(bench.N_RAYS, 3, device=dev)
params = [p for p in model.parameters()]
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for i in range(steps):
    model.background_color = torch.rand(3, device=dev)
    out = model(rays)
    loss = bench.masked_smooth_l1(out['comp_rgb'], target, out['rays_valid'])
    for p in params:
        p.grad = None
    loss.backward()
torch.cuda.synchronize()
print('done', model._fused.last_stats)
