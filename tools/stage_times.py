"""Per-kernel CUDA-event timings of the fused NeRF step at the bench workload (development aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nsr_b200 import synthetic
from nsr_b200.lib import lib
from nsr_b200.graph import GraphedStep

dev = torch.device('cuda:0')
n = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_RAYS
model = bench.build_model(dev)
if len(sys.argv) > 2:
    model._fused.mode = 'two_pass' if sys.argv[2] == 'two_pass' else 'per_ray'
    model._fused.bwd_kernel = 'rays' if sys.argv[2] == 'per_ray_bwd' else 'tiles'
rays = [torch.from_numpy(synthetic.sample_rays(n, seed=i)).to(dev) for i in range(4)]
target = torch.rand(n, 3, device=dev)
params = [p for p in model.parameters() if p.numel() > 0]
flush = torch.empty(64 * 1024 * 1024, device=dev)


from nsr_b200.losses import nerf_rgb_loss
FUSED_LOSS = os.environ.get('NSR_TORCH_LOSS', '0') != '1'


def loss_fn(out, batch):
    if FUSED_LOSS and 'acc_rgb' in out:
        return nerf_rgb_loss(out['acc_rgb'], out['opacity'], model.background_color, batch['rgb'])[0]
    return bench.masked_smooth_l1(out['comp_rgb'], batch['rgb'], out['rays_valid'])


gs = GraphedStep(model, loss_fn, n, batch_spec={'rgb': (3,)}, device=dev)
bg = torch.rand(3, device=dev)
for i in range(5):
    gs(rays[i % 4], rgb=target, background_color=bg)
torch.cuda.synchronize()
evs = []
for i in range(50):
    flush.fill_(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gs(rays[i % 4], rgb=target, background_color=bg); e1.record()
    evs.append((e0, e1))
torch.cuda.synchronize()
res = {'graph_step_ms': sum(a.elapsed_time(b) for a, b in evs) / len(evs), 'counts': gs.counts()}


def eager():
    model.background_color = bg
    out = model(rays[0])
    loss = loss_fn(out, {'rgb': target})
    for p in params:
        p.grad = None
    loss.backward()


for _ in range(3):
    eager()
lib.profile = {}
for _ in range(20):
    flush.fill_(1.0)
    eager()
torch.cuda.synchronize()
res['kernels_ms'] = {k: round(sum(a.elapsed_time(b) for a, b in v) / len(v), 5) for k, v in lib.profile.items()}
res['kernels_sum_ms'] = round(sum(sum(a.elapsed_time(b) for a, b in v) for v in lib.profile.values()) / 20, 5)
lib.profile = None
print(json.dumps(res))
