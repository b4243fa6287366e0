"""Kernel-level breakdown of one C3 (neus-blender) training step with torch.profiler (CUPTI): which CUDA kernels the step
launches, their device time, and how much of the wall clock is host launch overhead.  Development aid."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from torch.profiler import profile, ProfilerActivity

sys.argv = sys.argv[:1]
import importlib.util
spec = importlib.util.spec_from_file_location('neus_times_mod', os.path.join(os.path.dirname(os.path.abspath(__file__)), 'neus_times.py'))
src = open(spec.origin).read().split('\nres = {}')[0]
ns = {'__file__': spec.origin}
exec(compile(src, spec.origin, 'exec'), ns)
build, step, configs, D = ns['build'], ns['step'], ns['configs'], ns['D']

which = os.environ.get('NEUS_CFG', 'C3')
fn, n = (configs.neus_blender, 8192) if which == 'C3' else (configs.neus_dtu, 4096)
m, rays = build(fn, n)
target, mask = torch.rand(n, 3, device=D), (torch.rand(n, device=D) > 0.5).float()
for _ in range(3):
    step(m, rays, target, mask)
torch.cuda.synchronize()
STEPS = 5
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        step(m, rays, target, mask)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA or getattr(e, 'self_device_time_total', 0) > 0]
rows = sorted(((e.key, e.count / STEPS, e.self_device_time_total / STEPS) for e in prof.key_averages() if e.self_device_time_total > 0),
              key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f'# {which}: device time per step {tot/1e3:.3f} ms in {sum(r[1] for r in rows):.0f} kernels/memops')
for k, c, t in rows[:45]:
    print(f'{t:9.1f} us  x{c:5.1f}  {k[:110]}')
