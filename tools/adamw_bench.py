"""Launch-shape sweep of the fused AdamW kernel (development aid): one subprocess per NSR_ADAMW_VARIANT, CUDA events, L2 flushed."""
import os, subprocess, sys, json
HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    sys.path.insert(0, HERE)
    import torch
    from nsr_b200.optim import FusedAdamW
    D = torch.device('cuda:0')
    n = 12610160
    p = torch.nn.Parameter(torch.randn(n, device=D) * 0.01)
    half = torch.empty(n, dtype=torch.float16, device=D)

    class M:  # minimal fp16-copy owner
        params = p
        _half_key = None
        def _params_half(self):
            return half
    opt = FusedAdamW([p], lr=1e-2, betas=(0.9, 0.99), eps=1e-15, half_shadows={p: M()})
    p.grad = torch.randn(n, device=D) * 1e-3
    flush = torch.empty(64 * 1024 * 1024, device=D)
    for _ in range(3):
        opt.step()
    evs = []
    for i in range(20):
        flush.fill_(float(i))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); opt.step(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) for x, y in evs)
    ms = sum(ts) / len(ts)
    print(json.dumps({'variant': os.environ.get('NSR_ADAMW_VARIANT', 'default'), 'ms_mean': round(ms, 4), 'ms_min': round(ts[0], 4),
                      'GBps_mean': round(30.0 * n / ms / 1e6, 1)}))
else:
    for v in ['1,8', '1,0', '2,0', '2,8', '2,16', '4,0', '4,4', '4,8']:
        r = subprocess.run([sys.executable, __file__, 'child'], env=dict(os.environ, NSR_ADAMW_VARIANT=v), capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:])
