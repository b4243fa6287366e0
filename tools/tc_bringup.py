"""Bring-up of the tcgen05 MLP kernel: try both LBO/SBO descriptor interpretations, report which matches mma.sync, time both paths."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nsr_b200 import tcnn
from nsr_b200.lib import lib, ptr, stream

D = torch.device('cuda:0')
res = {}
for (n_in, n_out, nh) in [(32, 16, 1), (32, 3, 2)]:
    net = tcnn.Network(n_in, n_out, dict(otype='FullyFusedMLP', activation='ReLU', output_activation='None', n_neurons=64, n_hidden_layers=nh)).to(D)
    x = torch.randn(1 << 20, n_in, device=D).half().contiguous()
    ph = net._params_half()
    ref = torch.empty(x.shape[0], 16, dtype=torch.float16, device=D)
    lib.call('nsr_mlp_fwd', net.mlp.ref(), ptr(x), ptr(ph), ptr(ref), x.shape[0], stream())
    torch.cuda.synchronize()
    for variant in (0, 1):
        out = torch.zeros_like(ref)
        status = torch.zeros(1, dtype=torch.int32, device=D)
        try:
            lib.call('nsr_mlp_fwd_tc', net.mlp.ref(), ptr(x), ptr(ph), ptr(out), x.shape[0], variant, ptr(status), stream())
            torch.cuda.synchronize()
            err = (out[:, :n_out].float() - ref[:, :n_out].float()).abs().max().item()
            res[f'{n_in}-{nh}h-v{variant}'] = {'max_err': err, 'status': int(status.item())}
        except Exception as e:  # noqa
            res[f'{n_in}-{nh}h-v{variant}'] = {'error': str(e)[:200]}
            print(json.dumps(res)); sys.exit(0)

    def timeit(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20 * 1e3
    good = min((0, 1), key=lambda v: res[f'{n_in}-{nh}h-v{v}'].get('max_err', 1e9))
    res[f'{n_in}-{nh}h-us_mma_sync'] = timeit(lambda: lib.call('nsr_mlp_fwd', net.mlp.ref(), ptr(x), ptr(ph), ptr(ref), x.shape[0], stream()))
    res[f'{n_in}-{nh}h-us_tcgen05'] = timeit(lambda: lib.call('nsr_mlp_fwd_tc', net.mlp.ref(), ptr(x), ptr(ph), ptr(ref), x.shape[0], good, None, stream()))
print(json.dumps(res))
