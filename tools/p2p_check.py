"""2+ GPU check of the NVLink peer-memory gradient exchange (csrc/p2p.cu) against NCCL, with timings.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/p2p_check.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
from nsr_b200.parallel import make_grad_sync, GradSync, P2PGradSync

n_big = 12602992
params = [torch.nn.Parameter(torch.zeros(n_big, device=dev)), torch.nn.Parameter(torch.zeros(7168, device=dev)), torch.nn.Parameter(torch.zeros(3, device=dev))]
out = {'rank': rank, 'world': world}
try:
    for variant in ('p2p', 'multimem'):
        sync, desc = make_grad_sync(params, world)
        out['desc'] = desc
        if not isinstance(sync, P2PGradSync):
            break
        if variant == 'p2p':
            sync.multicast = 0
        elif not sync.multicast_available:
            out['multimem'] = 'no multicast mapping'
            break
        else:
            sync.multicast = int(sync.hdl.multicast_ptr)
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        grads = [torch.randn(p.shape, device=dev, generator=g) for p in params]
        ref = [x.clone() for x in grads]
        for r in ref:
            dist.all_reduce(r)
            r.div_(world)
        for p, x in zip(params, grads):
            p.grad = x.clone()
        sync.all_reduce_mean()
        torch.cuda.synchronize()
        sync.check()
        err = max(float((p.grad - r).abs().max()) for p, r in zip(params, ref))
        out[variant + '_max_err'] = err
        # timing: copy-in + barrier + reduce + barrier, vs NCCL all-reduce + scale
        nccl = GradSync(params, world)
        for name, fn in ((variant, sync.all_reduce_mean), ('nccl', nccl.all_reduce_mean)):
            ts = []
            for i in range(12):
                for p, x in zip(params, grads):
                    p.grad = x.clone()
                dist.barrier(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            out[name + '_ms'] = round(sorted(ts[2:])[len(ts[2:]) // 2], 4)
        # exchange alone (gradients already in the symmetric buffer: p.grad IS the view => no copy-in)
        ts = []
        for i in range(12):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); sync.all_reduce_mean(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out[variant + '_no_copy_ms'] = round(sorted(ts[2:])[len(ts[2:]) // 2], 4)
        sync.check()
        # three ranges on three channels, two of them on side streams (what bind_pipelined does), against the NCCL result
        import ctypes
        from nsr_b200.lib import lib, ptr
        for p, x in zip(params, grads):
            sync.view_of(p).copy_(x)
        torch.cuda.synchronize(); dist.barrier()
        cuts = [0, (sync.n // 3) // (4 * world) * (4 * world), (2 * sync.n // 3) // 64 * 64, sync.n]
        mcp = ctypes.c_void_p(sync.multicast) if sync.multicast else None
        side = [torch.cuda.Stream(), torch.cuda.Stream()]
        cur = torch.cuda.current_stream()
        for ch in (2, 1, 0):
            st = cur if ch == 0 else side[ch - 1]
            st.wait_stream(cur)
            lib.call('nsr_p2p_exchange_mean_range', sync._peer, sync._fpeer, mcp, ptr(sync.epoch2[ch]), ptr(sync.err), sync.rank, sync.world,
                     cuts[ch], cuts[ch + 1] - cuts[ch], ch, 1, ctypes.c_void_p(st.cuda_stream))
        for sd in side:
            cur.wait_stream(sd)
        torch.cuda.synchronize()
        sync.check()
        out[variant + '_ranges_max_err'] = max(float((sync.view_of(p) - r).abs().max()) for p, r in zip(params, ref))
except Exception as e:
    out['error'] = f'{type(e).__name__}: {e}'[:300]
print(json.dumps(out), flush=True)
sys.stdout.flush()
os._exit(0)
