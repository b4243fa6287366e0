"""End-to-end training on a synthetic scene with a known answer -- the system-level counterpart of the kernel parity tests.

There are no dataset files offline (the reference's only quantitative anchor is its README PSNR table on NeRF-Synthetic).  This tool
builds a small posed-image dataset by quadrature of an ANALYTIC radiance field (plain torch, independent of every nsr_b200 kernel), then
trains the drop-in model through the same pieces a `systems/nerf.py` / `systems/neus.py` training step uses:

    rays.training_batch (pixel -> ray)  ->  model(rays)  ->  loss  ->  backward  ->  FusedAdamW.step  ->  model.update_step (occupancy
    refresh every 16 steps)  ->  rays.RayBudget (dynamic ray count)

and reports held-out PSNR before / after, steps/s and rays/s as one JSON line.

    python tools/train_synthetic.py [--model nerf|neus] [--steps 2000] [--images 48] [--size 160] [--rays 4096]
    python tools/train_synthetic.py --dataset-only --device cpu     # build + describe the dataset (no CUDA needed)
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


# ---- analytic scene -----------------------------------------------------------------------------------------------------------
def scene(x):
    """x [.., 3] world -> (density [..], rgb [.., 3]): a soft ball of radius 0.6 with position-dependent colour on a grey slab"""
    r = x.norm(dim=-1)
    ball = torch.sigmoid((0.6 - r) / 0.02)
    slab = (torch.sigmoid((0.9 - x[..., 0].abs()) / 0.02) * torch.sigmoid((0.9 - x[..., 1].abs()) / 0.02)
            * torch.sigmoid((0.05 - (x[..., 2] + 0.75).abs()) / 0.01))
    density = 80.0 * torch.clamp(ball + slab, max=1.0)
    col_ball = 0.5 + 0.5 * torch.sin(4.0 * x + torch.tensor([0.0, 2.0, 4.0], device=x.device))
    col_slab = torch.full_like(col_ball, 0.55)
    w = (ball / (ball + slab + 1e-6))[..., None]
    return density, w * col_ball + (1 - w) * col_slab


@torch.no_grad()
def render_reference(rays, radius=1.5, n_samples=768, chunk=8192):
    """quadrature of the analytic field along each ray inside the [-radius, radius]^3 box, white background -> (rgb [n,3], opacity [n])"""
    out_rgb, out_op = [], []
    for s in range(0, rays.shape[0], chunk):
        o, d = rays[s:s + chunk, :3], rays[s:s + chunk, 3:]
        inv = 1.0 / torch.where(d.abs() < 1e-9, torch.full_like(d, 1e-9), d)
        t0, t1 = (-radius - o) * inv, (radius - o) * inv
        near = torch.minimum(t0, t1).amax(dim=-1).clamp(min=0.0)
        far = torch.maximum(t0, t1).amin(dim=-1)
        hit = far > near
        far = torch.where(hit, far, near)
        u = (torch.arange(n_samples, device=rays.device) + 0.5) / n_samples
        t = near[:, None] + (far - near)[:, None] * u[None]
        delta = ((far - near) / n_samples)[:, None]
        sigma, rgb = scene(o[:, None, :] + d[:, None, :] * t[..., None])
        alpha = 1.0 - torch.exp(-sigma * delta)
        trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha[:, :-1]], dim=1), dim=1)
        w = trans * alpha
        op = w.sum(dim=1)
        out_rgb.append((w[..., None] * rgb).sum(dim=1) + (1.0 - op)[:, None])
        out_op.append(op)
    return torch.cat(out_rgb), torch.cat(out_op)


def camera_rays(directions, c2w):
    """all pixels of one camera, plain torch (models/ray_utils.py:23-43 + F.normalize): the dataset builder stays kernel-free"""
    d = (directions[:, :, None, :] * c2w[None, None, :3, :3]).sum(-1).reshape(-1, 3)
    o = c2w[None, :3, 3].expand(d.shape)
    return torch.cat([o, F.normalize(d, dim=-1)], dim=-1)


def build_dataset(n_images, size, device, seed=0):
    from nsr_b200 import synthetic
    from nsr_b200.rays import get_ray_directions
    c2w = torch.from_numpy(synthetic.cameras(n=n_images + 4, seed=seed)).to(device)
    focal = synthetic.FOCAL * size / synthetic.IMG_W
    directions = get_ray_directions(size, size, focal, focal, size / 2, size / 2).to(device)
    images, masks = [], []
    for i in range(c2w.shape[0]):
        rgb, op = render_reference(camera_rays(directions, c2w[i]))
        images.append(rgb.reshape(size, size, 3))
        masks.append((op > 0.5).float().reshape(size, size))
    images, masks = torch.stack(images), torch.stack(masks)
    return {'directions': directions, 'c2w': c2w[:n_images], 'images': images[:n_images], 'masks': masks[:n_images],
            'c2w_test': c2w[n_images:], 'images_test': images[n_images:], 'size': size}


def psnr(a, b):
    return float(-10.0 * torch.log10(((a - b) ** 2).mean().clamp(min=1e-12)))


# ---- training -----------------------------------------------------------------------------------------------------------------
def evaluate(model, ds):
    from nsr_b200 import rays as nrays
    model.eval()
    vals = []
    with torch.no_grad():
        for i in range(ds['c2w_test'].shape[0]):
            batch = nrays.image_batch(ds['directions'], ds['c2w_test'], i)
            model.background_color = torch.ones(3, device=batch['rays'].device)
            out = model(batch['rays'])
            key = 'comp_rgb_full' if 'comp_rgb_full' in out else 'comp_rgb'
            vals.append(psnr(out[key].reshape(-1, 3).cpu(), ds['images_test'][i].reshape(-1, 3).cpu()))
    model.train()
    return float(np.mean(vals))


def train(args):
    from nsr_b200 import models, configs, rays as nrays
    from nsr_b200.optim import FusedAdamW
    from nsr_b200.lib import lib
    dev = torch.device(args.device)
    ds = build_dataset(args.images, args.size, dev)
    cfg = configs.nerf_blender() if args.model == 'nerf' else configs.neus_blender()
    torch.manual_seed(0)
    model = models.make(args.model, cfg).to(dev)
    model.train()
    opt = FusedAdamW.for_model(model, lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.1 ** (1.0 / args.steps))   # systems: lr decays to 10 % over the run
    budget = nrays.RayBudget(args.rays, 64, args.max_rays)           # target: rays * 64 kept samples per step
    gen = torch.Generator(device=dev).manual_seed(1)
    n_img, size = ds['c2w'].shape[0], ds['size']
    psnr0 = evaluate(model, ds)
    white = torch.ones(3, device=dev)
    launches0 = lib.launches
    torch.cuda.synchronize()
    t_start, n_rays_total, last_loss = time.time(), 0, float('nan')
    for step in range(args.steps):
        model.update_step(0, step)                                   # occupancy refresh (models/nerf.py:45-55, models/neus.py:79-111)
        n = budget.train_num_rays
        idx = torch.randint(0, n_img, (n,), device=dev, generator=gen)
        x = torch.randint(0, size, (n,), device=dev, generator=gen)
        y = torch.randint(0, size, (n,), device=dev, generator=gen)
        batch = nrays.training_batch(ds['directions'], ds['c2w'], ds['images'], ds['masks'], idx, x, y)
        model.background_color = white                               # the dataset is rendered on white
        out = model(batch['rays'])
        if args.model == 'nerf':                                      # systems/nerf.py:97: masked smooth-L1, without the boolean-mask sync
            v = out['rays_valid'].float()
            loss = (F.smooth_l1_loss(out['comp_rgb'], batch['rgb'], reduction='none') * v).sum() / (3.0 * v.sum().clamp(min=1.0))
            budget.observe(out['num_samples'])
        else:                                                         # systems/neus.py:98-113 (fused loss block)
            from nsr_b200.losses import neus_losses
            loss, _ = neus_losses(out, batch['rgb'], batch['fg_mask'], lambda_rgb_mse=10.0, lambda_eikonal=0.1, lambda_mask=0.1)
            budget.observe(out['num_samples_full'])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        sched.step()
        n_rays_total += n
        if step % 250 == 0 or step == args.steps - 1:
            last_loss = float(loss.detach())
            print(f'step {step:5d}  loss {last_loss:.5f}  rays/step {n}', file=sys.stderr)
    torch.cuda.synchronize()
    secs = time.time() - t_start
    res = {'tool': 'train_synthetic', 'model': args.model, 'steps': args.steps, 'images': n_img, 'image_size': size,
           'psnr_before': round(psnr0, 2), 'psnr_after': round(evaluate(model, ds), 2), 'final_loss': last_loss,
           'seconds': round(secs, 2), 'steps_per_s': round(args.steps / secs, 1), 'rays_per_s': round(n_rays_total / secs),
           'final_rays_per_step': budget.train_num_rays, 'our_kernel_launches': lib.launches - launches0,
           'occupied_fraction': float(model.occupancy_grid.binary.float().mean()) if hasattr(model.occupancy_grid, 'binary') else None}
    if args.export:
        mesh = model.export(configs_export())
        res['mesh'] = {'vertices': int(mesh['v_pos'].shape[0]), 'faces': int(mesh['t_pos_idx'].shape[0])}
    print(json.dumps(res))


def configs_export():
    from nsr_b200.config import Config
    return Config(dict(chunk_size=2097152, export_vertex_color=False))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--model', default='nerf', choices=['nerf', 'neus'])
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--images', type=int, default=48)
    ap.add_argument('--size', type=int, default=160)
    ap.add_argument('--rays', type=int, default=4096)
    ap.add_argument('--max-rays', type=int, default=8192)
    ap.add_argument('--device', default='cuda:0')
    ap.add_argument('--export', action='store_true', help='also extract the isosurface mesh at the end')
    ap.add_argument('--dataset-only', action='store_true')
    args = ap.parse_args()
    if args.dataset_only:
        ds = build_dataset(args.images, args.size, torch.device(args.device))
        print(json.dumps({'images': list(ds['images'].shape), 'mean_rgb': [round(float(v), 4) for v in ds['images'].mean(dim=(0, 1, 2))],
                          'object_fraction': round(float(ds['masks'].mean()), 4), 'test_images': int(ds['c2w_test'].shape[0])}))
        return
    train(args)


if __name__ == '__main__':
    main()
