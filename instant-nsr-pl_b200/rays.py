"""Pixel -> ray front end of a training / evaluation step, the drop-in for ``preprocess_data`` of the reference's
systems (systems/nerf.py:33-91, systems/neus.py:34-96) and for models/ray_utils.py.

``get_ray_directions`` / ``get_rays`` keep the reference's signatures (dataset classes call them once at load time: plain torch);
``training_batch`` / ``image_batch`` are the per-step part and run as ONE kernel (csrc/rays.cu) instead of ~12 torch kernels."""
import numpy as np
import torch

from .lib import lib, ptr, stream, check_cuda, contig


def get_ray_directions(W, H, fx, fy, cx, cy, use_pixel_centers=True):
    """models/ray_utils.py:9-20 -> [H, W, 3] camera-space directions"""
    c = 0.5 if use_pixel_centers else 0.0
    i, j = np.meshgrid(np.arange(W, dtype=np.float32) + c, np.arange(H, dtype=np.float32) + c, indexing='xy')
    i, j = torch.from_numpy(i), torch.from_numpy(j)
    return torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)


def get_rays(directions, c2w, keepdim=False):
    """models/ray_utils.py:23-43 (load-time helper; the per-step path is training_batch)"""
    assert directions.shape[-1] == 3
    if directions.ndim == 2:
        assert c2w.ndim == 3
        rays_d = (directions[:, None, :] * c2w[:, :3, :3]).sum(-1)
        rays_o = c2w[:, :, 3].expand(rays_d.shape)
    elif c2w.ndim == 2:
        rays_d = (directions[:, :, None, :] * c2w[None, None, :3, :3]).sum(-1)
        rays_o = c2w[None, None, :, 3].expand(rays_d.shape)
    else:
        rays_d = (directions[None, :, :, None, :] * c2w[:, None, None, :3, :3]).sum(-1)
        rays_o = c2w[:, None, None, :, 3].expand(rays_d.shape)
    if not keepdim:
        rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    return rays_o, rays_d


def _dataset_dims(directions, all_c2w, all_images):
    if directions.dim() not in (3, 4) or directions.shape[-1] != 3:
        raise ValueError(f'directions must be [H,W,3] or [N,H,W,3], got {tuple(directions.shape)}')
    H, W = directions.shape[-3], directions.shape[-2]
    if all_c2w.dim() != 3 or all_c2w.shape[1] not in (3, 4) or all_c2w.shape[2] != 4:
        raise ValueError(f'all_c2w must be [N,3,4] or [N,4,4], got {tuple(all_c2w.shape)}')
    n_images = all_c2w.shape[0]
    if directions.dim() == 4 and directions.shape[0] != n_images:
        raise ValueError('per-image directions and all_c2w disagree on the number of images')
    if all_images is not None and tuple(all_images.shape[:3]) != (n_images, H, W):
        raise ValueError(f'all_images must be [{n_images},{H},{W},C], got {tuple(all_images.shape)}')
    return H, W, n_images


def training_batch(directions, all_c2w, all_images, all_fg_masks, index, x, y, background_color=None, apply_mask=False):
    """One kernel for systems/nerf.py:42-56,66,78-85: -> dict(rays [n,6] = origin | unit direction, rgb [n,3], fg_mask [n]).
    index / x / y: int64 [n] drawn by the caller (torch.randint, as the reference does); directions [H,W,3] or [N,H,W,3]; all_c2w
    [N,3|4,4]; all_images [N,H,W,C>=3] fp32; all_fg_masks [N,H,W] fp32; apply_mask blends rgb with background_color [3]."""
    check_cuda(directions, all_c2w, all_images, all_fg_masks, index, x, y, what='rays.training_batch')
    H, W, n_images = _dataset_dims(directions, all_c2w, all_images)
    d, c, im, mk = contig(directions, torch.float32), contig(all_c2w, torch.float32), contig(all_images, torch.float32), \
        contig(all_fg_masks, torch.float32)
    idx, xx, yy = contig(index, torch.int64), contig(x, torch.int64), contig(y, torch.int64)
    n, dev = idx.shape[0], d.device
    if xx.shape[0] != n or yy.shape[0] != n:
        raise ValueError('index, x and y must have the same length')
    if apply_mask and background_color is None:
        raise ValueError('apply_mask needs background_color')
    bg = None if background_color is None else contig(background_color.to(dev), torch.float32)
    rays, rgb, fg = torch.empty(n, 6, device=dev), torch.empty(n, 3, device=dev), torch.empty(n, device=dev)
    if n == 0:
        return {'rays': rays, 'rgb': rgb, 'fg_mask': fg}
    lib.call('nsr_gather_rays', ptr(d), int(d.dim() == 4), ptr(c), c.shape[1], ptr(im), im.shape[-1], ptr(mk), ptr(idx), ptr(xx), ptr(yy), 0,
             ptr(bg), int(bool(apply_mask)), H, W, n_images, ptr(rays), ptr(rgb), ptr(fg), n, stream())
    return {'rays': rays, 'rgb': rgb, 'fg_mask': fg}


def image_batch(directions, all_c2w, index, all_images=None, all_fg_masks=None, background_color=None, apply_mask=False):
    """validation / test (systems/nerf.py:57-64,66): every pixel of image ``index`` in row-major order -> dict(rays [H*W,6]
    (+ rgb, fg_mask when the images are given))."""
    check_cuda(directions, all_c2w, all_images, all_fg_masks, what='rays.image_batch')
    H, W, n_images = _dataset_dims(directions, all_c2w, all_images)
    index = int(index)
    if not 0 <= index < n_images:
        raise IndexError(f'image {index} out of range [0, {n_images})')
    d, c = contig(directions, torch.float32), contig(all_c2w, torch.float32)
    im = None if all_images is None else contig(all_images, torch.float32)
    mk = None if all_fg_masks is None else contig(all_fg_masks, torch.float32)
    n, dev = H * W, d.device
    bg = None if background_color is None else contig(background_color.to(dev), torch.float32)
    rays = torch.empty(n, 6, device=dev)
    rgb = torch.empty(n, 3, device=dev) if im is not None else None
    fg = torch.empty(n, device=dev) if mk is not None else None
    lib.call('nsr_gather_rays', ptr(d), int(d.dim() == 4), ptr(c), c.shape[1], ptr(im), 0 if im is None else im.shape[-1], ptr(mk), None, None,
             None, index, ptr(bg), int(bool(apply_mask) and im is not None), H, W, n_images, ptr(rays), ptr(rgb), ptr(fg), n, stream())
    out = {'rays': rays}
    if rgb is not None:
        out['rgb'] = rgb
    if fg is not None:
        out['fg_mask'] = fg
    return out


class RayBudget:
    """``dynamic_ray_sampling`` of the reference's systems (systems/nerf.py:91-95, systems/neus.py:91-95): after every step

        train_num_rays <- min(int(train_num_rays * 0.9 + int(train_num_rays * (train_num_samples / num_samples)) * 0.1), max_rays)

    with train_num_samples = train_num_rays(0) * num_samples_per_ray.  The reference reads ``out['num_samples'].sum().item()`` -- a
    host sync in the middle of every step.  Here ``observe()`` queues a non-blocking copy of the device-side count into pinned host
    memory and ``update()`` folds in every observation whose copy has completed, so the controller runs one step behind the device
    instead of stalling it (``sync=True`` restores the reference's step-exact behaviour)."""

    def __init__(self, train_num_rays, num_samples_per_ray, max_train_num_rays, sync=False):
        self.train_num_rays = int(train_num_rays)
        self.train_num_samples = int(train_num_rays) * int(num_samples_per_ray)
        self.max_train_num_rays = int(max_train_num_rays)
        self.sync = bool(sync)
        self._pending = []  # (pinned host tensor, event or None)

    @staticmethod
    def rule(train_num_rays, train_num_samples, num_samples, max_train_num_rays):
        """one update of systems/nerf.py:93-95, literally"""
        target = int(train_num_rays * (train_num_samples / num_samples))
        return min(int(train_num_rays * 0.9 + target * 0.1), max_train_num_rays)

    def observe(self, num_samples):
        """num_samples: the step's sample count(s) as a tensor (out['num_samples'], or the *_full count of NeuS); any shape, summed."""
        total = num_samples.detach().sum().reshape(1).to(torch.int64)
        if total.is_cuda:
            host = torch.empty(1, dtype=torch.int64, pin_memory=True)
            host.copy_(total, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._pending.append((host, ev))
        else:
            self._pending.append((total.clone(), None))
        return self.update(wait=self.sync)   # folds in whatever has arrived (everything when sync): the value for the NEXT step

    def update(self, wait=False):
        """fold in the finished observations (all of them when ``wait``); returns the current train_num_rays"""
        while self._pending:
            host, ev = self._pending[0]
            if ev is not None:
                if wait:
                    ev.synchronize()
                elif not ev.query():
                    break
            self._pending.pop(0)
            n = int(host[0])
            if n > 0:  # the reference would divide by zero here; an empty step carries no information
                self.train_num_rays = self.rule(self.train_num_rays, self.train_num_samples, n, self.max_train_num_rays)
        return self.train_num_rays
