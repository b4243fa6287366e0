"""Fused training-step back end: the per-ray loss of systems/nerf.py:68-97 as one CUDA op.

    loss, comp_rgb = nerf_rgb_loss(out['acc_rgb'], out['opacity'], background_color, target_rgb)

equals ``F.smooth_l1_loss(comp_rgb[valid], target[valid])`` with ``comp_rgb = acc_rgb + bg * (1 - opacity)`` and
``valid = opacity > 0`` (the boolean-mask indexing of the reference forces a host sync; this does not).  Optional: the
models work with any torch loss; this op removes ~30 small kernels per step from the captured graph."""
import torch

import ctypes as _C

from .lib import lib, ptr, stream, check_cuda, contig, NeusLossT


class _NerfRgbLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, acc_rgb, opacity, bg, target):
        n = acc_rgb.shape[0]
        accum = torch.empty(4, device=acc_rgb.device)   # zeroed by the entry point; [2] = the loss
        comp = torch.empty_like(acc_rgb)
        lib.call('nsr_nerf_loss_fwd', ptr(acc_rgb), ptr(opacity), ptr(bg), ptr(target), ptr(comp), ptr(accum), n, stream())
        ctx.save_for_backward(acc_rgb, opacity, bg, target, accum)
        ctx.mark_non_differentiable(comp)
        return accum[2], comp

    @staticmethod
    def backward(ctx, g_loss, _g_comp):
        acc_rgb, opacity, bg, target, accum = ctx.saved_tensors
        n = acc_rgb.shape[0]
        g_acc = torch.empty_like(acc_rgb)
        g_op = torch.empty_like(opacity)
        gl = contig(g_loss.reshape(1), torch.float32)
        lib.call('nsr_nerf_loss_bwd', ptr(acc_rgb), ptr(opacity), ptr(bg), ptr(target), ptr(accum), ptr(gl), ptr(g_acc), ptr(g_op), n, stream())
        return g_acc, g_op, None, None


def nerf_rgb_loss(acc_rgb, opacity, background_color, target_rgb):
    """-> (loss scalar tensor, comp_rgb [N,3] detached)."""
    check_cuda(acc_rgb, opacity, background_color, target_rgb, what='nerf_rgb_loss')
    return _NerfRgbLoss.apply(contig(acc_rgb, torch.float32), contig(opacity, torch.float32), contig(background_color, torch.float32),
                              contig(target_rgb, torch.float32))


class _NeusLosses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, desc, comp_rgb, valid, target, opacity, fg_mask, sdf_grad, sdf, k_dev=None):
        n, k = comp_rgb.shape[0], (sdf_grad.shape[0] if sdf_grad is not None else (sdf.shape[0] if sdf is not None else 0))
        accum = torch.empty(8, device=comp_rgb.device)
        losses = torch.empty(7, device=comp_rgb.device)
        lib.call('nsr_neus_loss_fwd', _C.byref(desc), ptr(comp_rgb), ptr(valid), ptr(target), ptr(opacity), ptr(fg_mask), ptr(sdf_grad),
                 ptr(sdf), ptr(accum), ptr(losses), n, k, ptr(k_dev), stream())
        ctx.desc, ctx.k, ctx.k_dev = desc, k, k_dev
        ctx.save_for_backward(comp_rgb, valid, target, opacity, fg_mask, sdf_grad, sdf, accum)
        return losses[6], losses[:6].detach()

    @staticmethod
    def backward(ctx, g_total, _g_parts):
        comp_rgb, valid, target, opacity, fg_mask, sdf_grad, sdf, accum = ctx.saved_tensors
        n = comp_rgb.shape[0]
        g_rgb, g_op = torch.empty_like(comp_rgb), torch.empty_like(opacity)
        g_sg = torch.empty_like(sdf_grad) if (sdf_grad is not None and ctx.needs_input_grad[6]) else None
        g_s = torch.empty_like(sdf) if (sdf is not None and ctx.needs_input_grad[7]) else None
        gl = contig(g_total.reshape(1), torch.float32)
        lib.call('nsr_neus_loss_bwd', _C.byref(ctx.desc), ptr(comp_rgb), ptr(valid), ptr(target), ptr(opacity), ptr(fg_mask), ptr(sdf_grad),
                 ptr(sdf), ptr(accum), ptr(gl), ptr(g_rgb), ptr(g_op), ptr(g_sg), ptr(g_s), n, ctx.k, ptr(ctx.k_dev), stream())
        return None, g_rgb, None, None, g_op, None, g_sg, g_s, None


NEUS_LOSS_NAMES = ('rgb_mse', 'rgb_l1', 'eikonal', 'mask', 'opaque', 'sparsity')


def neus_losses(out, rgb, fg_mask=None, lambda_rgb_mse=10.0, lambda_rgb_l1=0.0, lambda_eikonal=0.1, lambda_mask=0.1, lambda_opaque=0.0,
                lambda_sparsity=0.0, sparsity_scale=1.0):
    """The loss block of systems/neus.py:98-121 as two CUDA kernels (one reduction, one gradient pass) instead of ~65 torch kernels and
    two boolean-mask host syncs.  ``out``: the 'neus' model's output dict (comp_rgb_full, rays_valid_full, opacity, sdf_grad_samples,
    sdf_samples; plus 'num_samples_dev' -- the device-side live sample count -- when the model ran in static-shape mode);
    ``rgb`` [N,3] target, ``fg_mask`` [N] (None = dataset without masks).  Defaults = configs/neus-blender.yaml:80-89.
    -> (total, parts) with parts[i] = the un-weighted loss NEUS_LOSS_NAMES[i] (for logging).  curvature / distortion terms
    (lambda 0 in every shipped config) stay with the caller."""
    comp, op = out['comp_rgb_full'], out['opacity']
    check_cuda(comp, op, rgb, what='neus_losses')
    d = NeusLossT(float(lambda_rgb_mse), float(lambda_rgb_l1), float(lambda_eikonal), float(lambda_mask if fg_mask is not None else 0.0),
                  float(lambda_opaque), float(lambda_sparsity), float(sparsity_scale))
    valid = contig(out['rays_valid_full'].reshape(-1), torch.bool).view(torch.uint8)
    sg, s = out.get('sdf_grad_samples'), out.get('sdf_samples')
    return _NeusLosses.apply(d, contig(comp, torch.float32), valid, contig(rgb, torch.float32), contig(op.reshape(-1), torch.float32),
                             None if fg_mask is None else contig(fg_mask.reshape(-1), torch.float32),
                             None if sg is None else contig(sg.reshape(-1, 3), torch.float32),
                             None if s is None else contig(s.reshape(-1), torch.float32), out.get('num_samples_dev'))
