"""Fused training-step back end: the per-ray loss of systems/nerf.py:68-97 as one CUDA op.

    loss, comp_rgb = nerf_rgb_loss(out['acc_rgb'], out['opacity'], background_color, target_rgb)

equals ``F.smooth_l1_loss(comp_rgb[valid], target[valid])`` with ``comp_rgb = acc_rgb + bg * (1 - opacity)`` and
``valid = opacity > 0`` (the boolean-mask indexing of the reference forces a host sync; this does not).  Optional: the
models work with any torch loss; this op removes ~30 small kernels per step from the captured graph."""
import torch

from .lib import lib, ptr, stream, check_cuda, contig


class _NerfRgbLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, acc_rgb, opacity, bg, target):
        n = acc_rgb.shape[0]
        accum = torch.zeros(2, device=acc_rgb.device)
        comp = torch.empty_like(acc_rgb)
        lib.call('nsr_nerf_loss_fwd', ptr(acc_rgb), ptr(opacity), ptr(bg), ptr(target), ptr(comp), ptr(accum), n, stream())
        ctx.save_for_backward(acc_rgb, opacity, bg, target, accum)
        ctx.mark_non_differentiable(comp)
        loss = accum[0] / torch.clamp(accum[1] * 3.0, min=1.0)
        return loss, comp

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
