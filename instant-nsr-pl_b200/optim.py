"""Fused AdamW for the flat parameter vectors of the drop-in modules.

The reference builds ``torch.optim.AdamW(lr=1e-2, betas=(0.9, 0.99), eps=1e-15)`` from its config (systems/utils.py:314-325,
configs/nerf-blender.yaml:74-79) -- a dense update over all 12.6 M table entries every step, ~8 table-sized torch kernels plus the
fp32->fp16 re-cast tiny-cuda-nn does in the next forward.  ``FusedAdamW`` is the same optimizer (same constructor arguments,
param groups, ``state_dict`` keys ``step / exp_avg / exp_avg_sq``) with one kernel per parameter tensor (``nsr_adamw_step``) that also
refreshes the fp16 copy our kernels read.

    opt = FusedAdamW.for_model(model, lr=1e-2, betas=(0.9, 0.99), eps=1e-15)     # or FusedAdamW(params, ...)
    loss.backward(); opt.step()

``capturable=True`` keeps learning rate and step count on the device so that ``step()`` can be captured into the CUDA graph of the
training step (``nsr_b200.graph.GraphedStep(post_backward=opt.step)``); call ``opt.sync_lr()`` outside the graph after a scheduler
changed ``group['lr']``.
"""
import ctypes as C

import torch

from .lib import lib, ptr, stream, AdamWT, NsrError


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, capturable=False, half_shadows=None):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError('FusedAdamW: invalid hyper-parameters')
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.capturable = bool(capturable)
        # parameter -> module owning an fp16 copy of it (tcnn-shaped modules): refreshed by the same kernel
        self._shadows = dict(half_shadows or {})
        self.grad_scale = None   # optional 1-element CUDA tensor: gradients are divided by it (GradScaler.get_scale())
        self.found_inf = None    # optional 1-element CUDA float tensor: != 0 skips the step (GradScaler semantics)

    @classmethod
    def for_model(cls, model, params=None, **kw):
        """Collects the fp16-copy owners (modules with ``params`` + ``_params_half``) of ``model`` automatically."""
        shadows = {}
        for mod in model.modules():
            if hasattr(mod, '_params_half') and isinstance(getattr(mod, 'params', None), torch.nn.Parameter):
                shadows[mod.params] = mod
        return cls(params if params is not None else [p for p in model.parameters() if p.requires_grad], half_shadows=shadows, **kw)

    def _init_state(self, p, group):
        st = self.state[p]
        if len(st) == 0:
            st['step'] = torch.zeros((), dtype=torch.float32, device=p.device) if self.capturable else 0
            st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if self.capturable:
                st['lr_step'] = torch.tensor([group['lr'], 0.0], dtype=torch.float32, device=p.device)
        return st

    def sync_lr(self):
        """capturable mode: push the groups' current ``lr`` to the device-side copies (call outside graph capture)."""
        for group in self.param_groups:
            for p in group['params']:
                st = self.state.get(p)
                if st and 'lr_step' in st:
                    st['lr_step'][0] = float(group['lr'])

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group['betas']
            for p in group['params']:
                if p.grad is None or p.numel() == 0:
                    continue
                if not p.is_cuda:
                    raise NotImplementedError('FusedAdamW: only CUDA parameters are supported; there is no CPU path')
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous():
                    raise NsrError('FusedAdamW: parameters and gradients must be contiguous fp32')
                if p.grad.is_sparse:
                    raise NsrError('FusedAdamW does not support sparse gradients')
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                st = self._init_state(p, group)
                h = AdamWT()
                h.lr, h.beta1, h.beta2, h.eps, h.weight_decay = group['lr'], b1, b2, group['eps'], group['weight_decay']
                h.inv_grad_scale = 1.0
                # a skipped update (found_inf != 0, GradScaler semantics) must not advance the bias-correction step count
                if self.capturable:
                    inc = 1.0 if self.found_inf is None else (self.found_inf.reshape(()) == 0).to(torch.float32)
                    st['lr_step'][1] += inc   # torch op: part of the captured graph
                    st['step'] += inc
                    h.step, dev_state = 0, st['lr_step']
                else:
                    if self.found_inf is None or float(self.found_inf) == 0.0:   # host mode: one read when a scaler is attached
                        st['step'] += 1
                    h.step, dev_state = max(int(st['step']), 1), None
                half = None
                mod = self._shadows.get(p)
                if mod is not None:
                    half = mod._params_half()   # current copy (allocated on first use); overwritten in place below
                g_in = g
                if self.grad_scale is not None:   # un-scale on the fly: fold 1/scale into the kernel when it is a host number
                    if torch.is_tensor(self.grad_scale):
                        g_in = g / self.grad_scale
                    else:
                        h.inv_grad_scale = 1.0 / float(self.grad_scale)
                lib.call('nsr_adamw_step', C.byref(h), ptr(p), ptr(g_in), ptr(st['exp_avg']), ptr(st['exp_avg_sq']), ptr(half),
                         ptr(dev_state), ptr(self.found_inf), p.numel(), stream())
                torch.autograd.graph.increment_version(p)   # the write happened behind autograd's back
                if mod is not None:   # the fp16 copy is already up to date: re-key the module's cache to the new version
                    mod._half_key = (p._version, p.data_ptr(), p.device)
        return loss


def _get_parameters(model, name):
    m = model
    for part in name.split('.'):
        m = getattr(m, part)
    if isinstance(m, torch.nn.Module):
        return list(m.parameters())
    if isinstance(m, torch.nn.Parameter):
        return [m]
    return []


def parse_optimizer(config, model):
    """systems/utils.py:314-325 (``parse_optimizer``, called from systems/base.py:119) with the same config section -- ``name``,
    ``args`` and the optional per-submodule ``params`` (configs/neus-blender.yaml:96-102: geometry / texture / variance ...) -- but
    ``name: AdamW`` builds the one-pass FusedAdamW (identical update rule, state_dict keys and param groups; it also keeps the fp16
    parameter copies of the tcnn-shaped modules fresh).  Any other name resolves to ``torch.optim`` as in the reference."""
    args = dict(config.get('args', {}) or {})
    if 'params' in config and config['params']:
        groups = [{'params': _get_parameters(model, name), 'name': name, **dict(group_args or {})} for name, group_args in config['params'].items()]
    else:
        groups = [p for p in model.parameters() if p.requires_grad]
    if config['name'] in ('AdamW', 'FusedAdamW'):
        args.pop('fused', None)
        args.pop('foreach', None)
        return FusedAdamW.for_model(model, params=groups, **args)
    return getattr(torch.optim, config['name'])(groups, **args)
