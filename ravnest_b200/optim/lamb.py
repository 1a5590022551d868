"""LAMB optimizer (You et al. 2019), pure-PyTorch reference implementation.

The reference's BERT example uses ``torch_optimizer.Lamb(lr, betas, eps, weight_decay, clamp_value,
adam, debias)`` (examples/bert/provider.py:6,:48-49), which is not installed here.  This class keeps
that constructor and update rule:

    m <- b1 m + (1-b1) g ;  v <- b2 v + (1-b2) g^2 ;  (optional bias correction)
    u <- m / (sqrt(v) + eps) + wd * w
    trust <- clamp(||w||, 0, clamp_value) / ||u||        (1 if either norm is 0, or adam=True)
    w <- w - lr * c * trust * u      with c = sqrt(1-b2^t)/(1-b1^t) if debias else 1

It is the numerical reference for the fused CUDA LAMB in ``ravnest_b200/ops/csrc/fused_optim.cu``.
"""
import math

import torch
from torch.optim import Optimizer


class Lamb(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, clamp_value=10.0,
                 adam=False, debias=False):
        if lr <= 0.0:
            raise ValueError('Invalid learning rate: {}'.format(lr))
        if eps < 0.0:
            raise ValueError('Invalid epsilon value: {}'.format(eps))
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError('Invalid betas: {}'.format(betas))
        if weight_decay < 0 or clamp_value < 0:
            raise ValueError('Invalid weight_decay / clamp_value')
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, clamp_value=clamp_value,
                        adam=adam, debias=debias)
        super().__init__(params, defaults)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                g = p.grad
                state = self.state[p]
                if not state:
                    state['step'] = 0
                    state['exp_avg'] = torch.zeros_like(p)
                    state['exp_avg_sq'] = torch.zeros_like(p)
                m, v = state['exp_avg'], state['exp_avg_sq']
                state['step'] += 1
                m.mul_(b1).add_(g, alpha=1 - b1)
                v.mul_(b2).addcmul_(g, g, value=1 - b2)
                if group['debias']:
                    c = math.sqrt(1 - b2 ** state['step']) / (1 - b1 ** state['step'])
                else:
                    c = 1.0
                update = m / v.sqrt().add(group['eps'])
                if group['weight_decay'] != 0:
                    update.add_(p, alpha=group['weight_decay'])
                w_norm = torch.linalg.vector_norm(p).clamp(0, group['clamp_value'])
                u_norm = torch.linalg.vector_norm(update)
                if group['adam']:
                    trust = torch.ones((), device=p.device)
                else:
                    trust = torch.where((w_norm > 0) & (u_norm > 0), w_norm / u_norm, torch.ones_like(w_norm))
                p.add_(update * trust, alpha=-group['lr'] * c)
        return loss
