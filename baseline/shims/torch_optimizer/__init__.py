"""Stand-in for ``torch_optimizer`` (not installable offline): only ``Lamb``, the optimizer of the reference's BERT
example (``examples/bert/provider.py:6,:48-49``).  Plain per-parameter PyTorch implementation of LAMB (You et al., 2019)
with the constructor of ``torch_optimizer.Lamb``: lr, betas, eps, weight_decay, clamp_value, adam, debias."""
import math

import torch
from torch.optim.optimizer import Optimizer

__all__ = ['Lamb']


class Lamb(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, clamp_value=10.0, adam=False,
                 debias=False):
        if lr <= 0.0 or eps < 0.0 or weight_decay < 0 or clamp_value < 0.0:
            raise ValueError('invalid Lamb hyper-parameter')
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError('invalid beta: {}'.format(betas))
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.clamp_value, self.adam, self.debias = clamp_value, adam, debias

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                grad = p.grad
                if grad.is_sparse:
                    raise RuntimeError('Lamb does not support sparse gradients')
                state = self.state[p]
                if len(state) == 0:
                    state['step'] = 0
                    state['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                exp_avg, exp_avg_sq = state['exp_avg'], state['exp_avg_sq']
                state['step'] += 1
                exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)
                exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
                if self.debias:
                    correction = math.sqrt(1 - beta2 ** state['step']) / (1 - beta1 ** state['step'])
                else:
                    correction = 1.0
                step_size = group['lr'] * correction
                weight_norm = torch.norm(p).clamp(0, self.clamp_value)
                update = exp_avg / exp_avg_sq.sqrt().add(group['eps'])
                if group['weight_decay'] != 0:
                    update.add_(p, alpha=group['weight_decay'])
                update_norm = torch.norm(update)
                if weight_norm == 0 or update_norm == 0:
                    trust = 1.0
                else:
                    trust = weight_norm / update_norm
                if self.adam:
                    trust = 1.0
                p.add_(update * (-step_size * trust))
        return loss
