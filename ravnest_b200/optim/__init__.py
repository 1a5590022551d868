"""Optimizers: pure-PyTorch LAMB + the fused flat-buffer family (CUDA).

``build_optimizer`` is what ``Compute`` calls: the user passes an optimizer CLASS plus kwargs (the
reference API, node.py:207-215).  On CPU — or for optimizer classes we do not recognise — the class is
instantiated over the stage's fp32 master parameters exactly like the reference does over its cloned
params.  On CUDA, SGD / Adam / AdamW / Lamb are lowered onto ``FusedFlatOptimizer`` whose step is ONE
kernel over the flat buckets: (DP all-reduce over NVLink peer memory) + update + bf16 re-publish.
"""
import torch

from .lamb import Lamb

__all__ = ['Lamb', 'build_optimizer']


def _kind_of(optimizer_cls):
    name = getattr(optimizer_cls, '__name__', '').lower()
    if optimizer_cls is torch.optim.SGD or name == 'sgd':
        return 'sgd'
    if optimizer_cls is torch.optim.AdamW or name == 'adamw':
        return 'adamw'
    if optimizer_cls is torch.optim.Adam or name == 'adam':
        return 'adam'
    if optimizer_cls is Lamb or name == 'lamb':
        return 'lamb'
    return None


def build_optimizer(optimizer_cls, params, store, dp_group=None, fused=None, average_grads=False):
    kind = _kind_of(optimizer_cls)
    want_fused = fused if fused is not None else (store.device.type == 'cuda' and kind is not None)
    if want_fused and kind is not None and store.device.type == 'cuda':
        from ..parallel.fused_optim import FusedFlatOptimizer
        opt = FusedFlatOptimizer(store, kind, params, dp_group=dp_group, average_grads=average_grads)
        return opt, opt
    return optimizer_cls(store.master_params, **params), None
