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
    """Only classes we KNOW are lowered onto the fused kernels (identity / subclass of torch's, our Lamb, or the
    ``torch_optimizer.Lamb`` the reference's BERT example uses) — a user class that merely shares a name is left alone."""
    if not isinstance(optimizer_cls, type):
        return None
    for cls, kind in ((torch.optim.SGD, 'sgd'), (torch.optim.AdamW, 'adamw'), (torch.optim.Adam, 'adam'), (Lamb, 'lamb')):
        if optimizer_cls is cls or (issubclass(optimizer_cls, cls) and optimizer_cls.step is cls.step):
            return kind
    if optimizer_cls.__name__ == 'Lamb' and optimizer_cls.__module__.split('.')[0] == 'torch_optimizer':
        return 'lamb'
    return None


def build_optimizer(optimizer_cls, params, store, dp_group=None, fused=None, average_grads=False, sync_grads=None):
    kind = _kind_of(optimizer_cls)
    want_fused = fused if fused is not None else (store.device.type == 'cuda' and kind is not None)
    if want_fused and kind is not None and store.device.type == 'cuda':
        from ..parallel.fused_optim import FusedFlatOptimizer
        try:
            opt = FusedFlatOptimizer(store, kind, params, dp_group=dp_group, average_grads=average_grads, sync_grads=sync_grads)
            return opt, opt
        except TypeError as err:            # amsgrad / maximize / unknown kwargs: the stock optimizer has them, the kernel does not
            import warnings
            warnings.warn('fused optimizer not used: {}'.format(err))
    return optimizer_cls(store.master_params, **params), None
