"""Periodic sharded checkpoints with resume (new; the reference only saves weights at the end,
SURVEY.md §5.4).  One file per node next to its stage artefact:
``<template_path>/ckpt_<tag>.pt`` = master params + optimizer state + counters + RNG."""
import os

import torch


def _path(node, tag):
    return os.path.join(node.template_path, 'ckpt_{}.pt'.format(tag))


def _philox_state(device):
    if device.type != 'cuda':
        return None
    from ..ops import native
    return native.rng_state(device).detach().to('cpu')


def _snapshot(obj):
    if isinstance(obj, torch.Tensor):
        return obj.detach().to('cpu', copy=True)
    if isinstance(obj, dict):
        return {k: _snapshot(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_snapshot(v) for v in obj)
    return obj


def save_checkpoint(node, tag='latest'):
    cs = node.compute_session
    with cs.lock:
        cs.sync_master()
        # everything is COPIED under the lock (``.to('cpu')`` aliases CPU tensors, ``optimizer.state_dict()`` returns live references):
        # the scheduler may update the stage while torch.save serialises
        state = {
            'master': cs.store.master.detach().to('cpu', copy=True),
            'names': list(cs.store.names),
            'buffers': {k: v.detach().to('cpu', copy=True) for k, v in node.model.named_buffers()},
            'optimizer': _snapshot(cs.optimizer.state_dict()) if cs.optimizer is not None else None,
            'lr_scheduler': cs.lr_scheduler.state_dict() if cs.lr_scheduler is not None else None,
            'counters': {'n_backwards': cs.n_backwards, 'n_updates': cs.n_updates, 'version': cs.store.version,
                         'forward_pass_id': node.forward_pass_id, 'latest_backward_id': node.latest_backward_id,
                         'node_n_backwards': node.n_backwards, 'n_forwards': node.n_forwards},
            'rng': {'cpu': torch.get_rng_state(),
                    'cuda': torch.cuda.get_rng_state(node.device) if node.device.type == 'cuda' else None,
                    'philox': _philox_state(node.device)},     # {seed, step} of the dropout kernels (ops/native.py rng_state)
        }
    path = _path(node, tag)
    tmp = path + '.tmp'
    torch.save(state, tmp)
    os.replace(tmp, path)
    return path


def load_checkpoint(node, tag='latest'):
    path = _path(node, tag)
    if not os.path.exists(path):
        return False
    state = torch.load(path, map_location='cpu', weights_only=False)
    cs = node.compute_session
    with cs.lock:
        assert state['names'] == cs.store.names, 'checkpoint belongs to a different stage split'
        cs.store.master.copy_(state['master'].to(cs.store.device))
        cs.store.publish()
        buffers = dict(node.model.named_buffers())
        for k, v in state['buffers'].items():
            if k in buffers:
                buffers[k].copy_(v.to(buffers[k].device))
        if state['optimizer'] is not None and cs.optimizer is not None:
            cs.optimizer.load_state_dict(state['optimizer'])
        if state['lr_scheduler'] is not None and cs.lr_scheduler is not None:
            cs.lr_scheduler.load_state_dict(state['lr_scheduler'])
        c = state['counters']
        cs.n_backwards, cs.n_updates = c['n_backwards'], c['n_updates']
        node.forward_pass_id = node.latest_backward_id = c['latest_backward_id']
        node.n_backwards, node.n_forwards = c['node_n_backwards'], c['latest_backward_id']
        torch.set_rng_state(state['rng']['cpu'])
        if state['rng']['cuda'] is not None and node.device.type == 'cuda':
            torch.cuda.set_rng_state(state['rng']['cuda'], node.device)
        if state['rng'].get('philox') is not None and node.device.type == 'cuda':
            from ..ops import native
            native.rng_state(node.device).copy_(state['rng']['philox'].to(node.device))
    return True
