"""User-facing training loop.

API parity: reference ``ravnest/trainer.py:6-127`` — constructor arguments (:37), ``train()``,
``evaluate()``, ``pred(data)``, ``prelim_checks()`` (subclasses call it first; docs/features.rst:28).
Only the root (or a solo node) drives data; stem / leaf nodes serve until the job ends.

Deliberate differences:
  * ``prelim_checks`` on stem/leaf blocks on an event and RETURNS ``False`` when the pipeline has
    been shut down, instead of spinning forever in ``while True: time.sleep(0)`` (:54-57);
  * the dtype handling of non-tensor batches is fixed (the reference's branch is inverted, :79-83);
  * ``pred`` returns the prediction on the root — the leaf's result travels back up the backward lane
    (``ActionTypes.INFERENCE_RESULT``); the reference returns ``None`` (:113).
"""
import time

import numpy as np
import torch

from .strings import ActionTypes, NodeTypes


class Trainer:
    def __init__(self, node=None, lr_scheduler=None, lr_scheduler_params=None, train_loader=None, val_loader=None,
                 val_freq=1, save=False, epochs=1, batch_size=64, step_size=1, inputs_dtype=None):
        self.node = node
        self.is_driver = node.node_type in (NodeTypes.ROOT, NodeTypes.SOLO)
        self.train_loader = train_loader
        self.val_loader = val_loader
        self.val_freq = val_freq
        self.save = save
        self.epochs = epochs
        self.batch_size = batch_size
        self.step_size = step_size
        self.n_forwards = 0
        self.inputs_dtype = inputs_dtype
        self.lr_scheduler = None
        self.train_seconds = None
        if self.is_driver and lr_scheduler is not None and node.optimizer is not None:
            self.lr_scheduler = lr_scheduler(node.optimizer, **(lr_scheduler_params or {}))

    # ------------------------------------------------------------------ helpers
    def prelim_checks(self):
        """Stem / leaf providers park here while the scheduler thread serves the pipeline."""
        if self.is_driver:
            return True
        self.node.wait_until_stopped()
        return False

    def _as_input(self, x):
        if not torch.is_tensor(x):
            x = torch.as_tensor(np.asarray(x))
        if self.inputs_dtype is not None and x.dtype != self.inputs_dtype:
            x = x.to(self.inputs_dtype)
        return x

    # ------------------------------------------------------------------ API
    def train(self):
        if not self.prelim_checks():
            return
        t0 = time.time()
        self.n_forwards = 0
        for epoch in range(self.epochs):
            for batch in self.train_loader:
                x = batch[0] if isinstance(batch, (tuple, list)) else batch
                self.node.forward_compute(tensors=self._as_input(x))
                self.n_forwards += 1
            if self.val_loader is not None and (epoch + 1) % max(1, self.val_freq if self.val_freq < self.epochs else 1) == 0:
                self.evaluate()
            self.node.wait_for_backwards()
            if self.lr_scheduler is not None:
                self.lr_scheduler.step()
        self.node.comm_session.parallel_ring_reduce()
        self.train_seconds = time.time() - t0
        if self.save:
            self.node.trigger_save_submodel()

    def finish(self, grace: float = 1.0, timeout=None) -> None:
        """End a provider script (new).  On the driving node (root / solo): wait for outstanding work, then stop the node — a
        stopping root sends SHUTDOWN down its pipeline; on stem / leaf nodes: block until that SHUTDOWN arrives.  Without this
        call providers keep serving forever, which is the reference's behaviour."""
        if self.is_driver:
            self.node.wait_for_backwards()
            time.sleep(grace)                  # let trailing inference / save messages reach the leaf
            self.node.stop()
        else:
            self.node.wait_until_stopped(timeout)

    def evaluate(self):
        if not self.is_driver or self.val_loader is None:
            return
        self.node.wait_for_backwards()
        for batch in self.val_loader:
            x = batch[0] if isinstance(batch, (tuple, list)) else batch
            self.node.no_grad_forward_compute(tensors=self._as_input(x), output_type=ActionTypes.VAL_ACCURACY)

    def pred(self, data):
        if not self.is_driver:
            return None
        x = torch.as_tensor(np.asarray(data), dtype=torch.float32) if isinstance(data, np.ndarray) else data
        return self.node.no_grad_forward_compute(tensors=x, output_type=ActionTypes.ACCURACY)
