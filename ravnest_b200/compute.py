"""Stage compute engine: forward / (recompute) / backward / optimizer step for one pipeline stage.

Behavioural parity with reference ``ravnest/compute.py`` (``Compute``, :10-540):
  * a forward pass is identified by ``forward_pass_id`` and remembers the weight VERSION it used
    (``fpid_to_version`` :23-26, :85-90);
  * backward of fpid runs against that same version — PipeDream weight stashing — while the update
    is applied to the newest weights (:133-206, :214-271);
  * ``update_frequency`` gradient accumulation and one loss line per optimizer update (:273-311);
  * RNG is replayed so a recomputed forward sees the same dropout masks (:63-68, :234-237).

Design differences (B200-first):
  * weights live in a flat versioned ``ParamStore``; stashing is pinning a ring slot, not cloning a
    state_dict through a Manager process;
  * two execution modes: ``recompute=True`` is the reference schedule (no-grad forward, re-run with
    grad before backward); ``recompute=False`` keeps the autograd graph of every in-flight fpid
    resident in HBM (180 GB/GPU makes the mandatory recompute a bad trade) — same numerics, one
    forward less per step on root/stem stages;
  * the weight version only advances when the weights actually change (every optimizer update),
    not on every backward (reference :191) — observable behaviour is identical, the per-backward
    full-model clone disappears;
  * optimizer step runs over flat buffers and (on CUDA) through the fused all-reduce+update kernel
    in ``ravnest_b200.parallel.fused_optim``.
"""
import os
import threading
from typing import Any, Callable, Dict, List, Optional, Sequence

import torch
from torch.func import functional_call

from .runtime.param_store import ParamStore, Slot


def _flatten_outputs(out: Any) -> List[torch.Tensor]:
    if isinstance(out, torch.Tensor):
        return [out]
    if isinstance(out, (tuple, list)):
        flat = []
        for o in out:
            flat.extend(_flatten_outputs(o))
        return flat
    if isinstance(out, dict):
        flat = []
        for o in out.values():
            flat.extend(_flatten_outputs(o))
        return flat
    return []


class Flight:
    """Book-keeping for one in-flight forward pass."""
    __slots__ = ('fpid', 'slot', 'inputs', 'outputs', 'rng', 'recomputed', 'graph')

    def __init__(self, fpid, slot, inputs, outputs, rng, graph=None):
        self.graph = graph
        self.fpid = fpid
        self.slot: Slot = slot
        self.inputs: List[torch.Tensor] = inputs
        self.outputs: Optional[List[torch.Tensor]] = outputs
        self.rng = rng
        self.recomputed = outputs is not None


class Compute:
    def __init__(self, model: torch.nn.Module, optimizer_cls=None, optimizer_params: Optional[dict] = None,
                 criterion: Optional[Callable] = None, device='cpu', update_frequency: int = 1,
                 cluster_length: int = 1, recompute: bool = True, compute_dtype: Optional[torch.dtype] = None,
                 loss_filename: Optional[str] = 'losses.txt', is_last: bool = False, dp_group=None,
                 fused_optimizer: Optional[bool] = None, lr_scheduler=None, lr_scheduler_params: Optional[dict] = None,
                 average_grads_over_dp: bool = False, use_cuda_graphs: bool = False, channels_last: bool = False):
        self.device = torch.device(device)
        self.model = model
        self.criterion = criterion
        self.update_frequency = max(1, int(update_frequency))
        self.recompute = bool(recompute)
        self.is_last = is_last
        self.loss_filename = loss_filename
        self.store = ParamStore(model, self.device, n_slots=cluster_length + 2, compute_dtype=compute_dtype)
        if self.device.type == 'cuda':
            from . import ops
            if ops.backend() == 'native' and os.environ.get('RAVNEST_WGRAD_STAGING', '1') != '0':
                self.store.enable_wgrad_staging()       # fp32 accumulator for convolution weight gradients, flushed per update
        self.flights: Dict[int, Flight] = {}
        self.n_backwards = 0
        self.n_updates = 0
        self._loss_acc: Optional[torch.Tensor] = None
        self._pending_losses: List[Any] = []
        self._loss_lock = threading.Lock()
        self.last_loss: Optional[torch.Tensor] = None
        self.dp_group = dp_group
        self.lock = threading.RLock()
        self.optimizer = None
        self.lr_scheduler = None
        self._fused = None
        self._average_grads = bool(average_grads_over_dp)
        self.use_graphs = bool(use_cuda_graphs)
        self.channels_last = bool(channels_last)
        self.graph_warmup = 3
        self._graph = None
        self._graph_seen = 0
        self._eager_backwards = 0
        self._graph_failed = False
        self._stage_graphs = []
        self._max_stage_graphs = cluster_length + 2
        self.run_next_step_eagerly = False  # one-shot: skip the CUDA graphs for the next stage step (K2 pull in flight)
        self.after_next_step = None         # one-shot callback after the next forward / loss step
        self.loss_readback = False          # copy every step's loss to pinned host memory (async D2H)
        self._loss_host = None
        self.loss_history: List[Any] = []
        if optimizer_cls is not None:
            self._build_optimizer(optimizer_cls, dict(optimizer_params or {}), fused_optimizer, average_grads_over_dp)
            if lr_scheduler is not None:
                self.lr_scheduler = lr_scheduler(self.optimizer, **dict(lr_scheduler_params or {}))

    # ------------------------------------------------------------------ optimizer plumbing
    def _build_optimizer(self, optimizer_cls, params, fused, average_grads):
        from .optim import build_optimizer
        # the DP group always reaches the fused optimizer (parameter averaging needs it too); the gradient all-reduce inside
        # step() is switched on only in gradient-mode DP
        self.optimizer, self._fused = build_optimizer(optimizer_cls, params, self.store, dp_group=self.dp_group,
                                                     fused=fused, average_grads=average_grads, sync_grads=average_grads)

    def optimizer_step(self) -> None:
        """Apply the accumulated gradient to the newest weights and publish a new version."""
        with torch.no_grad():
            self.store.flush_wgrad()                # staged convolution weight gradients -> gradient bucket (no-op without staging)
            if self._fused is not None:
                self._fused.step()              # all-reduce (+) update (+) cast-publish in one kernel
            else:
                if self.dp_group is not None and getattr(self, '_average_grads', False):
                    torch.distributed.all_reduce(self.store.grad, group=self.dp_group)
                    self.store.grad.div_(torch.distributed.get_world_size(self.dp_group))
                self.store.attach_master_grads()
                self.optimizer.step()
                self.store.publish()
                self.store.zero_grad()
        self.n_updates += 1

    def sync_master(self) -> None:
        """Make the WHOLE fp32 master (and optimizer moments) current on this rank.  No-op except with the two-shot fused optimizer,
        which keeps them sharded over the stage's data-parallel replicas (``parallel/fused_optim.py``)."""
        if self._fused is not None and hasattr(self._fused, 'sync_master'):
            self._fused.sync_master()

    # ------------------------------------------------------------------ RNG replay
    def _rng_capture(self):
        cpu = torch.get_rng_state()
        cuda = torch.cuda.get_rng_state(self.device) if self.device.type == 'cuda' else None
        return cpu, cuda

    def _rng_restore(self, state):
        cpu, cuda = state
        torch.set_rng_state(cpu)
        if cuda is not None:
            torch.cuda.set_rng_state(cuda, self.device)

    # ------------------------------------------------------------------ forward
    def _prepare_inputs(self, args: Sequence[Any], need_grad: bool) -> List[Any]:
        prepared = []
        for a in args:
            if isinstance(a, torch.Tensor):
                a = a.detach()
                if a.device != self.device:
                    a = a.to(self.device, non_blocking=True)
                if a.is_floating_point() and a.dtype != self.store.compute_dtype and a.dtype in (torch.float32, torch.float16, torch.bfloat16):
                    a = a.to(self.store.compute_dtype)
                if self.channels_last and a.dim() == 4 and a.is_floating_point() and a.shape[1] >= 8:
                    # (image batches — fewer than 8 channels — stay as they are: the stem's pack kernel reads any layout)
                    a = a.contiguous(memory_format=torch.channels_last)
                if need_grad and a.is_floating_point():
                    a.requires_grad_(True)
            prepared.append(a)
        return prepared

    def _call(self, slot: Slot, inputs: Sequence[Any]):
        return functional_call(self.model, slot.param_dict, tuple(inputs))

    def forward(self, fpid: int, args: Sequence[Any], input_grads: bool = True):
        """Training-mode forward of a root/stem stage.  Returns the (detached) stage output."""
        with self.lock:
            self.model.train()
            slot = self.store.acquire()
            if self.recompute:
                inputs = self._prepare_inputs(args, need_grad=False)
                rng = self._rng_capture()
                with torch.no_grad():
                    out = self._call(slot, inputs)
                self.flights[fpid] = Flight(fpid, slot, inputs, None, rng)
                return out
            eager_once, self.run_next_step_eagerly = self.run_next_step_eagerly, False
            if self.use_graphs and self.device.type == 'cuda' and not eager_once:
                gs = self._stage_graph_for(args, input_grads)
                if gs is not None:
                    out = gs.run_forward(args)
                    self.flights[fpid] = Flight(fpid, slot, [], [], None, graph=gs)
                    self._fire_after_step()
                    return out
            inputs = self._prepare_inputs(args, need_grad=input_grads)
            with torch.enable_grad():
                out = self._call(slot, inputs)
            self.flights[fpid] = Flight(fpid, slot, inputs, _flatten_outputs(out), None)
            self._fire_after_step()
            return _detach_tree(out)

    def _fire_after_step(self) -> None:
        hook, self.after_next_step = self.after_next_step, None
        if hook is not None:
            hook()

    def _stage_graph_for(self, args, input_grads):
        """A free GraphedStage for this input signature (captured lazily after the eager warm-up)."""
        from .runtime.graphs import GraphedStage, _signature
        if self._graph_failed:
            return None
        sig = _signature(args, None)
        for gs in self._stage_graphs:
            if not gs.busy and gs.signature == sig:
                return gs
        self._graph_seen += 1
        # capture records forward AND backward: both must have run eagerly first (cuDNN autotuning, lazily created workspaces
        # and allocator growth are illegal inside a capture).  Deep pipelines see several forwards before the first backward
        # comes back, so counting forwards alone is not enough (4-stage ResNet: capture invalidated on every non-leaf stage).
        if (self._graph_seen <= self.graph_warmup or self._eager_backwards < 2 or len(self._stage_graphs) >= self._max_stage_graphs):
            return None
        try:
            # capture every instance the schedule can ever need (one per in-flight pass) NOW, inside the warm-up, instead of
            # lazily when the in-flight depth first reaches a new maximum (that used to land a ~30 ms capture in steady state)
            prepared = self._prepare_inputs(args, need_grad=False)
            first = None
            while len(self._stage_graphs) < self._max_stage_graphs:
                gs = GraphedStage(self, prepared, input_grads)
                self._stage_graphs.append(gs)
                first = first or gs
            return first
        except Exception as err:
            self._graph_failed = True
            import warnings
            warnings.warn('CUDA-graph capture of the stage failed, running eagerly: {!r}'.format(err))
            torch.cuda.synchronize(self.device)
            return None

    def recompute_forward(self, fpid: int, input_grads: bool = True) -> None:
        """Re-run fpid's forward WITH grad on its stashed weight version and RNG stream
        (reference compute.py:214-271)."""
        fl = self.flights[fpid]
        if fl.recomputed:
            return
        inputs = []
        for a in fl.inputs:
            if isinstance(a, torch.Tensor) and a.is_floating_point() and input_grads:
                a = a.detach().requires_grad_(True)
            inputs.append(a)
        devices = [self.device] if self.device.type == 'cuda' else []
        with torch.random.fork_rng(devices=devices):
            self._rng_restore(fl.rng)
            self.model.train()
            with torch.enable_grad():
                out = self._call(fl.slot, inputs)
        fl.inputs = inputs
        fl.outputs = _flatten_outputs(out)
        fl.recomputed = True

    # ------------------------------------------------------------------ backward
    def backward(self, fpid: int, grads: Dict[int, torch.Tensor], update: Optional[bool] = None) -> List[Optional[torch.Tensor]]:
        """Backward of a root/stem stage.  ``grads`` maps output index -> dL/d(output).  Returns the
        gradient for every stage input (``None`` for non-differentiable inputs)."""
        with self.lock:
            fl = self.flights[fpid]
            if fl.graph is not None:
                in_grads = fl.graph.run_backward(grads)
                self.store.release(fl.slot)
                del self.flights[fpid]
                self._after_backward(update)
                return in_grads
            if not fl.recomputed:
                self.recompute_forward(fpid)
            outs, gs = [], []
            for idx, o in enumerate(fl.outputs):
                g = grads.get(idx)
                if g is None or not o.requires_grad:
                    continue
                if g.device != o.device:
                    g = g.to(o.device, non_blocking=True)
                if g.dtype != o.dtype:
                    g = g.to(o.dtype)
                outs.append(o)
                gs.append(g)
            if outs:
                torch.autograd.backward(outs, gs)
            self._eager_backwards += 1
            in_grads = [a.grad if isinstance(a, torch.Tensor) and a.requires_grad else None for a in fl.inputs]
            self.store.release(fl.slot)
            del self.flights[fpid]
            self._after_backward(update)
            return in_grads

    def _after_backward(self, update: Optional[bool]) -> bool:
        self.n_backwards += 1
        if update is None:
            update = self.n_backwards % self.update_frequency == 0
        if update and self.optimizer is not None:
            self.optimizer_step()
        return bool(update)

    # ------------------------------------------------------------------ leaf / solo
    def loss_step(self, args: Sequence[Any], targets: Any, update: Optional[bool] = None, input_grads: bool = True):
        """Leaf (or solo) step: forward WITH grad, criterion, backward, maybe update
        (reference compute.py:273-311).  Returns (loss tensor, input grads).

        On CUDA the step is captured into a CUDA graph after ``graph_warmup`` eager steps (see
        runtime/graphs.py); anything un-capturable in the user's criterion falls back to eager."""
        with self.lock:
            self.model.train()
            slot = self.store.current
            det = in_grads = None
            eager_once, self.run_next_step_eagerly = self.run_next_step_eagerly, False
            if self.use_graphs and self.device.type == 'cuda' and not eager_once:
                det, in_grads = self._graphed_loss_step(args, targets, input_grads, slot)
            if det is None:
                inputs = self._prepare_inputs(args, need_grad=input_grads)
                with torch.enable_grad():
                    out = self._call(slot, inputs)
                    loss = self.criterion(out, targets)
                loss.backward()
                in_grads = [a.grad if isinstance(a, torch.Tensor) and a.requires_grad else None for a in inputs]
                det = loss.detach()
            self._fire_after_step()
            self.last_loss = det
            if self.loss_readback and det.is_cuda:
                if self._loss_host is None:
                    self._loss_host = torch.empty(64, dtype=torch.float32, pin_memory=True)
                    self._loss_slot = 0
                self._loss_host[self._loss_slot % 64].copy_(det.float(), non_blocking=True)
                self._loss_slot += 1
            self._loss_acc = det.clone() if self._loss_acc is None else self._loss_acc.add_(det)
            updated = self._after_backward(update)
            if updated:
                self._queue_loss_line()
            return det, in_grads

    def _graphed_loss_step(self, args, targets, input_grads, slot):
        from .runtime.graphs import GraphedLossStep
        g = self._graph
        if g is not None and g.matches(args, targets, slot):
            return g.run(args, targets)
        self._graph_seen += 1
        if self._graph_seen <= self.graph_warmup or self._graph_failed:
            return None, None
        try:
            inputs = self._prepare_inputs(args, need_grad=False)
            self._graph = GraphedLossStep(self, inputs, targets, input_grads)
            return self._graph.run(inputs, targets)
        except Exception as err:          # un-capturable criterion / op: stay eager, say so once
            self._graph_failed = True
            self._graph = None
            import warnings
            warnings.warn('CUDA-graph capture of the loss step failed, running eagerly: {!r}'.format(err))
            torch.cuda.synchronize(self.device)
            return None, None

    def _queue_loss_line(self) -> None:
        """One line per optimizer update = loss summed over the accumulation window
        (reference compute.py:290-301) — but copied to pinned host memory asynchronously instead of
        a blocking ``.item()`` on the scheduler thread."""
        acc, self._loss_acc = self._loss_acc, None
        if acc is None or self.loss_filename is None:
            return
        if acc.device.type == 'cuda':
            host = torch.empty((), dtype=torch.float32, pin_memory=True)
            host.copy_(acc.float(), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._pending_losses.append((host, ev))
        else:
            self._pending_losses.append((acc.float(), None))
        if len(self._pending_losses) >= 64:
            self.flush_losses(block=False)

    def flush_losses(self, block: bool = True) -> None:
        """Write the loss lines whose device-to-host copies have landed.  Callable from any thread (scheduler, a stop() triggered
        by a SHUTDOWN message, the user's main thread): the queue is drained under its own lock."""
        if self.loss_filename is None:
            return
        with self._loss_lock:
            ready = []
            while self._pending_losses:
                host, ev = self._pending_losses[0]
                if ev is not None and not ev.query():
                    if not block:
                        break
                    ev.synchronize()
                ready.append(round(float(host), 4))
                self._pending_losses.pop(0)
            if ready:
                with open(self.loss_filename, 'a') as fh:
                    fh.write(''.join('{}\n'.format(v) for v in ready))

    # ------------------------------------------------------------------ inference
    def no_grad_forward(self, args: Sequence[Any]):
        with self.lock:
            was = self.model.training
            self.model.eval()
            try:
                inputs = self._prepare_inputs(args, need_grad=False)
                with torch.no_grad():
                    return self._call(self.store.current, inputs)
            finally:
                self.model.train(was)

    # ------------------------------------------------------------------ state
    def current_version(self) -> int:
        return self.store.version

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """Newest weights under their original parameter names (+ buffers)."""
        self.store.bind_module()
        return self.model.state_dict()

    def latest_weights(self, first: Optional[str] = None, last: Optional[str] = None) -> Dict[str, torch.Tensor]:
        """Master copy of parameters ``first..last`` (served by get_latest_weights, endpoints.py:145)."""
        self.sync_master()
        names = self.store.names
        i = names.index(first) if first in names else 0
        j = names.index(last) if last in names else len(names) - 1
        with torch.no_grad():
            return {n: self.store.master_views[k].detach().to('cpu') for k, n in enumerate(names) if i <= k <= j}

    def load_weights(self, tensors: Dict[str, torch.Tensor]) -> None:
        """Adopt externally produced weights (parameter averaging / latest-weights pull) as a new version."""
        with self.lock:
            self.store.load_master_from(tensors)
            self.store.publish()


def _detach_tree(out: Any) -> Any:
    if isinstance(out, torch.Tensor):
        return out.detach()
    if isinstance(out, tuple):
        return tuple(_detach_tree(o) for o in out)
    if isinstance(out, list):
        return [_detach_tree(o) for o in out]
    if isinstance(out, dict):
        return {k: _detach_tree(v) for k, v in out.items()}
    return out
