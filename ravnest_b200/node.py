"""Node runtime: one pipeline stage (one GPU on a B200 box) and its scheduler.

Public surface kept from reference ``ravnest/node.py`` (``Node``, :23-782): constructor arguments
(:61-62), ``forward_compute :370``, ``no_grad_forward_compute :399``, ``wait_for_backwards :702``,
``trigger_save_submodel :712``, ``update_with_latest_weights :726``, ``reset :732``, attributes
``node_type`` / ``optimizer`` / ``comm_session`` used by ``Trainer``; action handlers are still
dispatched by name (:345, :362): ``root_forward``, ``forward``, ``no_grad_forward``, ``backward``,
``find_loss``, ``val_accuracy``, ``accuracy``, ``prediction``, ``save_submodel``.

Redesign:
  * ONE process per node/GPU; no ``mp.Manager`` server, no spawned gRPC process, no pickling between
    receiver and scheduler.  A ``Mailbox`` (condition variables) feeds one scheduler thread;
  * transport is pluggable: ``grpc`` (wire-compatible control plane + CPU path) or ``dist``
    (torch.distributed: NCCL p2p for tensors over NVLink, gloo for headers);
  * bounded staleness, backward-first scheduling, ``update_frequency`` / ``reduce_factor``
    thresholds, epoch-change propagation and the root/leaf data-ordering contract follow the
    reference;
  * ``solo`` nodes (1-stage replicas — unsupported upstream) train with the same API;
  * reference bugs that are NOT reproduced: ``prediction`` works (node.py:686 calls a missing
    method), saving keeps ``submod.pt`` restartable (node.py:722 deletes it), ``reduce_factor=None``
    no longer needs ``labels`` on every node (node.py:180-181), leaf republishes weights after
    averaging (node.py:621-624).
"""
import os
import shutil
import threading
import time
from typing import Any, Dict, List, Optional

import torch

from .communication import Communication, StagePlan, make_message
from .compute import Compute
from .runtime.mailbox import BACKWARD, FORWARD, Mailbox
from .strings import ActionTypes, NodeStatus, NodeTypes
from .utils.config import load_node_json_configs
from .utils.serialization import load_pickle, load_submodel, save_submodel


class Node:
    def __init__(self, name=None, model=None, optimizer=None, optimizer_params=None, lr_scheduler=None,
                 lr_scheduler_params=None, lr_step_on_epoch_change=True, criterion=None, update_frequency=1,
                 reduce_factor=None, labels=None, test_labels=None, device=torch.device('cpu'),
                 loss_filename='losses.txt', compression=False, average_optim=False, **kwargs):
        self.name = name
        self.base_dir = kwargs.get('base_dir') or os.environ.get('RAVNEST_HOME', '.')
        cfg = load_node_json_configs(node_name=name, base_dir=self.base_dir)
        self.config = cfg
        self.node_id = cfg['node_id']
        self.node_type = cfg['node_type']
        self.local_address = '{}:{}'.format(cfg['local_host'], cfg['local_port'])
        self.template_path = os.path.join(self.base_dir, cfg['template_path'])
        self.submod_file = cfg['submod_file']                 # 'submod_<k>'
        self.rank = cfg['rank']
        self.ring_size = cfg['ring_size']
        self.cluster_length = cfg['cluster_length']
        self.device = torch.device(device)
        if self.device.type == 'cuda' and self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self.compression = compression
        self.average_optim = average_optim
        self.update_frequency = max(1, int(update_frequency))
        self.lr_step_on_epoch_change = lr_step_on_epoch_change
        self.loss_filename = os.path.join(self.base_dir, loss_filename) if loss_filename and not os.path.isabs(loss_filename) else loss_filename
        self.val_accuracy_file = os.path.join(self.base_dir, 'val_accuracies.txt')
        self.labels_to_device = kwargs.get('labels_to_device', False)
        self.prefetch_h2d = bool(kwargs.get('prefetch_h2d', False))
        self.checkpoint_every = int(kwargs.get('checkpoint_every', 0) or 0)      # backward passes between automatic checkpoints
        self.inference_timeout = float(kwargs.get('inference_timeout', 120.0))
        self._stop_lock, self._stopping, self._stopped = threading.Lock(), False, threading.Event()
        self._h2d_stream = None
        self.status = NodeStatus.IDLE
        if kwargs.get('reset', True):
            self.reset()

        # ---- stage module + templates -------------------------------------------------------
        if model is None:
            model = load_submodel(os.path.join(self.template_path, 'submod.pt'))
        dtype = kwargs.get('dtype')
        model = model.to(device=self.device, dtype=dtype) if dtype is not None else model.to(self.device)
        channels_last = kwargs.get('channels_last')
        if channels_last is None:       # NHWC is the native layout of the conv kernels (and of cuDNN's tensor-core paths)
            channels_last = self.device.type == 'cuda' and any(isinstance(m, torch.nn.Conv2d) for m in model.modules())
        if channels_last:
            model = model.to(memory_format=torch.channels_last)
        self.channels_last = bool(channels_last)
        self.model = model
        is_root = self.node_type in (NodeTypes.ROOT, NodeTypes.SOLO)
        in_t = load_pickle(os.path.join(self.template_path, '{}_input.pkl'.format(self.submod_file)))
        out_t = load_pickle(os.path.join(self.template_path, '{}_output.pkl'.format(self.submod_file)))
        mi_path = os.path.join(self.template_path, 'model_inputs.pkl')
        mi_t = load_pickle(mi_path) if is_root and os.path.exists(mi_path) else None
        self.plan = StagePlan(self.submod_file, in_t, out_t, mi_t, cfg.get('model_input_names'))

        # ---- data / labels (leaf + solo own the criterion; reference node.py:144-163) ---------
        self.is_last = self.node_type in (NodeTypes.LEAF, NodeTypes.SOLO)
        self.criterion = criterion if self.is_last else None
        self.labels = labels if self.is_last else None
        if isinstance(self.labels, torch.Tensor):
            self.labels = _TensorLabels(self.labels, kwargs.get('label_batch_size', kwargs.get('batch_size', 64)))
        self.test_labels = test_labels if self.is_last else None
        self._labels_iter = None
        self._test_labels_iter = None
        self._labels_pos = 0
        n_labels = None
        if labels is not None:
            try:
                n_labels = len(labels)
            except TypeError:
                n_labels = None
        if reduce_factor is None and not n_labels:
            # a node without labels (the typical root) cannot derive "once per epoch": no periodic averaging on this node's own
            # clock — it still averages at the end of Trainer.train() — and, above all, no pipeline drain per update
            if cfg['ring_size'] > 1 and kwargs.get('dp_mode', 'average') == 'average':
                import warnings
                warnings.warn('{}: reduce_factor not given and no labels to derive it from; pass the same reduce_factor to every node '
                              'of the cluster to average periodically'.format(name))
            reduce_factor = 1 << 40
        elif reduce_factor is None:
            reduce_factor = n_labels
        self.reduce_factor = reduce_factor
        self.reduce_threshold = self.update_frequency * max(1, int(reduce_factor))

        # ---- transport ------------------------------------------------------------------------
        self.mailbox = Mailbox(capacity=kwargs.get('mailbox_capacity', max(2, self.cluster_length + 1)))
        transport = kwargs.get('transport')
        if transport is None:
            transport = 'dist' if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 'grpc'
        self.dp_group = None
        self.dp_mode = kwargs.get('dp_mode', 'average')       # 'average' (reference) | 'gradient'
        if transport == 'dist':
            from .runtime.dist_transport import DistTransport
            self.transport = DistTransport(self, **kwargs.get('transport_options', {}))
            self.dp_group = self.transport.dp_group
        elif transport == 'grpc':
            from .runtime.grpc_transport import GrpcTransport
            fwd = (cfg['forward_target_host'], cfg['forward_target_port']) if cfg.get('forward_target_host') is not None else None
            bwd = (cfg['backward_target_host'], cfg['backward_target_port']) if cfg.get('backward_target_host') is not None else None
            self.transport = GrpcTransport(name, cfg['local_host'], cfg['local_port'], self.mailbox, fwd, bwd,
                                           weights_provider=self._serve_latest_weights)
        else:
            self.transport = transport                          # injected (tests)

        # ---- compute engine -------------------------------------------------------------------
        recompute = kwargs.get('recompute')
        if recompute is None:
            recompute = self.device.type != 'cuda'              # HBM-resident activations on B200
        self.compute_session = Compute(
            model=self.model, optimizer_cls=optimizer, optimizer_params=optimizer_params or {},
            criterion=self.criterion, device=self.device, update_frequency=self.update_frequency,
            cluster_length=self.cluster_length, recompute=recompute, compute_dtype=dtype,
            loss_filename=self.loss_filename if self.is_last else None, is_last=self.is_last,
            dp_group=self.dp_group,
            fused_optimizer=kwargs.get('fused_optimizer'), lr_scheduler=lr_scheduler,
            lr_scheduler_params=lr_scheduler_params or {}, average_grads_over_dp=self.dp_mode == 'gradient',
            use_cuda_graphs=kwargs.get('use_cuda_graphs', self.device.type == 'cuda'), channels_last=self.channels_last)
        self.optimizer = self.compute_session.optimizer
        self.lr_scheduler = self.compute_session.lr_scheduler
        self.comm_session = Communication(
            name=name, compute=self.compute_session, transport=self.transport, rank=self.rank, ring_size=self.ring_size,
            ring_ids=cfg.get('ring_ids'), param_addresses=cfg.get('param_addresses'), average_optim=average_optim,
            dp_group=self.dp_group)

        # ---- scheduler state ------------------------------------------------------------------
        self.forward_pass_id = 0
        self.latest_backward_id = 0
        self.n_backwards = 0
        self.n_forwards = 0
        self._cv = threading.Condition()
        self._errors: List[BaseException] = []
        self._val_hits: List[Any] = []
        self._predictions: List[Any] = []
        self._prediction_cv = threading.Condition()
        self._saved = threading.Event()
        self._stop = threading.Event()
        self.metrics = {'forward': 0, 'backward': 0, 'updates': 0, 'reduces': 0, 'checkpoints': 0}
        self._resume_tag = kwargs.get('resume')           # e.g. 'auto': continue from this node's newest automatic checkpoint
        from .utils.tracing import StageTracer
        from .runtime.watchdog import FaultInjector
        self.tracer = StageTracer(name, self.device, enabled=kwargs.get('trace', True),
                                  path=os.path.join(self.base_dir, '{}_metrics.jsonl'.format(name)) if kwargs.get('metrics_file') else None)
        self.faults = FaultInjector(kwargs.get('fault_spec'))
        self.heartbeat = None
        self.compute_stream = torch.cuda.Stream(self.device) if self.device.type == 'cuda' else None
        if self.compute_stream is not None:
            self.compute_stream.wait_stream(torch.cuda.current_stream(self.device))   # init work is done
        self.comm_session.stream_ctx = self.on_compute_stream     # collectives called from the user's thread join the compute stream
        if self._resume_tag:
            self.resumed = self.restore(self._resume_tag)      # False when this node has no such checkpoint (fresh start)
        self.start()

    # ------------------------------------------------------------------ lifecycle
    def start(self) -> None:
        self.transport.start()
        self._worker = threading.Thread(target=self._scheduler_loop, name='{}-sched'.format(self.name), daemon=True)
        self._worker.start()

    def start_heartbeat(self, period: float = 1.0, on_failure=None):
        """Ping pipeline neighbours over the gRPC control plane (reference defines Ping but never uses it)."""
        from .runtime.watchdog import Heartbeat
        cfg = self.config
        peers = [(cfg[k + '_target_host'], cfg[k + '_target_port']) for k in ('forward', 'backward') if cfg.get(k + '_target_host') is not None]
        self.heartbeat = Heartbeat(self.transport, peers, period=period, on_failure=on_failure)
        self.heartbeat.start()
        return self.heartbeat

    def wait_until_stopped(self, timeout: Optional[float] = None) -> bool:
        """Block until this node has been stopped — by ``stop()`` here or by the SHUTDOWN message a stopping root sends down
        its pipeline.  Provider scripts of stem / leaf nodes end with this call."""
        if not self._stop.wait(timeout):
            return False
        return self._stopped.wait(30.0 if timeout is None else timeout)      # the stopping thread finished its hand-offs

    def stop(self) -> None:
        """Idempotent and safe to call from several threads (user code, the SHUTDOWN handler)."""
        with self._stop_lock:
            if self._stopping:
                return
            self._stopping = True
        if self.heartbeat is not None:
            self.heartbeat.stop()
        self.tracer.dump()
        self._stop.set()
        self.mailbox.close()
        try:
            self.compute_session.flush_losses()
        finally:
            try:
                self.transport.stop()          # also tells the next stage that this node is leaving
            finally:
                self._stopped.set()

    def _scheduler_loop(self) -> None:
        """Reference ``check_load_forward_buffer`` (node.py:330-367): backward lane first, then
        forward; the leaf remaps FORWARD->FIND_LOSS and NO_GRAD_FORWARD->output_type."""
        if self.device.type == 'cuda':
            torch.cuda.set_device(self.device)
            # dedicated compute stream: collectives / barriers issued by the user's main thread on the
            # legacy default stream can never stall stage compute (and vice versa)
            torch.cuda.set_stream(self.compute_stream)
        while not self._stop.is_set():
            t_wait = time.time()
            item = self.mailbox.get(timeout=0.25)
            self.tracer.idle(time.time() - t_wait)
            if item is None:
                continue
            _lane, msg = item
            if '_ready_event' in msg:
                from .runtime.dist_transport import DistTransport
                DistTransport.adopt(msg, self.device, self.tracer)
            action = msg['action']
            if self.faults.rules and self.faults.check(action, msg.get('forward_pass_id', -1)) == 'drop':
                continue
            if self.is_last and action == ActionTypes.FORWARD:
                action = ActionTypes.FIND_LOSS
            elif self.is_last and action == ActionTypes.NO_GRAD_FORWARD and self.node_type != NodeTypes.SOLO:
                action = msg.get('output_type') or ActionTypes.VAL_ACCURACY
            if action == ActionTypes.SHUTDOWN:
                threading.Thread(target=self.stop, daemon=True).start()
                break
            try:
                with self.tracer.phase(action, msg.get('forward_pass_id', -1)):
                    getattr(self, action)(msg)
            except BaseException as err:      # surface on the caller thread instead of dying silently
                with self._cv:
                    self._errors.append(err)
                    self._cv.notify_all()
                raise
            self.status = NodeStatus.IDLE

    def on_compute_stream(self):
        """Context for main-thread entry points that touch the master weights / version slots: they run on the scheduler's compute
        stream, i.e. ordered after every queued stage step and optimizer update (and before the next one)."""
        return torch.cuda.stream(self.compute_stream) if self.compute_stream is not None else _NullCtx()

    def _check(self) -> None:
        if self._errors:
            raise RuntimeError('{} scheduler failed'.format(self.name)) from self._errors[0]

    # ------------------------------------------------------------------ root-side API
    def forward_compute(self, tensors=None, **kwargs):
        """Submit one training mini-batch (root / solo).  Blocks only to keep the number of in-flight
        passes within ``cluster_length`` (bounded staleness, reference node.py:384-388)."""
        self._check()
        with self._cv:
            self._cv.wait_for(lambda: self.forward_pass_id - self.latest_backward_id <= self.cluster_length or self._errors)
        self._check()
        if (self.ring_size > 1 and self.dp_mode == 'average' and self.n_forwards
                and self.n_forwards % self.reduce_threshold == 0):
            self.wait_for_backwards()       # flush before a parameter-averaging point (reference node.py:387-388)
        args = self._pack_inputs(tensors, kwargs)
        msg = make_message(ActionTypes.ROOT_FORWARD, self.forward_pass_id, {}, args=args)
        if self.prefetch_h2d and self.device.type == 'cuda':
            self._stage_inputs_async(msg)
        self.forward_pass_id += 1
        self.n_forwards += 1
        self.mailbox.put(FORWARD, msg)

    def _stage_inputs_async(self, msg) -> None:
        """Opt-in (``prefetch_h2d=True``): start the host-to-device copy of this batch NOW, on a dedicated copy stream, from the
        submitting thread — it overlaps the compute of the passes already in flight instead of running at the head of this
        pass on the compute stream (ResNet-50 batch 100: 60 MB, ~1.1 ms over PCIe 5).  The scheduler waits on the event."""
        if self._h2d_stream is None:
            self._h2d_stream = torch.cuda.Stream(self.device)
        with torch.cuda.stream(self._h2d_stream):
            staged = [a.to(self.device, non_blocking=True) if isinstance(a, torch.Tensor) and a.device.type == 'cpu' else a
                      for a in msg['args']]
            ev = torch.cuda.Event()
            ev.record(self._h2d_stream)
        msg['args'] = staged
        msg['_ready_event'] = ev
        msg['_tensors'] = [t for t in staged if isinstance(t, torch.Tensor) and t.is_cuda]

    def _pack_inputs(self, tensors, kwargs) -> List[Any]:
        """Positional + keyword inputs -> one list in the MODEL's signature order (``plan.stage_args`` maps it onto the stage)."""
        if tensors is None:
            positional = []
        elif isinstance(tensors, (list, tuple)):
            positional = list(tensors)
        else:
            positional = [tensors]
        names = self.plan.input_names
        if kwargs:
            if not names:
                raise ValueError('keyword inputs need model_input_names in the plan')
            by_name = {}
            for k, v in kwargs.items():
                k = k[2:-1] if k.startswith('l_') and k.endswith('_') and k[2:-1] in names else k   # 'l_<name>_' accepted
                by_name[k] = v
            for n in names[len(positional):]:
                if n not in by_name:
                    raise ValueError('missing model input {!r}; expected {}'.format(n, names))
                positional.append(by_name[n])
        return positional

    def no_grad_forward_compute(self, tensors=None, output_type=None, **kwargs):
        """Inference pass through the pipeline (reference node.py:399-428).  ``output_type`` selects
        what the leaf does with the result: 'val_accuracy' | 'accuracy' | 'prediction'."""
        self._check()
        model_args = self._pack_inputs(tensors, kwargs)
        args = self.plan.stage_args(model_args)
        with self.on_compute_stream():       # same stream as the scheduler: ordered after every queued optimizer update
            if self.node_type == NodeTypes.SOLO:
                out = self.compute_session.no_grad_forward(args)
                return self._consume_inference(out, output_type)
            self.status = NodeStatus.FORWARD
            out = self.compute_session.no_grad_forward(args)
            data = self.plan.root_payload(out, model_args, self.compression)
            self.transport.send(FORWARD, make_message(ActionTypes.NO_GRAD_FORWARD, -1, data, output_type=output_type))
        self.status = NodeStatus.IDLE
        if output_type in (ActionTypes.ACCURACY, ActionTypes.PREDICTION):
            # the leaf sends its result back up the backward lane (the reference's Trainer.pred returns None, trainer.py:113)
            return self.pop_prediction(timeout=self.inference_timeout)
        return None

    def wait_for_backwards(self, timeout: Optional[float] = None) -> None:
        """Drain the pipeline: every submitted forward has been answered by its backward
        (reference node.py:702-710)."""
        with self._cv:
            ok = self._cv.wait_for(lambda: self.latest_backward_id >= self.forward_pass_id or self._errors, timeout=timeout)
        self._check()
        if not ok:
            raise TimeoutError('{}: backwards did not drain'.format(self.name))

    # ------------------------------------------------------------------ action handlers
    def root_forward(self, msg) -> None:
        self.status = NodeStatus.FORWARD
        fpid, model_args = msg['forward_pass_id'], msg['args']        # model signature order (see _pack_inputs)
        args = self.plan.stage_args(model_args)                         # this stage's placeholder order
        if self.node_type == NodeTypes.SOLO:
            targets = self._next_labels()
            self.compute_session.loss_step(args, targets[0], input_grads=False)
            self._after_backward(fpid, epoch_change=targets[1])
            return
        out = self.compute_session.forward(fpid, args, input_grads=False)
        prepared = self.compute_session.flights[fpid].inputs
        inputs = self.plan.model_order(prepared, model_args) if prepared else model_args
        data = self.plan.root_payload(out, inputs, self.compression)
        self.transport.send(FORWARD, make_message(ActionTypes.FORWARD, fpid, data))
        self.metrics['forward'] += 1

    def forward(self, msg) -> None:
        """Stem forward (reference node.py:469-490)."""
        self.status = NodeStatus.FORWARD
        fpid, data = msg['forward_pass_id'], msg['data']
        args = self.plan.gather_args(data)
        out = self.compute_session.forward(fpid, args)
        self.plan.add_outputs(data, out, self.compression)
        self.transport.send(FORWARD, make_message(ActionTypes.FORWARD, fpid, data))
        self.metrics['forward'] += 1

    def find_loss(self, msg) -> None:
        """Leaf: forward with grad, criterion, backward, send input grads upstream
        (reference node.py:575-629)."""
        self.status = NodeStatus.BACKWARD
        fpid, data = msg['forward_pass_id'], msg['data']
        args = self.plan.gather_args(data)
        targets, epoch_change = self._next_labels()
        _loss, in_grads = self.compute_session.loss_step(args, targets)
        grads = self.plan.add_input_grads({}, in_grads, self.compression)
        self.transport.send(BACKWARD, make_message(ActionTypes.BACKWARD, fpid, grads, epoch_change=epoch_change))
        self._after_backward(fpid, epoch_change=epoch_change, own_scheduler_done=True)

    def backward(self, msg) -> None:
        """Root / stem backward (reference node.py:511-573)."""
        self.status = NodeStatus.BACKWARD
        fpid, data = msg['forward_pass_id'], msg['data']
        out_grads = self.plan.take_output_grads(data)
        in_grads = self.compute_session.backward(fpid, out_grads)
        if self.node_type != NodeTypes.ROOT:
            grads = self.plan.add_input_grads(data, in_grads, self.compression)
            self.transport.send(BACKWARD, make_message(ActionTypes.BACKWARD, fpid, grads, epoch_change=msg.get('epoch_change', False)))
        self._after_backward(fpid, epoch_change=msg.get('epoch_change', False))

    def _after_backward(self, fpid: int, epoch_change: bool = False, own_scheduler_done: bool = False) -> None:
        cs = self.compute_session
        updated = cs.n_backwards % self.update_frequency == 0
        if cs.lr_scheduler is not None:
            if self.lr_step_on_epoch_change:
                if epoch_change:
                    cs.lr_scheduler.step()
            elif updated:
                cs.lr_scheduler.step()
        self.n_backwards += 1
        self.metrics['backward'] += 1
        if self.checkpoint_every and self.n_backwards % self.checkpoint_every == 0:
            # periodic sharded checkpoint (new): every node writes its own shard after the same backward pass,
            # so the set of 'auto' files of a cluster describes one consistent point of the asynchronous schedule per stage
            self.checkpoint('auto')
            self.metrics['checkpoints'] += 1
        if self.ring_size > 1 and self.dp_mode == 'average' and self.n_backwards % self.reduce_threshold == 0:
            self.status = NodeStatus.REDUCING
            self.comm_session.parallel_ring_reduce()
            self.metrics['reduces'] += 1
        with self._cv:
            self.latest_backward_id = fpid + 1
            self._cv.notify_all()

    # ------------------------------------------------------------------ labels (leaf / solo)
    def _next_labels(self):
        """Next item of ``labels`` plus an epoch-change flag (reference node.py:579-587).  ``labels``
        may be a DataLoader / iterable (restarted on exhaustion) or a tensor indexed batch-wise."""
        epoch_change = False
        if self._labels_iter is None:
            self._labels_iter = iter(self.labels)
        try:
            item = next(self._labels_iter)
        except StopIteration:
            self._labels_iter = iter(self.labels)
            item = next(self._labels_iter)
            epoch_change = True
        if self.labels_to_device:
            item = _to_device(item, self.device)
        return item, epoch_change

    def _next_test_labels(self):
        if self._test_labels_iter is None:
            self._test_labels_iter = iter(self.test_labels)
        try:
            return next(self._test_labels_iter), False
        except StopIteration:
            self._test_labels_iter = iter(self.test_labels)
            return next(self._test_labels_iter), True

    # ------------------------------------------------------------------ inference actions
    def no_grad_forward(self, msg) -> None:
        """Stem relay of an inference pass (reference node.py:492-509)."""
        data = msg['data']
        args = self.plan.gather_args(data)
        out = self.compute_session.no_grad_forward(args)
        self.plan.add_outputs(data, out, self.compression)
        self.transport.send(FORWARD, make_message(ActionTypes.NO_GRAD_FORWARD, -1, data, output_type=msg.get('output_type')))

    def _leaf_inference(self, msg):
        args = self.plan.gather_args(msg['data'])
        return self.compute_session.no_grad_forward(args)

    def _consume_inference(self, out, output_type):
        if output_type == ActionTypes.VAL_ACCURACY:
            return self._score_validation(out)
        if output_type == ActionTypes.ACCURACY:
            return self._score_accuracy(out)
        return out

    def _score_validation(self, out):
        """argmax vs next(test_labels); epoch mean appended to val_accuracies.txt (node.py:631-667)."""
        targets, wrapped = self._next_test_labels() if self.test_labels is not None else ((None, None), False)
        y = targets[1] if isinstance(targets, (tuple, list)) else targets
        logits = out[0] if isinstance(out, (tuple, list)) else out
        pred = logits.argmax(dim=-1).to('cpu')
        y = y.argmax(dim=-1) if (isinstance(y, torch.Tensor) and y.dim() == pred.dim() + 1) else y
        acc = (pred == y.to('cpu')).float().mean().item()
        self._val_hits.append(acc)
        n_batches = len(self.test_labels) if hasattr(self.test_labels, '__len__') else None
        if n_batches is not None and len(self._val_hits) >= n_batches:
            mean = sum(self._val_hits) / len(self._val_hits)
            self._val_hits = []
            with open(self.val_accuracy_file, 'a') as fh:
                fh.write('{}\n'.format(mean))
            self.last_val_accuracy = mean
        return acc

    def _score_accuracy(self, out):
        logits = out[0] if isinstance(out, (tuple, list)) else out
        return logits.argmax(dim=-1).to('cpu')

    def val_accuracy(self, msg) -> None:
        self._score_validation(self._leaf_inference(msg))

    def accuracy(self, msg) -> None:
        self._publish_inference(self._score_accuracy(self._leaf_inference(msg)))

    def prediction(self, msg) -> None:
        """Raw outputs of an inference pass (the reference's handler is broken, node.py:683-690)."""
        self._publish_inference(_to_device(self._leaf_inference(msg), 'cpu'))

    def _publish_inference(self, result) -> None:
        """Leaf: keep the result (``pop_prediction``) and send it to the root, which is where ``Trainer.pred`` waits for it."""
        with self._prediction_cv:
            self._predictions.append(result)
            self._prediction_cv.notify_all()
        if self.node_type == NodeTypes.LEAF:
            self.transport.send(BACKWARD, make_message(ActionTypes.INFERENCE_RESULT, -1, {'result': result}))

    def inference_result(self, msg) -> None:
        """Backward-lane relay of an inference result: stems pass it on, the root hands it to the waiting caller."""
        if self.node_type == NodeTypes.ROOT:
            with self._prediction_cv:
                self._predictions.append(msg['data']['result'])
                self._prediction_cv.notify_all()
        else:
            self.transport.send(BACKWARD, make_message(ActionTypes.INFERENCE_RESULT, -1, msg['data']))

    def pop_prediction(self, timeout: Optional[float] = None):
        with self._prediction_cv:
            if not self._prediction_cv.wait_for(lambda: self._predictions, timeout=timeout):
                raise TimeoutError('no prediction arrived')
            return self._predictions.pop(0)

    # ------------------------------------------------------------------ save / restore / sync
    def _save_local(self) -> str:
        """Write ``<template_path>/<submod_k>.pt`` (reference node.py:692-724) and keep ``submod.pt``
        current so the node can restart."""
        cs = self.compute_session
        cs.flush_losses()
        path = os.path.join(self.template_path, '{}.pt'.format(self.submod_file))
        with cs.lock, self.on_compute_stream():
            cs.sync_master()
            # save the MASTER weights (fp32 even when the stage computes in bf16): point the module's parameters at copies of
            # them for the duration of the pickle, then back at the live compute-dtype version (stage steps use their own
            # parameter dicts, never the module's attributes, so nothing running is disturbed)
            lookup = dict(self.model.named_parameters())
            with torch.no_grad():
                for name, master in cs.store.master_state().items():
                    lookup[name].data = master.detach().clone()
            try:
                save_submodel(self.model, path)        # tensors keep their device tag; load_submodel maps them to the CPU
            finally:
                cs.store.bind_module()
        shutil.copyfile(path, os.path.join(self.template_path, 'submod.pt'))
        self._saved.set()
        return path

    def save_submodel(self, msg) -> None:
        self._save_local()
        if self.node_type not in (NodeTypes.LEAF, NodeTypes.SOLO):
            self.transport.send(FORWARD, make_message(ActionTypes.SAVE_SUBMODEL, -1, {}))

    def trigger_save_submodel(self) -> None:
        self.wait_for_backwards()
        self._save_local()
        if self.node_type == NodeTypes.ROOT:
            self.transport.send(FORWARD, make_message(ActionTypes.SAVE_SUBMODEL, -1, {}))
            if hasattr(self.transport, 'flush'):
                self.transport.flush(timeout=60)

    def _serve_latest_weights(self, first, last):
        return self.compute_session.latest_weights(first, last)

    def update_with_latest_weights(self, leader: int = 0) -> None:
        """(Re)join: adopt the newest parameters of this stage (reference node.py:726-730).

        gRPC transport: stream them from the peer(s) of the next cluster (``get_latest_weights`` RPC).
        dist transport on CUDA: K2 — the DP-group leader publishes its weight bucket in symmetric memory and the
        next forward's GEMMs read their weight tiles straight from the leader's HBM over NVLink, landing them
        locally on the way (``parallel/broadcast.py``).  Collective: every replica of the stage must call it."""
        cs = self.compute_session
        if self.dp_group is not None and self.device.type == 'cuda' and cs.store.compute_dtype == torch.bfloat16:
            from .parallel.broadcast import WeightBroadcaster
            self.wait_for_backwards() if self.node_type in (NodeTypes.ROOT, NodeTypes.SOLO) else None
            if getattr(self, '_broadcaster', None) is None:
                self._broadcaster = WeightBroadcaster(cs.store, self.dp_group, leader)
            with cs.lock, self.on_compute_stream():
                self._broadcaster.publish()
                self._broadcaster.arm()
                # the fp32 master travels exactly (NCCL, behind the armed bf16 pull): the follower's next optimizer step starts
                # from the leader's master, not from a bf16-rounded copy of it
                torch.distributed.broadcast(cs.store.master, src=torch.distributed.get_global_rank(self.dp_group, leader),
                                            group=self.dp_group)
                self._pending_pull = True
                # the pull completes by itself: the next stage step runs eagerly (the armed GEMMs land the weights), then
                # whatever was not consumed is copied and the pull is closed — no second user call
                cs.run_next_step_eagerly = True
                cs.after_next_step = self.finish_weight_pull
            return
        if self.dp_group is not None:
            with cs.lock, self.on_compute_stream():
                torch.distributed.broadcast(cs.store.master, src=torch.distributed.get_global_rank(self.dp_group, leader),
                                            group=self.dp_group)
                cs.store.publish()
            return
        weights = self.comm_session.get_latest_weights()
        self.compute_session.load_weights(weights)

    def finish_weight_pull(self) -> None:
        """Complete a K2 pull after the first forward has consumed (and landed) the GEMM weights.  Called automatically after the
        first stage step that follows ``update_with_latest_weights``; calling it by hand is harmless."""
        if getattr(self, '_pending_pull', False):
            self._pending_pull = False
            with self.compute_session.lock, self.on_compute_stream():
                self._broadcaster.finish()

    def checkpoint(self, tag: str = 'latest') -> str:
        from .utils.checkpoint import save_checkpoint
        with self.on_compute_stream():
            return save_checkpoint(self, tag)

    def restore(self, tag: str = 'latest') -> bool:
        from .utils.checkpoint import load_checkpoint
        with self.on_compute_stream():
            return load_checkpoint(self, tag)

    def reset(self) -> None:
        """Clear run outputs of a previous job (reference node.py:732-746)."""
        for path in (self.loss_filename, getattr(self, 'val_accuracy_file', None)):
            if path and os.path.exists(path) and self.config['node_type'] in (NodeTypes.LEAF, NodeTypes.SOLO):
                os.remove(path)
        aux = os.path.join(self.base_dir, '{}_aux'.format(self.name))
        if os.path.isdir(aux):
            shutil.rmtree(aux)


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _TensorLabels:
    """Plain-tensor labels (examples/sorter/provider.py:24): yield ``(None, labels[i:i+bs])`` so a
    criterion written as ``targets[1]`` works for loaders and tensors alike."""

    def __init__(self, tensor, batch_size):
        self.tensor, self.batch_size = tensor, int(batch_size)

    def __len__(self):
        return (len(self.tensor) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for i in range(0, len(self.tensor), self.batch_size):
            yield None, self.tensor[i:i + self.batch_size]


def _to_device(obj, device):
    if isinstance(obj, torch.Tensor):
        return obj.to(device, non_blocking=True)
    if isinstance(obj, dict):
        return {k: _to_device(v, device) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_device(v, device) for v in obj)
    if hasattr(obj, 'to'):
        try:
            return obj.to(device)
        except Exception:
            return obj
    return obj
