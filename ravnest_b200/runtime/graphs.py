"""CUDA-graph capture of stage steps.

The reference's inner loop is Python + one library kernel per op; at BERT-base / batch 8x512 a B200
finishes each kernel faster than the CPU can launch the next one (measured: ~1160 launches and 20 ms
per step eager, device time == wall time).  Capturing the whole forward + criterion + backward of a
stage into ONE graph removes the launch bound: per step the host does an input copy and a replay.

``GraphedLossStep`` covers leaf / solo stages (forward WITH grad, criterion, backward into the flat
gradient bucket).  Warm-up runs are REAL training steps executed eagerly; the capture itself executes
nothing, so the step being captured is performed by the first replay.
"""
from typing import Any, Callable, List, Optional, Sequence

import torch


def _tree_map(fn: Callable, obj: Any) -> Any:
    if isinstance(obj, torch.Tensor):
        return fn(obj)
    if isinstance(obj, dict):
        return {k: _tree_map(fn, v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_tree_map(fn, v) for v in obj)
    return obj


def _tree_tensors(obj: Any, out: List[torch.Tensor]) -> List[torch.Tensor]:
    if isinstance(obj, torch.Tensor):
        out.append(obj)
    elif isinstance(obj, dict):
        for v in obj.values():
            _tree_tensors(v, out)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _tree_tensors(v, out)
    return out


def _signature(args: Sequence[Any], targets: Any):
    sig = []
    for t in _tree_tensors([list(args), targets], []):
        sig.append((tuple(t.shape), 'float' if t.is_floating_point() else t.dtype))
    return tuple(sig)


class GraphedLossStep:
    def __init__(self, compute, args: Sequence[Any], targets: Any, input_grads: bool):
        self.compute = compute
        self.device = compute.device
        self.signature = _signature(args, targets)
        self.input_grads = input_grads
        dev = self.device

        def stage(t):
            return torch.empty(t.shape, dtype=t.dtype, device=dev).copy_(t, non_blocking=True)

        self.static_args = [stage(a) if isinstance(a, torch.Tensor) else a for a in args]
        for a in self.static_args:
            if isinstance(a, torch.Tensor) and a.is_floating_point() and input_grads:
                a.requires_grad_(True)
        self.static_targets = _tree_map(stage, targets)
        self._arg_tensors = [a for a in self.static_args if isinstance(a, torch.Tensor)]
        self._target_tensors = _tree_tensors(self.static_targets, [])
        self.graph = torch.cuda.CUDAGraph()
        stream = torch.cuda.current_stream(dev)
        cap = torch.cuda.Stream(dev)
        cap.wait_stream(stream)
        compute.model.train()
        slot = compute.store.current
        with torch.cuda.graph(self.graph, stream=cap, capture_error_mode='thread_local'):
            with torch.enable_grad():
                out = compute._call(slot, self.static_args)
                loss = compute.criterion(out, self.static_targets)
            loss.backward()
            self.static_loss = loss.detach()
        stream.wait_stream(cap)
        self.slot = slot
        self.static_in_grads = [a.grad if isinstance(a, torch.Tensor) and a.requires_grad else None for a in self.static_args]

    def matches(self, args, targets, slot) -> bool:
        return slot is self.slot and _signature(args, targets) == self.signature

    def run(self, args: Sequence[Any], targets: Any):
        src_args = [a for a in args if isinstance(a, torch.Tensor)]
        for dst, src in zip(self._arg_tensors, src_args):
            dst.detach().copy_(src, non_blocking=True)
        for dst, src in zip(self._target_tensors, _tree_tensors(targets, [])):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        grads = [g.clone() if g is not None else None for g in self.static_in_grads]
        return self.static_loss.clone(), grads


class GraphedStage:
    """Forward and backward of a root/stem stage as two CUDA graphs sharing one memory pool.

    One instance == one in-flight pass: it owns a private copy of the weights (that copy IS the stashed
    version — PipeDream weight stashing becomes "don't refresh the copy until the pass has done its
    backward"), static input/output buffers and the saved activations inside the graph pool.  A stage keeps
    ``cluster_length + 1`` of them, so HBM holds every in-flight pass (resident-activation mode).
    """

    def __init__(self, compute, args: Sequence[Any], input_grads: bool):
        self.compute = compute
        dev = compute.device
        store = compute.store
        self.signature = _signature(args, None)
        self.busy = False
        self.weights_version = -1
        # private weight copy + leaves that accumulate into the stage's flat gradient bucket
        self.wbuf = torch.empty_like(store.current.flat)
        self.params = {}
        for k, name in enumerate(store.names):
            self.params[name] = store.make_leaf(self.wbuf, k)

        def stage(t):
            return torch.empty(t.shape, dtype=t.dtype, device=dev).copy_(t, non_blocking=True)

        self.static_args = [stage(a) if isinstance(a, torch.Tensor) else a for a in args]
        for a in self.static_args:
            if isinstance(a, torch.Tensor) and a.is_floating_point() and input_grads:
                a.requires_grad_(True)
        self._arg_tensors = [a for a in self.static_args if isinstance(a, torch.Tensor)]
        self.refresh_weights()
        stream = torch.cuda.current_stream(dev)
        cap = torch.cuda.Stream(dev)
        cap.wait_stream(stream)
        compute.model.train()
        from torch.func import functional_call
        self.fwd_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.fwd_graph, stream=cap, capture_error_mode='thread_local'):
            with torch.enable_grad():
                out = functional_call(compute.model, self.params, tuple(self.static_args))
        self.out_tree = out
        self.static_outs = _tree_tensors(out, [])
        self.grad_outs = [torch.zeros_like(o) if o.requires_grad else None for o in self.static_outs]
        roots = [o for o in self.static_outs if o.requires_grad]
        seeds = [g for g in self.grad_outs if g is not None]
        self.bwd_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.bwd_graph, stream=cap, pool=self.fwd_graph.pool(), capture_error_mode='thread_local'):
            if roots:
                torch.autograd.backward(roots, seeds)
        stream.wait_stream(cap)
        self.static_in_grads = [a.grad if isinstance(a, torch.Tensor) and a.requires_grad else None for a in self.static_args]

    def refresh_weights(self) -> None:
        store = self.compute.store
        if self.weights_version != store.version:
            self.wbuf.copy_(store.current.flat, non_blocking=True)
            self.weights_version = store.version

    def run_forward(self, args: Sequence[Any]):
        self.busy = True
        self.refresh_weights()
        for dst, src in zip(self._arg_tensors, [a for a in args if isinstance(a, torch.Tensor)]):
            dst.detach().copy_(src, non_blocking=True)
        self.fwd_graph.replay()
        return _tree_map(lambda t: t.detach().clone(), self.out_tree)

    def run_backward(self, grads) -> List[Optional[torch.Tensor]]:
        for idx, buf in enumerate(self.grad_outs):
            if buf is None:
                continue
            g = grads.get(idx)
            if g is None:
                buf.zero_()
            else:
                buf.copy_(g, non_blocking=True)
        self.bwd_graph.replay()
        out = [g.clone() if g is not None else None for g in self.static_in_grads]
        self.busy = False
        return out
