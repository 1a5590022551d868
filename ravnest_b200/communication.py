"""Payload plans and cross-cluster collectives.

Parity: reference ``ravnest/communication.py`` — payload builders ``create_forward_payload :98-123``,
``create_backward_payload :79-96``, ``create_no_grad_forward_payload :333-348``; collectives
``parallel_ring_reduce :125-158``, ``single_ring_reduce :160-277``; ``get_latest_weights :279-290``;
fp16 "compression" ``utils.py:184-194``.

What changed:
  * payload routing is computed ONCE from the templates (``StagePlan``) instead of being re-derived
    from nested dicts on every message;
  * the ring all-reduce exists twice: (a) the wire-compatible per-parameter ring over the gRPC
    endpoints (reduce-scatter + all-gather, chunks split along each parameter's largest axis, the
    reference algorithm) for CPU / off-box nodes, and (b) the GPU path — one flat-bucket all-reduce
    per stage over NVLink (fused P2P kernel, or NCCL as the baseline tier);
  * iterations are tagged, so there is no ``reduce_iteration`` polling (the RPC still answers).
"""
import threading
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from .strings import ActionTypes

FP16_MAX = 65504.0


# ----------------------------------------------------------------------------------------------
# compression (reference utils.py:184-194)
# ----------------------------------------------------------------------------------------------
def compress_tensor(t: torch.Tensor) -> torch.Tensor:
    if t.dtype == torch.float32:
        if t.is_cuda:
            from . import ops
            if ops.backend() == 'native':
                from .ops import native
                return native.pack_for_wire(t, 'fp16')          # clamp + cast in one pass (K12)
        return t.clamp(-FP16_MAX, FP16_MAX).to(torch.float16)
    if t.dtype == torch.float64:
        return t.to(torch.float32)
    return t


def decompress_tensor(t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    return t.to(dtype) if t.is_floating_point() and t.dtype != dtype else t


# ----------------------------------------------------------------------------------------------
# static routing plan for one stage
# ----------------------------------------------------------------------------------------------
class StagePlan:
    """Routing derived from ``submod_<k>_input.pkl`` / ``_output.pkl`` / ``model_inputs.pkl``."""

    def __init__(self, name: str, input_template: Dict[int, Dict[str, Any]], output_template: Dict[int, Dict[str, Any]],
                 model_inputs_template: Optional[Dict[int, Dict[str, Any]]] = None, input_names: Optional[List[str]] = None):
        self.name = name
        self.inputs: List[Tuple[str, int]] = []
        for pos in sorted(input_template):
            (producer, ref), = input_template[pos].items()
            idx = 0 if ref == 'placeholder:tensor' else int(ref)
            self.inputs.append((producer, idx))
        self.output_targets: Dict[int, List[str]] = {int(k): list(v.get('target', [])) for k, v in output_template.items()}
        self.model_inputs_targets = None
        if model_inputs_template is not None:
            self.model_inputs_targets = {int(k): list(v.get('target', [])) for k, v in model_inputs_template.items()}
        self.input_names = list(input_names or [])

    # ---- forward ---------------------------------------------------------------------------
    def stage_args(self, model_inputs: Sequence[Any]) -> List[Any]:
        """Root / solo stage: the model inputs (in the MODEL's signature order, which is how ``Node.forward_compute`` packs
        positional and keyword inputs) re-ordered into this stage's placeholder order.  The splitter numbers a stage's
        placeholders by first use, so e.g. BERT's root takes (input_ids, token_type_ids, attention_mask) while the model
        signature is (input_ids, attention_mask, token_type_ids)."""
        if not all(producer == 'model_inputs' for producer, _ in self.inputs):
            raise ValueError('{} is not fed by model inputs only'.format(self.name))
        return [model_inputs[idx] for _, idx in self.inputs]

    def model_order(self, stage_inputs: Sequence[Any], model_inputs: Sequence[Any]) -> List[Any]:
        """Inverse of ``stage_args`` for relaying: model-order list that prefers the stage's already prepared (device) tensors."""
        merged = list(model_inputs)
        for (producer, idx), value in zip(self.inputs, stage_inputs):
            if producer == 'model_inputs' and idx < len(merged) and value is not None:
                merged[idx] = value
        return merged

    def gather_args(self, data: Dict[str, Dict[int, Dict[str, Any]]], consume: bool = True) -> List[Any]:
        """Pull this stage's positional args out of the relay dict; strike our name off each target
        list and drop entries nobody downstream needs (reference compute.py:329-407)."""
        args = []
        for producer, idx in self.inputs:
            entry = data[producer][idx]
            args.append(entry['data'])
            if consume and self.name in entry['target']:
                entry['target'] = [t for t in entry['target'] if t != self.name]
        if consume:
            for producer in list(data):
                for idx in list(data[producer]):
                    if not data[producer][idx]['target']:
                        del data[producer][idx]
                if not data[producer]:
                    del data[producer]
        return args

    def add_outputs(self, data: Dict, outputs: Any, compression: bool = False) -> Dict:
        outs = [outputs] if isinstance(outputs, torch.Tensor) else list(outputs)
        mine = {}
        for idx, targets in self.output_targets.items():
            if not targets:
                continue
            t = outs[idx].detach()
            mine[idx] = {'data': compress_tensor(t) if compression else t, 'target': list(targets)}
        if mine:
            data[self.name] = mine
        return data

    def root_payload(self, outputs: Any, model_inputs: Sequence[Any], compression: bool = False) -> Dict:
        """Forward payload of the root: its outputs plus raw model inputs later stages consume
        (reference communication.py:118-122)."""
        data: Dict[str, Dict[int, Dict[str, Any]]] = {}
        relay = {}
        for idx, targets in (self.model_inputs_targets or {}).items():
            rest = [t for t in targets if t != self.name]
            if rest and idx < len(model_inputs):
                relay[idx] = {'data': model_inputs[idx], 'target': rest}
        if relay:
            data['model_inputs'] = relay
        return self.add_outputs(data, outputs, compression)

    # ---- backward --------------------------------------------------------------------------
    def take_output_grads(self, data: Dict[str, Dict[int, torch.Tensor]]) -> Dict[int, torch.Tensor]:
        return data.pop(self.name, {})

    def add_input_grads(self, data: Dict[str, Dict[int, torch.Tensor]], in_grads: Sequence[Optional[torch.Tensor]],
                        compression: bool = False) -> Dict:
        """Merge this stage's input gradients into the relay dict, summing when several consumers
        of the same tensor contribute (reference node.py:535-540)."""
        for (producer, idx), g in zip(self.inputs, in_grads):
            if g is None or producer == 'model_inputs':
                continue
            g = compress_tensor(g) if compression else g
            slot = data.setdefault(producer, {})
            if idx in slot:
                prev = slot[idx]
                slot[idx] = prev.to(g.device) + g if prev.dtype == g.dtype else prev.to(g.device, g.dtype) + g
            else:
                slot[idx] = g
        return data


# ----------------------------------------------------------------------------------------------
# collectives
# ----------------------------------------------------------------------------------------------
def split_along_largest_axis(t: torch.Tensor, n: int) -> List[torch.Tensor]:
    """``tensor_split`` along the largest dimension (reference utils.py:157-165)."""
    if t.dim() == 0:
        return [t.reshape(1)] + [t.new_zeros(0) for _ in range(n - 1)]
    axis = max(range(t.dim()), key=lambda d: t.shape[d])
    return list(torch.tensor_split(t, n, dim=axis))


def _cat_like(chunks: List[torch.Tensor], like: torch.Tensor) -> torch.Tensor:
    if like.dim() == 0:
        return chunks[0].reshape(())
    axis = max(range(like.dim()), key=lambda d: like.shape[d])
    return torch.cat(chunks, dim=axis)


class Communication:
    """Cross-cluster (data-parallel) collectives for one node."""

    def __init__(self, name: str, compute, transport, rank: int = 0, ring_size: int = 1,
                 ring_ids: Optional[Dict[int, str]] = None, param_addresses: Optional[Dict[str, str]] = None,
                 average_optim: bool = False, dp_group=None, timeout: float = 600.0):
        self.name = name
        self.compute = compute
        self.transport = transport
        self.rank = int(rank)
        self.ring_size = int(ring_size)
        self.average_optim = average_optim
        self.dp_group = dp_group
        self.timeout = timeout
        names = compute.store.names if compute is not None else []
        self.ring_param_keys = self._ranges(names, {int(k): v for k, v in (ring_ids or {}).items() if v is not None})
        self.param_address_mapping: Dict[str, Tuple[str, int]] = {}
        for addr, keys in self._ranges(names, dict((a, k) for a, k in (param_addresses or {}).items())).items():
            host, port = addr.rsplit(':', 1)
            for k in keys:
                self.param_address_mapping[k] = (host, int(port))
        self.n_reduces = 0
        self.stream_ctx = None        # set by Node: context that enters the stage's compute stream

    @staticmethod
    def _ranges(names: List[str], firsts: Dict[Any, str]) -> Dict[Any, List[str]]:
        """{key: first_param_name} -> {key: [names from that parameter up to the next key's first]}."""
        marks = sorted(((names.index(first), key) for key, first in firsts.items() if first in names))
        out: Dict[Any, List[str]] = {}
        for i, (start, key) in enumerate(marks):
            stop = marks[i + 1][0] if i + 1 < len(marks) else len(names)
            out[key] = names[start:stop]
        return out

    # ---- public API (names kept from the reference) -----------------------------------------
    def parallel_ring_reduce(self) -> None:
        """Average parameters (and optionally optimizer state) with the same stage of every other
        cluster.  No-op for a single cluster (reference communication.py:126)."""
        if self.ring_size <= 1 or self.compute is None:
            return
        import contextlib
        ctx = self.stream_ctx() if self.stream_ctx is not None else contextlib.nullcontext()
        with self.compute.lock, ctx:
            if self.dp_group is not None:
                self._flat_average()
            else:
                threads = [threading.Thread(target=self._guard, args=(rid, keys)) for rid, keys in self.ring_param_keys.items()]
                self._errors: List[BaseException] = []
                for t in threads:
                    t.start()
                for t in threads:
                    t.join()
                if self._errors:
                    raise self._errors[0]
                self.compute.store.publish()
        self.n_reduces += 1

    def _guard(self, ring_id, keys):
        try:
            self.single_ring_reduce(ring_id, keys)
        except BaseException as err:  # surfaced by parallel_ring_reduce
            self._errors.append(err)

    # ---- GPU / torch.distributed path --------------------------------------------------------
    def _flat_average(self) -> None:
        import torch.distributed as dist
        store = self.compute.store
        fused = getattr(self.compute, '_fused', None)
        world = dist.get_world_size(self.dp_group)
        if fused is not None and hasattr(fused, 'average_parameters') and fused.world == world:
            fused.average_parameters(include_state=self.average_optim)
            return
        dist.all_reduce(store.master, group=self.dp_group)
        store.master.div_(world)
        if self.average_optim and self.compute.optimizer is not None:
            for state in self.compute.optimizer.state.values():
                for v in state.values():
                    if isinstance(v, torch.Tensor) and v.is_floating_point() and v.numel() > 1:
                        dist.all_reduce(v, group=self.dp_group)
                        v.div_(world)
        store.publish()

    # ---- gRPC ring (reference algorithm) -----------------------------------------------------
    def _optim_state_entries(self, keys: Sequence[str]) -> Dict[str, torch.Tensor]:
        entries = {}
        opt = self.compute.optimizer
        if not self.average_optim or opt is None:
            return entries
        store = self.compute.store
        for key in keys:
            p = store.master_params[store.index_of(key)]
            for sname, sval in opt.state.get(p, {}).items():
                if isinstance(sval, torch.Tensor) and sval.is_floating_point() and sval.shape == p.shape:
                    entries['optim::{}::{}'.format(sname, key)] = sval
        return entries

    def single_ring_reduce(self, ring_id: int, keys: Sequence[str]) -> None:
        n, r = self.ring_size, self.rank
        store = self.compute.store
        rings = self.transport.rings
        tensors: Dict[str, torch.Tensor] = {k: store.master_views[store.index_of(k)] for k in keys}
        tensors.update(self._optim_state_entries(keys))
        owner = {name: self.param_address_mapping[name.split('::')[-1]] for name in tensors}
        chunks = {name: [c.detach().to('cpu', copy=True) for c in split_along_largest_axis(t, n)] for name, t in tensors.items()}
        names = list(tensors)

        def exchange(phase: str, it: int, send_pos: int) -> Dict[str, torch.Tensor]:
            by_dest: Dict[Tuple[str, int], Dict[str, torch.Tensor]] = {}
            for name in names:
                by_dest.setdefault(owner[name], {})[name] = chunks[name][send_pos]
            senders = [threading.Thread(target=self.transport.send_ring_chunk, args=(phase, dest, ring_id, it, payload))
                       for dest, payload in by_dest.items()]
            for t in senders:
                t.start()
            got = rings.collect(phase, ring_id, it, names, timeout=self.timeout)
            for t in senders:
                t.join()
            rings.advance(phase, ring_id)
            return got

        with torch.no_grad():
            for it in range(n - 1):                       # reduce-scatter
                send_pos = (r - it) % n
                recv_pos = (r - it - 1) % n
                got = exchange('reduce', it, send_pos)
                for name in names:
                    chunks[name][recv_pos] = chunks[name][recv_pos] + got[name]
            for it in range(n - 1):                       # all-gather
                send_pos = (r - it + 1) % n
                recv_pos = (r - it) % n
                got = exchange('gather', it, send_pos)
                for name in names:
                    chunks[name][recv_pos] = got[name]
            for name, t in tensors.items():
                t.copy_(_cat_like(chunks[name], t).div_(n).to(t.device, t.dtype))
        rings.reset(ring_id)

    # ---- latest weights (reference communication.py:279-330) ---------------------------------
    def get_latest_weights(self) -> Dict[str, torch.Tensor]:
        """Pull the newest copy of every parameter this node owns from the peer(s) of the next
        cluster that hold it; used to (re)join a running job."""
        store = self.compute.store
        by_dest: Dict[Tuple[str, int], List[str]] = {}
        for name in store.names:
            if name in self.param_address_mapping:
                by_dest.setdefault(self.param_address_mapping[name], []).append(name)
        merged: Dict[str, torch.Tensor] = {}
        for dest, keys in by_dest.items():
            merged.update(self.transport.fetch_latest_weights(dest, keys[0], keys[-1]))
        return merged


def make_message(action: str, fpid: int, data: Dict, **extra) -> Dict[str, Any]:
    msg = {'action': action, 'forward_pass_id': fpid, 'data': data}
    msg.update(extra)
    return msg


__all__ = ['Communication', 'StagePlan', 'compress_tensor', 'decompress_tensor', 'make_message',
           'split_along_largest_axis', 'ActionTypes']
