"""Fused flat-buffer optimizers (K1): DP all-reduce over peer memory + update + weight publish.

``FusedFlatOptimizer`` IS a ``torch.optim.Optimizer`` (one param group over the stage's fp32 master
parameters) so user lr-schedulers work unchanged (reference node.py:213-215), but ``step()`` is one
launch of ``rvn_fused_allreduce_update`` (SGD / Adam / AdamW) or two of ``rvn_fused_allreduce_lamb1`` +
``rvn_fused_lamb2`` (LAMB) over the flat buckets of ``runtime.param_store.ParamStore``.

With a DP group the gradient bucket lives in symmetric memory; every replica's kernel reads all
replicas' gradients straight over NVLink (one-shot all-reduce: (P-1)/P of the bucket crosses the
link once, no reduce-scatter/all-gather round trips, no separate scale / cast / copy kernels) and a
device-side flag barrier replaces NCCL.  ``average_parameters`` is the same trick for the reference's
periodic parameter averaging (``parallel_ring_reduce``).
"""
import math
import os
from typing import Optional

import torch
import torch.distributed as dist

from ..ops import _ext

_CHUNK = 16384          # elements per LAMB chunk (never straddles a parameter)


class FusedFlatOptimizer(torch.optim.Optimizer):
    def __init__(self, store, kind: str, params: dict, dp_group=None, average_grads: bool = False,
                 sync_grads: Optional[bool] = None):
        """``dp_group``: the stage's replicas.  ``sync_grads`` (default: ``average_grads``): all-reduce the gradient bucket inside
        ``step()`` (gradient-mode DP).  With ``sync_grads=False`` the group is only used by ``average_parameters()`` — the reference's
        periodic parameter averaging — and ``step()`` is a purely local update."""
        self.kind = kind
        self.store = store
        p = dict(params)
        if kind == 'sgd':
            defaults = dict(lr=p.pop('lr', 1e-3), momentum=p.pop('momentum', 0.0), dampening=p.pop('dampening', 0.0),
                            weight_decay=p.pop('weight_decay', 0.0), nesterov=p.pop('nesterov', False))
        elif kind in ('adam', 'adamw'):
            defaults = dict(lr=p.pop('lr', 1e-3), betas=p.pop('betas', (0.9, 0.999)), eps=p.pop('eps', 1e-8),
                            weight_decay=p.pop('weight_decay', 0.01 if kind == 'adamw' else 0.0))
        elif kind == 'lamb':
            defaults = dict(lr=p.pop('lr', 1e-3), betas=p.pop('betas', (0.9, 0.999)), eps=p.pop('eps', 1e-6),
                            weight_decay=p.pop('weight_decay', 0.0), clamp_value=p.pop('clamp_value', 10.0),
                            adam=p.pop('adam', False), debias=p.pop('debias', False))
        else:
            raise ValueError(kind)
        for k in ('maximize', 'differentiable', 'amsgrad'):
            if p.pop(k, False):          # different numerics: refuse instead of silently ignoring (the caller falls back to torch's optimizer)
                raise TypeError('fused {} does not implement {}=True'.format(kind, k))
        for k in ('foreach', 'fused', 'capturable'):
            p.pop(k, None)
        if p:
            raise TypeError('unsupported optimizer arguments for fused {}: {}'.format(kind, sorted(p)))
        dev = store.device
        n = store.total
        self.steps = 0
        self.epoch = 0
        self.dp_group = dp_group
        self.average_grads = average_grads
        self.sync_grads = bool(average_grads if sync_grads is None else sync_grads)
        self.grad_sym = self.sig_sym = self.avg_sym = self.land_sym = self.norm_sym = self.master_sym = None
        self.two_shot = False
        self._master_stale = False
        self.rank, self.world = 0, 1
        if dp_group is not None and dist.get_world_size(dp_group) > 1:
            from .symmetric import SymmetricBuffer
            self.sig_sym = SymmetricBuffer(64, torch.int32, dev, dp_group)
            self.rank, self.world = self.sig_sym.rank, self.sig_sym.world
            if self.sync_grads:
                self.grad_sym = SymmetricBuffer(n, store.compute_dtype, dev, dp_group)
                store.rebind_grad(self.grad_sym.tensor)
                # K1 v2, two-shot: each rank reduces / updates only its shard and pushes the new weights to everybody's landing
                # buffer — (N-1)/N of the bucket per direction instead of (N-1) x bucket inbound, and 1/N of the optimizer's HBM
                # traffic.  Default from 3 replicas up (at 2 the link traffic is the same and one-shot has one barrier less);
                # RAVNEST_K1_TWOSHOT=0/1 overrides.  The fp32 master lives in symmetric memory so that any rank can refresh the
                # shards it does not own with one-sided peer copies (``sync_master``: checkpoints, saves, parameter averaging).
                want = os.environ.get('RAVNEST_K1_TWOSHOT')
                self.two_shot = (self.world >= 3) if want is None else (want == '1')
                if self.two_shot:
                    self.land_sym = SymmetricBuffer(n, store.compute_dtype, dev, dp_group)
                    self.master_sym = SymmetricBuffer(n, torch.float32, dev, dp_group, zero=False)
                    store.rebind_master(self.master_sym.tensor)
                    dist.barrier(group=dp_group)
                    per = (n // self.world + 63) // 64 * 64
                    self.shards = [(min(n, r * per), min(n, (r + 1) * per)) for r in range(self.world)]
                    self.shard = self.shards[self.rank]
        super().__init__(store.master_params, defaults)
        need_m = kind != 'sgd' or defaults['momentum'] != 0
        need_v = kind != 'sgd'
        self.m_sym = self.v_sym = None
        if self.two_shot:                  # sharded moments are peer-readable too (state_dict gathers them one-sidedly)
            from .symmetric import SymmetricBuffer
            self.m_sym = SymmetricBuffer(n, torch.float32, dev, dp_group) if need_m else None
            self.v_sym = SymmetricBuffer(n, torch.float32, dev, dp_group) if need_v else None
        self.exp_avg = (self.m_sym.tensor if self.m_sym is not None else torch.zeros(n, dtype=torch.float32, device=dev)) if need_m else torch.zeros(0, device=dev)
        self.exp_avg_sq = (self.v_sym.tensor if self.v_sym is not None else torch.zeros(n, dtype=torch.float32, device=dev)) if need_v else torch.zeros(0, device=dev)
        self.done_counter = torch.zeros(4, dtype=torch.int32, device=dev)
        if kind == 'lamb':
            self._build_chunks()
            if self.two_shot:
                # shards follow chunk boundaries (chunks never straddle a parameter): contiguous chunk ranges of ~equal element count
                from .symmetric import SymmetricBuffer
                ends = self.chunk_end.cpu().tolist()
                n_chunks = len(ends)
                cut = [0]
                for r in range(1, self.world):
                    target = n * r // self.world
                    cut.append(next((i for i, e in enumerate(ends) if e >= target), n_chunks - 1) + 1)
                cut.append(n_chunks)
                cut = [min(max(c, 0), n_chunks) for c in cut]
                for i in range(1, len(cut)):
                    cut[i] = max(cut[i], cut[i - 1])
                self.chunk_range = (cut[self.rank], cut[self.rank + 1])
                starts = self.chunk_start.cpu().tolist()
                edge = [starts[c] if c < n_chunks else n for c in cut]
                edge[0] = 0
                self.shards = [(edge[r], edge[r + 1]) for r in range(self.world)]
                self.shard = self.shards[self.rank]
                self.norm_sym = SymmetricBuffer(2 * len(store.numels), torch.float32, dev, dp_group)
                self.norms = self.norm_sym.tensor
                if any(self.shards[r][1] <= self.shards[r][0] or cut[r + 1] <= cut[r] for r in range(self.world)):
                    self.two_shot = False          # fewer chunks than replicas (toy models): every rank takes the same decision
        self._C = _ext.load()

    # ------------------------------------------------------------------ helpers
    def _build_chunks(self):
        st = self.store
        cp, cs, ce = [], [], []
        for i, (off, n) in enumerate(zip(st.offsets, st.numels)):
            for s in range(off, off + n, _CHUNK):
                cp.append(i)
                cs.append(s)
                ce.append(min(off + n, s + _CHUNK))
        dev = st.device
        self.chunk_param = torch.tensor(cp, dtype=torch.int32, device=dev)
        self.chunk_start = torch.tensor(cs, dtype=torch.int64, device=dev)
        self.chunk_end = torch.tensor(ce, dtype=torch.int64, device=dev)
        self.norms = torch.zeros(2 * len(st.numels), dtype=torch.float32, device=dev)

    def _peers(self):
        if self.grad_sym is None:
            return [self.store.grad.data_ptr()], [], 0
        return self.grad_sym.ptrs, self.sig_sym.ptrs, self.rank

    def _args(self, group):
        kind = self.kind
        t = self.steps
        lr = float(group['lr'])
        b1, b2 = group.get('betas', (0.0, 0.0))
        c1 = c2 = 1.0
        if kind in ('adam', 'adamw'):
            c1, c2 = 1.0 - b1 ** t, math.sqrt(1.0 - b2 ** t)
        if kind == 'lamb' and group['debias']:
            lr = lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
        scale = (1.0 / self.world) if (self.average_grads and self.grad_sym is not None and self.world > 1) else 1.0
        f = [lr, b1, b2, group.get('eps', 0.0), group['weight_decay'], group.get('momentum', 0.0), group.get('dampening', 0.0),
             scale, c1, c2, group.get('clamp_value', 0.0)]
        flags = [int(group.get('nesterov', False)), int(t == 1), int(kind == 'adamw'), int(group.get('adam', False))]
        return f, flags

    # ------------------------------------------------------------------ API
    @torch.no_grad()
    def step(self, closure=None):
        self.steps += 1
        self.epoch += 1
        group = self.param_groups[0]
        f, flags = self._args(group)
        grads, sigs, rank = self._peers()
        st = self.store
        bf16_grad = st.grad.dtype == torch.bfloat16

        def writer(slot):
            if self.two_shot:
                out_bf16 = slot.flat.dtype == torch.bfloat16
                if self.kind == 'lamb':
                    c0, c1 = self.chunk_range
                    self._C.fused_twoshot_lamb(grads, sigs, self.land_sym.ptrs, self.norm_sym.ptrs, rank, f, flags, st.master, self.exp_avg,
                                               self.exp_avg_sq, bf16_grad, out_bf16, self.chunk_param, self.chunk_start, self.chunk_end,
                                               c0, c1 - c0, self.norms, self.epoch, self.done_counter)
                else:
                    self._C.fused_twoshot_update(0 if self.kind == 'sgd' else 1, grads, sigs, self.land_sym.ptrs, rank, f, flags, st.master,
                                                 self.exp_avg, self.exp_avg_sq, bf16_grad, out_bf16, self.shard[0], self.shard[1],
                                                 self.epoch, self.done_counter)
                # every rank's shard has landed (the kernel's exit barrier): adopt the gathered weights as the new version.  The
                # fp32 master outside the own shard is refreshed from them lazily (``sync_master``) — nothing on the step path reads it
                slot.flat.copy_(self.land_sym.tensor)
                self._master_stale = True
                return
            if self.kind == 'lamb':
                self._C.fused_lamb(grads, sigs, rank, f, flags, st.master, self.exp_avg, self.exp_avg_sq, slot.flat, bf16_grad,
                                   self.chunk_param, self.chunk_start, self.chunk_end, self.norms, self.epoch, self.done_counter)
            else:
                self._C.fused_update(0 if self.kind == 'sgd' else 1, grads, sigs, rank, f, flags, st.master, self.exp_avg,
                                     self.exp_avg_sq, slot.flat, bf16_grad, self.epoch, self.done_counter)

        st.publish(writer)
        st.zero_grad()

    @torch.no_grad()
    def sync_master(self):
        """Two-shot mode keeps the fp32 master (and moments) sharded: a rank only updates its own shard.  Refresh the shards owned
        by the peers with one-sided copies out of their symmetric master (race-free between two steps: the step kernel's exit
        barrier means every peer has finished its update, and nobody starts the next one before this rank arrives at the entry
        barrier).  Call on the stage's compute stream before anything reads the WHOLE master: checkpoint, save, averaging."""
        if not (self.two_shot and self._master_stale):
            return
        for r, (s0, e0) in enumerate(self.shards):
            if r != self.rank and e0 > s0:
                self.store.master[s0:e0].copy_(self.master_sym.peer_tensor(r)[s0:e0])
                for sym, buf in ((self.m_sym, self.exp_avg), (self.v_sym, self.exp_avg_sq)):
                    if sym is not None:
                        buf[s0:e0].copy_(sym.peer_tensor(r)[s0:e0])
        self._master_stale = False

    @torch.no_grad()
    def average_parameters(self, include_state: bool = False):
        """Parameter averaging across the DP group in one kernel (reference parallel_ring_reduce)."""
        if self.world <= 1:
            return
        self.sync_master()
        from .symmetric import SymmetricBuffer
        st = self.store
        if self.avg_sym is None:
            self.avg_sym = SymmetricBuffer(st.total, torch.float32, st.device, self.dp_group)
        bufs = [st.master] + ([self.exp_avg, self.exp_avg_sq] if include_state else [])
        for i, buf in enumerate(bufs):
            if buf.numel() != st.total:
                continue
            self.avg_sym.tensor.copy_(buf)
            self.epoch += 1

            def writer(slot, buf=buf):
                self._C.fused_param_average(self.avg_sym.ptrs, self.sig_sym.ptrs, self.rank, buf, slot.flat, self.epoch,
                                            self.done_counter)
            if i == 0:
                st.publish(writer)
            else:
                scratch = torch.empty_like(st.current.flat)
                self._C.fused_param_average(self.avg_sym.ptrs, self.sig_sym.ptrs, self.rank, buf, scratch, self.epoch,
                                            self.done_counter)

    # state_dict / load_state_dict keep the flat moments (checkpoint/resume)
    def state_dict(self):
        self.sync_master()
        sd = super().state_dict()
        sd['flat'] = {'exp_avg': self.exp_avg.cpu(), 'exp_avg_sq': self.exp_avg_sq.cpu(), 'steps': self.steps}
        return sd

    def load_state_dict(self, sd):
        sd = dict(sd)                       # never mutate the caller's dict
        flat = sd.pop('flat', None)
        super().load_state_dict(sd)
        if flat is not None:
            self.exp_avg.copy_(flat['exp_avg'].to(self.exp_avg.device))
            self.exp_avg_sq.copy_(flat['exp_avg_sq'].to(self.exp_avg_sq.device))
            self.steps = flat['steps']
