"""Flat, versioned parameter storage for one pipeline stage.

The reference implements PipeDream-style weight stashing by cloning the full ``state_dict`` after
every backward and swapping ``param.data`` around the recompute (``ravnest/compute.py:47-51``,
``:191-199``, ``:218-221``, ``:261``, ``:530-540``; ``utils.py:119-125``).  Here a stage owns

  * ``master``  : one flat fp32 buffer (the optimizer's copy; reference "clone params", node.py:207)
  * ``slots``   : a ring of flat buffers in the compute dtype, one per live weight version
  * ``grad``    : one flat gradient buffer every version's leaves accumulate into

so stashing a version is "keep using slot i", publishing a new version is one fused cast/copy into
a free slot (or the fused all-reduce+optimizer kernel writing it directly), and the flat buffers
are exactly the buckets the NVLink collectives operate on.  Offsets are padded to 128 bytes so the
vectorised kernels (16 B / thread, multimem v4) never straddle parameters.
"""
from collections import OrderedDict
from typing import Dict, List, Optional

import torch

_ALIGN_BYTES = 128


class Slot:
    __slots__ = ('index', 'flat', 'leaves', 'param_dict', 'version', 'refs')

    def __init__(self, index, flat, leaves, param_dict):
        self.index = index
        self.flat = flat
        self.leaves = leaves            # List[Tensor] leaf views (requires_grad) in layout order
        self.param_dict = param_dict    # name -> leaf, for functional_call
        self.version = -1
        self.refs = 0


class ParamStore:
    def __init__(self, module: torch.nn.Module, device, n_slots: int = 3, compute_dtype: Optional[torch.dtype] = None,
                 master_dtype: torch.dtype = torch.float32):
        self.module = module
        self.device = torch.device(device)
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        self.names: List[str] = [n for n, _ in named]
        self.shapes = [tuple(p.shape) for _, p in named]
        self.numels = [p.numel() for _, p in named]
        self.compute_dtype = compute_dtype or (named[0][1].dtype if named else torch.float32)
        self.master_dtype = master_dtype
        align = 64      # elements: 128 B for bf16, 256 B for fp32
        self.offsets: List[int] = []
        off = 0
        for n in self.numels:
            self.offsets.append(off)
            off += (n + align - 1) // align * align
        self.total = max(off, align)
        self.master = torch.zeros(self.total, dtype=master_dtype, device=self.device)
        self.grad = torch.zeros(self.total, dtype=self.compute_dtype, device=self.device)
        # spatial convolution weights live in OHWI (channels_last) order inside every flat buffer: the implicit-GEMM kernels read
        # them K-major without a per-step permute copy and their OHWI weight gradients add into the bucket as a dense stream
        self.channels_last = [len(s) == 4 and s[2] * s[3] > 1 for s in self.shapes]
        self.master_views = self.views_of(self.master)
        self.grad_views = self.views_of(self.grad)
        self.wgrad_acc: Optional[torch.Tensor] = None      # fp32 accumulator of the convolution weight gradients (see below)
        with torch.no_grad():
            for view, (_, p) in zip(self.master_views, named):
                view.copy_(p.detach().to(self.device, master_dtype))
        self.slots: List[Slot] = []
        for i in range(max(2, n_slots)):
            self.slots.append(self._make_slot(i))
        self.version = 0
        self.current: Slot = self.slots[0]
        self._publish_into(self.current)
        self.current.version = 0
        # optimizer-facing master parameters (leaf tensors sharing the flat master storage)
        self.master_params = [torch.nn.Parameter(v, requires_grad=True) for v in self.master_views]
        self._master_grad: Optional[torch.Tensor] = None
        self.bind_module()

    # ------------------------------------------------------------------ construction helpers
    def view_of(self, flat: torch.Tensor, k: int) -> torch.Tensor:
        """Parameter ``k`` as a (possibly channels_last-strided) view of a flat buffer laid out like this store."""
        o, n, s = self.offsets[k], self.numels[k], self.shapes[k]
        seg = flat[o:o + n]
        if self.channels_last[k]:
            return seg.view(s[0], s[2], s[3], s[1]).permute(0, 3, 1, 2)
        return seg.view(s)

    def views_of(self, flat: torch.Tensor) -> List[torch.Tensor]:
        return [self.view_of(flat, k) for k in range(len(self.shapes))]

    def make_leaf(self, flat: torch.Tensor, k: int) -> torch.Tensor:
        """Parameter ``k`` of a weight buffer laid out like this store, as an autograd leaf that accumulates into the stage's flat
        gradient bucket (and, for convolution weights, into the fp32 staging accumulator when that is enabled)."""
        leaf = self.view_of(flat, k).detach().requires_grad_(True)
        leaf.grad = self.grad_views[k]            # every version accumulates into the one flat gradient buffer
        if self.wgrad_acc is not None and len(self.shapes[k]) == 4:
            leaf._rvn_wgrad_acc = self.wgrad_acc[self.offsets[k]:self.offsets[k] + self.numels[k]]
        return leaf

    def _make_slot(self, index: int) -> Slot:
        flat = torch.zeros(self.total, dtype=self.compute_dtype, device=self.device)
        leaves, pdict = [], OrderedDict()
        for k, name in enumerate(self.names):
            leaf = self.make_leaf(flat, k)
            leaves.append(leaf)
            pdict[name] = leaf
        return Slot(index, flat, leaves, pdict)

    # ------------------------------------------------------------------ fp32 staging of convolution weight gradients
    def enable_wgrad_staging(self) -> bool:
        """Give every 4-D (convolution) weight a slice of ONE persistent fp32 accumulator laid out like the gradient bucket.  The
        split-K weight-gradient kernels add their partial tiles straight into it (bulk reductions at the L2) across all micro-steps
        of an update; ``flush_wgrad`` moves the sums into the gradient bucket once per optimizer step.  Replaces, per convolution
        and micro-step, a zero-fill of an fp32 buffer and a cast-add into the bucket (ResNet-50: 2 x 52 launches, 0.4 ms per step),
        and keeps the accumulation over micro-steps in fp32.  CUDA + native ops only; returns whether it is on."""
        if self.wgrad_acc is not None:
            return True
        if self.device.type != 'cuda' or not any(len(s) == 4 for s in self.shapes):
            return False
        self.wgrad_acc = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        for slot in self.slots:
            for k, leaf in enumerate(slot.leaves):
                if len(self.shapes[k]) == 4:
                    leaf._rvn_wgrad_acc = self.wgrad_acc[self.offsets[k]:self.offsets[k] + self.numels[k]]
        return True

    def flush_wgrad(self) -> None:
        """grad += accumulator; accumulator = 0.  Must run before anything consumes the gradient bucket."""
        if self.wgrad_acc is None:
            return
        from ..ops import native
        native.wgrad_flush(self.wgrad_acc, self.grad)

    def _publish_into(self, slot: Slot) -> None:
        with torch.no_grad():
            slot.flat.copy_(self.master)        # fused cast fp32 -> compute dtype (one kernel)

    # ------------------------------------------------------------------ version ring
    def acquire(self) -> Slot:
        """Pin the current version for an in-flight forward pass."""
        self.current.refs += 1
        return self.current

    def release(self, slot: Slot) -> None:
        slot.refs -= 1
        assert slot.refs >= 0

    def free_slot(self) -> Slot:
        for s in self.slots:
            if s is not self.current and s.refs == 0:
                return s
        s = self._make_slot(len(self.slots))    # ring too small for the staleness in flight: grow it
        self.slots.append(s)
        return s

    def publish(self, writer=None) -> Slot:
        """Make a new weight version current.  ``writer(slot)`` may fill the slot itself (fused
        kernels write the compute-dtype weights directly); default is cast-copy from master."""
        slot = self.free_slot() if self.current.refs > 0 else self.current
        if writer is None:
            self._publish_into(slot)
        else:
            writer(slot)
        self.version += 1
        slot.version = self.version
        self.current = slot
        return slot

    def live_versions(self) -> Dict[int, int]:
        return {s.version: s.refs for s in self.slots if s.refs > 0 or s is self.current}

    # ------------------------------------------------------------------ gradients / master sync
    def master_grad(self) -> torch.Tensor:
        """Flat gradient in master dtype (alias of ``grad`` when dtypes match)."""
        if self.compute_dtype == self.master_dtype:
            return self.grad
        if self._master_grad is None:
            self._master_grad = torch.zeros(self.total, dtype=self.master_dtype, device=self.device)
        self._master_grad.copy_(self.grad)
        return self._master_grad

    def attach_master_grads(self) -> None:
        flat = self.master_grad()
        for k, p in enumerate(self.master_params):
            p.grad = self.view_of(flat, k)

    def rebind_grad(self, new_flat: torch.Tensor) -> None:
        """Move the gradient bucket into externally allocated (e.g. symmetric / peer-mapped) memory."""
        assert new_flat.numel() == self.total and new_flat.dtype == self.grad.dtype
        new_flat.copy_(self.grad)
        self.grad = new_flat
        self.grad_views = self.views_of(self.grad)
        for slot in self.slots:
            for leaf, g in zip(slot.leaves, self.grad_views):
                leaf.grad = g
        self._master_grad = None

    def rebind_master(self, new_flat: torch.Tensor) -> None:
        """Move the fp32 master into externally allocated (symmetric / peer-mapped) memory.  Must happen before an optimizer is
        built over ``master_params``."""
        assert new_flat.numel() == self.total and new_flat.dtype == self.master.dtype
        with torch.no_grad():
            new_flat.copy_(self.master)
        self.master = new_flat
        self.master_views = self.views_of(self.master)
        self.master_params = [torch.nn.Parameter(v, requires_grad=True) for v in self.master_views]

    def zero_grad(self, discard_staged: bool = False) -> None:
        """Zero the gradient bucket.  The fp32 staging accumulator is zero at rest (``flush_wgrad`` leaves it so, and every
        optimizer step flushes first); ``discard_staged`` also drops convolution gradients staged since the last flush."""
        self.grad.zero_()
        if discard_staged and self.wgrad_acc is not None:
            self.wgrad_acc.zero_()

    # ------------------------------------------------------------------ module <-> store
    def bind_module(self) -> None:
        """Point the module's own Parameters at the current version (state_dict / save / eval)."""
        lookup = dict(self.module.named_parameters())
        for name, leaf in self.current.param_dict.items():
            lookup[name].data = leaf.detach()

    def load_master_from(self, tensors: Dict[str, torch.Tensor]) -> None:
        with torch.no_grad():
            for name, view in zip(self.names, self.master_views):
                if name in tensors:
                    view.copy_(tensors[name].to(self.device, self.master_dtype))

    def master_state(self) -> "OrderedDict[str, torch.Tensor]":
        return OrderedDict((n, v) for n, v in zip(self.names, self.master_views))

    def index_of(self, name: str) -> int:
        return self.names.index(name)

    def range_of(self, first: str, last: str) -> slice:
        """Flat element range covering parameters ``first..last`` inclusive (DP ring ranges)."""
        i, j = self.names.index(first), self.names.index(last)
        return slice(self.offsets[i], self.offsets[j] + self.numels[j])
