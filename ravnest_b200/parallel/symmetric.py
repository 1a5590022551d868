"""Symmetric (peer-mapped) device buffers for in-kernel collectives.

``torch.distributed._symmetric_memory`` is used purely as plumbing: it allocates with cuMem, exchanges
the handles between the ranks of a group and maps every peer's buffer (and, on NVSwitch, a multicast
object) into this process.  The kernels in ``ops/csrc/fused_optim.cu`` then load/store those peer
addresses directly.
"""
from typing import List, Optional

import torch
import torch.distributed as dist


def symmetric_available() -> bool:
    try:
        import torch.distributed._symmetric_memory  # noqa: F401
        return torch.cuda.is_available()
    except Exception:
        return False


class SymmetricBuffer:
    """A flat tensor whose storage is mapped into every rank of ``group``."""

    def __init__(self, numel: int, dtype: torch.dtype, device: torch.device, group, zero: bool = True):
        import torch.distributed._symmetric_memory as symm
        self.group = group
        self.tensor = symm.empty(numel, dtype=dtype, device=device)
        if zero:
            self.tensor.zero_()
        torch.cuda.synchronize(device)
        self.handle = symm.rendezvous(self.tensor, group)
        self.rank = int(self.handle.rank)
        self.world = int(self.handle.world_size)
        self.ptrs: List[int] = [int(p) for p in self.handle.buffer_ptrs]
        try:
            self.multicast_ptr: Optional[int] = int(self.handle.multicast_ptr) or None
        except Exception:
            self.multicast_ptr = None
        dist.barrier(group=group)       # nobody touches a peer mapping before everybody has zeroed theirs

    def peer_tensor(self, rank: int) -> torch.Tensor:
        return self.handle.get_buffer(rank, self.tensor.shape, self.tensor.dtype)
