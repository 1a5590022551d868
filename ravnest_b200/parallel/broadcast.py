"""K2 — parameter broadcast from the stage leader fused with the first GEMM that consumes the parameters.

Reference: ``update_with_latest_weights`` -> ``get_latest_weights`` RPC streaming the whole pickled state dict over
gRPC (ravnest/node.py:726-730, communication.py:279-330, endpoints.py:145-154) before any compute can start.

Here the leader of a stage's DP group publishes its current flat weight bucket into symmetric memory.  A follower
does NOT copy it first: it *arms* the pull, and the next forward pass runs every Linear's GEMM with the B operand
(the weight) TMA-loaded straight from the leader's HBM over NVLink; CTAs of the first M block write each weight tile
through to the follower's local bucket while the tensor cores consume it (``rvn_gemm_bf16_tcgen05``, ``b_copy``).
After that one forward the follower's weights are local again.  Non-GEMM parameters (LayerNorm, biases, embeddings —
a few hundred KB..MB) are fetched with ordinary peer copies at arm time.
"""
from typing import Dict, Optional

import torch
import torch.distributed as dist

from ..ops import native
from .symmetric import SymmetricBuffer


class WeightBroadcaster:
    def __init__(self, store, dp_group, leader_rank_in_group: int = 0):
        self.store = store
        self.group = dp_group
        self.leader = leader_rank_in_group
        self.sym = SymmetricBuffer(store.total, store.compute_dtype, store.device, dp_group, zero=False)
        self.rank = self.sym.rank
        self.armed = 0

    def publish(self) -> None:
        """Everybody calls this at the sync point; the leader's current weights become visible to its peers."""
        if self.rank == self.leader:
            self.sym.tensor.copy_(self.store.current.flat)
        torch.cuda.current_stream(self.store.device).synchronize()
        dist.barrier(group=self.group)

    def arm(self, slot=None) -> int:
        """Follower: route the next forward's Linear GEMMs through the leader's copy.  Returns the number of armed weights."""
        if self.rank == self.leader:
            return 0
        st = self.store
        slot = slot or st.current
        peer_flat = self.sym.peer_tensor(self.leader)
        n = 0
        with torch.no_grad():
            for k, (name, shape, leaf) in enumerate(zip(st.names, st.shapes, slot.leaves)):
                src = st.view_of(peer_flat, k)
                if len(shape) == 2 and shape[1] % 8 == 0 and st.compute_dtype == torch.bfloat16 and not name.endswith('embeddings.weight'):
                    native.arm_peer_weight(leaf, src)          # consumed (and landed locally) by the first GEMM that uses it
                    n += 1
                else:
                    leaf.detach().copy_(src)
        self.armed = n
        return n

    def finish(self) -> None:
        """Anything still armed (weights the forward never touched) is copied the plain way.  The fp32 master is NOT rebuilt from
        the bf16 weights (that would round it): ``Node.update_with_latest_weights`` broadcasts it exactly.  The version counter
        moves so pass-private weight copies (``GraphedStage.wbuf``) refresh."""
        for leaf, src in native.drain_peer_weights():
            leaf.detach().copy_(src)
        self.store.version += 1
        self.store.current.version = self.store.version
