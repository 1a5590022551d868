"""2-GPU worker for K2: a follower runs its first forward with weights TMA-loaded from the leader's HBM and must end
with (a) the same outputs as the leader and (b) the leader's weights in its local bucket.  Also times
broadcast-fused-GEMM vs broadcast-then-GEMM."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ravnest_b200 import nn as rnn  # noqa: E402
from ravnest_b200.ops import native  # noqa: E402
from ravnest_b200.parallel.broadcast import WeightBroadcaster  # noqa: E402
from ravnest_b200.runtime.dist_transport import init_distributed  # noqa: E402
from ravnest_b200.runtime.param_store import ParamStore  # noqa: E402


class Stack(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = rnn.Linear(768, 3072, activation='gelu')
        self.b = rnn.Linear(3072, 768)
        self.ln = rnn.LayerNorm(768)

    def forward(self, x):
        return self.ln(self.b(self.a(x)))


def main():
    rank, world, dev = init_distributed()
    group = dist.new_group(list(range(world)))
    torch.manual_seed(1234 + rank)                      # different weights on every rank
    mod = Stack().to(dev, torch.bfloat16)
    store = ParamStore(mod, dev, n_slots=2, compute_dtype=torch.bfloat16)
    bc = WeightBroadcaster(store, group, leader_rank_in_group=0)
    torch.manual_seed(7)
    x = torch.randn(4096, 768, device=dev).bfloat16()
    from torch.func import functional_call
    bc.publish()
    armed = bc.arm()
    with torch.no_grad():
        y = functional_call(mod, store.current.param_dict, (x,))
    bc.finish()
    torch.cuda.synchronize()
    # every rank now holds rank 0's weights and produced rank 0's output
    ref_flat = store.current.flat.clone()
    dist.broadcast(ref_flat, src=0, group=group)
    ys = y.clone()
    dist.broadcast(ys, src=0, group=group)
    assert torch.equal(store.current.flat, ref_flat), 'weights differ from the leader after the fused pull'
    assert float((y.float() - ys.float()).abs().max()) < 1e-2, 'outputs differ from the leader'
    assert rank == 0 or (armed == 2 and native.calls['gemm_fwd_bcast'] == 2)
    # ---- timing: fused (B over NVLink inside the GEMM) vs copy-then-GEMM, follower side, device timed
    w_local = store.current.leaves[0]                              # a.weight [3072, 768]
    peer = bc.sym.peer_tensor(0)[store.offsets[0]:store.offsets[0] + store.numels[0]].view(3072, 768)
    C = native.C()

    def fused():
        flags, epoch = native.bcast_flags(w_local.device)
        return C.gemm(x, False, peer, False, None, 0, None, False, 0, 1, None, w_local.detach(), flags, epoch)

    def unfused():
        w_local.detach().copy_(peer)
        return C.gemm(x, False, w_local.detach(), False, None, 0, None, False, 0)

    def local_only():
        return C.gemm(x, False, w_local.detach(), False, None, 0, None, False, 0)

    res = {}
    for name, fn in (('fused_pull_gemm', fused), ('copy_then_gemm', unfused), ('local_gemm', local_only)):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            dist.barrier(group=group)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        t = torch.tensor([sorted(ts)[len(ts) // 2]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        res[name + '_ms'] = float(t)
    if rank == 0:
        nbytes = 3072 * 768 * 2
        res['weight_bytes'] = nbytes
        res['nvlink_floor_ms'] = nbytes / 770e9 * 1e3
        print('BCAST_GPU_OK', json.dumps(res))
        os.makedirs('gpurun_out', exist_ok=True)
        json.dump(res, open('gpurun_out/bcast_fused.json', 'w'), indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
