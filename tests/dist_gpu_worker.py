"""2+ GPU worker: fused P2P all-reduce + optimizer vs NCCL all_reduce + torch optimizer (launched by torchrun)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import ravnest_b200 as rv  # noqa: E402
from ravnest_b200.optim import build_optimizer  # noqa: E402
from ravnest_b200.runtime.dist_transport import init_distributed  # noqa: E402
from ravnest_b200.runtime.param_store import ParamStore  # noqa: E402


class Mod(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Parameter(torch.randn(1024, 777))
        self.b = torch.nn.Parameter(torch.randn(4099))
        self.c = torch.nn.Parameter(torch.randn(512, 512, 3))


def main():
    rank, world, dev = init_distributed()
    group = dist.new_group(list(range(world)))
    worst = 0.0
    for kind, cls, params in (('sgd', torch.optim.SGD, dict(lr=0.05, momentum=0.9, weight_decay=5e-4)),
                              ('adam', torch.optim.Adam, dict(lr=1e-2)),
                              ('lamb', rv.Lamb, dict(lr=1e-2, weight_decay=0.01, debias=True, clamp_value=10000.0))):
        for dtype in (torch.bfloat16, torch.float32):
            torch.manual_seed(0)
            mod, ref = Mod().to(dev), Mod().to(dev)
            ref.load_state_dict(mod.state_dict())
            store = ParamStore(mod.to(dtype), dev, n_slots=3, compute_dtype=dtype)
            opt, fused = build_optimizer(cls, dict(params), store, dp_group=group, fused=True, average_grads=True)
            ropt = cls(ref.parameters(), **params)
            for step in range(3):
                torch.manual_seed(100 * step + rank)
                for p, view in zip(ref.parameters(), store.grad_views):
                    g = (torch.randn_like(p) * 0.1).to(dtype)
                    view.copy_(g)
                    gsum = g.float().clone()
                    dist.all_reduce(gsum, group=group)
                    p.grad = gsum / world
                fused.step()
                ropt.step()
            torch.cuda.synchronize()
            # the published compute-dtype weights are complete on every rank (two-shot: gathered by the kernel itself) ...
            for p, leaf in zip(ref.parameters(), store.current.leaves):
                err = float((leaf.detach().float() - p).norm() / p.norm())
                assert err < (6e-3 if dtype == torch.bfloat16 else 1e-4), ('published', kind, dtype, err, getattr(fused, 'two_shot', None))
            # ... and so is the fp32 master once the shards of the peers have been pulled (no-op in one-shot mode)
            fused.sync_master()
            torch.cuda.synchronize()
            for k, (p, view) in enumerate(zip(ref.parameters(), store.master_views)):
                err = float((view - p).norm() / p.norm())
                worst = max(worst, err)
                if err >= (3e-3 if dtype == torch.bfloat16 else 1e-4) and os.environ.get('RVN_K1_DEBUG'):
                    d = (view - p).reshape(-1).abs()
                    bad = (d > 1e-2).nonzero().reshape(-1)
                    print('rank', rank, 'param', k, 'offset', store.offsets[k], 'numel', p.numel(), 'bad range', int(bad.min()), int(bad.max()), 'n bad', bad.numel(),
                          'shards', getattr(fused, 'shards', None), flush=True)
                    continue
                assert err < (3e-3 if dtype == torch.bfloat16 else 1e-4), (kind, dtype, err)
            # parameter averaging kernel (reference parallel_ring_reduce semantics); the barrier keeps this rank's fill_ away from
            # peers that are still pulling its master shard (sync_master is one-sided)
            torch.cuda.synchronize()
            dist.barrier(group=group)
            store.master.fill_(float(rank + 1))
            if getattr(fused, 'two_shot', False):
                fused._master_stale = False
            fused.average_parameters()
            torch.cuda.synchronize()
            real = torch.cat([v.reshape(-1) for v in store.master_views])
            assert torch.allclose(real, torch.full_like(real, (world + 1) / 2)), (kind, float(real[0]))
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print('DIST_GPU_OK worst_rel_err', float(t), 'two_shot', os.environ.get('RAVNEST_K1_TWOSHOT', 'auto'), 'world', world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
