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
            for p, view in zip(ref.parameters(), store.master_views):
                err = float((view - p).norm() / p.norm())
                worst = max(worst, err)
                assert err < (3e-3 if dtype == torch.bfloat16 else 1e-4), (kind, dtype, err)
            # parameter averaging kernel (reference parallel_ring_reduce semantics)
            store.master.fill_(float(rank + 1))
            fused.average_parameters()
            torch.cuda.synchronize()
            real = torch.cat([v.reshape(-1) for v in store.master_views])
            assert torch.allclose(real, torch.full_like(real, (world + 1) / 2)), (kind, float(real[0]))
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print('DIST_GPU_OK worst_rel_err', float(t))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
