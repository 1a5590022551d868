import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from ravnest_b200.optim import build_optimizer
from ravnest_b200.runtime.dist_transport import init_distributed
from ravnest_b200.runtime.param_store import ParamStore

rank, world, dev = init_distributed()
group = dist.new_group(list(range(world)))
torch.manual_seed(0)
mod = torch.nn.Linear(1000, 1000, bias=False).to(dev)
ref = torch.nn.Parameter(mod.weight.detach().clone())
store = ParamStore(mod, dev, n_slots=3, compute_dtype=torch.float32)
opt, fused = build_optimizer(torch.optim.SGD, dict(lr=0.1), store, dp_group=group, fused=True, average_grads=True)
ropt = torch.optim.SGD([ref], lr=0.1)
print(rank, 'two_shot', fused.two_shot, 'shards', fused.shards, 'n', store.total, flush=True)
for step in range(2):
    torch.manual_seed(100 * step + rank)
    g = torch.randn_like(ref)
    store.grad_views[0].copy_(g)
    gs = g.clone(); dist.all_reduce(gs, group=group); ref.grad = gs / world
    fused.step(); ropt.step()
torch.cuda.synchronize()
flat_ref = ref.detach().reshape(-1)
def report(tag):
    m = store.master[:flat_ref.numel()]
    errs = [float((m[s:e] - flat_ref[s:min(e, flat_ref.numel())]).abs().max()) if e > s else 0 for s, e in fused.shards]
    w = store.current.flat[:flat_ref.numel()]
    werrs = [float((w[s:e] - flat_ref[s:min(e, flat_ref.numel())]).abs().max()) if e > s else 0 for s, e in fused.shards]
    print(rank, tag, 'master err per shard', errs, 'published err per shard', werrs, 'stale', fused._master_stale, flush=True)
report('before sync')
fused.sync_master(); torch.cuda.synchronize()
report('after sync')
for r in range(world):
    pt = fused.master_sym.peer_tensor(r)
    print(rank, 'peer', r, 'ptr', pt.data_ptr(), 'own ptr', store.master.data_ptr(), float(pt[:4].sum()), flush=True)
dist.barrier(); dist.destroy_process_group()
