"""K1 micro-benchmark under torchrun: fused P2P all-reduce + optimizer update + bf16 re-publish (one launch) vs NCCL all_reduce
+ torch optimizer + cast.  CUDA events, max over ranks, L2-exceeding buckets.  Prints one JSON line per case and writes
gpurun_out/k1_bench_n<N>.json.  Roofline = max(HBM time of the update traffic, NVLink time of the (N-1) peer buckets each GPU pulls)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ravnest_b200.optim import build_optimizer  # noqa: E402
from ravnest_b200.runtime.dist_transport import init_distributed  # noqa: E402
from ravnest_b200.runtime.param_store import ParamStore  # noqa: E402

HBM_GBS = float(os.environ.get('RVN_HBM_GBS', 7000))        # measured copy bandwidth (MEASURED_PEAKS.json if present)
NVL_GBS = 900.0                                              # NVLink 5 per direction

try:
    HBM_GBS = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(iters):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    t = torch.tensor([ev[0].elapsed_time(ev[1]) / iters], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank, world, dev = init_distributed()
    group = dist.new_group(list(range(world)))
    rows = []
    for n_params in (55_000_000, 110_000_000):
        import ravnest_b200 as rv
        for kind, cls, params in (('sgd', torch.optim.SGD, dict(lr=0.05, momentum=0.9)), ('adamw', torch.optim.AdamW, dict(lr=1e-4, weight_decay=0.01)),
                                  ('lamb', rv.Lamb, dict(lr=1e-3, weight_decay=0.01))):
            lin = torch.nn.Linear(n_params // 1000, 1000, bias=False).to(dev).to(torch.bfloat16)
            store = ParamStore(lin, dev, n_slots=2, compute_dtype=torch.bfloat16)
            opt, fused = build_optimizer(cls, dict(params), store, dp_group=group, fused=True, average_grads=True)
            store.grad.normal_()
            t_fused = timed(fused.step)
            # library path: NCCL all-reduce of the bf16 bucket, fp32 master optimizer step, cast-publish
            ref_master = [torch.nn.Parameter(store.master.clone())]
            ropt = cls(ref_master, **params)
            bucket = store.grad.clone()
            wout = torch.empty_like(bucket)

            def lib():
                dist.all_reduce(bucket, group=group)
                ref_master[0].grad = bucket.float().div_(world)
                ropt.step()
                wout.copy_(ref_master[0].detach())
            t_lib = timed(lib)
            n = store.total
            state = {'sgd': 1, 'adamw': 2, 'lamb': 2}[kind]
            hbm_bytes = n * (2 + 4 * 2 + state * 4 * 2 + 2 + (12 if kind == 'lamb' else 0))   # grad in, master rw, state rw, bf16 out (+ LAMB phase 2 re-read)
            nvl_bytes = n * 2 * (world - 1)                           # every peer bucket crosses the link once per GPU
            two_shot = bool(getattr(fused, 'two_shot', False))
            if two_shot:                # shard of the gradients in, shard of the weights out; optimizer traffic only for the own shard
                nvl_bytes = n * 2 * (world - 1) / world
                hbm_bytes = hbm_bytes / world + n * 2 * 2          # + landing buffer -> version slot copy
            roof_ms = max(hbm_bytes / HBM_GBS, nvl_bytes / NVL_GBS) / 1e6
            # the ALGORITHMIC floor whatever the implementation: reduce-scatter + sharded update + all-gather
            algo_ms = max((n * (2 + 4 * 2 + state * 4 * 2) / world + n * 2) / HBM_GBS, n * 2 * (world - 1) / world / NVL_GBS) / 1e6
            row = dict(kernel=('rvn_fused_twoshot_' if two_shot else 'rvn_fused_allreduce_') + ('lamb1+lamb2' if kind == 'lamb' else 'update'), mode='two-shot' if two_shot else 'one-shot',
                       optimizer=kind, n_gpus=world, params=n, fused_ms=round(t_fused, 4), nccl_plus_torch_ms=round(t_lib, 4),
                       speedup=round(t_lib / t_fused, 2), roofline_ms=round(roof_ms, 4), fraction_of_roofline=round(roof_ms / t_fused, 3),
                       algorithmic_floor_ms=round(algo_ms, 4), fraction_of_algorithmic_floor=round(algo_ms / t_fused, 3),
                       nvlink_gbs_per_gpu=round(nvl_bytes / t_fused / 1e6, 1), hbm_gbs=round(hbm_bytes / t_fused / 1e6, 1))
            rows.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
            del store, opt, fused, ropt, ref_master, bucket, wout, lin
            torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs('gpurun_out', exist_ok=True)
        json.dump(rows, open('gpurun_out/k1_bench_n{}_{}.json'.format(world, 'twoshot' if rows and rows[-1]['mode'] == 'two-shot' else 'oneshot'), 'w'), indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
