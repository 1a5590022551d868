#!/usr/bin/env python
"""Headline benchmark: BERT-base pre-training samples/sec (bf16) — or ResNet-50 images/sec — through
the ravnest_b200 engine on N GPUs of one B200 box.

Contract (driver):  python bench.py --gpus N --steps K --warmup W   (N>1 launched under torchrun)
prints ONE JSON line on rank 0.  ``--impl reference`` reports why the upstream package cannot run here;
``--impl baseline`` runs the same schedule on stock cuBLAS/cuDNN/ATen ops + NCCL collectives + unfused
torch optimizer (the baseline tier of BASELINE.md §3).

Configs (BASELINE.json): BERT-base bf16, batch 8x512 per replica, LAMB, update_frequency 16
(examples/bert/provider.py): N=1 solo, N=2 -> 2 stages x 1 replica, N=4 -> 2x2, N=8 -> 2x4.
ResNet-50 (num_classes=200), batch 100x3x224x224 per replica, SGD-momentum: N=1 solo, 2 -> 2x1,
4 -> 4x1, 8 -> 4x2.  Synthetic data of that shape, random-init weights.
"""
import argparse
import json
import os
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='native', choices=['native', 'baseline', 'reference'])
    ap.add_argument('--model', default='bert', choices=['bert', 'resnet50', 'inception', 'sorter', 'cnn'])
    ap.add_argument('--ref-cpu', action='store_true', help='reference arm on CPU (plumbing check of the arm itself; not a benchmark)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--also', default=None, help='comma-separated further models timed after --model in the same job; their results are nested '
                    'under "also" (default: resnet50 next to bert — the two halves of the BASELINE.json headline; "" for none)')
    ap.add_argument('--update-frequency', type=int, default=None)
    ap.add_argument('--topology', default=None, help='SxR override (stages x replicas), e.g. 1x2 to exercise pure DP')
    ap.add_argument('--batch', type=int, default=0, help='per-replica batch override (sweeps; the result is marked as a non-headline config)')
    ap.add_argument('--balance', default='measured', choices=['measured', 'compute', 'params'], help='stage-balancing policy of the planner')
    ap.add_argument('--proportions', default=None, help='comma-separated parameter shares per stage (reference policy) instead of the compute-balanced plan')
    ap.add_argument('--no-prefetch-h2d', dest='prefetch_h2d', action='store_false', help='copy host inputs on the compute stream at the head of their '
                    'pass instead of on a copy stream at submit time (Node(prefetch_h2d=True) is the default here: ResNet-50 e2e +6 %%)')
    ap.set_defaults(prefetch_h2d=True)
    ap.add_argument('--no-graphs', action='store_true', help='disable CUDA-graph capture of the stage step')
    ap.add_argument('--rebalance', type=int, default=1, help='rounds of profile-guided re-planning before the warm-up (pipelines only): the plan '
                    'is cut again with cost shares corrected by the measured per-stage step time')
    ap.add_argument('--layers', type=int, default=None, help='debug only: shrink the model (result marked invalid)')
    ap.add_argument('--diagnose', type=int, default=0, help='after the timed regions: profile this many steady-state steps on every rank and '
                    'write gpurun_out/diag_<model>_n<N>_rank<r>.txt (per-stream busy time, gaps, top kernels, native call counters)')
    return ap.parse_args()


def run_reference(args):
    """The unmodified upstream package (baseline/_ref) through its own clusterize / Node / Trainer: baseline/ref_bench.py."""
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import ref_bench
    return ref_bench.run(args, (SyntheticBert, SyntheticImages), ClockSampler)


class ClockSampler(threading.Thread):
    """Sample SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown',
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: 'hw_thermal_slowdown',
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: 'sw_thermal_slowdown',
                     nv.nvmlClocksThrottleReasonSwPowerCap: 'sw_power_cap',
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: 'hw_power_brake'}
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(self.period)
        except Exception as err:        # NVML missing: report it rather than fake clocks
            self.reasons.add('nvml_unavailable:{}'.format(type(err).__name__))

    def summary(self):
        self._halt.set()
        self.join(2)
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


def topology_for(model, n):
    if model == 'bert':
        return {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (2, 4)}[n]
    if model == 'inception':                      # BASELINE.json config #4: 3-stage pipeline (use --gpus 3 via torchrun, or a grid)
        return {1: (1, 1), 2: (2, 1), 3: (3, 1), 4: (2, 2), 6: (3, 2), 8: (2, 4)}[n]
    if model == 'sorter':                         # config #5: pure pipeline, one stage per GPU
        return (n, 1)
    return {1: (1, 1), 2: (2, 1), 4: (4, 1), 8: (4, 2)}[n]


def build_model(args, meta_only=False):
    import torch
    from ravnest_b200 import models
    if meta_only:
        return None, None, None, {'bert': 8, 'inception': 64, 'sorter': 64, 'cnn': 64}.get(args.model, 100)
    if args.model == 'bert':
        cfg = models.BertConfig()
        if args.layers:
            cfg.num_hidden_layers = args.layers
        model = models.BertForPreTraining(cfg)
        B, S = 8, 512
        example = {'input_ids': torch.randint(0, cfg.vocab_size, (2, S)), 'attention_mask': torch.ones(2, S, dtype=torch.long),
                   'token_type_ids': torch.zeros(2, S, dtype=torch.long)}
        return model, (), example, B
    if args.model == 'inception':
        return models.Inception3(), (torch.rand(2, 3, 32, 32),), {}, 64
    if args.model == 'sorter':
        cfg = models.GPTConfig(model_type='gpt-nano', vocab_size=3, block_size=11)
        return models.GPT(cfg), (torch.randint(0, 3, (2, 11)),), {}, 64
    model = models.resnet50(num_classes=200)
    return model, (torch.rand(2, 3, 224, 224),), {}, 100


class SyntheticBert:
    """Fixed pool of synthetic BERT batches (pinned host memory or device-resident)."""

    def __init__(self, n, batch, seq, vocab, device=None, seed=0):
        import torch
        g = torch.Generator().manual_seed(seed)
        self.items = []
        for _ in range(n):
            ids = torch.randint(0, vocab, (batch, seq), generator=g)
            mask = torch.ones(batch, seq, dtype=torch.long)
            tt = torch.zeros(batch, seq, dtype=torch.long)
            labels = torch.full((batch, seq), -100, dtype=torch.long)
            pick = torch.rand(batch, seq, generator=g) < 0.15
            labels[pick] = ids[pick]
            sop = torch.randint(0, 2, (batch,), generator=g)
            item = {'input_ids': ids, 'attention_mask': mask, 'token_type_ids': tt, 'labels': labels, 'sentence_order_label': sop}
            if device is not None:
                item = {k: v.to(device) for k, v in item.items()}
            elif torch.cuda.is_available():
                item = {k: v.pin_memory() for k, v in item.items()}
            self.items.append(item)

    def __len__(self):
        return len(self.items)

    def __iter__(self):
        while True:
            for it in self.items:
                yield it

    def input_bytes(self):
        it = self.items[0]
        return sum(v.numel() * v.element_size() for v in it.values())


class SyntheticImages:
    def __init__(self, n, batch, classes, device=None, seed=0, dtype=None, size=224, tokens=None):
        import torch
        g = torch.Generator().manual_seed(seed)
        self.items = []
        for _ in range(n):
            if tokens is not None:                 # GPT-Sorter: [batch, 11] int64 inputs, targets with -1 on the input half
                x = torch.randint(0, classes, (batch, tokens), generator=g)
                y = torch.sort(x[:, :6], dim=1)[0]
                y = torch.cat([torch.full((batch, 5), -1, dtype=torch.long), y], dim=1)
            else:
                x = torch.randn(batch, 3, size, size, generator=g)
                y = torch.randint(0, classes, (batch,), generator=g)
            if device is not None:
                x, y = x.to(device), y.to(device)
            elif torch.cuda.is_available():
                x, y = x.pin_memory(), y.pin_memory()
            self.items.append((x, y))

    def __len__(self):
        return len(self.items)

    def __iter__(self):
        while True:
            for it in self.items:
                yield it

    def input_bytes(self):
        x, y = self.items[0]
        return x.numel() * x.element_size() + y.numel() * y.element_size()


def main():
    args = parse()
    if args.also is None:
        args.also = 'resnet50' if (args.model == 'bert' and args.impl != 'reference' and not args.layers and not args.batch
                                   and not args.topology and not args.diagnose) else ''
    if args.impl == 'reference':
        return run_reference(args)
    import torch
    import torch.distributed as dist
    import ravnest_b200 as rv
    from ravnest_b200 import ops
    from ravnest_b200.runtime.dist_transport import init_distributed

    world = int(os.environ.get('WORLD_SIZE', 1))
    if world != args.gpus:
        if args.gpus > 1 and 'RANK' not in os.environ:
            # convenience: re-launch ourselves under torchrun
            import subprocess
            cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
                   '--master-addr', '127.0.0.1', '--master-port', str(29500 + os.getpid() % 400)] + sys.argv
            return subprocess.call(cmd)
    if not torch.cuda.is_available():
        print(json.dumps({'error': 'no CUDA device visible; bench.py needs a B200'}))
        return 1
    import faulthandler
    # a hung multi-GPU run must be cheap and diagnosable: dump every thread's stack and exit instead of burning the box
    watchdog_s = int(os.environ.get('RVN_BENCH_WATCHDOG', '150'))

    def milestone(text):
        """Re-arm the watchdog (each phase gets the full budget) and, with RAVNEST_TRACE_STARTUP=1, log progress on stderr."""
        faulthandler.dump_traceback_later(watchdog_s, exit=True)
        if os.environ.get('RAVNEST_TRACE_STARTUP', '0') == '1':
            print('[rank {} t={:.1f}] {}'.format(os.environ.get('RANK', '0'), time.time() % 10000, text), file=sys.stderr, flush=True)

    milestone('imports done')
    rank, world, dev = init_distributed()
    milestone('process group up')
    gloo = dist.new_group(backend='gloo') if world > 1 else None
    if args.impl == 'baseline':
        ops.set_backend('torch')
    rv.set_seed(1234 + 0)
    torch.backends.cudnn.deterministic = False
    torch.backends.cudnn.benchmark = True

    results = {}
    models_to_run = [args.model] + [m for m in (args.also.split(',') if args.also else []) if m and m != args.model]
    for i, name in enumerate(models_to_run):
        sub = argparse.Namespace(**vars(args))
        sub.model = name
        if i > 0:                       # secondary models: same K / W, no per-model overrides
            sub.update_frequency = None
            sub.topology = sub.proportions = None
            sub.batch = 0
            sub.diagnose = 0
        results[name] = bench_one(sub, rank, world, dev, gloo, milestone)
        _release_gpu_objects(dev)
    if rank == 0:
        result = results[args.model]
        extra = {k: v for k, v in results.items() if k != args.model and v is not None}
        if extra:
            result['also'] = extra      # the other half of the headline metric (BASELINE.json: ResNet-50 images/s AND BERT samples/s)
        print(json.dumps(result), flush=True)
    faulthandler.cancel_dump_traceback_later()
    # the measurements are out: whatever happens during teardown (a peer that is already gone, a slow communicator destroy)
    # must neither hang the box nor turn a finished run into a failure
    killer = threading.Timer(90.0, lambda: os._exit(0))
    killer.daemon = True
    killer.start()
    if world > 1:
        torch.cuda.synchronize(dev)
        dist.barrier(group=gloo)
        dist.destroy_process_group()
    return 0


def _release_gpu_objects(dev):
    """Collect dead Node / CUDA-graph objects now.  Left to the cyclic GC they can be destroyed at an arbitrary later moment — e.g. inside
    the next stage-graph capture, where freeing a graph's memory pool is an illegal operation and invalidates the capture
    (cudaErrorStreamCaptureIsolation, seen once on the second re-planning round)."""
    import gc
    import torch
    gc.collect()
    torch.cuda.synchronize(dev)
    torch.cuda.empty_cache()


def bench_one(args, rank, world, dev, gloo, milestone):
    """Plan, build and time ONE model on the already initialised process group; returns the result dict on rank 0."""
    import torch
    import torch.distributed as dist
    import ravnest_b200 as rv
    fused = args.impl == 'native'
    stages, replicas = topology_for(args.model, world)
    if args.topology:
        stages, replicas = (int(v) for v in args.topology.lower().split('x'))
        assert stages * replicas == world
    home = os.path.join(os.environ.get('RAVNEST_BENCH_HOME') or os.path.join(tempfile.gettempdir(), 'rvn_bench_{}'.format(os.environ.get('MASTER_PORT', '0'))), args.model)
    os.environ['RAVNEST_HOME'] = home
    _, _, _, batch = build_model(args, meta_only=True)
    if args.batch:
        batch = args.batch
    user_props = [float(v) for v in args.proportions.split(',')] if args.proportions else None
    plan_state = {'proportions': user_props, 'rounds': 0, 'history': []}

    def plan(props, balance):
        """(Re)write the cluster plan (rank 0) and wait for it (everybody)."""
        if rank == 0:
            os.makedirs(home, exist_ok=True)
            model, ex_args, ex_kwargs, _ = build_model(args)
            ex_tensors = [t for t in list(ex_args) + list(ex_kwargs.values()) if hasattr(t, 'shape')]
            cost_scale = batch / max(1, ex_tensors[0].shape[0])          # the plan is traced on a 2-sample example of the real shape
            rv.clusterize(model, example_args=ex_args, example_kwargs=ex_kwargs, topology={'stages': stages, 'replicas': replicas},
                          node_configs=rv.default_gpu_configs(world), base_dir=home, proportions=props,
                          balance=balance, cost_scale=cost_scale, plan_device=dev)
            del model
        milestone('plan written' if rank == 0 else 'waiting for the plan')
        if world > 1:
            dist.barrier(group=gloo)

    # stages are balanced by measured step time (balance='measured'), not by parameter bytes, unless --proportions pins the reference policy
    plan(user_props, 'params' if user_props else args.balance)
    milestone('building node')

    cfg = rv.load_node_json_configs('node_{}'.format(rank), base_dir=home)
    replica = cfg['rank']
    fused = args.impl == 'native'
    if args.model == 'bert':
        upd = args.update_frequency or 16
        pool_dev = SyntheticBert(4, batch, 512, 30522, device=dev, seed=replica)
        pool_host = SyntheticBert(4, batch, 512, 30522, device=None, seed=replica)

        def criterion(outputs, targets):
            t = targets
            if not t['labels'].is_cuda:
                t = {k: v.to(dev, non_blocking=True) for k, v in t.items() if k in ('labels', 'sentence_order_label')}
            return rv.models.bert_pretraining_loss(outputs, t['labels'], t['sentence_order_label']) / upd

        node_kwargs = dict(optimizer=rv.Lamb, optimizer_params={'lr': 1.76e-3, 'eps': 1e-6, 'weight_decay': 0.01,
                                                                'clamp_value': 10000.0, 'debias': True},
                           update_frequency=upd, dtype=torch.bfloat16)
        feed = lambda node, item: node.forward_compute(input_ids=item['input_ids'], attention_mask=item['attention_mask'],
                                                       token_type_ids=item['token_type_ids'])
        metric, unit = 'bert_base_pretraining_samples_per_sec', 'samples/s'
    elif args.model == 'sorter':
        upd = args.update_frequency or 1
        pool_dev = SyntheticImages(4, batch, 3, device=dev, seed=replica, tokens=11)
        pool_host = SyntheticImages(4, batch, 3, device=None, seed=replica, tokens=11)

        def criterion(outputs, targets):
            return torch.nn.functional.cross_entropy(outputs.reshape(-1, outputs.size(-1)).float(), targets[1].reshape(-1).to(dev, non_blocking=True),
                                                     ignore_index=-1)

        node_kwargs = dict(optimizer=torch.optim.Adam, optimizer_params={}, update_frequency=upd, dtype=torch.bfloat16)
        feed = lambda node, item: node.forward_compute(tensors=item[0])
        metric, unit = 'gpt_sorter_samples_per_sec', 'samples/s'
    else:
        upd = args.update_frequency or 1
        size, classes = (32, 10) if args.model == 'inception' else (224, 200)
        pool_dev = SyntheticImages(2, batch, classes, device=dev, seed=replica, size=size)
        pool_host = SyntheticImages(2, batch, classes, device=None, seed=replica, size=size)
        loss_fct = torch.nn.CrossEntropyLoss()

        def criterion(outputs, targets):
            return loss_fct(outputs.float(), targets[1].to(dev, non_blocking=True))

        node_kwargs = dict(optimizer=torch.optim.SGD, optimizer_params={'lr': 0.01, 'momentum': 0.9, 'weight_decay': 5e-4},
                           update_frequency=upd, dtype=torch.bfloat16)
        feed = lambda node, item: node.forward_compute(tensors=item[0])
        metric, unit = ('inception_v3_images_per_sec' if args.model == 'inception' else 'resnet50_images_per_sec'), 'images/s'

    labels = {'dev': iter(pool_dev), 'host': iter(pool_host)}

    class Labels:                   # leaf/solo label source switches together with the input source
        mode = 'dev'

        def __iter__(self):
            return self

        def __next__(self):
            return next(labels[Labels.mode])

        def __len__(self):
            return 1 << 30

    def make_node():
        return rv.Node(name='node_{}'.format(rank), criterion=criterion, labels=Labels(), device=dev, base_dir=home,
                       reduce_factor=1 << 30, dp_mode='gradient', fused_optimizer=fused, loss_filename=None,
                       recompute=False, use_cuda_graphs=not args.no_graphs, prefetch_h2d=args.prefetch_h2d, **node_kwargs)

    node = make_node()
    driver = node.node_type in ('root', 'solo')

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(group=gloo)
        torch.cuda.synchronize(dev)

    def run(steps, mode, readback):
        Labels.mode = mode
        node.compute_session.loss_readback = readback
        src = iter(pool_dev if mode == 'dev' else pool_host)
        if driver:
            for _ in range(steps):
                feed(node, next(src))
            node.wait_for_backwards()
        sync()

    def timed(steps, mode, readback):
        sync()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream = node.compute_stream
        start.record(stream)
        t0 = time.time()
        run(steps, mode, readback)
        end.record(stream)
        torch.cuda.synchronize(dev)
        wall = time.time() - t0
        ms = torch.tensor([start.elapsed_time(end), wall * 1e3], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1])

    milestone('node up, warm-up')
    # untimed pre-warm in addition to the W warm-up steps: CUDA-graph capture of a stage needs graph_warmup eager forwards, two
    # eager backwards and then one capture per in-flight pass (cluster_length + 2 instances) — none of that may land in the
    # timed region whatever W the caller picks
    prewarm = 12 + 2 * stages
    run(prewarm, 'dev', False)

    def stage_busy(steps=8):
        """Device time every stage spends in its own handlers per step (max over the replicas of a stage)."""
        node.tracer.resolve(block=True)
        e0 = node.tracer.totals_ms.get('exposed_comm', 0.0)
        b0 = sum(v for k, v in node.tracer.totals_ms.items() if k != 'exposed_comm')
        run(steps, 'dev', False)
        node.tracer.resolve(block=True)
        own = (sum(v for k, v in node.tracer.totals_ms.items() if k != 'exposed_comm') - b0) / steps
        own -= (node.tracer.totals_ms.get('exposed_comm', 0.0) - e0) / steps
        mine = (int(cfg.get('stage_index', 0)), own)
        every = [None] * world
        dist.all_gather_object(every, mine, group=gloo)
        per_stage = [0.0] * stages
        for st, ms in every:
            per_stage[st] = max(per_stage[st], ms)
        return per_stage

    # profile-guided re-planning: the planner's per-node estimates miss what only the running pipeline shows (graph replay,
    # boundary copies, clocks); cut the plan again with shares corrected by the measured stage times (rv.rebalanced_proportions)
    while stages > 1 and user_props is None and plan_state['rounds'] < args.rebalance:
        busy_now = stage_busy()
        plan_state['history'].append([round(v, 3) for v in busy_now])
        if max(busy_now) <= 1.06 * (sum(busy_now) / len(busy_now)):
            break
        cur = plan_state['proportions'] or [1.0 / stages] * stages
        plan_state['proportions'] = rv.rebalanced_proportions(cur, busy_now)
        plan_state['rounds'] += 1
        milestone('re-planning with shares {}'.format([round(v, 3) for v in plan_state['proportions']]))
        sync()
        if driver:
            node.stop()
        node._stop.wait(20)
        time.sleep(0.2)
        sync()
        node = None
        _release_gpu_objects(dev)       # the old stage graphs must be gone BEFORE anything new is captured
        plan(plan_state['proportions'], 'measured')
        node = make_node()
        run(prewarm, 'dev', False)
    run(max(3, args.warmup), 'dev', False)
    milestone('warm-up done, timing')
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    node.tracer.resolve(block=True)
    exposed0 = node.tracer.totals_ms.get('exposed_comm', 0.0)
    busy0 = sum(v for k, v in node.tracer.totals_ms.items() if k != 'exposed_comm')
    updates0 = node.compute_session.n_updates
    dev_ms, dev_wall = timed(args.steps, 'dev', False)
    updates_in_window = node.compute_session.n_updates - updates0
    clocks = sampler.summary()
    node.tracer.resolve(block=True)
    own_exposed = (node.tracer.totals_ms.get('exposed_comm', 0.0) - exposed0) / args.steps
    exposed = torch.tensor([own_exposed], device=dev)
    if world > 1:
        dist.all_reduce(exposed, op=dist.ReduceOp.MAX)
    exposed_ms = float(exposed)
    # device time each rank's stage spent in its own handlers (forward / backward / update ...) per step: the balance of the plan
    busy = (sum(v for k, v in node.tracer.totals_ms.items() if k != 'exposed_comm') - busy0) / args.steps - own_exposed
    busy_all = [None] * world
    if world > 1:
        dist.all_gather_object(busy_all, round(busy, 3), group=gloo)
    else:
        busy_all = [round(busy, 3)]
    result = None
    e2e = None
    if not args.no_e2e:
        run(3, 'host', True)
        e2e_ms, e2e_wall = timed(args.steps, 'host', True)
        e2e = {'value': replicas * batch * args.steps / (max(e2e_ms, e2e_wall) / 1e3), 'unit': unit,
               'h2d_bytes_per_step': pool_host.input_bytes(), 'd2h_bytes_per_step': 4, 'ms_per_step': max(e2e_ms, e2e_wall) / args.steps}
    # the loss every leaf / solo rank computed last (after the timed regions): evidence that the step trains, not just runs
    cs = node.compute_session
    mine = float(cs.last_loss.float()) if getattr(cs, 'last_loss', None) is not None else None
    last_losses = [None] * world
    if world > 1:
        dist.all_gather_object(last_losses, mine, group=gloo)
    else:
        last_losses = [mine]
    # data-parallel replicas of a stage must hold identical master weights after the timed steps: proof that the fused all-reduce +
    # optimizer kernel (K1) did its job on every replica (max |w_i - w_j| over the stage's DP group, worst stage reported)
    replica_diff = None
    if node.dp_group is not None:
        with node.on_compute_stream():
            cs.sync_master()                    # two-shot K1 keeps the fp32 master sharded over the replicas
            m = cs.store.master
            hi, lo = m.clone(), m.clone()
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=node.dp_group)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=node.dp_group)
            worst = torch.stack([(hi - lo).abs().max(), m.abs().max()])
        torch.cuda.synchronize(dev)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        replica_diff = {'max_abs_diff': float(worst[0]), 'max_abs_weight': float(worst[1]), 'updates_checked': cs.n_updates}
        del hi, lo
    launches = count_launches(node, feed, pool_dev, driver, sync, Labels) if rank == 0 or world > 1 else None
    if args.diagnose:
        diagnose(args, node, feed, pool_dev, driver, sync, Labels, rank, world)
    if rank == 0:
        value = replicas * batch * args.steps / (dev_ms / 1e3)
        result = {
            'metric': metric, 'value': value, 'unit': unit, 'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
            'ms_per_step': dev_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16', 'data': 'synthetic (random tokens/images of the benchmark shape; random-init weights)',
            'impl': args.impl,
            'config': {'model': {'bert': 'bert-base-uncased BertForPreTraining (110.1M)', 'resnet50': 'resnet50(num_classes=200) (23.9M)',
                                 'inception': 'Inception3 CIFAR variant (21.6M)', 'sorter': 'minGPT gpt-nano (85.7K)'}[args.model],
                       'global_batch': replicas * batch, 'per_replica_batch': batch, 'seq_len': 512 if args.model == 'bert' else None,
                       'parallelism': 'pp{}xdp{}'.format(stages, replicas), 'optimizer': {'bert': 'LAMB', 'sorter': 'Adam'}.get(args.model, 'SGD-momentum'),
                       'update_frequency': upd, 'schedule': 'async 1F1B, weight stashing, activations resident (no recompute)', 'cuda_graphs': not args.no_graphs, 'prewarm_steps_before_warmup': prewarm,
                       'stage_shares': plan_state['proportions'], 'replan_rounds': plan_state['rounds'], 'stage_busy_before_replan_ms': plan_state['history'],
                       'l2': 'working set (weights+activations, >1 GB/step) exceeds the 126 MB L2; no explicit flush'},
            'clocks': clocks, 'e2e': e2e,
            'gpu_launches': (launches or {}).get('ours_per_step', 0) * args.steps if isinstance(launches, dict) else 0,
            'gpu_launches_detail': launches, 'wall_ms_per_step': dev_wall / args.steps,
            'exposed_comm_ms_per_step': exposed_ms, 'stage_busy_ms_per_rank': busy_all,
            'last_loss_per_rank': [None if v is None else round(v, 4) for v in last_losses],
            'optimizer_updates_in_window': updates_in_window, 'replica_weights': replica_diff,
        }
        if replica_diff is not None and replica_diff['max_abs_diff'] > 1e-6 * max(1.0, replica_diff['max_abs_weight']):
            result['invalid'] = 'data-parallel replicas diverged: {}'.format(replica_diff)
        if args.layers:
            result['invalid'] = 'debug run with --layers'
        if args.batch:
            result['invalid'] = 'sweep run with --batch (not the BASELINE.json batch size)'
    sync()
    if driver:
        node.stop()
    node._stop.wait(20)
    time.sleep(0.2)
    if world > 1:
        dist.barrier(group=gloo)
    node = None
    return result


def diagnose(args, node, feed, pool, driver, sync, Labels, rank, world):
    """Steady-state timeline summary of this rank: where the step's device time goes and where the stream sits idle."""
    import torch
    from collections import defaultdict
    from torch.profiler import ProfilerActivity, profile
    from ravnest_b200.ops import native
    Labels.mode = 'dev'
    steps = args.diagnose
    src = iter(pool)
    sync()
    before = dict(native.calls)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        if driver:
            for _ in range(steps):
                feed(node, next(src))
            node.wait_for_backwards()
        sync()
    after = dict(native.calls)
    per_stream = defaultdict(list)
    by_name = defaultdict(lambda: [0, 0.0])
    try:            # kineto events carry the stream id (FunctionEvent does not)
        raw = [(e.device_resource_id(), e.start_ns() / 1e3, (e.start_ns() + e.duration_ns()) / 1e3, e.name())
               for e in prof.profiler.kineto_results.events()
               if str(e.device_type()).endswith('CUDA') and e.duration_ns() > 0]
    except Exception:
        raw = [(-1, ev.time_range.start, ev.time_range.end, ev.name) for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
    for stream, t0, t1, name in raw:
        per_stream[stream].append((t0, t1, name))
        by_name[name][0] += 1
        by_name[name][1] += (t1 - t0)
    lines = ['rank {} of {}  node_type {}  steps {}'.format(rank, world, node.node_type, steps)]
    for stream, evs in sorted(per_stream.items(), key=lambda kv: -sum(b - a for a, b, _ in kv[1])):
        evs.sort()
        busy = sum(b - a for a, b, _ in evs)
        span = evs[-1][1] - evs[0][0]
        gaps = sorted(((evs[i + 1][0] - evs[i][1], evs[i][2][:60], evs[i + 1][2][:60]) for i in range(len(evs) - 1)), reverse=True)
        big = [g for g in gaps if g[0] > 20]
        lines.append('stream {}: {} kernels, busy {:.3f} ms/step, span {:.3f} ms/step, idle {:.3f} ms/step in {} gaps > 20 us'.format(
            stream, len(evs), busy / 1e3 / steps, span / 1e3 / steps, sum(g[0] for g in big) / 1e3 / steps, len(big)))
        for g in big[:12]:
            lines.append('    gap {:8.1f} us  after {}  before {}'.format(g[0], g[1], g[2]))
    lines.append('top kernels (count/step, us/step):')
    for name, (n, t) in sorted(by_name.items(), key=lambda kv: -kv[1][1])[:45]:
        lines.append('  {:7.1f} {:9.1f}  {}'.format(n / steps, t / steps, name[:110]))
    lines.append('native calls per step: ' + json.dumps({k: round((after.get(k, 0) - before.get(k, 0)) / steps, 2) for k in after
                                                         if after.get(k, 0) != before.get(k, 0)}))
    lines.append('native fallbacks: ' + json.dumps(dict(getattr(native, 'fallbacks', {}))))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'diag_{}_n{}_rank{}.txt'.format(args.model, world, rank)), 'w') as fh:
        fh.write('\n'.join(lines) + '\n')


def count_launches(node, feed, pool, driver, sync, Labels):
    """Kernels whose name starts with ``rvn_`` (ours) launched by ONE step on this rank, via the
    torch profiler — outside the timed region."""
    import torch
    from torch.profiler import ProfilerActivity, profile
    Labels.mode = 'dev'
    try:
        sync()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            if driver:
                feed(node, next(iter(pool)))
                node.wait_for_backwards()
            sync()
        mine = total = 0
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                total += 1
                if 'rvn_' in ev.name:
                    mine += 1
        from ravnest_b200.ops import native_ops_loaded
        return {'ours_per_step': mine, 'all_per_step': total, 'native_calls_traced': native_ops_loaded()}
    except Exception as err:
        return {'error': repr(err)}


if __name__ == '__main__':
    sys.exit(main())
