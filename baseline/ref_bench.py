"""Reference arm of ``bench.py`` (``--impl reference``): drive the UNMODIFIED upstream package (``baseline/_ref``, see
``install_reference.py``) through its own public API — ``ravnest.clusterize`` → ``ravnest.Node`` → ``Trainer.train()`` —
for the same metric / model / batch / optimizer / topology / step count as the native arm.

Nothing of ravnest_b200 (models, kernels, engine, planner) is on this path.  What this file adds around the reference:

* a provider script per node, modelled on the reference's ``examples/{bert,resnet50}/provider.py`` (synthetic data instead
  of the on-disk datasets — there is no network);
* ``node_data/node_configs.json`` sized so that the reference's genetic clusterer (unseeded, several zero-penalty optima)
  lands on the benchmark topology: RAM per node is chosen from the reference's own ``get_memory_reqs`` so that a cluster
  needs exactly S nodes, and Python's ``random`` is seeded with the first seed for which the reference's
  ``genetic_algorithm`` returns R clusters of S nodes;
* one provider process per GPU (``CUDA_VISIBLE_DEVICES`` = local rank, so every provider sees its GPU as ``cuda`` exactly
  like the reference's one-machine-per-provider examples); at N=1 the smallest cluster the reference supports (root + leaf,
  ``ravnest/operations/utils.py:536-543``) shares the single GPU;
* timing: W warm-up steps through ``Trainer.train()``, a barrier, then K steps through ``Trainer.train()``; every provider
  brackets the timed region with CUDA events on the stream the reference computes on (the default stream); the result is
  the max over providers.  The reference's path is end to end by construction (host loader → ``.to(device)`` → … →
  ``loss.item()``), so ``e2e`` equals ``value``.

Precision is the reference's: fp32 (torch defaults, no autocast — it has no reduced-precision compute path).
"""
import contextlib
import json
import os
import random
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, '_ref')
SHIMS = os.path.join(HERE, 'shims')


def _reap():
    """The reference spawns a Manager process and a gRPC server process per node as daemons and never joins them; this harness leaves
    through os._exit (daemon threads of the reference spin forever), which skips multiprocessing's atexit cleanup — so kill every
    descendant explicitly, or they outlive the benchmark holding the GPU and the stdout pipe."""
    try:
        import psutil
        for child in psutil.Process().children(recursive=True):
            try:
                child.kill()
            except psutil.Error:
                pass
    except Exception:
        pass


def _leave(code=0):
    _reap()
    os._exit(code)


def unavailable(why):
    print(json.dumps({'impl': 'reference', 'unavailable': why}), flush=True)
    return 0


def _topology(model, n):
    """(stages, replicas, nodes) — identical to the native arm; at N=1 the reference's minimum cluster (2 nodes) on one GPU."""
    if n == 1:
        return 2, 1, 2
    if model == 'bert':
        s, r = {2: (2, 1), 4: (2, 2), 8: (2, 4)}[n]
    elif model == 'inception':
        s, r = {2: (2, 1), 3: (3, 1), 4: (2, 2), 6: (3, 2), 8: (2, 4)}[n]
    elif model == 'sorter':
        s, r = n, 1
    else:
        s, r = {2: (2, 1), 4: (4, 1), 8: (4, 2)}[n]
    return s, r, n


class CyclicLoader:
    """DataLoader stand-in: ``length`` batches per epoch, cycling over a small pool of host tensors."""

    def __init__(self, pool, length, as_pairs):
        self.pool, self.length, self.as_pairs = pool, length, as_pairs

    def __len__(self):
        return self.length

    def __iter__(self):
        for i in range(self.length):
            item = self.pool[i % len(self.pool)]
            yield item


def _build_model(model_name, device):
    import torch
    if model_name == 'bert':
        from transformers import BertConfig, BertForPreTraining
        config = BertConfig()
        if os.environ.get('RVN_REF_TINY') == '1':      # plumbing check on CPU only; the result is marked invalid
            config = BertConfig(num_hidden_layers=4, hidden_size=128, num_attention_heads=2, intermediate_size=256, vocab_size=1000)
        config.return_dict = False
        # the explicit matmul-softmax attention of the transformers release the reference was written against: its additive
        # float mask is what crosses the stage cut (the reference calls requires_grad_() on every received tensor,
        # ravnest/compute.py:436, which a boolean SDPA mask cannot take)
        config._attn_implementation = 'eager'
        model = BertForPreTraining(config)
        batch, seq = 8, 512
        g = torch.Generator().manual_seed(0)
        example_kwargs = {'input_ids': torch.randint(0, config.vocab_size, (batch, seq), generator=g).to(device),
                          'attention_mask': torch.ones(batch, seq, dtype=torch.long, device=device),
                          'token_type_ids': torch.zeros(batch, seq, dtype=torch.long, device=device)}
        return model.to(device), (), example_kwargs, batch
    if model_name == 'resnet50':
        from torchvision.models import resnet50
        return resnet50(num_classes=200).to(device), (torch.rand((100, 3, 224, 224), device=device),), {}, 100
    if model_name == 'inception':
        sys.path.insert(0, REF)
        from models import inception_v3
        return inception_v3().to(device), (torch.rand((64, 3, 32, 32), device=device),), {}, 64
    if model_name == 'sorter':
        sys.path.insert(0, REF)
        from examples.sorter.mingpt.model_without_padding_mask import GPT
        cfg = GPT.get_default_config()
        cfg.model_type = 'gpt-nano'
        cfg.vocab_size = 3
        cfg.block_size = 11
        return GPT(cfg).to(device), (torch.randint(0, 2, (64, 11), dtype=torch.int64, device=device),), {}, 64
    if model_name == 'cnn':
        sys.path.insert(0, REF)
        from models import CNN_Net
        return CNN_Net().to(device), (torch.rand((64, 1, 8, 8), device=device),), {}, 64
    raise ValueError(model_name)


def _plan(model_name, stages, replicas, n_nodes, base_port, device, log):
    """rank 0: node_configs.json + the reference's clusterize() in the current directory."""
    import torch
    import ravnest
    from ravnest.operations import utils as ref_planner
    from ravnest.operations.genetic import genetic_algorithm
    model, ex_args, ex_kwargs, _ = _build_model(model_name, device)
    os.makedirs('node_data', exist_ok=True)
    # the reference's own size estimate (MB) decides how many nodes a cluster needs
    with contextlib.redirect_stdout(log):
        if ex_kwargs:
            size_mb = ref_planner.get_memory_reqs(model=model, input_data=dict(ex_kwargs))
        elif model_name == 'sorter':
            size_mb = ref_planner.get_memory_reqs(model=model, input_data=ex_args[0])
        else:
            size_mb = ref_planner.get_memory_reqs(model=model, input_size=ex_args[0].shape)
    ram_gb = size_mb / (stages - 0.5) / 1024.0            # (S-1)*ram <= size < S*ram  ->  exactly S nodes per cluster
    configs = {str(i): {'IP': '127.0.0.1:{}'.format(base_port + i), 'benchmarks': {'ram': ram_gb, 'bandwidth': 10}}
               for i in range(n_nodes)}
    with open('node_data/node_configs.json', 'w') as fh:
        json.dump(configs, fh)
    pool, _ = ref_planner.spawn_node_pool(mode='load_from_configs')
    seed = None
    for cand in range(200):
        random.seed(cand)
        with contextlib.redirect_stdout(log):
            clusters = genetic_algorithm(pool, size_mb)
        sizes = sorted(len(v) for v in clusters.values())
        if sizes == [stages] * replicas:
            seed = cand
            break
    if seed is None:
        raise RuntimeError('the reference GA never produced {}x{} in 200 seeds'.format(stages, replicas))
    random.seed(seed)
    with contextlib.redirect_stdout(log):
        if model_name == 'sorter':
            ravnest.clusterize(model=model, example_args=ex_args, pass_data=True)
        elif ex_kwargs:
            ravnest.clusterize(model=model, example_kwargs=ex_kwargs)
        else:
            ravnest.clusterize(model=model, example_args=ex_args)
    del model
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
    return {'ga_seed': seed, 'model_size_mb': size_mb, 'ram_gb_per_node': round(ram_gb, 3)}


def run(args, datasets, ClockSampler):
    """Entry point used by ``bench.py --impl reference``.  ``datasets`` = (SyntheticBert, SyntheticImages) of bench.py."""
    try:
        if not os.path.exists(os.path.join(REF, 'ravnest', 'protos', 'server_pb2_grpc.py')):
            sys.path.insert(0, HERE)
            import install_reference
            install_reference.install()
    except Exception as err:
        if int(os.environ.get('RANK', '0')) == 0:
            return unavailable('baseline/_ref is not populated and cannot be installed here: {}'.format(err))
        return 0

    n_gpus = args.gpus
    stages, replicas, n_nodes = _topology(args.model, n_gpus)
    child = os.environ.get('RVN_REF_CHILD') == '1'
    if n_gpus == 1 and not child and 'RVN_REF_PORT' not in os.environ:
        os.environ['RVN_REF_PORT'] = str(20000 + os.getpid() % 20000)
    if n_gpus == 1:
        rank = 1 if child else 0
        local_rank = 0
        master_port = int(os.environ['RVN_REF_PORT'])
    else:
        rank, local_rank = int(os.environ['RANK']), int(os.environ.get('LOCAL_RANK', os.environ['RANK']))
        master_port = int(os.environ.get('MASTER_PORT', '29500'))
        if int(os.environ.get('WORLD_SIZE', '1')) != n_gpus:
            return unavailable('launch with torchrun --nproc-per-node {}'.format(n_gpus)) if rank == 0 else 0
    use_cuda = not args.ref_cpu
    if use_cuda:
        os.environ['CUDA_VISIBLE_DEVICES'] = str(local_rank)      # before torch initialises CUDA: every provider sees `cuda`
    paths = [SHIMS, REF, os.path.join(REF, 'examples', 'bert')]
    os.environ['PYTHONPATH'] = os.pathsep.join(paths + [p for p in os.environ.get('PYTHONPATH', '').split(os.pathsep) if p])
    for p in reversed(paths):
        sys.path.insert(0, p)
    workdir = os.path.join(os.environ.get('TMPDIR', '/tmp'), 'rvn_refbench_{}'.format(master_port))
    os.makedirs(workdir, exist_ok=True)
    helper = None
    if n_gpus == 1 and not child:
        env = dict(os.environ, RVN_REF_CHILD='1')
        helper = subprocess.Popen([sys.executable, os.path.join(os.path.dirname(HERE), 'bench.py')] + sys.argv[1:], env=env)
    os.chdir(workdir)

    import torch
    import torch.distributed as dist
    if use_cuda and not torch.cuda.is_available():
        if rank == 0:
            unavailable('no CUDA device visible')
        return 0
    device = torch.device('cuda') if use_cuda else torch.device('cpu')
    if n_gpus == 1:       # two providers on one GPU, launched by us
        dist.init_process_group('gloo', init_method='tcp://127.0.0.1:{}'.format(master_port + 1), rank=rank, world_size=n_nodes)
    else:                 # torchrun: its agent store
        dist.init_process_group('gloo')
    # ----- everything the reference prints goes to a log file; stdout carries the one JSON line -------------------
    log_path = os.path.join(workdir, 'provider_{}.log'.format(rank))
    real_stdout = os.dup(1)
    sys.stdout.flush()
    log_fd = os.open(log_path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    os.dup2(log_fd, 1)
    log = open(log_path, 'a')

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + '\n').encode())

    import threading

    def thread_died(hook_args):                        # the reference has no error handling: a dead worker thread = a silent hang
        import traceback
        traceback.print_exception(hook_args.exc_type, hook_args.exc_value, hook_args.exc_traceback, file=log)
        log.flush()
        with open(os.path.join(workdir, 'FAILED_{}'.format(rank)), 'w') as fh:
            fh.write('provider {}: {!r}'.format(rank, hook_args.exc_value))

    def watchdog():                                    # any provider's failure ends every provider (rank 0 reports it)
        import glob
        limit = time.time() + float(os.environ.get('RVN_REF_TIMEOUT', '1500'))
        while True:
            time.sleep(0.5)
            marks = glob.glob(os.path.join(workdir, 'FAILED_*'))
            if marks or time.time() > limit:
                why = open(marks[0]).read() if marks else 'no result within RVN_REF_TIMEOUT'
                if rank == 0:
                    emit({'impl': 'reference', 'unavailable': 'reference run failed: {} (logs: {})'.format(why, workdir)})
                _leave(0)

    for stale in [f for f in os.listdir(workdir) if f.startswith('FAILED_')] if rank == 0 else []:
        os.remove(os.path.join(workdir, stale))
    threading.excepthook = thread_died
    threading.Thread(target=watchdog, daemon=True).start()
    try:
        result = _provider(args, datasets, ClockSampler, rank, local_rank, n_nodes, stages, replicas, master_port, device, log)
    except BaseException as err:                       # a crashed provider must not hang its peers
        import traceback
        traceback.print_exc(file=log)
        log.flush()
        with open(os.path.join(workdir, 'FAILED_{}'.format(rank)), 'w') as fh:
            fh.write('provider {}: {!r}'.format(rank, err))
        time.sleep(5)                                  # the watchdog reports and exits
        _leave(0)
    if rank == 0 and result is not None:
        emit(result)
    if helper is not None:
        try:
            helper.wait(30)
        except subprocess.TimeoutExpired:
            helper.kill()
    # the reference spawns a Manager and a gRPC server process per node and never joins them
    _leave(0)


def _provider(args, datasets, ClockSampler, rank, local_rank, n_nodes, stages, replicas, master_port, device, log):
    import torch
    import torch.distributed as dist
    SyntheticBert, SyntheticImages = datasets
    use_cuda = device.type == 'cuda'
    steps, warmup = args.steps, max(3, args.warmup)
    base_port = 30000 + (master_port % 2000) * 10
    plan_info = [None]
    if rank == 0:
        plan_info[0] = _plan(args.model, stages, replicas, n_nodes, base_port, device, log)
    dist.barrier()
    import ravnest                                      # noqa: F401  (the unmodified reference)
    from ravnest import Node, Trainer, set_seed
    set_seed(42)
    with open('node_data/nodes/node_{}.json'.format(rank)) as fh:
        meta = json.load(fh)
    replica = meta['rank']
    node_type = meta['node_type']
    upd = args.update_frequency or (16 if args.model == 'bert' else 1)
    total = warmup + steps

    # ---- provider script (cf. reference examples/bert/provider.py, examples/resnet50/provider.py) ------------------
    if args.model == 'bert':
        from torch_optimizer import Lamb
        from bert_trainer import BERT_Trainer           # the reference's own custom trainer (examples/bert/bert_trainer.py)
        vocab = 1000 if os.environ.get('RVN_REF_TINY') == '1' else 30522
        pool = SyntheticBert(4, 8, 512, vocab, device=None, seed=replica).items
        batch = 8
        loss_fct = torch.nn.CrossEntropyLoss()

        def loss_fn(y_pred, targets):
            labels = targets['labels'].to(device)
            sentence_order_label = targets['sentence_order_label'].to(device)
            masked_lm_loss = loss_fct(y_pred[0].view(-1, vocab), labels.view(-1))
            sentence_order_loss = loss_fct(y_pred[1].view(-1, 2), sentence_order_label.view(-1))
            total_loss = masked_lm_loss + sentence_order_loss
            total_loss.div_(upd)
            return total_loss

        node_kwargs = dict(optimizer=Lamb, optimizer_params={'lr': 0.00176, 'eps': 1e-6, 'debias': True, 'weight_decay': 0.01,
                                                             'clamp_value': 10000.0}, update_frequency=upd)
        make_trainer = lambda node, loader: BERT_Trainer(node=node, train_loader=loader, epochs=1)
        input_bytes = sum(pool[0][k].numel() * pool[0][k].element_size() for k in ('input_ids', 'attention_mask', 'token_type_ids'))
        as_pairs = False
        metric, unit = 'bert_base_pretraining_samples_per_sec', 'samples/s'
        opt_name, model_desc = 'LAMB', 'transformers BertForPreTraining(BertConfig()) (110.1M)'
    else:
        size, classes, batch = {'resnet50': (224, 200, 100), 'inception': (32, 10, 64), 'sorter': (0, 3, 64), 'cnn': (8, 10, 64)}[args.model]
        if args.model == 'sorter':
            pool = SyntheticImages(4, batch, 3, device=None, seed=replica, tokens=11).items
            ce = torch.nn.CrossEntropyLoss(ignore_index=-1)

            def loss_fn(preds, targets):
                return ce(preds.view(-1, preds.size(-1)), targets[1].to(device).view(-1))
            node_kwargs = dict(optimizer=torch.optim.Adam, optimizer_params={}, update_frequency=upd)
            opt_name = 'Adam'
        else:
            pool = SyntheticImages(2, batch, classes, device=None, seed=replica, size=size).items
            if args.model == 'cnn':                    # 1x8x8 digits (plumbing check only)
                pool = [(x[:, :1].contiguous(), y) for x, y in pool]
            loss_fct = torch.nn.CrossEntropyLoss()

            def loss_fn(preds, targets):
                return loss_fct(preds, targets[1].to(device))
            node_kwargs = dict(optimizer=torch.optim.SGD, optimizer_params={'lr': 0.01, 'momentum': 0.9, 'weight_decay': 0.0005},
                               update_frequency=upd)
            opt_name = 'SGD-momentum'
        make_trainer = lambda node, loader: Trainer(node=node, train_loader=loader, epochs=1, batch_size=batch)
        input_bytes = pool[0][0].numel() * pool[0][0].element_size()
        as_pairs = True
        metric, unit = {'resnet50': ('resnet50_images_per_sec', 'images/s'), 'inception': ('inception_v3_images_per_sec', 'images/s'),
                        'sorter': ('gpt_sorter_samples_per_sec', 'samples/s'), 'cnn': ('cnn_samples_per_sec', 'samples/s')}[args.model]
        model_desc = {'resnet50': 'torchvision resnet50(num_classes=200) (23.9M)', 'inception': 'reference models.py Inception3 (21.6M)',
                      'sorter': 'reference minGPT gpt-nano (85.7K)', 'cnn': 'reference CNN_Net'}[args.model]

    labels = CyclicLoader(pool, total, as_pairs)
    # DP cadence.  BERT (update_frequency 16): reduce_factor=1 = parameter averaging across replicas every 16 backwards — the
    # cadence of the native arm's per-update gradient all-reduce.  update_frequency 1 models: the reference's example default
    # (reduce_factor=None -> once per epoch, here once per W+K steps, plus the averaging at the end of Trainer.train()); a
    # per-step cadence would put the reference's `wait_for_backwards` poll (time.sleep(1), ravnest/node.py:387-388,:709-710)
    # on every step and measure that sleep instead of the system.
    reduce_factor = 1 if upd > 1 else None
    node = Node(name='node_{}'.format(rank), criterion=loss_fn, labels=labels, device=device, reduce_factor=reduce_factor,
                loss_filename='losses_{}.txt'.format(rank), **node_kwargs)
    dist.barrier()
    is_root = node_type == 'root'

    def phase(n_steps, upto):
        """Root: one Trainer.train() over n_steps batches; the others wait until their own backward count reaches `upto`."""
        if is_root:
            make_trainer(node, CyclicLoader(pool, n_steps, as_pairs)).train()
        else:
            while node.n_backwards < upto:
                time.sleep(0.0005)

    def sync():
        if use_cuda:
            torch.cuda.synchronize()
        dist.barrier()

    phase(warmup, warmup)
    sync()
    sampler = None
    if rank == 0 and use_cuda:
        sampler = ClockSampler(local_rank)
        sampler.start()
    if use_cuda:
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
    t0 = time.time()
    phase(steps, total)
    if use_cuda:
        end.record()
        torch.cuda.synchronize()
        dev_ms = start.elapsed_time(end)
    else:
        dev_ms = (time.time() - t0) * 1e3
    wall_ms = (time.time() - t0) * 1e3
    clocks = sampler.summary() if sampler is not None else None
    times = [None] * n_nodes
    dist.all_gather_object(times, (dev_ms, wall_ms, node_type, replica))
    loss_line = None
    try:
        with open('losses_{}.txt'.format(rank)) as fh:
            lines = fh.read().split()
            loss_line = float(lines[-1]) if lines else None
    except OSError:
        pass
    losses = [None] * n_nodes
    dist.all_gather_object(losses, loss_line)
    dist.barrier()
    if rank != 0:
        return None
    ms = max(t[0] for t in times)
    value = replicas * batch * steps / (ms / 1e3)
    return {
        'metric': metric, 'value': value, 'unit': unit, 'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup,
        'ms_per_step': ms / steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'fp32 (the reference has no reduced-precision compute path)',
        'data': 'synthetic (random tokens/images of the benchmark shape; random-init weights)',
        'impl': 'reference',
        'config': {'model': model_desc, 'global_batch': replicas * batch, 'per_replica_batch': batch,
                   'seq_len': 512 if args.model == 'bert' else None,
                   'parallelism': 'pp{}xdp{}'.format(stages, replicas) + (' (2 providers share the single GPU: the reference needs >= 2 nodes per cluster)' if args.gpus == 1 else ''),
                   'optimizer': opt_name, 'update_frequency': upd,
                   'schedule': 'reference: async pipeline, weight stashing, recompute; periodic parameter averaging (ring over gRPC), reduce_factor={}'.format(reduce_factor),
                   'transport': 'gRPC/TCP on localhost, pickled CPU tensors (the reference has no other transport)',
                   'l2': 'working set exceeds L2; no explicit flush', 'planner': plan_info[0]},
        'clocks': clocks,
        'e2e': {'value': value, 'unit': unit, 'h2d_bytes_per_step': input_bytes, 'd2h_bytes_per_step': 4,
                'ms_per_step': ms / steps,
                'note': 'the reference path is end to end by construction: host loader -> .to(device) every step, loss.item() every step; '
                        'activations additionally cross PCIe twice per hop'},
        'gpu_launches': 0,
        **({'invalid': 'RVN_REF_TINY plumbing run'} if os.environ.get('RVN_REF_TINY') == '1' else {}),
        'wall_ms_per_step': max(t[1] for t in times) / steps,
        'per_provider_ms': [{'node': i, 'type': t[2], 'replica': t[3], 'device_ms': round(t[0], 2)} for i, t in enumerate(times)],
        'last_loss_per_rank': losses,
    }
