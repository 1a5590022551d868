"""Worker used by the multi-process tests / examples: one rank == one node."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.utils.data import DataLoader  # noqa: E402

import ravnest_b200 as rv  # noqa: E402
from ravnest_b200.runtime.dist_transport import init_distributed  # noqa: E402


def digits_like(n=256, seed=0):
    g = torch.Generator().manual_seed(seed)
    y = torch.randint(0, 10, (n,), generator=g)
    x = torch.randn(n, 1, 8, 8, generator=g) * 0.3
    for i in range(n):          # class-dependent pattern so the task is learnable
        x[i, 0, y[i] % 8, :] += 2.0
        x[i, 0, :, (y[i] * 3) % 8] += 1.0
    return x, torch.nn.functional.one_hot(y, 10).float()


def loss_fn(preds, targets):
    return torch.nn.functional.mse_loss(preds, targets[1].to(preds.device))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--epochs', type=int, default=3)
    ap.add_argument('--dp-mode', default='average')
    ap.add_argument('--recompute', type=int, default=1)
    ap.add_argument('--out', default='result.json')
    ap.add_argument('--val', type=int, default=0)
    ap.add_argument('--optimizer', default='adam')
    ap.add_argument('--dump', type=int, default=0, help='write the final master parameters to master.<rank>.pt')
    ap.add_argument('--replan', type=int, default=0, help='after the run: stop every node, cut the plan again with rebalanced shares, build new nodes in the SAME processes and train again')
    args = ap.parse_args()
    rank, world, dev = init_distributed()
    rv.set_seed(42)
    x, y = digits_like()
    cfg = rv.load_node_json_configs('node_{}'.format(rank))
    replica = cfg['rank']
    shard = slice(replica * 128, replica * 128 + 128) if cfg['ring_size'] > 1 else slice(0, 256)
    data = list(zip(x[shard], y[shard]))
    loader = DataLoader(data, batch_size=32, shuffle=False)
    opt_cls, opt_params = (torch.optim.SGD, {'lr': 0.05}) if args.optimizer == 'sgd' else (torch.optim.Adam, {'lr': 2e-3})
    node = rv.Node(name='node_{}'.format(rank), optimizer=opt_cls, optimizer_params=opt_params,
                   criterion=loss_fn, labels=loader, test_labels=loader, device=dev, reduce_factor=2,
                   dp_mode=args.dp_mode, recompute=bool(args.recompute), loss_filename='losses_{}.txt'.format(rank))
    trainer = rv.Trainer(node=node, train_loader=loader, val_loader=loader if args.val else None, val_freq=1, epochs=args.epochs,
                         inputs_dtype=torch.float32)
    trainer.train()
    if trainer.is_driver:
        node.wait_for_backwards()
        time.sleep(0.2)
        node.stop()
    node._stop.wait(30)
    time.sleep(0.3)
    node.compute_session.flush_losses()
    if args.replan:
        # profile-guided re-planning inside a running job (bench.py does this on GPUs): new stage boundaries, new Node objects, same
        # process group and communicators
        from ravnest_b200 import models
        first_run_backwards = node.n_backwards
        torch.distributed.barrier()
        stages = cfg['cluster_length']
        if rank == 0:
            shares = rv.rebalanced_proportions([1.0 / stages] * stages, [1.0 + 0.5 * i for i in range(stages)])
            rv.set_seed(43)
            rv.clusterize(models.CNN_Net(), example_args=(torch.rand(32, 1, 8, 8),), topology={'stages': stages, 'replicas': cfg['ring_size']},
                          node_configs=rv.default_gpu_configs(world), proportions=shares, balance='compute', base_dir=os.environ.get('RAVNEST_HOME', '.'))
        torch.distributed.barrier()
        node = rv.Node(name='node_{}'.format(rank), optimizer=opt_cls, optimizer_params=opt_params,
                       criterion=loss_fn, labels=loader, test_labels=loader, device=dev, reduce_factor=2,
                       dp_mode=args.dp_mode, recompute=bool(args.recompute), loss_filename='losses2_{}.txt'.format(rank))
        trainer = rv.Trainer(node=node, train_loader=loader, epochs=args.epochs, inputs_dtype=torch.float32)
        trainer.train()
        if trainer.is_driver:
            node.wait_for_backwards()
            time.sleep(0.2)
            node.stop()
        node._stop.wait(30)
        time.sleep(0.3)
        node.compute_session.flush_losses()
        assert node.n_backwards == first_run_backwards, (node.n_backwards, first_run_backwards)
    if args.dump:
        torch.save({n: v.clone() for n, v in zip(node.compute_session.store.names, node.compute_session.store.master_views)},
                   'master.{}.pt'.format(rank))
    summary = {'rank': rank, 'node_type': node.node_type, 'n_backwards': node.n_backwards,
               'version': node.compute_session.current_version(),
               'checksum': float(node.compute_session.store.master.double().sum()),
               'header_stats': dict(getattr(node.transport, 'header_stats', {}))}
    with open('{}.{}'.format(args.out, rank), 'w') as fh:
        json.dump(summary, fh)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
