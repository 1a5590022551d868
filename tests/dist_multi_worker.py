"""torchrun worker for the remaining BASELINE.json pipeline configs on CPU/gloo: `--model inception` (3-stage pipeline, image
batches, integer class targets) and `--model sorter` (pure pipeline of the minGPT sorter, plain-tensor labels)."""
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


def sort_data(n=128, seed=0):
    g = torch.Generator().manual_seed(seed)
    inp = torch.randint(0, 3, (n, 6), generator=g)
    sol = torch.sort(inp, dim=1)[0]
    cat = torch.cat([inp, sol], dim=1)
    x, y = cat[:, :-1].clone(), cat[:, 1:].clone()
    y[:, :5] = -1
    return x, y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--model', required=True)
    ap.add_argument('--epochs', type=int, default=2)
    args = ap.parse_args()
    rank, world, dev = init_distributed()
    rv.set_seed(11)
    if args.model == 'sorter':
        x, y = sort_data()
        loader = DataLoader(list(zip(x, y)), shuffle=False, batch_size=32)
        crit = lambda out, tgt: torch.nn.functional.cross_entropy(out.reshape(-1, out.size(-1)), tgt[1].reshape(-1), ignore_index=-1)  # noqa: E731
        node = rv.Node(name='node_{}'.format(rank), optimizer=torch.optim.Adam, optimizer_params={'lr': 3e-3}, criterion=crit, labels=y,
                       label_batch_size=32, device=dev, recompute=False, loss_filename='losses_{}.txt'.format(rank))
        trainer = rv.Trainer(node=node, train_loader=loader, epochs=args.epochs)
    else:
        g = torch.Generator().manual_seed(3)
        xs, ys = torch.randn(24, 3, 32, 32, generator=g), torch.randint(0, 10, (24,), generator=g)
        loader = DataLoader(list(zip(xs, ys)), shuffle=False, batch_size=8)
        ce = torch.nn.CrossEntropyLoss()
        node = rv.Node(name='node_{}'.format(rank), optimizer=torch.optim.SGD, optimizer_params={'lr': 0.01, 'momentum': 0.9}, labels=loader,
                       criterion=lambda out, tgt: ce(out, tgt[1]), device=dev, recompute=False, loss_filename='losses_{}.txt'.format(rank))
        trainer = rv.Trainer(node=node, train_loader=loader, epochs=args.epochs, inputs_dtype=torch.float32)
    trainer.train()
    if trainer.is_driver:
        node.wait_for_backwards()
        time.sleep(0.2)
        node.stop()
    node._stop.wait(60)
    time.sleep(0.3)
    node.compute_session.flush_losses()
    with open('result.json.{}'.format(rank), 'w') as fh:
        json.dump({'rank': rank, 'node_type': node.node_type, 'n_backwards': node.n_backwards,
                   'finite': bool(torch.isfinite(node.compute_session.store.master).all())}, fh)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
