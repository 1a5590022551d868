"""Plumbing example #1: the small CNN on scikit-learn's 8 x 8 digits, three providers on localhost over the gRPC transport.

    python cluster_formation.py cnn                      # plan: 3 nodes -> one cluster of 3 stages
    python examples/cnn/provider.py --all                # start node_0, node_1, node_2 as three processes, wait for them
    python examples/cnn/provider.py --node node_1        # ... or start one provider per terminal / host

Settings of the reference example (examples/cnn/provider.py:17-62): Adam, MSE against one-hot targets on the softmax output,
batch 64, 100 epochs, validation through the pipeline, sub-models saved at the end.  The reference asks the user to copy the
script once per node and edit the name; here the node name is an argument.
"""
import argparse
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
from torch.utils.data import DataLoader, TensorDataset  # noqa: E402

import ravnest_b200 as ravnest  # noqa: E402

BATCH = 64


class Digits:
    """60 / 40 split of the 1797 digit images as [N, 1, 8, 8] float tensors with one-hot float targets."""

    def __init__(self, seed: int = 1):
        from sklearn.datasets import load_digits
        raw = load_digits()
        images = torch.tensor(raw.data, dtype=torch.float32).view(-1, 1, 8, 8)
        targets = torch.nn.functional.one_hot(torch.tensor(raw.target), 10).float()
        order = torch.randperm(len(images), generator=torch.Generator().manual_seed(seed))
        cut = int(0.6 * len(images))
        self.train = TensorDataset(images[order[:cut]], targets[order[:cut]])
        self.val = TensorDataset(images[order[cut:]], targets[order[cut:]])

    def loaders(self):
        # the root iterates inputs and the leaf iterates targets: both must see the SAME order, hence the fixed generator
        shuffle = torch.Generator().manual_seed(42)
        return (DataLoader(self.train, batch_size=BATCH, shuffle=True, generator=shuffle),
                DataLoader(self.val, batch_size=BATCH, shuffle=False))


def mse_on_probabilities(outputs, targets):
    """``targets`` is the label loader's batch as yielded: (inputs, one-hot labels)."""
    return torch.nn.functional.mse_loss(outputs, targets[1])


def run_node(name: str, epochs: int) -> None:
    ravnest.set_seed(42)
    train_loader, val_loader = Digits().loaders()
    node = ravnest.Node(name=name, optimizer=torch.optim.Adam, criterion=mse_on_probabilities, labels=train_loader,
                        test_labels=val_loader, device=torch.device('cpu'))
    trainer = ravnest.Trainer(node=node, train_loader=train_loader, val_loader=val_loader, val_freq=64, epochs=epochs,
                              batch_size=BATCH, inputs_dtype=torch.float32, save=True)
    trainer.train()
    trainer.evaluate()
    trainer.finish()                          # root: stop the pipeline (SHUTDOWN travels downstream); stem / leaf: wait for it


def main() -> int:
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    ap.add_argument('node', nargs='?', default=None, help='node name (positional form, e.g. node_0)')
    ap.add_argument('--node', dest='node_opt', default=None)
    ap.add_argument('--all', action='store_true', help='launch node_0 .. node_2 as sub-processes')
    ap.add_argument('--epochs', type=int, default=100)
    args = ap.parse_args()
    if args.all:
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), '--node', 'node_{}'.format(i), '--epochs', str(args.epochs)])
                 for i in range(3)]
        return max(p.wait() for p in procs)
    run_node(args.node_opt or args.node or 'node_0', args.epochs)
    return 0


if __name__ == '__main__':
    sys.exit(main())
