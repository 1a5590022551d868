"""Inception-V3 (CIFAR variant) on synthetic CIFAR-shaped data, 3-stage pipeline
(reference examples/inception_v3/provider.py: batch 64, SGD, 10 epochs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
from torch.utils.data import DataLoader, TensorDataset  # noqa: E402

from examples._gpu_common import launch  # noqa: E402
from ravnest_b200 import Node, Trainer, set_seed  # noqa: E402

set_seed(42)
ROOT = os.environ.get('CIFAR10_ROOT')
if ROOT:                                              # torchvision's CIFAR-10 when a copy is on disk (no download offline)
    import torchvision
    import torchvision.transforms as T
    tf = T.Compose([T.ToTensor(), T.Normalize((0.4914, 0.4822, 0.4465), (0.2470, 0.2435, 0.2616))])
    data = torchvision.datasets.CIFAR10(ROOT, train=True, download=False, transform=tf)
    val_data = torchvision.datasets.CIFAR10(ROOT, train=False, download=False, transform=tf)
else:
    N = int(os.environ.get('SAMPLES', 6400))
    data = val_data = TensorDataset(torch.randn(N, 3, 32, 32), torch.randint(0, 10, (N,)))
loss_fct = torch.nn.CrossEntropyLoss()


def loaders():
    g = torch.Generator().manual_seed(42)
    return DataLoader(data, batch_size=64, shuffle=True, generator=g, drop_last=True), DataLoader(val_data, batch_size=64)


def make_node(name, dev):
    train, val = loaders()
    return Node(name=name, optimizer=torch.optim.SGD, optimizer_params={'lr': 0.01, 'momentum': 0.9, 'weight_decay': 0.0005},
                criterion=lambda preds, targets: loss_fct(preds.float(), targets[1].to(preds.device)), labels=train, test_labels=val,
                device=dev, dtype=torch.bfloat16)


def make_trainer(node):
    train, val = loaders()
    return Trainer(node=node, train_loader=train, val_loader=val, epochs=10, batch_size=64, save=True)


if __name__ == '__main__':
    launch(make_node, make_trainer)
