"""ResNet-50 (num_classes=200) on synthetic Tiny-ImageNet-shaped data, 4 stages x 2 replicas on 8 GPUs
(reference examples/resnet50/provider.py: batch 100, SGD lr .01 momentum .9 wd 5e-4, 5 epochs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
from torch.utils.data import DataLoader, TensorDataset  # noqa: E402

from examples._gpu_common import launch  # noqa: E402
from ravnest_b200 import Node, Trainer, set_seed  # noqa: E402

set_seed(42)
ROOT = os.environ.get('TINY_IMAGENET_ROOT')
if ROOT:                                              # the real dataset when a copy is on disk
    from examples.resnet50.TinyImageNet import TinyImageNet
    data, val_data = TinyImageNet(ROOT, 'train'), TinyImageNet(ROOT, 'val')
else:                                                 # no network in the build image: synthetic tensors of the same shape
    N = int(os.environ.get('SAMPLES', 2000))
    data = val_data = TensorDataset(torch.randn(N, 3, 224, 224), torch.randint(0, 200, (N,)))
loss_fct = torch.nn.CrossEntropyLoss()


def loaders():
    g = torch.Generator().manual_seed(42)
    return DataLoader(data, batch_size=100, shuffle=True, generator=g, drop_last=True), DataLoader(val_data, batch_size=100)


def make_node(name, dev):
    train, val = loaders()
    return Node(name=name, optimizer=torch.optim.SGD, optimizer_params={'lr': 0.01, 'momentum': 0.9, 'weight_decay': 0.0005},
                criterion=lambda preds, targets: loss_fct(preds.float(), targets[1].to(preds.device)), labels=train, test_labels=val,
                device=dev, dtype=torch.bfloat16, dp_mode='gradient')


def make_trainer(node):
    train, val = loaders()
    return Trainer(node=node, train_loader=train, val_loader=val, epochs=5, batch_size=100, save=True)


if __name__ == '__main__':
    launch(make_node, make_trainer)
