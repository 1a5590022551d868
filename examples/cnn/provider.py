"""CNN on sklearn digits, 3 providers on localhost (plumbing config #1).  Run `python cluster_formation.py cnn`
first, then start this script three times:  python examples/cnn/provider.py node_0 | node_1 | node_2
(reference examples/cnn/provider.py: copy to provider_N.py and edit the name)."""
import sys

import numpy as np
import torch
from sklearn import datasets
from sklearn.model_selection import train_test_split
from torch.utils.data import DataLoader

from ravnest_b200 import Node, Trainer, set_seed

set_seed(42)


def loaders():
    data = datasets.load_digits()
    y = np.eye(10, dtype=np.float32)[data.target]
    X_train, X_test, y_train, y_test = train_test_split(data.data.astype(np.float32), y, test_size=0.4, random_state=1)
    X_train, X_test = X_train.reshape(-1, 1, 8, 8), X_test.reshape(-1, 1, 8, 8)
    g = torch.Generator().manual_seed(42)       # root and leaf must iterate in the same order (docs/train.rst:223-229)
    train = DataLoader(list(zip(X_train, torch.tensor(y_train))), generator=g, shuffle=True, batch_size=64)
    val = DataLoader(list(zip(X_test, torch.tensor(y_test))), shuffle=False, batch_size=64)
    return train, val


def loss_fn(preds, targets):
    return torch.nn.functional.mse_loss(preds, targets[1])


if __name__ == '__main__':
    name = sys.argv[1] if len(sys.argv) > 1 else 'node_0'
    train_loader, val_loader = loaders()
    node = Node(name=name, optimizer=torch.optim.Adam, device=torch.device('cpu'), criterion=loss_fn, labels=train_loader,
                test_labels=val_loader)
    trainer = Trainer(node=node, train_loader=train_loader, val_loader=val_loader, val_freq=64, epochs=100, batch_size=64,
                      inputs_dtype=torch.float32, save=True)
    trainer.train()
    trainer.evaluate()
