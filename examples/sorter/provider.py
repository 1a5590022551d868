"""GPT-Sorter: minGPT gpt-nano learning to sort (reference examples/sorter/provider.py: Adam, CE ignore_index -1,
batch 64, 1 epoch).  `python cluster_formation.py sorter` then run once per node name, or use a grid on GPUs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
from torch.utils.data import DataLoader  # noqa: E402

from examples.sorter.dataset import SortDataset  # noqa: E402
from ravnest_b200 import Node, Trainer, set_seed  # noqa: E402

set_seed(42)


def sorter_criterion(outputs, targets):
    return torch.nn.functional.cross_entropy(outputs.reshape(-1, outputs.size(-1)).float(), targets[1].reshape(-1).to(outputs.device),
                                             ignore_index=-1)


if __name__ == '__main__':
    name = sys.argv[1] if len(sys.argv) > 1 else 'node_0'
    data = SortDataset('train', n=int(os.environ.get('SAMPLES', 6400)))
    loader = DataLoader(data, shuffle=False, batch_size=64)
    node = Node(name=name, optimizer=torch.optim.Adam, criterion=sorter_criterion, labels=loader, device=torch.device('cpu'))
    trainer = Trainer(node=node, train_loader=loader, save=True, epochs=1, batch_size=64)
    trainer.train()
    trainer.finish()                # root: stop the pipeline; stem / leaf: return once the root has done so
