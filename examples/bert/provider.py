"""BERT-base pre-training (MLM + SOP), bf16, 2 stages x 4 replicas on 8 GPUs, LAMB, update_frequency 16,
linear warm-up (reference examples/bert/provider.py:20-75) on synthetic token data (no dataset files offline).

  python cluster_formation.py bert 2x4
  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/bert/provider.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
from torch.utils.data import DataLoader, Dataset  # noqa: E402

from examples._gpu_common import launch  # noqa: E402
from examples.bert.bert_trainer import BERT_Trainer  # noqa: E402
from ravnest_b200 import Lamb, Node, models, set_seed  # noqa: E402

set_seed(42)
update_frequency, epochs, vocab = 16, int(os.environ.get('EPOCHS', 1)), 30522


class SyntheticMLM(Dataset):
    def __init__(self, n=512, seq=512):
        self.n, self.seq = n, seq

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(i)
        ids = torch.randint(0, vocab, (self.seq,), generator=g)
        labels = torch.full((self.seq,), -100, dtype=torch.long)
        pick = torch.rand(self.seq, generator=g) < 0.15
        labels[pick] = ids[pick]
        return {'input_ids': ids, 'attention_mask': torch.ones(self.seq, dtype=torch.long), 'token_type_ids': torch.zeros(self.seq, dtype=torch.long),
                'labels': labels, 'sentence_order_label': torch.randint(0, 2, (), generator=g)}


def loader():
    g = torch.Generator().manual_seed(42)
    return DataLoader(SyntheticMLM(), batch_size=8, generator=g, shuffle=True, drop_last=True)


def loss_fn(y_pred, targets):
    dev = y_pred[0].device
    total = models.bert_pretraining_loss(y_pred, targets['labels'].to(dev), targets['sentence_order_label'].to(dev))
    return total / update_frequency


def linear_warmup(optimizer, num_warmup_steps, num_training_steps):
    def factor(step):
        if step < num_warmup_steps:
            return step / max(1, num_warmup_steps)
        return max(0.0, (num_training_steps - step) / max(1, num_training_steps - num_warmup_steps))
    return torch.optim.lr_scheduler.LambdaLR(optimizer, factor)


def make_node(name, dev):
    train = loader()
    steps = (len(train) + update_frequency - 1) // update_frequency * epochs
    return Node(name=name, optimizer=Lamb, optimizer_params={'lr': 0.00176, 'eps': 1e-6, 'debias': True, 'weight_decay': 0.01, 'clamp_value': 10000.0},
                update_frequency=update_frequency, lr_scheduler=linear_warmup,
                lr_scheduler_params={'num_warmup_steps': 5000, 'num_training_steps': steps}, lr_step_on_epoch_change=False,
                criterion=loss_fn, labels=train, device=dev, dtype=torch.bfloat16, dp_mode='gradient')


def make_trainer(node):
    return BERT_Trainer(node=node, train_loader=loader(), epochs=epochs)


if __name__ == '__main__':
    launch(make_node, make_trainer)
