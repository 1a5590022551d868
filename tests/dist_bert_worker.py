"""torchrun worker: small BERT pre-training over the torch.distributed transport (gloo on CPU), keyword inputs, dict targets,
the additive key bias relayed through a stem, LAMB with gradient accumulation — the path bench.py drives on GPUs."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.utils.data import DataLoader, Dataset  # noqa: E402

import ravnest_b200 as rv  # noqa: E402
from ravnest_b200 import models  # noqa: E402
from ravnest_b200.runtime.dist_transport import init_distributed  # noqa: E402

VOCAB, SEQ = 128, 16


class Tokens(Dataset):
    def __len__(self):
        return 24

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(i)
        ids = torch.randint(1, VOCAB, (SEQ,), generator=g)
        mask = torch.ones(SEQ, dtype=torch.long)
        if os.environ.get("RAGGED", "1") == "1":
            mask[SEQ - (i % 5):] = 0                   # ragged padding
        labels = torch.full((SEQ,), -100, dtype=torch.long)
        labels[::4] = ids[::4]
        return {'input_ids': ids, 'attention_mask': mask, 'token_type_ids': torch.zeros(SEQ, dtype=torch.long), 'labels': labels,
                'sentence_order_label': torch.tensor(i % 2)}


def loss_fn(pred, targets):
    return models.bert_pretraining_loss(pred, targets['labels'], targets['sentence_order_label']) / 2


class KwTrainer(rv.Trainer):
    def train(self):
        if not self.prelim_checks():
            return
        for _ in range(self.epochs):
            for batch in self.train_loader:
                self.node.forward_compute(input_ids=batch['input_ids'], token_type_ids=batch['token_type_ids'],
                                          attention_mask=batch['attention_mask'])
            self.node.wait_for_backwards()


def main():
    rank, world, dev = init_distributed()
    rv.set_seed(7)
    loader = DataLoader(Tokens(), batch_size=4, shuffle=False)
    node = rv.Node(name='node_{}'.format(rank), optimizer=rv.Lamb, optimizer_params={'lr': 5e-3, 'weight_decay': 0.01}, update_frequency=2,
                   criterion=loss_fn, labels=loader, device=dev, recompute=False, loss_filename='losses_{}.txt'.format(rank))
    trainer = KwTrainer(node=node, train_loader=loader, epochs=3)
    trainer.train()
    if trainer.is_driver:
        node.wait_for_backwards()
        time.sleep(0.2)
        node.stop()
    node._stop.wait(30)
    time.sleep(0.3)
    node.compute_session.flush_losses()
    with open('result.json.{}'.format(rank), 'w') as fh:
        json.dump({'rank': rank, 'node_type': node.node_type, 'n_backwards': node.n_backwards,
                   'finite': bool(torch.isfinite(node.compute_session.store.master).all()),
                   'n_inputs': len(node.plan.inputs) if hasattr(node.plan, 'inputs') else -1}, fh)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
