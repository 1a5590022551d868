"""Trainer for models that take keyword inputs (dict batches), as the reference's BERT example needs one
(examples/bert/bert_trainer.py, docs/features.rst "custom trainer").

``KeywordBatchTrainer`` forwards a configurable subset of every batch's keys as keyword arguments of the model; the remaining
keys (labels ...) never leave the host that owns the label loader.  ``BERT_Trainer`` is that trainer bound to BERT's three inputs.
"""
from typing import Iterable, Sequence

import ravnest_b200 as ravnest


class KeywordBatchTrainer(ravnest.Trainer):
    input_keys: Sequence[str] = ()

    def __init__(self, node=None, train_loader=None, epochs=1, input_keys: Iterable[str] = None, **trainer_kwargs):
        super().__init__(node=node, train_loader=train_loader, epochs=epochs, **trainer_kwargs)
        if input_keys is not None:
            self.input_keys = tuple(input_keys)

    def _submit(self, batch) -> None:
        self.node.forward_compute(**{key: batch[key] for key in self.input_keys})

    def train(self):
        if not self.prelim_checks():            # stem / leaf nodes park here until the root announces the end of training
            return
        remaining = self.epochs
        while remaining > 0:
            for batch in self.train_loader:
                self._submit(batch)
            self.node.wait_for_backwards()      # epoch boundary: drain the pipeline before the next pass over the data
            remaining -= 1


class BERT_Trainer(KeywordBatchTrainer):
    input_keys = ('input_ids', 'token_type_ids', 'attention_mask')
