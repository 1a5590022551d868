"""Custom Trainer that feeds keyword inputs (reference examples/bert/bert_trainer.py:3-17, docs/features.rst:12-59)."""
import ravnest_b200 as ravnest


class BERT_Trainer(ravnest.Trainer):
    def __init__(self, node=None, train_loader=None, epochs=1):
        super().__init__(node=node, train_loader=train_loader, epochs=epochs)

    def train(self):
        if not self.prelim_checks():
            return
        for _epoch in range(self.epochs):
            for batch in self.train_loader:
                self.node.forward_compute(input_ids=batch['input_ids'], token_type_ids=batch['token_type_ids'],
                                          attention_mask=batch['attention_mask'])
            self.node.wait_for_backwards()      # to be called at the end of every epoch
