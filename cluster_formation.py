"""Planner driver: pick a model block, run `python cluster_formation.py [cnn|resnet50|inception|sorter|bert] [SxR]`.

Same role as the reference's cluster_formation.py:1-65 (one block per model) but selectable from the
command line instead of by (un)commenting code.  Reads node_data/node_configs.json unless a grid `SxR`
(stages x replicas on one B200 box) is given.
"""
import sys

import torch

from ravnest_b200 import clusterize, default_gpu_configs, models, set_seed

set_seed(42)
which = sys.argv[1] if len(sys.argv) > 1 else 'cnn'
grid = sys.argv[2] if len(sys.argv) > 2 else None
kwargs = {}
if grid:
    s, r = (int(v) for v in grid.lower().split('x'))
    kwargs = dict(topology={'stages': s, 'replicas': r}, node_configs=default_gpu_configs(s * r))

if which == 'cnn':
    clusterize(model=models.CNN_Net(), example_args=(torch.rand((64, 1, 8, 8)),), verbose=True, **kwargs)
elif which == 'resnet50':
    clusterize(model=models.resnet50(num_classes=200), example_args=(torch.rand((4, 3, 224, 224)),), verbose=True, **kwargs)
elif which == 'inception':
    clusterize(model=models.inception_v3(), example_args=(torch.rand((64, 3, 32, 32)),), verbose=True, **kwargs)
elif which == 'sorter':
    cfg = models.GPTConfig(model_type='gpt-nano', vocab_size=3, block_size=11)
    clusterize(model=models.GPT(cfg), example_args=(torch.randint(0, 2, (64, 11), dtype=torch.int64),), pass_data=True, verbose=True, **kwargs)
elif which == 'bert':
    B, S = 2, 512
    clusterize(model=models.BertForPreTraining(models.BertConfig()),
               example_kwargs={'input_ids': torch.randint(0, 30522, (B, S)), 'attention_mask': torch.ones(B, S, dtype=torch.long),
                               'token_type_ids': torch.zeros(B, S, dtype=torch.long)}, verbose=True, **kwargs)
else:
    raise SystemExit('unknown model {!r}'.format(which))
