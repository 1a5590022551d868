"""Single-GPU check of the root/stem stage path (Compute.forward + Compute.backward, eager and graphed): which wgrad path runs
(accumulate-into-bucket vs materialise + AccumulateGrad), and the per-pass device time of a half-BERT stage."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ravnest_b200 as rv  # noqa: E402
from ravnest_b200 import models, ops  # noqa: E402
from ravnest_b200.compute import Compute  # noqa: E402
from ravnest_b200.ops import native  # noqa: E402

dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
ops.set_backend('native')
cfg = models.BertConfig(num_hidden_layers=6)
full = models.BertForPreTraining(cfg)


class Half(torch.nn.Module):            # embeddings + 6 layers -> hidden states
    def __init__(self, m):
        super().__init__()
        self.bert = m.bert

    def forward(self, ids, mask, tt):
        out = self.bert(ids, mask, tt)
        return out[0] if isinstance(out, (tuple, list)) else out


stage = Half(full).to(dev, torch.bfloat16)
for graphs in (False, True):
    eng = Compute(stage, optimizer_cls=rv.Lamb, optimizer_params={'lr': 1e-3}, device=dev, compute_dtype=torch.bfloat16, loss_filename=None,
                  is_last=False, recompute=False, update_frequency=16, cluster_length=2, use_cuda_graphs=graphs)
    B, S = 8, 512
    ids = torch.randint(0, 30522, (B, S), device=dev)
    mask = torch.ones(B, S, dtype=torch.long, device=dev)
    tt = torch.zeros(B, S, dtype=torch.long, device=dev)
    for step in range(14):
        before = dict(native.calls)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = eng.forward(step, [ids, mask, tt], input_grads=False)
        out2 = eng.forward(step + 100, [ids, mask, tt], input_grads=False)       # two passes in flight, like the pipeline
        h = out[0] if isinstance(out, (tuple, list)) else out
        eng.backward(step, {0: torch.randn_like(h)})
        eng.backward(step + 100, {0: torch.randn_like(h)})
        b.record()
        torch.cuda.synchronize()
        d = {k: native.calls[k] - before.get(k, 0) for k in native.calls if native.calls[k] != before.get(k, 0)}
        print('graphs', graphs, 'step', step, 'ms/2 passes', round(a.elapsed_time(b), 3), {k: d[k] for k in d if 'wgrad' in k or 'colsum' in k})
