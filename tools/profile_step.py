"""Per-kernel device-time table of ONE eager BERT-base training step (solo stage) -> gpurun_out/step_profile.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import ravnest_b200 as rv  # noqa: E402
from ravnest_b200 import models, ops  # noqa: E402
from ravnest_b200.compute import Compute  # noqa: E402

impl = sys.argv[1] if len(sys.argv) > 1 else 'native'
ops.set_backend('torch' if impl == 'baseline' else 'native')
dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
model = models.BertForPreTraining().to(dev, torch.bfloat16)
B, S = 8, 512
ids = torch.randint(0, 30522, (B, S), device=dev)
mask = torch.ones(B, S, dtype=torch.long, device=dev)
tt = torch.zeros(B, S, dtype=torch.long, device=dev)
labels = torch.where(torch.rand(B, S, device=dev) < 0.15, ids, torch.full_like(ids, -100))
sop = torch.randint(0, 2, (B,), device=dev)
eng = Compute(model, optimizer_cls=rv.Lamb, optimizer_params={'lr': 1e-3, 'weight_decay': 0.01}, device=dev,
              criterion=lambda o, t: models.bert_pretraining_loss(o, t[0], t[1]), compute_dtype=torch.bfloat16, loss_filename=None,
              is_last=True, recompute=False, update_frequency=16, fused_optimizer=impl != 'baseline')
for _ in range(3):
    eng.loss_step([ids, mask, tt], (labels, sop), input_grads=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.loss_step([ids, mask, tt], (labels, sop), input_grads=False)
    torch.cuda.synchronize()
os.makedirs('gpurun_out', exist_ok=True)
with open('gpurun_out/step_profile_{}.txt'.format(impl), 'w') as fh:
    fh.write(prof.key_averages().table(sort_by='cuda_time_total', row_limit=60, max_name_column_width=90))
print(open('gpurun_out/step_profile_{}.txt'.format(impl)).read()[-6000:])
