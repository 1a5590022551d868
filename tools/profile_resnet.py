"""Per-kernel device-time table of one ResNet-50 (bs 100, bf16, channels_last) training step -> gpurun_out/resnet_profile_<impl>.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from ravnest_b200 import models, ops  # noqa: E402
from ravnest_b200.compute import Compute  # noqa: E402

impl = sys.argv[1] if len(sys.argv) > 1 else 'native'
ops.set_backend('torch' if impl == 'baseline' else 'native')
dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
torch.backends.cudnn.benchmark = True
model = models.resnet50(num_classes=200).to(dev, torch.bfloat16).to(memory_format=torch.channels_last)
x = torch.randn(100, 3, 224, 224, device=dev)
y = torch.randint(0, 200, (100,), device=dev)
lossf = torch.nn.CrossEntropyLoss()
eng = Compute(model, optimizer_cls=torch.optim.SGD, optimizer_params={'lr': 0.01, 'momentum': 0.9, 'weight_decay': 5e-4}, device=dev,
              criterion=lambda o, t: lossf(o.float(), t), compute_dtype=torch.bfloat16, loss_filename=None, is_last=True, recompute=False,
              fused_optimizer=impl != 'baseline', channels_last=True)
for _ in range(4):
    eng.loss_step([x], y, input_grads=False)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    eng.loss_step([x], y, input_grads=False)
e.record()
torch.cuda.synchronize()
print('eager ms/step', s.elapsed_time(e) / 5, 'images/s', 100 * 5 / s.elapsed_time(e) * 1e3)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.loss_step([x], y, input_grads=False)
    torch.cuda.synchronize()
os.makedirs('gpurun_out', exist_ok=True)
with open('gpurun_out/resnet_profile_{}.txt'.format(impl), 'w') as fh:
    fh.write(prof.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=100))
print(open('gpurun_out/resnet_profile_{}.txt'.format(impl)).read()[:200])
from ravnest_b200.ops import native  # noqa: E402
print('native calls', dict(native.calls), 'fallbacks', dict(native.fallbacks))
