import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ravnest_b200.ops import _ext
C = _ext.load()
torch.manual_seed(0)
for (N, Cc, H, W, k, s, p, is_max) in [(1, 8, 5, 5, 3, 1, 1, False), (1, 8, 5, 5, 3, 1, 1, True), (2, 192, 17, 17, 3, 1, 1, False), (2, 64, 8, 8, 3, 2, 1, False)]:
    x = torch.randn(N, H, W, Cc, device='cuda').bfloat16()
    outs = C.pool_fwd(x, k, s, p, is_max)
    y = outs[0]
    dy = torch.randn_like(y)
    dx = C.pool_bwd(dy, outs[1] if is_max else None, H, W, k, s, p, is_max)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    fn = torch.nn.functional.max_pool2d if is_max else torch.nn.functional.avg_pool2d
    yr = fn(xr, k, s, p)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    ref = xr.grad.permute(0, 2, 3, 1)
    print((N, Cc, H, W, k, s, p, is_max), 'fwd err', float((y.float() - yr.permute(0, 2, 3, 1)).abs().max()), 'bwd err', float((dx.float() - ref).abs().max()),
          'ratio', float(dx.float().norm() / ref.norm()))
    if Cc == 8 and not is_max:
        print(dx.float()[0, :, :, 0]); print(ref[0, :, :, 0])
