import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ravnest_b200.ops import _ext
C = _ext.load()
torch.manual_seed(0)
F = torch.nn.functional
for (N, Cc, H, W, k, s, p) in [(1, 8, 5, 5, 3, 1, 1), (2, 192, 17, 17, 3, 1, 1), (2, 64, 8, 8, 3, 2, 1)]:
    x = torch.randn(N, H, W, Cc, device='cuda').bfloat16()
    y = C.pool_fwd(x, k, s, p, False)[0]
    dy = torch.randn_like(y)
    dx = C.pool_bwd(dy, None, H, W, k, s, p, False).float()
    refs = {}
    for name, prep in (('cuda_cl', lambda t: t.float().permute(0, 3, 1, 2)), ('cuda_nchw', lambda t: t.float().permute(0, 3, 1, 2).contiguous()),
                       ('cpu_nchw', lambda t: t.float().cpu().permute(0, 3, 1, 2).contiguous())):
        xr = prep(x).requires_grad_(True)
        F.avg_pool2d(xr, k, s, p).backward(prep(dy))
        refs[name] = xr.grad.permute(0, 2, 3, 1).cuda()
    print((N, Cc, H, W, k, s, p), {n: round(float((dx - r).abs().max()), 4) for n, r in refs.items()},
          'cl-vs-nchw', float((refs['cuda_cl'] - refs['cuda_nchw']).abs().max()), 'nchw-vs-cpu', float((refs['cuda_nchw'] - refs['cpu_nchw']).abs().max()))
