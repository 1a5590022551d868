"""Minimal repro for the fused-MLP backward crash: same GEMM calls, first on the main thread, then via autograd."""
import faulthandler
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.enable()
import torch  # noqa: E402

from ravnest_b200 import ops  # noqa: E402
from ravnest_b200.ops import _ext  # noqa: E402

C = _ext.load()
dev = 'cuda'
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dy = (torch.randn(M, 768, device=dev) * 0.1).bfloat16()
w2 = (torch.randn(768, 3072, device=dev) * 0.03).bfloat16()
z = torch.randn(M, 3072, device=dev).bfloat16()
for bn in (256, 128, 0):
    out = C.gemm(dy, False, w2, True, None, 1, z, False, bn, 1)[0]
    torch.cuda.synchronize()
    print('direct dgrad+dact ok bn', bn, float(out.float().abs().mean()), flush=True)
g = torch.zeros(768, 3072, device=dev).bfloat16()
h = torch.randn(M, 3072, device=dev).bfloat16()
C.gemm(dy, True, h, True, None, 0, g, False, 0, 2, g)
torch.cuda.synchronize()
print('direct wgrad accumulate ok', float(g.float().abs().mean()), flush=True)
x = (torch.randn(4, M // 4, 768, device=dev)).bfloat16().requires_grad_(True)
w1 = (torch.randn(3072, 768, device=dev) * 0.03).bfloat16().requires_grad_(True)
b1 = torch.zeros(3072, device=dev).bfloat16().requires_grad_(True)
w2p = (torch.randn(768, 3072, device=dev) * 0.03).bfloat16().requires_grad_(True)
b2 = torch.zeros(768, device=dev).bfloat16().requires_grad_(True)
for it in range(2):
    y = ops.mlp(x, w1, b1, w2p, b2, 'gelu')
    y.backward(torch.randn_like(y))
    torch.cuda.synchronize()
    print('autograd mlp ok iter', it, flush=True)
