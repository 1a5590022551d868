"""Launch a few instances of one kernel family at the benchmark shapes (for `ncu -k regex:... python tools/ncu_targets.py <target>`).
Targets: fmha, conv, bn, ln, ce, optim, pool, gemm_epi, embed."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ravnest_b200 import ops  # noqa: E402
from ravnest_b200.ops import _ext, native  # noqa: E402

C = _ext.load()
dev = 'cuda'
target = sys.argv[1]
torch.manual_seed(0)


def r(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).bfloat16()


for it in range(2):
    if target == 'fmha':                                   # BERT-base: B=8, S=512, 12 heads of 64, padding mask + dropout
        qkv = r(8, 512, 2304, scale=0.5).requires_grad_(True)
        mask = torch.ones(8, 512, device=dev)
        mask[:, 480:] = 0
        y = ops.attention_qkv(qkv, 12, mask, False, 0.1)
        y.backward(r(8, 512, 768))
    elif target == 'conv':                                 # ResNet-50 layer2 3x3 (batch 100, 28x28, 128 ch) + layer1 3x3
        for (n, c, h, k, s) in [(100, 128, 28, 128, 1), (100, 64, 56, 64, 1), (100, 256, 28, 256, 2)]:
            x = r(n, c, h, h).contiguous(memory_format=torch.channels_last).requires_grad_(True)
            w = r(k, c, 3, 3, scale=0.05).contiguous(memory_format=torch.channels_last).requires_grad_(True)
            y = ops.conv2d(x, w, None, s, 1)
            y.backward(torch.ones_like(y))
    elif target == 'bn':                                   # ResNet-50 layer1 output BN: [100, 256, 56, 56] with residual + ReLU
        x = r(100, 256, 56, 56).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        res = r(100, 256, 56, 56).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        g, b = r(256).requires_grad_(True), r(256).requires_grad_(True)
        rm, rv = torch.zeros(256, device=dev), torch.ones(256, device=dev)
        y = ops.batch_norm_act(x, g, b, rm, rv, res, True, True, 0.1, 1e-5)
        y.backward(torch.ones_like(y))
    elif target == 'ln':
        x, res = r(4096, 768).requires_grad_(True), r(4096, 768).requires_grad_(True)
        g, b = r(768).requires_grad_(True), r(768).requires_grad_(True)
        ops.layer_norm(x, g, b, 1e-12, res, 0.1).backward(r(4096, 768))
    elif target == 'ce':
        logits = r(4096, 30528).requires_grad_(True)
        tgt = torch.randint(0, 30522, (4096,), device=dev)
        ops.cross_entropy(logits, tgt).backward()
    elif target == 'pool':
        x = r(100, 64, 112, 112).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        y = ops.max_pool2d(x, 3, 2, 1)
        y.backward(torch.ones_like(y))
    elif target == 'gemm_epi':                             # FFN: GELU + act' store forward, multiply-by-act' dgrad, accumulate wgrad
        x = r(4096, 768).requires_grad_(True)
        w1, b1 = r(3072, 768, scale=0.03).requires_grad_(True), r(3072).requires_grad_(True)
        w2, b2 = r(768, 3072, scale=0.03).requires_grad_(True), r(768).requires_grad_(True)
        for p in (w1, b1, w2, b2):
            p.grad = torch.zeros_like(p)
        ops.mlp(x, w1, b1, w2, b2, 'gelu').backward(r(4096, 768))
    elif target == 'embed':                                # BERT embeddings: gather, owner-computes backward, dropout
        w = r(30522, 768, scale=0.02).requires_grad_(True)
        w.grad = torch.zeros_like(w)
        idx = torch.randint(0, 30522, (8, 512), device=dev)
        y = ops.dropout(ops.embedding(idx, w, 0), 0.1, True)
        y.backward(r(8, 512, 768))
    elif target == 'gemm':                                 # BERT shapes through the tile / split-K model: fwd (bn256, bn192), dgrad, accumulate wgrad
        x, w_qkv, w_ffn2 = r(4096, 768), r(2304, 768, scale=0.03), r(768, 3072, scale=0.03)
        C.gemm(x, False, w_qkv, False, None, 0, None, False, 0)                     # 4096 x 2304 x 768  (bn256)
        h = r(4096, 3072)
        C.gemm(h, False, w_ffn2, False, None, 0, None, False, 0)                    # 4096 x 768 x 3072  (bn192)
        dy = r(4096, 768)
        C.gemm(dy, False, w_ffn2, True, None, 0, None, False, 0)                    # dgrad 4096 x 3072 x 768
        g = torch.zeros(768, 768, dtype=torch.bfloat16, device=dev)
        C.gemm(dy, True, x, True, None, 0, g, False, 0, 2, g)                       # wgrad 768 x 768 x 4096: split-K + TMA reduce-add
        g2 = torch.zeros(2304, 768, dtype=torch.bfloat16, device=dev)
        C.gemm(r(4096, 2304), True, x, True, None, 0, g2, False, 0, 2, g2)
    elif target == 'gemm2':                                # CTA-pair kernel on the LM-head shape and a large square
        C.gemm_2cta(r(4096, 768), False, r(30528, 768, scale=0.03), False, None, 0, 256)
        C.gemm_2cta(r(8192, 8192, scale=0.1), False, r(8192, 8192, scale=0.1), False, None, 0, 256)
        C.gemm(r(8192, 8192, scale=0.1), False, r(8192, 8192, scale=0.1), False, None, 0, None, False, 256)
    elif target == 'optim':
        from ravnest_b200.parallel.fused_optim import FusedFlatOptimizer
        from ravnest_b200.runtime.param_store import ParamStore
        lin = torch.nn.Linear(4096, 4096 * 6, bias=False).to(dev)           # ~100M parameters: BERT-base sized bucket
        store = ParamStore(lin, dev, n_slots=2, compute_dtype=torch.bfloat16)
        opt = FusedFlatOptimizer(store, 'adamw', dict(lr=1e-4, weight_decay=0.01))
        store.grad.normal_()
        opt.step()
    torch.cuda.synchronize()
print('done', target, dict(native.calls))
