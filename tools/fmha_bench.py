"""FMHA micro-benchmark at the BERT shape (B8 S512 h12 d64, key-padding mask + dropout 0.1): rvn fwd / bwd vs torch SDPA (cuDNN / flash),
CUDA-event timed with an L2 flush between iterations.  Writes gpurun_out/fmha_bench.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ravnest_b200 import ops  # noqa: E402

dev = 'cuda'
ops.set_backend('native')
B, S, heads, D = 8, 512, 12, 64
H = heads * D
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, reps=15):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return sorted(ts)[len(ts) // 2]


res = {}
for p_drop in (0.0, 0.1):
    qkv = (torch.randn(B, S, 3 * H, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True)
    mask = torch.ones(B, S, dtype=torch.long, device=dev)
    mask[:, S - 37:] = 0
    out = ops.attention_qkv(qkv, heads, mask, False, p_drop)
    dout = torch.randn_like(out)
    res['rvn_fwd_us_p{}'.format(p_drop)] = round(timed(lambda: ops.attention_qkv(qkv, heads, mask, False, p_drop)), 1)

    def fwd_bwd():
        o = ops.attention_qkv(qkv, heads, mask, False, p_drop)
        o.backward(dout)
        qkv.grad = None
    res['rvn_fwd_bwd_us_p{}'.format(p_drop)] = round(timed(fwd_bwd), 1)
    q, k, v = [t.reshape(B, S, heads, D).transpose(1, 2).contiguous().detach().requires_grad_(True) for t in qkv.detach().chunk(3, dim=-1)]
    do2 = torch.randn(B, heads, S, D, device=dev, dtype=torch.bfloat16)

    def sdpa():
        return torch.nn.functional.scaled_dot_product_attention(q, k, v, dropout_p=p_drop)

    def sdpa_fb():
        o = sdpa()
        o.backward(do2)
        q.grad = k.grad = v.grad = None
    res['sdpa_fwd_us_p{}'.format(p_drop)] = round(timed(sdpa), 1)
    res['sdpa_fwd_bwd_us_p{}'.format(p_drop)] = round(timed(sdpa_fb), 1)
print(json.dumps(res))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/fmha_bench.json', 'w'))
