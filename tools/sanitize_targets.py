"""Tiny instances of every tcgen05 / TMA kernel for compute-sanitizer (memcheck / racecheck / synccheck): small enough that the tool's
~100x slowdown stays within a minute, large enough to exercise every role (producer, MMA issuer, epilogue, split-K, tails)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ravnest_b200 import ops  # noqa: E402
from ravnest_b200.ops import _ext  # noqa: E402

C = _ext.load()
dev = 'cuda'
ops.set_backend('native')
torch.manual_seed(0)


def r(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).bfloat16()


def check(name, out, ref, tol=3e-2):
    err = float((out.float() - ref.float()).norm() / (ref.float().norm() + 1e-9))
    print('{:28s} rel_err {:.2e} {}'.format(name, err, 'ok' if err < tol else 'MISMATCH'), flush=True)
    assert err < tol, name


# GEMM: every layout, three tile widths, N / K tails, bias + GELU epilogue, split-K accumulate through TMA reduce-add
for (M, N, K) in [(256, 192, 128), (136, 200, 72)]:
    a, b, bias = r(M, K, scale=0.5), r(N, K, scale=0.1), r(N)
    ref = a.float() @ b.float().t()
    for bn in (128, 192, 256):
        check('gemm_nt bn{} {}x{}x{}'.format(bn, M, N, K), C.gemm(a, False, b, False, None, 0, None, False, bn)[0], ref)
    check('gemm_nt bias+gelu', C.gemm(a, False, b, False, bias, 1, None, False, 128)[0], torch.nn.functional.gelu(ref + bias.float()))
    bt = r(K, N, scale=0.1)
    check('gemm_nn (dgrad layout)', C.gemm(a, False, bt, True, None, 0, None, False, 0)[0], a.float() @ bt.float())
    at = r(K, M, scale=0.5)
    check('gemm_tt (wgrad layout)', C.gemm(at, True, bt, True, None, 0, None, False, 0)[0], at.float().t() @ bt.float())
    g = r(M, N)
    g0 = g.clone()
    C.gemm(at, True, bt, True, None, 0, g, False, 128, 2, g, None, None, 0, 2)
    check('gemm_tt split-K reduce-add', g, g0.float() + at.float().t() @ bt.float())
# 2-CTA GEMM (cta_group::2)
a, b = r(512, 128, scale=0.5), r(256, 128, scale=0.1)
check('gemm_2cta bn256', C.gemm_2cta(a, False, b, False, None, 0, 256), a.float() @ b.float().t())
check('gemm_2cta bn128', C.gemm_2cta(a, False, b, False, None, 0, 128), a.float() @ b.float().t())
# convolution forward / dgrad / wgrad (bulk reduce-add epilogue), strided pointwise
for (n, c, h, k, ks, s, p) in [(2, 64, 8, 64, 3, 1, 1), (2, 64, 9, 128, 1, 2, 0), (2, 128, 8, 64, 3, 2, 1)]:
    x = r(n, c, h, h).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = r(k, c, ks, ks, scale=0.05).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = ops.conv2d(x, w, None, s, p)
    gy = r(*y.shape).contiguous(memory_format=torch.channels_last)
    y.backward(gy)
    xr, wr = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xr, wr, None, s, p)
    yr.backward(gy.float())
    check('conv {}x{} s{} fwd'.format(ks, ks, s), y, yr)
    check('conv {}x{} s{} dgrad'.format(ks, ks, s), x.grad, xr.grad)
    check('conv {}x{} s{} wgrad'.format(ks, ks, s), w.grad, wr.grad)
# attention forward / backward (dQ by bulk reduce-add), padding mask, small-head padding path
for (B, S, heads, D, causal) in [(1, 160, 2, 64, False), (2, 11, 3, 16, True)]:
    H = heads * D
    qkv = r(B, S, 3 * H, scale=0.8).requires_grad_(True)
    mask = torch.ones(B, S, dtype=torch.long, device=dev)
    mask[:, S - 3:] = 0
    m = None if causal else mask
    out = ops.attention_qkv(qkv, heads, m, causal, 0.0)
    go = r(B, S, H)
    out.backward(go)
    ri = qkv.detach().float().requires_grad_(True)
    ref = ops.reference_attention_qkv(ri, heads, m, causal, 0.0)
    ref.backward(go.float())
    check('fmha D{} fwd'.format(D), out, ref)
    check('fmha D{} bwd'.format(D), qkv.grad, ri.grad, 5e-2)
torch.cuda.synchronize()
print('SANITIZE_TARGETS_OK')
