"""Hot-op dispatch: hand-written sm_100a kernels on CUDA, plain PyTorch reference elsewhere.

``backend()`` is ``'native'`` by default: CUDA tensors go to the kernels in ``ops/csrc`` (loaded
from the in-tree extension ``ravnest_b200/ops/rvn_C.so``) and a missing extension is a hard error —
a GPU box must never fall back silently.  ``set_backend('torch')`` selects the stock
cuBLAS/cuDNN/ATen path: that is the BASELINE TIER the product is measured against.

Every op has ``reference_<op>`` (fp32-accurate PyTorch math) used by the CPU path and by the GPU
numerics tests.
"""
import math
import os
from typing import Optional

import torch
import torch.nn.functional as F

_BACKEND = os.environ.get('RAVNEST_OPS_BACKEND', 'native')


def set_backend(name: str) -> None:
    global _BACKEND
    assert name in ('native', 'torch')
    _BACKEND = name


def backend() -> str:
    return _BACKEND


def _native(*tensors) -> bool:
    return _BACKEND == 'native' and all(t is None or t.is_cuda for t in tensors)


# ----------------------------------------------------------------------------------------------
# reference implementations
# ----------------------------------------------------------------------------------------------
def _act(y, activation):
    if activation is None:
        return y
    if activation == 'gelu':
        return F.gelu(y)
    if activation == 'gelu_tanh':
        return F.gelu(y, approximate='tanh')
    if activation == 'relu':
        return F.relu(y)
    raise ValueError(activation)


def reference_linear(x, weight, bias=None, activation=None):
    return _act(F.linear(x, weight, bias), activation)


def reference_layer_norm(x, weight, bias, eps=1e-5, residual=None, dropout_p=0.0):
    if dropout_p > 0.0:
        x = F.dropout(x, dropout_p, True)
    if residual is not None:
        x = x + residual
    return F.layer_norm(x, (x.shape[-1],), weight, bias, eps)


def reference_embedding(idx, weight):
    return F.embedding(idx, weight)


def reference_attention_qkv(qkv, heads, key_padding_mask=None, causal=False, dropout_p=0.0):
    """qkv: [B, S, 3H] packed (q | k | v); returns [B, S, H]."""
    B, S, H3 = qkv.shape
    H = H3 // 3
    D = H // heads
    q, k, v = qkv.split(H, dim=-1)
    q = q.view(B, S, heads, D).transpose(1, 2)
    k = k.view(B, S, heads, D).transpose(1, 2)
    v = v.view(B, S, heads, D).transpose(1, 2)
    att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(D))
    if causal:
        mask = torch.ones(S, S, dtype=torch.bool, device=qkv.device).tril()
        att = att.masked_fill(~mask, float('-inf'))
    if key_padding_mask is not None:
        if key_padding_mask.is_floating_point():
            att = att + key_padding_mask[:, None, None, :].to(att.dtype)
        else:
            att = att.masked_fill(~key_padding_mask.bool()[:, None, None, :], float('-inf'))
    att = torch.softmax(att.float(), dim=-1).to(qkv.dtype)
    if dropout_p > 0.0:
        att = F.dropout(att, dropout_p, True)
    y = att @ v
    return y.transpose(1, 2).reshape(B, S, H)


def reference_cross_entropy(logits, target, ignore_index=-100):
    return F.cross_entropy(logits.float(), target, ignore_index=ignore_index)


# ----------------------------------------------------------------------------------------------
# public ops
# ----------------------------------------------------------------------------------------------
def linear(x, weight, bias=None, activation=None):
    if _native(x, weight):
        from . import native
        return native.linear(x, weight, bias, activation)
    return reference_linear(x, weight, bias, activation)


def layer_norm(x, weight, bias, eps=1e-5, residual=None, dropout_p=0.0):
    """``LN(dropout(x) + residual)``"""
    if _native(x, weight):
        from . import native
        return native.layer_norm(x, weight, bias, eps, residual, dropout_p)
    return reference_layer_norm(x, weight, bias, eps, residual, dropout_p)


def mlp(x, w1, b1, w2, b2, activation='gelu'):
    """``act(x W1^T + b1) W2^T + b2``"""
    if _native(x, w1):
        from . import native
        return native.mlp(x, w1, b1, w2, b2, activation)
    return reference_linear(reference_linear(x, w1, b1, activation), w2, b2, None)


def embedding(idx, weight, padding_idx=None):
    if _native(idx, weight):
        from . import native
        return native.embedding(idx, weight, padding_idx)
    return F.embedding(idx, weight, padding_idx)


def dropout(x, p=0.5, training=True):
    """Dropout whose mask is a pure function of (seed, step, position) on the native path: nothing is saved for backward."""
    if not training or p <= 0.0:
        return x
    if _native(x):
        from . import native
        return native.dropout(x, p)
    return F.dropout(x, p, True)


def key_padding_bias(mask):
    """Attention mask [B, S] (any dtype, non-zero = attend — the HF convention) -> additive fp32 key bias (0 / -inf)."""
    if mask is None:
        return None
    return torch.zeros(mask.shape, dtype=torch.float32, device=mask.device).masked_fill_(mask == 0, float('-inf'))


def attention_qkv(qkv, heads, key_padding_mask=None, causal=False, dropout_p=0.0):
    """``key_padding_mask``: [B, S] integer / bool mask (non-zero = attend), or — any FLOATING dtype — the additive bias made by
    ``key_padding_bias`` (0 = attend, -inf = padding; a relayed bias arrives in the next stage's compute dtype)."""
    if _native(qkv):
        from . import native
        return native.attention_qkv(qkv, heads, key_padding_mask, causal, dropout_p)
    return reference_attention_qkv(qkv, heads, key_padding_mask, causal, dropout_p)


def cross_entropy(logits, target, ignore_index=-100):
    if _native(logits):
        from . import native
        return native.cross_entropy(logits, target, ignore_index)
    return reference_cross_entropy(logits, target, ignore_index)


def conv1x1_fork(x, weight, bn_stats=False):
    """``(conv1x1(x), x)``: the pointwise convolution that opens a residual block plus the block's identity branch.  On the native
    path the convolution's data gradient is accumulated INTO the gradient arriving over the skip connection by the dgrad GEMM's
    reduce-add epilogue (no separate add pass; ``ops/native.py: _Conv1x1Fork``) — the skip must therefore end in an op whose
    backward returns a freshly written tensor (``batch_norm_act(..., residual, relu=True)`` does)."""
    cout, cin = weight.shape[0], weight.shape[1]
    if (_native(x, weight) and x.dim() == 4 and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and cin % 8 == 0
            and cout % 8 == 0 and weight.shape[2] == 1 and weight.shape[3] == 1):
        from . import native
        return native.conv1x1_fork(x, weight, bn_stats)
    return conv2d(x, weight), x


def conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, bn_stats=False):
    """2-D convolution.  On CUDA (bf16, channels_last) pointwise convolutions run on the tcgen05 GEMM — in NHWC a 1x1
    convolution IS ``[N*H*W, Cin] x [Cout, Cin]^T`` with zero data movement (strided 1x1: the implicit-GEMM kernel, whose TMA
    element strides do the sub-sampling) — and square or rectangular kh x kw convolutions (stride 1 / 2, per-axis padding) on the
    implicit-GEMM kernel (forward, stride-1 dgrad, split-K wgrad with bulk reductions).  Image stems (Cin < 8, e.g. RGB) run on the same kernel over an 8-channel
    zero-bordered pack of the batch, read through an overlapping-stride tensor map.  Remaining library cases (cuDNN): other
    Cin % 8 != 0, stride-2 data gradients, groups > 1, dilation — counted in ``ops.native.fallbacks``.  ``bn_stats=True`` (a training
    BatchNorm follows): the producing kernel's epilogue also collects the per-channel sum and sum of squares of the output and hands
    them to ``batch_norm_act`` with the tensor, which then skips its statistics pass."""
    if _native(x, weight):
        from . import native
        return native.conv2d(x, weight, bias, stride, padding, dilation, groups, bn_stats)
    return F.conv2d(x, weight, bias, stride, padding, dilation, groups)


def batch_norm_act(x, weight, bias, running_mean, running_var, residual=None, relu=False, training=True, momentum=0.1, eps=1e-5,
                   num_batches_tracked=None):
    """``relu(batch_norm(x) + residual)`` (the tail of every ResNet / Inception conv block) in one pass over the activation.
    ``num_batches_tracked`` (optional int64 buffer) is incremented in training mode, inside the statistics kernel on the native path."""
    if _native(x, weight) and x.dim() == 4 and x.dtype == torch.bfloat16 and x.shape[1] % 8 == 0:
        from . import native
        return native.batch_norm_act(x, weight, bias, running_mean, running_var, residual, relu, training, momentum, eps,
                                     num_batches_tracked)
    if training and num_batches_tracked is not None:
        num_batches_tracked.add_(1)
    y = F.batch_norm(x, running_mean, running_var, weight, bias, training, momentum, eps)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y


def _sq(v):
    return v if isinstance(v, int) else (v[0] if len(set(v)) == 1 else None)


def max_pool2d(x, kernel_size, stride=None, padding=0):
    stride = kernel_size if stride is None else stride
    k, s, p = _sq(kernel_size), _sq(stride), _sq(padding)
    if _native(x) and None not in (k, s, p) and x.dim() == 4 and x.dtype == torch.bfloat16 and x.shape[1] % 8 == 0 and k <= 15:
        from . import native
        return native.pool2d(x, k, s, p, True)
    return F.max_pool2d(x, kernel_size, stride, padding)


def avg_pool2d(x, kernel_size, stride=None, padding=0):
    stride = kernel_size if stride is None else stride
    k, s, p = _sq(kernel_size), _sq(stride), _sq(padding)
    if _native(x) and None not in (k, s, p) and x.dim() == 4 and x.dtype == torch.bfloat16 and x.shape[1] % 8 == 0 and k <= 15:
        from . import native
        return native.pool2d(x, k, s, p, False)
    return F.avg_pool2d(x, kernel_size, stride, padding)


def native_ops_loaded():
    """Names of native kernels that have actually been launched in this process."""
    try:
        from . import native
        return dict(native.calls)
    except Exception:
        return {}


__all__ = ['native_ops_loaded', 'set_backend', 'backend', 'linear', 'mlp', 'conv2d', 'batch_norm_act', 'max_pool2d', 'avg_pool2d', 'layer_norm', 'embedding', 'dropout', 'attention_qkv', 'key_padding_bias', 'cross_entropy',
           'reference_linear', 'reference_layer_norm', 'reference_embedding', 'reference_attention_qkv',
           'reference_cross_entropy']
