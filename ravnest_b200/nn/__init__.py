"""Leaf modules whose CUDA path is the hand-written sm_100a kernels in ``ravnest_b200/ops``.

They are drop-in for their ``torch.nn`` namesakes (same parameter names/shapes, so state dicts are
interchangeable with torchvision / HF / minGPT checkpoints) and are treated as FX leaves by the
planner (``operations/splitter.py``).  On CPU they run the plain PyTorch reference of the same math —
that reference is also what the GPU numerics tests compare the kernels against.

Call sites covered: K3 Linear (+bias/GELU/residual epilogues), K5 LayerNorm
(+residual), K6 attention (padding / causal mask), K10 Embedding, K9 cross-entropy.
"""
import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops


class Linear(nn.Linear):
    """``y = act(x W^T + b)``; ``activation`` in {None, 'gelu', 'gelu_tanh', 'relu'} is fused into the
    GEMM epilogue on CUDA (tcgen05 GEMM, ops/csrc/gemm_sm100.cu)."""

    def __init__(self, in_features, out_features, bias=True, activation: Optional[str] = None, device=None, dtype=None):
        super().__init__(in_features, out_features, bias=bias, device=device, dtype=dtype)
        self.activation = activation

    def forward(self, x):
        return ops.linear(x, self.weight, self.bias, self.activation)

    def extra_repr(self):
        return super().extra_repr() + ', activation={}'.format(self.activation)


class LayerNorm(nn.LayerNorm):
    """LayerNorm over the last dim with an optional fused residual: ``LN(x + residual)``."""

    def forward(self, x, residual=None):
        return ops.layer_norm(x, self.weight, self.bias, self.eps, residual)


class Embedding(nn.Embedding):
    def forward(self, idx):
        return ops.embedding(idx, self.weight)


class Dropout(nn.Dropout):
    pass


class SelfAttention(nn.Module):
    """Multi-head self-attention core: fused QKV projection -> FMHA -> output projection is left to
    the caller.  Parameter layout matches HF BERT's ``attention.self`` when ``fused_qkv=False``
    (``query``/``key``/``value`` Linear) and minGPT's ``c_attn`` when ``fused_qkv=True``."""

    def __init__(self, hidden, heads, dropout=0.0, causal=False, fused_qkv=False):
        super().__init__()
        assert hidden % heads == 0
        self.hidden, self.heads, self.head_dim = hidden, heads, hidden // heads
        self.causal = causal
        self.dropout_p = dropout
        self.fused_qkv = fused_qkv
        if fused_qkv:
            self.c_attn = Linear(hidden, 3 * hidden)
        else:
            self.query = Linear(hidden, hidden)
            self.key = Linear(hidden, hidden)
            self.value = Linear(hidden, hidden)

    def forward(self, x, key_padding_mask=None):
        """x: [B, S, H]; key_padding_mask: [B, S] with 1 = attend, 0 = padding (HF convention)."""
        B, S, _ = x.shape
        if self.fused_qkv:
            qkv = self.c_attn(x)
        else:
            w = torch.cat([self.query.weight, self.key.weight, self.value.weight], dim=0)
            b = torch.cat([self.query.bias, self.key.bias, self.value.bias], dim=0)
            qkv = ops.linear(x, w, b, None)
        out = ops.attention_qkv(qkv, self.heads, key_padding_mask, self.causal,
                                self.dropout_p if self.training else 0.0)
        return out


class GELU(nn.Module):
    def __init__(self, approximate='none'):
        super().__init__()
        self.approximate = approximate

    def forward(self, x):
        return F.gelu(x, approximate=self.approximate)


__all__ = ['Linear', 'LayerNorm', 'Embedding', 'Dropout', 'SelfAttention', 'GELU']
