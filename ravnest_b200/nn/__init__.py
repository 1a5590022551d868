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
        if self.max_norm is not None or self.sparse or self.scale_grad_by_freq:
            return super().forward(idx)
        return ops.embedding(idx, self.weight, self.padding_idx)


class Dropout(nn.Dropout):
    """``nn.Dropout`` on the Philox kernel (CUDA): no mask tensor; recompute and backward replay the forward's mask."""

    def forward(self, x):
        if self.inplace:
            return super().forward(x)
        return ops.dropout(x, self.p, self.training)


class SelfAttention(nn.Module):
    """Multi-head self-attention core: one fused QKV projection (a single [3H, H] parameter, so no per-step
    ``cat`` of q/k/v weights) -> FMHA kernel.  The output projection is left to the caller.

    ``qkv_name`` is the attribute the fused Linear is registered under ('c_attn' for minGPT checkpoints,
    'qkv' for BERT).  HF BERT checkpoints with separate ``query``/``key``/``value`` tensors are merged on load."""

    def __init__(self, hidden, heads, dropout=0.0, causal=False, fused_qkv=True, qkv_name='qkv'):
        super().__init__()
        assert hidden % heads == 0
        self.hidden, self.heads, self.head_dim = hidden, heads, hidden // heads
        self.causal = causal
        self.dropout_p = dropout
        self.qkv_name = qkv_name
        self.add_module(qkv_name, Linear(hidden, 3 * hidden))
        self._register_load_state_dict_pre_hook(self._merge_hf_qkv)

    def _merge_hf_qkv(self, state_dict, prefix, *_):
        for kind in ('weight', 'bias'):
            keys = [prefix + '{}.{}'.format(n, kind) for n in ('query', 'key', 'value')]
            if all(k in state_dict for k in keys):
                state_dict[prefix + '{}.{}'.format(self.qkv_name, kind)] = torch.cat([state_dict.pop(k) for k in keys], dim=0)

    @property
    def qkv_proj(self):
        return getattr(self, self.qkv_name)

    def split_qkv_state(self):
        """HF-style {query,key,value}.{weight,bias} views of the fused parameter (export helper)."""
        w, b = self.qkv_proj.weight, self.qkv_proj.bias
        H = self.hidden
        return {'{}.{}'.format(n, k): t[i * H:(i + 1) * H] for i, n in enumerate(('query', 'key', 'value')) for k, t in (('weight', w), ('bias', b))}

    def forward(self, x, key_padding_mask=None):
        """x: [B, S, H]; key_padding_mask: [B, S] with 1 = attend, 0 = padding (HF convention)."""
        qkv = self.qkv_proj(x)
        return ops.attention_qkv(qkv, self.heads, key_padding_mask, self.causal, self.dropout_p if self.training else 0.0)


class Conv2d(nn.Conv2d):
    """Drop-in ``nn.Conv2d`` (same parameters / state dict).  Pointwise convolutions of channels_last bf16 activations run
    on the tcgen05 GEMM, spatial ones on the implicit-GEMM kernel, the rest on cuDNN (``ops.conv2d``).

    ``fork=True`` (bias-free stride-1 pointwise convolutions only) makes ``forward`` return ``(conv(x), x)``: the convolution
    that opens a residual block hands the block its identity branch, and in backward adds its data gradient into the skip
    connection's gradient inside the GEMM epilogue (``ops.conv1x1_fork``).  ``bn_stats=True`` declares that a BatchNorm consumes the
    output: in training the producing kernel's epilogue then also collects the per-channel batch statistics (``ops.conv2d``)."""

    def __init__(self, *args, fork=False, bn_stats=False, **kwargs):
        super().__init__(*args, **kwargs)
        self.fork = bool(fork)
        self.bn_stats = bool(bn_stats)      # a BatchNorm consumes the output: let the kernel's epilogue collect its batch statistics
        if self.fork and (self.kernel_size != (1, 1) or self.stride != (1, 1) or self.padding != (0, 0) or self.bias is not None
                          or self.groups != 1):
            raise ValueError('fork=True needs a bias-free, stride-1, ungrouped 1x1 convolution')

    def forward(self, x):
        # (getattr: stage modules pickled before these hints existed still load and run)
        stats = getattr(self, 'bn_stats', False) and self.training
        if getattr(self, 'fork', False):
            return ops.conv1x1_fork(x, self.weight, stats)
        if isinstance(self.padding, str) or self.padding_mode != 'zeros':
            return super().forward(x)
        return ops.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups, stats)


class BatchNorm2d(nn.BatchNorm2d):
    """``nn.BatchNorm2d`` (same parameters / buffers / state dict) with the ReLU and the residual add of the surrounding block
    fused in: ``forward(x, residual=None)`` computes ``act(BN(x) + residual)``; ``relu=True`` selects the activation."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, relu=False, **kw):
        super().__init__(num_features, eps=eps, momentum=momentum, **kw)
        self.relu = relu

    def forward(self, x, residual=None):
        training = self.training or not self.track_running_stats
        nbt = self.num_batches_tracked if (self.training and self.track_running_stats) else None
        return ops.batch_norm_act(x, self.weight, self.bias, self.running_mean, self.running_var, residual, self.relu, training,
                                  self.momentum if self.momentum is not None else 0.1, self.eps, num_batches_tracked=nbt)


class KeyPaddingBias(nn.Module):
    """[B, S] attention mask (1 = attend) -> the additive fp32 key bias the attention kernel consumes, computed ONCE per stage
    forward instead of once per layer (12 x 4 tiny launches per BERT step).  Parameter-free, so state dicts are unaffected;
    a planner leaf, so a stage cut after it relays the bias instead of the integer mask."""
    _ravnest_leaf = True

    def forward(self, mask):
        return ops.key_padding_bias(mask)


class MaxPool2d(nn.MaxPool2d):
    def forward(self, x):
        if self.dilation not in (1, (1, 1)) or self.ceil_mode or self.return_indices:
            return super().forward(x)
        return ops.max_pool2d(x, self.kernel_size, self.stride, self.padding)


class AvgPool2d(nn.AvgPool2d):
    def forward(self, x):
        if self.ceil_mode or not self.count_include_pad or self.divisor_override is not None:
            return super().forward(x)
        return ops.avg_pool2d(x, self.kernel_size, self.stride, self.padding)


class GELU(nn.Module):
    def __init__(self, approximate='none'):
        super().__init__()
        self.approximate = approximate

    def forward(self, x):
        return F.gelu(x, approximate=self.approximate)


__all__ = ['Linear', 'LayerNorm', 'Embedding', 'Dropout', 'SelfAttention', 'KeyPaddingBias', 'Conv2d', 'BatchNorm2d', 'MaxPool2d', 'AvgPool2d', 'GELU']
