"""CUDA path of the hot ops: autograd Functions over the sm_100a kernels in ``rvn_C``.

Layout conventions of ``rvn_C.gemm(A, a_mn, B, b_mn, bias, act, dact_in, store_preact, block_n)``:
``D[M,N] = epi(A' B'^T)`` with A' = A ([M,K]) or A^T when ``a_mn`` (A stored [K,M]); B' = B ([N,K]) or
B^T when ``b_mn`` (B stored [K,N]).  So for ``y = x W^T``:
    forward  y  = gemm(x,  False, W,  False)           (both K-major)
    dgrad    dx = gemm(dy, False, W,  True)            (W read MN-major: no transpose materialised)
    wgrad    dW = gemm(dy, True,  x,  True)            (both MN-major)
"""
import math
from collections import Counter

import os

import torch
import torch.nn.functional as F

from . import _ext

_ACT = {None: 0, 'gelu': 1, 'gelu_tanh': 2, 'relu': 3}
calls = Counter()          # op name -> number of native launches (smoke() / bench report these)
fallbacks = Counter()      # op name -> number of times constraints forced the torch path


def C():
    return _ext.load()


def _tma_ok(t: torch.Tensor) -> bool:
    return (t.dim() == 2 and t.dtype == torch.bfloat16 and t.stride(1) == 1 and t.stride(0) % 8 == 0
            and t.data_ptr() % 16 == 0 and t.size(0) > 0 and t.size(1) > 0)


def _as2d(x: torch.Tensor) -> torch.Tensor:
    x2 = x.reshape(-1, x.shape[-1])
    if not _tma_ok(x2):
        x2 = x2.contiguous()
    return x2


_PEER_WEIGHTS = {}          # data_ptr of a local weight leaf -> (leaf, peer view holding the fresh values)


def arm_peer_weight(leaf, peer_view):
    _PEER_WEIGHTS[leaf.data_ptr()] = (leaf, peer_view)


def drain_peer_weights():
    items = list(_PEER_WEIGHTS.values())
    _PEER_WEIGHTS.clear()
    return items


_BCAST = {}


def bcast_flags(device):
    """Per-device landing flags of the broadcast-fused GEMM (one per weight tile) and a monotonically increasing
    epoch, so the flags never need clearing between calls."""
    key = (device.type, device.index)
    if key not in _BCAST:
        _BCAST[key] = [torch.zeros(1 << 16, dtype=torch.int32, device=device), 0]
    _BCAST[key][1] += 1
    return _BCAST[key][0], _BCAST[key][1]


def _forward_gemm(x2, weight, bias, act, need_pre):
    """x2 W^T (+bias, act).  If ``weight`` is armed for a leader pull (K2), B is read from the leader's HBM over NVLink
    and landed in ``weight`` by the same kernel."""
    armed = _PEER_WEIGHTS.pop(weight.data_ptr(), None) if _PEER_WEIGHTS else None
    if armed is not None:
        calls['gemm_fwd_bcast'] += 1
        flags, epoch = bcast_flags(weight.device)
        return C().gemm(x2, False, armed[1], False, bias, act, None, need_pre, 0, 1, None, weight.detach(), flags, epoch)
    return C().gemm(x2, False, weight, False, bias, act, None, need_pre, 0)


def _main_grad(weight):
    """The flat-bucket gradient view of a ParamStore leaf (or None): wgrad accumulates into it in the GEMM epilogue
    instead of materialising dW and paying a separate AccumulateGrad pass (Megatron-style gradient accumulation fusion)."""
    g = weight.grad
    if g is not None and g.dtype == torch.bfloat16 and g.dim() == 2 and _tma_ok(g) and ACCUMULATE_INTO_GRAD:
        return g
    return None


ACCUMULATE_INTO_GRAD = True


def _wgrad(dy2, x2, weight):
    g = _main_grad(weight)
    if g is not None:
        C().gemm(dy2, True, x2, True, None, 0, g, False, 0, 2, g)       # g += dy^T x
        calls['gemm_wgrad_acc'] += 1
        return None
    calls['gemm_wgrad'] += 1
    return C().gemm(dy2, True, x2, True, None, 0, None, False, 0)[0]


def _vec_grad(param):
    """The flat-bucket gradient view of a 1-D bf16 ParamStore leaf (bias / norm scale), or None."""
    g = param.grad
    if ACCUMULATE_INTO_GRAD and g is not None and g.dtype == torch.bfloat16 and g.dim() == 1 and g.is_contiguous():
        return g
    return None


def _bias_grad(dy2, bias):
    """db = column sums of dy, added straight into the gradient bucket by the reducing kernel when the leaf carries one."""
    calls['colsum'] += 1
    g = _vec_grad(bias) if bias.dtype == torch.bfloat16 else None
    if g is not None:
        C().column_sum(dy2, g)
        return None
    return C().column_sum(dy2).to(bias.dtype)


def _pad_rows(dy2, N):
    if _tma_ok(dy2):
        return dy2
    dy2 = dy2.contiguous()
    if _tma_ok(dy2):
        return dy2
    pad = torch.zeros(dy2.shape[0], (N + 7) // 8 * 8, dtype=dy2.dtype, device=dy2.device)
    pad[:, :N] = dy2
    return pad[:, :N]


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, activation):
        x2 = _as2d(x)
        act = _ACT[activation]
        need_pre = act != 0 and (x.requires_grad or weight.requires_grad)
        outs = _forward_gemm(x2, weight, bias, act, need_pre)
        calls['gemm_fwd'] += 1
        ctx.act = act
        ctx.x_shape = x.shape
        ctx.save_for_backward(x2, weight, bias, outs[1] if need_pre else None)
        return outs[0].reshape(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, weight, bias, z = ctx.saved_tensors
        N = weight.shape[0]
        dy2 = dy.reshape(-1, N)
        if ctx.act != 0:
            dyc = dy2 if dy2.is_contiguous() else dy2.contiguous()
            zc = z if z.is_contiguous() else z.contiguous()
            dy2 = C().act_bwd(dyc, zc, ctx.act)
            calls['act_bwd'] += 1
        dy2 = _pad_rows(dy2, N)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = C().gemm(dy2, False, weight, True, None, 0, None, False, 0)[0].reshape(ctx.x_shape)
            calls['gemm_dgrad'] += 1
        if ctx.needs_input_grad[1]:
            dw = _wgrad(dy2, x2, weight)
        if bias is not None and ctx.needs_input_grad[2]:
            db = _bias_grad(dy2, bias)
        return dx, dw, db, None


class _MLP(torch.autograd.Function):
    """y = act(x W1^T + b1) W2^T + b2 with the activation gradient fused into the dgrad GEMM epilogue
    (dz = (dy W2) * act'(z) in one kernel: no separate activation-backward pass, dh never hits HBM twice)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, activation):
        x2 = _as2d(x)
        act = _ACT[activation]
        h, z = _forward_gemm(x2, w1, b1, act, 2)        # z := act'(pre-activation): the backward epilogue is one multiply
        y = _forward_gemm(h, w2, b2, 0, False)[0]
        calls['gemm_fwd'] += 2
        ctx.act = act
        ctx.x_shape = x.shape
        ctx.save_for_backward(x2, w1, b1, w2, b2, h, z)
        return y.reshape(*x.shape[:-1], w2.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w1, b1, w2, b2, h, z = ctx.saved_tensors
        dy2 = _pad_rows(dy.reshape(-1, w2.shape[0]), w2.shape[0])
        dz = C().gemm(dy2, False, w2, True, None, ctx.act, z, False, 0, 3)[0]      # (dy W2) * act'(z), z holds act'(z)
        calls['gemm_dgrad_dact'] += 1
        dw2 = _wgrad(dy2, h, w2) if ctx.needs_input_grad[3] else None
        db2 = _bias_grad(dy2, b2) if (b2 is not None and ctx.needs_input_grad[4]) else None
        dx = None
        if ctx.needs_input_grad[0]:
            dx = C().gemm(dz, False, w1, True, None, 0, None, False, 0)[0].reshape(ctx.x_shape)
            calls['gemm_dgrad'] += 1
        dw1 = _wgrad(dz, x2, w1) if ctx.needs_input_grad[1] else None
        db1 = _bias_grad(dz, b1) if (b1 is not None and ctx.needs_input_grad[2]) else None
        return dx, dw1, db1, dw2, db2, None


def mlp(x, w1, b1, w2, b2, activation='gelu'):
    ok = (x.dtype == torch.bfloat16 and x.shape[-1] % 8 == 0 and w1.shape[0] % 8 == 0 and w2.shape[0] % 8 == 0
          and _tma_ok(w1) and _tma_ok(w2) and x.numel() > 0)
    if not ok:
        fallbacks['mlp'] += 1
        return linear(linear(x, w1, b1, activation), w2, b2, None)
    return _MLP.apply(x, w1, b1, w2, b2, activation)


def linear(x, weight, bias=None, activation=None):
    K = x.shape[-1]
    if x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16 or K % 8 != 0 or not _tma_ok(weight) or x.numel() == 0:
        fallbacks['linear'] += 1
        from . import reference_linear
        return reference_linear(x, weight, bias, activation)
    if bias is not None and (bias.dtype != torch.bfloat16 or not bias.is_contiguous()):
        bias = bias.to(torch.bfloat16).contiguous()
    return _Linear.apply(x, weight, bias, activation)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, residual, dropout_p):
        xc = x.contiguous()
        rc = residual.contiguous() if residual is not None else None
        rng = next_rng_snapshot(xc.device) if dropout_p > 0.0 else None
        outs = C().layernorm_fwd(xc, rc, weight, bias, eps, rng, dropout_p)
        calls['layernorm_fwd'] += 1
        y, mean, rstd = outs[0], outs[1], outs[2]
        ctx.has_res = residual is not None
        ctx.dropout_p = dropout_p
        ctx.save_for_backward(outs[3] if len(outs) > 3 else xc, weight, bias, mean, rstd, rng)
        return y

    @staticmethod
    def backward(ctx, dy):
        s, weight, bias, mean, rstd, rng = ctx.saved_tensors
        gg, gb = _vec_grad(weight), _vec_grad(bias)
        acc = gg is not None and gb is not None and ctx.needs_input_grad[1] and ctx.needs_input_grad[2]
        dsum, dx, dg, db = C().layernorm_bwd(dy.contiguous(), s, weight, mean, rstd, rng, ctx.dropout_p,
                                             gg if acc else None, gb if acc else None)
        calls['layernorm_bwd'] += 1
        return dx, dg, db, None, (dsum if ctx.has_res else None), None


def layer_norm(x, weight, bias, eps=1e-5, residual=None, dropout_p=0.0):
    """``LN(dropout(x) + residual)`` — dropout and the residual add are fused into the normalisation kernel."""
    H = x.shape[-1]
    if x.dtype != torch.bfloat16 or H % 8 != 0 or H > 4096 or weight.dtype != torch.bfloat16:
        fallbacks['layer_norm'] += 1
        from . import reference_layer_norm
        return reference_layer_norm(x, weight, bias, eps, residual, dropout_p)
    if residual is not None and residual.shape != x.shape:
        residual = residual.expand_as(x)
    return _LayerNorm.apply(x, weight, bias, eps, residual, float(dropout_p))


class _CrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, ignore_index):
        loss, lse, stats = C().ce_fwd(logits, target, ignore_index)
        calls['ce_fwd'] += 1
        ctx.ignore_index = ignore_index
        ctx.save_for_backward(logits, target, lse, stats)
        return loss

    @staticmethod
    def backward(ctx, g):
        logits, target, lse, stats = ctx.saved_tensors
        d = C().ce_bwd(logits, target, lse, stats, g.reshape(1).float().contiguous(), ctx.ignore_index)
        calls['ce_bwd'] += 1
        return d, None, None


def cross_entropy(logits, target, ignore_index=-100):
    ok = (logits.dim() == 2 and logits.dtype == torch.bfloat16 and logits.stride(1) == 1 and logits.stride(0) % 8 == 0
          and logits.data_ptr() % 16 == 0 and logits.shape[1] >= 64)
    if not ok:
        fallbacks['cross_entropy'] += 1
        return F.cross_entropy(logits.float(), target, ignore_index=ignore_index)
    return _CrossEntropy.apply(logits, target.contiguous(), ignore_index)


_RNG = {}


def rng_state(device) -> torch.Tensor:
    """Device-resident {seed, step} used by every dropout-bearing kernel (Philox counter RNG).  ``step`` is bumped
    by one tiny device op per consumer call, so CUDA-graph replays keep drawing fresh masks and a backward /
    recompute replays its forward's mask from the snapshot it saved."""
    key = (device.type, device.index)
    if key not in _RNG:
        seed = int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF
        _RNG[key] = torch.tensor([seed, 0], dtype=torch.int64, device=device)
    return _RNG[key]


def next_rng_snapshot(device) -> torch.Tensor:
    """Snapshot {seed, step} for one dropout consumer and advance ``step`` — one single-thread launch (was clone + add_)."""
    state = rng_state(device)
    if device.type == 'cuda' and _ext.available():
        return C().rng_next(state)
    snap = state.clone()
    state[1:].add_(1)
    return snap


class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, heads, kbias, causal, dropout_p, scale=None):
        qc = qkv.contiguous()
        scale = float(scale) if scale is not None else 1.0 / math.sqrt(qc.shape[-1] // 3 // heads)
        rng = next_rng_snapshot(qc.device) if dropout_p > 0.0 else None
        out, lse = C().fmha_fwd(qc, heads, kbias, causal, scale, rng, dropout_p)
        calls['fmha_fwd'] += 1
        ctx.heads, ctx.causal, ctx.scale, ctx.dropout_p = heads, causal, scale, dropout_p
        ctx.save_for_backward(qc, out, lse, kbias, rng)
        return out

    @staticmethod
    def backward(ctx, d_out):
        qc, out, lse, kbias, rng = ctx.saved_tensors
        dqkv = C().fmha_bwd(qc, out, d_out.contiguous(), lse, ctx.heads, kbias, ctx.causal, ctx.scale, rng, ctx.dropout_p)
        calls['fmha_bwd'] += 1
        return dqkv, None, None, None, None, None


def attention_qkv(qkv, heads, key_padding_mask=None, causal=False, dropout_p=0.0):
    """Fused attention over packed qkv [B, S, 3H] -> [B, S, H] (tcgen05 FMHA, ops/csrc/fmha_sm100.cu)."""
    B, S, H3 = qkv.shape
    H = H3 // 3
    D = H // heads
    if D == 64 and qkv.dtype == torch.bfloat16:
        kbias = None
        if key_padding_mask is not None:
            if key_padding_mask.is_floating_point():
                kbias = key_padding_mask.float().contiguous()             # already the additive bias (ops.key_padding_bias), any precision
            else:
                kbias = torch.zeros(B, S, dtype=torch.float32, device=qkv.device).masked_fill_(~key_padding_mask.bool(), float('-inf'))
        return _Attention.apply(qkv, heads, kbias, causal, float(dropout_p))
    if D < 64 and D % 8 == 0 and qkv.dtype == torch.bfloat16:
        # small heads (minGPT gpt-nano: 16; 32, 48): zero-pad every head to the kernel's 64 lanes.  q.k is unchanged by zero lanes
        # (the softmax scale stays 1/sqrt(D)), the padded lanes of P.V are zero and are dropped again — the same tcgen05 kernel,
        # no library attention on any of the five benchmark configs.
        kbias = None
        if key_padding_mask is not None:
            if key_padding_mask.is_floating_point():
                kbias = key_padding_mask.float().contiguous()
            else:
                kbias = torch.zeros(B, S, dtype=torch.float32, device=qkv.device).masked_fill_(~key_padding_mask.bool(), float('-inf'))
        padded = F.pad(qkv.view(B, S, 3 * heads, D), (0, 64 - D)).view(B, S, 3 * heads * 64)
        out = _Attention.apply(padded, heads, kbias, causal, float(dropout_p), 1.0 / math.sqrt(D))
        calls['fmha_padded_head'] += 1
        return out.view(B, S, heads, 64)[..., :D].reshape(B, S, H)
    # other head dims: library SDPA
    q, k, v = qkv.view(B, S, 3, heads, D).permute(2, 0, 3, 1, 4)
    mask = None
    if key_padding_mask is not None:
        mask = (key_padding_mask == 0 if key_padding_mask.is_floating_point() else key_padding_mask.bool())[:, None, None, :]
        if causal:
            mask = mask & torch.ones(S, S, dtype=torch.bool, device=qkv.device).tril()
    out = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=dropout_p, is_causal=causal and mask is None)
    fallbacks['attention'] += 1
    return out.transpose(1, 2).reshape(B, S, H)


NATIVE_EMBEDDING = True
NATIVE_DROPOUT = True


class _Embedding(torch.autograd.Function):
    """Row gather forward; deterministic owner-computes scatter backward that adds straight into the table's gradient
    bucket (ops/csrc/embed_misc.cu) — no sort, no atomics, no AccumulateGrad pass over the 30522 x 768 table."""

    @staticmethod
    def forward(ctx, idx, weight, padding_idx):
        calls['embedding_fwd'] += 1
        ctx.save_for_backward(idx, weight)
        ctx.padding_idx = padding_idx
        return C().embedding_fwd(idx, weight)

    @staticmethod
    def backward(ctx, dy):
        idx, weight = ctx.saved_tensors
        calls['embedding_bwd'] += 1
        g = weight.grad
        V, H = weight.shape
        if V <= 64 and dy.dtype == torch.bfloat16 and H % 8 == 0:
            # tiny tables (token types, small vocabularies): every row has thousands of contributors, which would serialise on
            # one owner CTA — dW = onehot(idx)^T dy on the tensor cores instead (A read MN-major: [n, V8] one-hot, K = n)
            v8 = (V + 7) // 8 * 8
            flat = idx.reshape(-1)
            onehot = torch.zeros(flat.numel(), v8, dtype=torch.bfloat16, device=dy.device)
            onehot.scatter_(1, flat.unsqueeze(1), 1.0)
            if ctx.padding_idx >= 0:
                onehot[:, ctx.padding_idx] = 0
            dw = C().gemm(onehot, True, dy.reshape(-1, H).contiguous(), True, None, 0, None, False, 0)[0][:V]
            calls['embedding_bwd_gemm'] += 1
            if ACCUMULATE_INTO_GRAD and g is not None and g.dtype == dw.dtype and g.shape == dw.shape:
                g.add_(dw)
                return None, None, None
            return None, dw, None
        if ACCUMULATE_INTO_GRAD and g is not None and g.dtype == dy.dtype and g.is_contiguous() and g.shape == weight.shape:
            C().embedding_bwd(idx, dy, weight.shape[0], ctx.padding_idx, g)
            return None, None, None
        return None, C().embedding_bwd(idx, dy, weight.shape[0], ctx.padding_idx, None), None


def embedding(idx, weight, padding_idx=None):
    ok = (NATIVE_EMBEDDING and idx.dtype == torch.int64 and weight.dtype in (torch.bfloat16, torch.float32) and weight.is_contiguous()
          and 0 < idx.numel() <= 16384 and weight.dim() == 2)
    if not ok:
        fallbacks['embedding'] += 1
        return F.embedding(idx, weight, padding_idx)
    return _Embedding.apply(idx, weight, -1 if padding_idx is None else int(padding_idx))


class _Dropout(torch.autograd.Function):
    """Philox dropout without a mask tensor: backward (and a recompute) regenerate the forward's mask from the saved
    16-byte {seed, step} snapshot (ops/csrc/embed_misc.cu, philox.cuh)."""

    @staticmethod
    def forward(ctx, x, p):
        rng = next_rng_snapshot(x.device)
        calls['dropout_fwd'] += 1
        ctx.p = p
        ctx.save_for_backward(rng)
        return C().dropout_apply(x.contiguous(), rng, p)

    @staticmethod
    def backward(ctx, dy):
        (rng,) = ctx.saved_tensors
        calls['dropout_bwd'] += 1
        return C().dropout_apply(dy.contiguous(), rng, ctx.p), None


def dropout(x, p):
    vec = 8 if x.dtype == torch.bfloat16 else 4
    if not NATIVE_DROPOUT or x.dtype not in (torch.bfloat16, torch.float32) or x.numel() == 0 or x.numel() % vec != 0 or p >= 1.0:
        fallbacks['dropout'] += 1
        return F.dropout(x, p, True)
    if x.dim() == 4 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last):
        # the mask only depends on the storage offset: run on the NHWC view so no layout copy is made
        y = _Dropout.apply(x.permute(0, 2, 3, 1), float(p))
        return y.permute(0, 3, 1, 2)
    return _Dropout.apply(x, float(p))


def pack_for_wire(t, mode='fp16'):
    """fp32 activation -> fp16 (clamped) / bf16 in one pass (K12)."""
    calls['pack'] += 1
    return C().pack_f32(t.contiguous(), 0 if mode == 'fp16' else 1)


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _conv_grad_acc(weight, dw_f32_ohwi):
    """Add an fp32 OHWI weight gradient into the leaf's flat-bucket gradient (one fused cast+add launch) when the bucket view has
    the same physical order; otherwise hand a parameter-shaped gradient back to autograd."""
    g = weight.grad
    if ACCUMULATE_INTO_GRAD and g is not None:
        g_ohwi = g.permute(0, 2, 3, 1)
        if g_ohwi.is_contiguous():
            g_ohwi.add_(dw_f32_ohwi.view(g_ohwi.shape))
            calls['conv_wgrad_acc'] += 1
            return None
    return dw_f32_ohwi.view(weight.shape[0], weight.shape[2], weight.shape[3], weight.shape[1]).to(weight.dtype).permute(0, 3, 1, 2)


def _conv_wgrad_into(weight, acc, dy, x, R, S, stride, pad, dil, pad_w=-1):
    """Weight gradient of a convolution, delivered where the leaf wants it: ``acc`` is the leaf's slice of the stage's persistent
    fp32 accumulator (``ParamStore.enable_wgrad_staging``) — the split-K kernel's bulk reductions add straight into it and ONE
    flush per optimizer step moves everything into the gradient bucket; without it, a zeroed fp32 buffer + a cast-add per layer."""
    if acc is not None and WGRAD_STAGING and ACCUMULATE_INTO_GRAD:
        C().conv_wgrad_nhwc(dy, x, R, S, stride, pad, dil, pad_w, acc)
        calls['conv_wgrad_staged'] += 1
        return None
    return _conv_grad_acc(weight, C().conv_wgrad_nhwc(dy, x, R, S, stride, pad, dil, pad_w))


def _wgrad_acc_of(weight):
    return getattr(weight, '_rvn_wgrad_acc', None)


class _ConvNHWC(torch.autograd.Function):
    """Implicit-GEMM convolution (ops/csrc/conv_sm100.cu) on NHWC activations: forward, data gradient (stride 1 directly, stride 2
    as four parity phases) and the split-K weight gradient all run on the tcgen05 kernels.
    ``weight`` is the [O, I, kh, kw] parameter itself; stored channels_last (ParamStore does) its OHWI view is free."""

    @staticmethod
    def forward(ctx, x_nhwc, weight, bias, stride, pad, dil, relu, pad_w=-1, stats=None):
        w_ohwi = weight.permute(0, 2, 3, 1)
        if not w_ohwi.is_contiguous():
            w_ohwi = w_ohwi.contiguous()
            calls['conv_weight_permute'] += 1
        pad_w = pad if pad_w < 0 else pad_w
        y = C().conv_nhwc(x_nhwc, w_ohwi, bias, stride, pad, dil, relu, False, 0, 0, pad_w, stats)
        calls['conv_fwd'] += 1
        ctx.cfg = (stride, pad, dil, relu)
        ctx.pad_w = pad_w
        ctx.has_bias = bias is not None
        ctx.wgrad_acc = _wgrad_acc_of(weight)
        ctx.save_for_backward(x_nhwc, weight, w_ohwi, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, w, y = ctx.saved_tensors
        stride, pad, dil, relu = ctx.cfg
        dy = dy.contiguous()
        if relu:
            dy = dy * (y > 0)
        dx = dw = db = None
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        native_dx = need_dx and stride == 1
        pad_w = ctx.pad_w
        if native_dx:
            dx = C().conv_nhwc(dy, w, None, 1, pad, dil, False, True, x.shape[1], x.shape[2], pad_w)
            calls['conv_dgrad'] += 1
        elif need_dx and stride == 2 and dil == 1 and STRIDE2_DGRAD:
            dx = _dgrad_s2(dy, w, x.shape[1], x.shape[2], pad, pad_w)
            native_dx = dx is not None
        if need_dw and NATIVE_CONV_WGRAD:
            dw = _conv_wgrad_into(weight, ctx.wgrad_acc, dy, x, w.shape[1], w.shape[2], stride, pad, dil, pad_w)
            calls['conv_wgrad'] += 1
            need_dw = False
        if need_dw or (need_dx and not native_dx) or ctx.has_bias:
            # library path on NCHW-shaped channels_last views (zero-copy)
            gx, gw, gb = torch.ops.aten.convolution_backward(
                dy.permute(0, 3, 1, 2), x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), [w.shape[0]] if ctx.has_bias else None,
                [stride, stride], [pad, pad_w], [dil, dil], False, [0, 0], 1, [need_dx and not native_dx, need_dw, ctx.has_bias])
            fallbacks['conv_dgrad_s2' if not need_dw else 'conv_wgrad'] += 1
            if need_dx and not native_dx:
                dx = gx.permute(0, 2, 3, 1).contiguous()
            if need_dw:
                dw = gw
            db = gb
        return dx, dw, db, None, None, None, None, None, None


_S2_ODD_HEIGHT_OK = True


def _dgrad_s2(dy, w, H, W, pad, pad_w):
    """Data gradient of a stride-2 convolution as four parity phases of the implicit-GEMM kernel (conv_sm100.cu: rvn_conv_dgrad_s2);
    every phase writes dx through a tensor map that steps over every second pixel.  For an odd input height the image stride is not
    a multiple of that map's row stride; should the driver's encoder refuse such a map, those shapes go back to the library path."""
    global _S2_ODD_HEIGHT_OK
    if H % 2 and not _S2_ODD_HEIGHT_OK:
        return None
    try:
        dx = C().conv_dgrad_s2(dy, w, H, W, pad, pad_w)
    except RuntimeError as exc:
        if H % 2 == 0 or 'conv_dgrad_s2 (code 1' not in str(exc):
            raise
        _S2_ODD_HEIGHT_OK = False
        return None
    calls['conv_dgrad_s2'] += 1
    return dx


class _Conv1x1(torch.autograd.Function):
    """Pointwise convolution on NHWC activations: forward and dgrad are plain GEMMs over the channel dimension; the weight
    gradient reduces over ALL pixels (K = N*H*W, up to 3e5 for ResNet's first stage) into a handful of output tiles, so it runs
    on the split-K wgrad kernel (pixel patches as the reduction dimension, fp32 atomics) instead of the TT GEMM."""

    @staticmethod
    def forward(ctx, x_nhwc, weight, bias, stats=None):
        n, h, w, c = x_nhwc.shape
        w2 = weight.reshape(weight.shape[0], c)
        y2 = C().gemm(x_nhwc.reshape(n * h * w, c), False, w2, False, bias, 0, None, False, 0, col_stats=stats)[0]
        calls['conv1x1_as_gemm'] += 1
        ctx.has_bias = bias is not None
        ctx.wgrad_acc = _wgrad_acc_of(weight)
        ctx.save_for_backward(x_nhwc, weight, w2)
        return y2.view(n, h, w, w2.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x, weight, w2 = ctx.saved_tensors
        n, h, w, c = x.shape
        dy = dy.contiguous()
        dy2 = dy.view(n * h * w, w2.shape[0])
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = C().gemm(dy2, False, w2, True, None, 0, None, False, 0)[0].view(n, h, w, c)
            calls['gemm_dgrad'] += 1
        if ctx.needs_input_grad[1]:
            dw = _conv_wgrad_into(weight, ctx.wgrad_acc, dy, x, 1, 1, 1, 0, 1)
            calls['conv_wgrad'] += 1
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _bias_grad(dy2, w2)
        return dx, dw, db, None


class _Conv1x1Fork(torch.autograd.Function):
    """``(conv1x1(x), x)``: the pointwise convolution that opens a residual block, together with the block's identity branch.
    Autograd would otherwise meet the two gradients of ``x`` — the convolution's data gradient and the one arriving over the skip
    connection — with a separate add kernel (19 per ResNet-50 step, 26 us each on the 100 x 56 x 56 x 256 maps); here the dgrad GEMM
    adds its tile into the skip gradient through its TMA reduce-add epilogue (``D += dy W``), no extra pass over the activation.
    The skip gradient is updated IN PLACE: use it only where that tensor has a single owner (the fused BN+add+ReLU backward writes a
    fresh one), as ``nn.Conv2d(fork=True)`` in ``models/resnet.py`` does."""

    @staticmethod
    def forward(ctx, x_nhwc, weight, stats=None):
        n, h, w, c = x_nhwc.shape
        w2 = weight.reshape(weight.shape[0], c)
        y2 = C().gemm(x_nhwc.reshape(n * h * w, c), False, w2, False, None, 0, None, False, 0, col_stats=stats)[0]
        calls['conv1x1_as_gemm'] += 1
        ctx.wgrad_acc = _wgrad_acc_of(weight)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x_nhwc, weight, w2)
        return y2.view(n, h, w, w2.shape[0]), x_nhwc            # an input returned as-is: autograd hands out an alias

    @staticmethod
    def backward(ctx, dy, dskip):
        x, weight, w2 = ctx.saved_tensors
        n, h, w, c = x.shape
        dx = dw = None
        if dy is not None:
            dy = dy.contiguous()
            dy2 = dy.view(n * h * w, w2.shape[0])
            if ctx.needs_input_grad[0]:
                if (dskip is not None and CONV_FORK_ACCUMULATE and dskip.is_contiguous() and dskip.dtype == torch.bfloat16
                        and dskip.data_ptr() % 16 == 0 and dskip.data_ptr() != dy.data_ptr()):
                    d2 = dskip.view(n * h * w, c)
                    C().gemm(dy2, False, w2, True, None, 0, d2, False, 0, 2, d2)        # dskip += dy W
                    calls['gemm_dgrad_fork_acc'] += 1
                    dx = dskip
                else:
                    dx = C().gemm(dy2, False, w2, True, None, 0, None, False, 0)[0].view(n, h, w, c)
                    calls['gemm_dgrad'] += 1
                    if dskip is not None:
                        dx = dx + dskip
            if ctx.needs_input_grad[1]:
                dw = _conv_wgrad_into(weight, ctx.wgrad_acc, dy, x, 1, 1, 1, 0, 1)
                calls['conv_wgrad'] += 1
        elif ctx.needs_input_grad[0]:
            dx = dskip
        return dx, dw, None


class _ConvSmallCin(torch.autograd.Function):
    """The image stem: a kh x kw convolution (kw <= 8) over 1 - 8 input channels, which TMA cannot tile directly (a pixel is 2 - 8
    bytes; the unit works in 16-byte granules).  ``stem_pack`` writes a zero-bordered 8-channel copy of the batch; the convolution
    then runs on the implicit-GEMM kernel as a kh x 1 convolution over a VIRTUAL 64-channel tensor whose "pixel" w is the 8 packed
    pixels w .. w+7 of a row — a strided view with overlapping pixels (W stride 8 elements, extent 64), so the im2col along the
    row is done by the TMA unit and the weight is the [Cout, kh, 8 x 8] zero-padded filter row.  Forward and the split-K weight
    gradient only (images carry no gradient)."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, ph, pw, stats=None):
        n, cin, h, w = x.shape
        cout, _, kh, kw = weight.shape
        ho = (h + 2 * ph - kh) // stride + 1
        wo = (w + 2 * pw - kw) // stride + 1
        hp = h + 2 * ph
        wp = max(w + 2 * pw, (wo - 1) * stride + 8)
        xp = C().stem_pack(x, ph, pw, hp, wp)
        calls['stem_pack'] += 1
        hv, wv = (ho - 1) * stride + kh, (wo - 1) * stride + 1
        xv = _stem_view(xp, hv, wv)
        wrow = torch.zeros(cout, kh, 8, 8, dtype=torch.bfloat16, device=x.device)
        wrow[:, :, :kw, :cin] = weight.permute(0, 2, 3, 1)
        y = C().conv_nhwc(xv, wrow.view(cout, kh, 1, 64), bias, stride, 0, 1, False, False, 0, 0, 0, stats)
        calls['conv_stem'] += 1
        ctx.cfg = (stride, kh, kw, cin)
        ctx.has_bias = bias is not None
        ctx.save_for_backward(xv, weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        xv, weight, bias = ctx.saved_tensors
        stride, kh, kw, cin = ctx.cfg
        dy = dy.contiguous()
        dw = db = None
        if ctx.needs_input_grad[1]:
            dwrow = C().conv_wgrad_nhwc(dy, xv, kh, 1, stride, 0, 1, 0)                 # [Cout, kh, 1, 64] fp32
            calls['conv_stem_wgrad'] += 1
            d_ohwi = dwrow.view(weight.shape[0], kh, 8, 8)[:, :, :kw, :cin]             # -> [Cout, kh, kw, Cin]
            g = weight.grad
            if ACCUMULATE_INTO_GRAD and g is not None:
                g.permute(0, 2, 3, 1).add_(d_ohwi)
            else:
                dw = d_ohwi.permute(0, 3, 1, 2).to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _bias_grad(dy.view(-1, dy.shape[-1]), bias)
        return None, dw, db, None, None, None, None


_STEM_OVERLAP_OK = True


def _stem_view(xp, hv, wv):
    """[N, hv, wv, 64] view of the packed [N, Hp, Wp, 8] batch: "pixel" w = packed pixels w .. w+7 (overlapping strides)."""
    n, hp, wp, _ = xp.shape
    xv = xp.as_strided((n, hv, wv, 64), (hp * wp * 8, wp * 8, 8, 1))
    if not _STEM_OVERLAP_OK:
        # a driver that refuses overlapping tensor-map strides: materialise the row-im2col once (8x the packed bytes, same kernels)
        calls['conv_stem_materialized'] += 1
        return xv.contiguous()
    return xv


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _with_bn_sums(y, sums):
    """Attach the per-channel (sum, sum of squares) the producing kernel's epilogue collected to the activation object:
    ``batch_norm_act`` on this very tensor then skips its statistics pass (BatchNorm = finalize + apply)."""
    if sums is not None:
        y._rvn_bn_sums = sums
        calls['conv_bn_stats'] += 1
    return y


def _new_bn_sums(weight, bn_stats, gemm_rows=0):
    """[2, Cout] fp32 accumulator for the producing epilogue, or None.  ``gemm_rows``: the convolution runs as a plain GEMM with
    that many rows; policy 'skip_pair' leaves BatchNorm its own statistics pass where that GEMM is routed to the CTA-pair kernel
    (short-K pointwise convolutions are epilogue bound there: the fused statistics cost more than the separate kernel)."""
    if not (bn_stats and BN_STATS_FROM_CONV):
        return None
    if gemm_rows and BN_STATS_GEMM_POLICY != 'all':
        if BN_STATS_GEMM_POLICY == 'none':
            return None
        if BN_STATS_GEMM_POLICY == 'skip_pair' and C().gemm_prefers_pair(gemm_rows, weight.shape[0], weight.shape[1]):
            return None
    return torch.zeros(2 * weight.shape[0], dtype=torch.float32, device=weight.device)


def conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, bn_stats=False):
    """NCHW-shaped, channels_last-strided tensors in and out (what torch modules exchange); inside, NHWC views feed either the
    tcgen05 GEMM (pointwise convs: in NHWC a 1x1 conv IS ``[N*H*W, Cin] x [Cout, Cin]^T``) or the implicit-GEMM conv kernel.
    ``bn_stats``: a training BatchNorm follows — the epilogue also sums y and y^2 per channel (see ``_with_bn_sums``)."""
    kh, kw = weight.shape[2], weight.shape[3]
    sh, sw = _pair(stride)
    dh, dw_ = _pair(dilation)
    if isinstance(padding, str):
        fallbacks['conv2d'] += 1
        return F.conv2d(x, weight, bias, stride, padding, dilation, groups)
    ph, pw = _pair(padding)
    cin, cout = weight.shape[1], weight.shape[0]
    base_ok = (groups == 1 and x.dim() == 4 and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16
               and cin % 8 == 0 and cout % 8 == 0)
    strided = sh != 1 or sw != 1
    if (SMALL_CIN_STEM and groups == 1 and x.dim() == 4 and cin < 8 and kw <= 8 and cout % 8 == 0 and sh == sw and sh in (1, 2)
            and dh == 1 and dw_ == 1 and weight.dtype == torch.bfloat16 and x.dtype in (torch.bfloat16, torch.float32)
            and not x.requires_grad and (bias is None or bias.dtype == torch.bfloat16)):
        # image stems (RGB / grey-scale): 8-channel zero-bordered pack + overlapping-stride TMA im2col, see _ConvSmallCin
        sums = _new_bn_sums(weight, bn_stats)
        y = _stem_apply(x, weight, bias, sh, ph, pw, sums)
        return _with_bn_sums(y.permute(0, 3, 1, 2), sums)
    if base_ok and kh == 1 and kw == 1 and ph == 0 and pw == 0 and not (strided and STRIDED_1X1_IMPLICIT and sh == sw and sh == 2):
        # (strided pointwise convs — ResNet's three downsample projections — take the implicit-GEMM kernel below: its TMA box load
        #  strides over the input pixels itself, where the sub-sampling slice here costs a strided copy forward and a fill + strided
        #  scatter backward, ~100 us each on the 100 x 256 x 56 x 56 map)
        if strided:
            x = x[:, :, ::sh, ::sw]
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        sums = _new_bn_sums(weight, bn_stats and bias is None, x.shape[0] * x.shape[2] * x.shape[3])
        y = _Conv1x1.apply(x.permute(0, 2, 3, 1), weight, bias, sums)     # zero-copy NHWC view in, NHWC out
        return _with_bn_sums(y.permute(0, 3, 1, 2), sums)                 # NCHW-shaped, channels_last-strided
    if base_ok and sh == sw and dh == dw_ and sh in (1, 2) and IMPLICIT_GEMM_CONV and (kh == kw and ph == pw or RECTANGULAR_CONV):
        # square and rectangular (Inception's 1x7 / 7x1 / 1x3 / 3x1, separate vertical / horizontal padding) kernels
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        sums = _new_bn_sums(weight, bn_stats)
        y = _ConvNHWC.apply(x.permute(0, 2, 3, 1), weight, bias, sh, ph, dh, False, pw, sums)
        return _with_bn_sums(y.permute(0, 3, 1, 2), sums)
    fallbacks['conv2d'] += 1
    return F.conv2d(x, weight, bias, stride, padding, dilation, groups)


def _stem_apply(x, weight, bias, stride, ph, pw, sums=None):
    global _STEM_OVERLAP_OK
    try:
        return _ConvSmallCin.apply(x, weight, bias, stride, ph, pw, sums)
    except RuntimeError as exc:
        if not _STEM_OVERLAP_OK or 'conv_forward' not in str(exc):
            raise
        _STEM_OVERLAP_OK = False          # tensor-map encoder rejected the overlapping view: use the materialised one from now on
        if sums is not None:
            sums.zero_()
        return _ConvSmallCin.apply(x, weight, bias, stride, ph, pw, sums)


def conv1x1_fork(x, weight, bn_stats=False):
    """``(conv1x1(x), x)`` for the block-opening pointwise convolution of a residual block (NCHW-shaped channels_last bf16)."""
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    sums = _new_bn_sums(weight, bn_stats, x.shape[0] * x.shape[2] * x.shape[3])
    y, skip = _Conv1x1Fork.apply(x.permute(0, 2, 3, 1), weight, sums)
    return _with_bn_sums(y.permute(0, 3, 1, 2), sums), skip.permute(0, 3, 1, 2)


def wgrad_flush(acc, grad):
    """grad += acc; acc = 0 — the stage's fp32 convolution-weight-gradient accumulator into its flat gradient bucket."""
    C().wgrad_flush(acc, grad)
    calls['wgrad_flush'] += 1


IMPLICIT_GEMM_CONV = True
NATIVE_CONV_WGRAD = True
SMALL_CIN_STEM = True
STRIDE2_DGRAD = os.environ.get('RAVNEST_STRIDE2_DGRAD', '1') != '0'
CONV_FORK_ACCUMULATE = True
WGRAD_STAGING = True
BN_STATS_FROM_CONV = os.environ.get('RAVNEST_BN_STATS_FROM_CONV', '1') != '0'      # BatchNorm statistics from the producer's epilogue
BN_STATS_GEMM_POLICY = os.environ.get('RAVNEST_BN_STATS_GEMM', 'all')              # pointwise (GEMM) convolutions: all | skip_pair | none
STRIDED_1X1_IMPLICIT = True
RECTANGULAR_CONV = True
BN_RELU_BITMASK = True


class _BatchNormAct(torch.autograd.Function):
    """y = relu(BN(x) + residual) on NHWC activations (ops/csrc/norm_nhwc.cu).  gamma / beta / running statistics are consumed in
    the module's dtype; the finalize kernel updates the running statistics and ``num_batches_tracked``; the parameter gradients
    are accumulated straight into the flat gradient bucket when the leaves carry one (no cast / add launches per layer)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, relu, eps, running_mean, running_var, batches_tracked, momentum, sums=None):
        outs = C().batchnorm_fwd(x, residual, weight, bias, None, None, eps, relu, True, running_mean, running_var,
                                 batches_tracked, momentum, BN_RELU_BITMASK and relu, sums)
        if sums is not None:
            calls['batchnorm_stats_from_producer'] += 1
        y, mean, invstd = outs[0], outs[1], outs[2]
        mask = outs[3] if len(outs) > 3 else None          # ReLU sign bits: backward reads 1 byte per 8 elements instead of y
        calls['batchnorm_fwd'] += 1
        ctx.relu, ctx.has_res = relu, residual is not None
        ctx.save_for_backward(x, y if (relu and mask is None) else None, weight, bias, mean, invstd, mask)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, weight, bias, mean, invstd, mask = ctx.saved_tensors
        gg, gb = weight.grad, bias.grad
        acc = (ACCUMULATE_INTO_GRAD and gg is not None and gb is not None and gg.dtype == weight.dtype and gb.dtype == weight.dtype
               and gg.is_contiguous() and gb.is_contiguous() and ctx.needs_input_grad[1] and ctx.needs_input_grad[2])
        dx, dres, dgamma, dbeta = C().batchnorm_bwd(dy.contiguous(), y, x, weight, mean, invstd, ctx.relu, ctx.has_res and ctx.relu,
                                                    gg if acc else None, gb if acc else None, mask)
        calls['batchnorm_bwd'] += 1
        if ctx.has_res and not ctx.relu:
            dres = dy
        return dx, dgamma, dbeta, (dres if ctx.has_res else None), None, None, None, None, None, None, None


def batch_norm_act(x, weight, bias, running_mean, running_var, residual=None, relu=False, training=True, momentum=0.1, eps=1e-5,
                   num_batches_tracked=None):
    """x / residual / result: NCHW-shaped channels_last bf16 tensors.  Training mode updates the running statistics."""
    n, c, h, w = x.shape
    sums = getattr(x, '_rvn_bn_sums', None) if (training and BN_STATS_FROM_CONV) else None     # collected by the producing epilogue
    if sums is not None and sums.numel() != 2 * c:
        sums = None
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    x2 = x.permute(0, 2, 3, 1)
    r2 = None
    if residual is not None:
        if not residual.is_contiguous(memory_format=torch.channels_last):
            residual = residual.contiguous(memory_format=torch.channels_last)
        r2 = residual.permute(0, 2, 3, 1)
    if training:
        y2 = _BatchNormAct.apply(x2, weight, bias, r2, relu, eps, running_mean, running_var,
                                 num_batches_tracked if running_mean is not None else None, momentum, sums)
        if running_mean is None and num_batches_tracked is not None:
            num_batches_tracked.add_(1)
    else:
        invstd = torch.rsqrt(running_var.float() + eps)
        y2 = C().batchnorm_fwd(x2, r2, weight, bias, running_mean.float(), invstd, eps, relu, False, None, None, None, 0.0)[0]
        calls['batchnorm_eval'] += 1
    return y2.permute(0, 3, 1, 2)


class _PoolNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, stride, pad, is_max):
        outs = C().pool_fwd(x, k, stride, pad, is_max)
        calls['pool_fwd'] += 1
        ctx.cfg = (x.shape[1], x.shape[2], k, stride, pad, is_max)
        ctx.save_for_backward(outs[1] if is_max else None)
        return outs[0]

    @staticmethod
    def backward(ctx, dy):
        (arg,) = ctx.saved_tensors
        H, W, k, stride, pad, is_max = ctx.cfg
        calls['pool_bwd'] += 1
        return C().pool_bwd(dy.contiguous(), arg, H, W, k, stride, pad, is_max), None, None, None, None


def pool2d(x, kernel_size, stride, padding, is_max):
    """Square max / average pooling of NCHW-shaped channels_last bf16 tensors on the NHWC kernels."""
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    y = _PoolNHWC.apply(x.permute(0, 2, 3, 1), int(kernel_size), int(stride), int(padding), bool(is_max))
    return y.permute(0, 3, 1, 2)
