"""minGPT-style decoder (GPT-Sorter, benchmark config #5), FX-traceable.

Architecture / name parity: reference ``examples/sorter/mingpt/model_without_padding_mask.py``
(``GPT``, :119-199): ``transformer.{wte,wpe,drop,h[i].{ln_1,attn.{c_attn,c_proj},ln_2,mlp.{c_fc,c_proj}},ln_f}``
+ ``lm_head`` (no bias); tanh-GELU (:31); causal attention with dropout (:83-107).  ``gpt-nano`` =
3 layers / 3 heads / 48 dims (:189) -> 85 728 parameters at vocab 3, block 11.

The reference needs ``ArangeForFx`` / ``MaskedFillForFx`` helper leaves (:33-49) to get through
tracing; here positions and the causal mask live inside leaf modules (``_Positions``,
``nn.SelfAttention(causal=True)``), and attention is the fused FMHA kernel on CUDA.
"""
from dataclasses import dataclass

import torch
from torch import nn

from .. import nn as rnn

PRESETS = {
    'gpt-nano': dict(n_layer=3, n_head=3, n_embd=48),
    'gpt-micro': dict(n_layer=4, n_head=4, n_embd=128),
    'gpt-mini': dict(n_layer=6, n_head=6, n_embd=192),
    'gpt2': dict(n_layer=12, n_head=12, n_embd=768),
}


@dataclass
class GPTConfig:
    model_type: str = 'gpt-nano'
    vocab_size: int = 3
    block_size: int = 11
    n_layer: int = None
    n_head: int = None
    n_embd: int = None
    embd_pdrop: float = 0.1
    resid_pdrop: float = 0.1
    attn_pdrop: float = 0.1

    def resolve(self):
        if self.n_layer is None:
            for k, v in PRESETS[self.model_type].items():
                setattr(self, k, v)
        return self


class _Positions(nn.Module):
    """tok_emb + pos_emb with the arange kept out of the traced graph."""
    _ravnest_leaf = True

    def __init__(self, cfg):
        super().__init__()
        self.wte = rnn.Embedding(cfg.vocab_size, cfg.n_embd)
        self.wpe = rnn.Embedding(cfg.block_size, cfg.n_embd)

    def forward(self, idx):
        pos = torch.arange(idx.shape[1], device=idx.device)
        return self.wte(idx) + self.wpe(pos).unsqueeze(0)


class CausalSelfAttention(nn.Module):
    _ravnest_leaf = True

    def __init__(self, cfg):
        super().__init__()
        self.core = rnn.SelfAttention(cfg.n_embd, cfg.n_head, cfg.attn_pdrop, causal=True, qkv_name='c_attn')
        self.c_attn = self.core.c_attn
        self.c_proj = rnn.Linear(cfg.n_embd, cfg.n_embd)
        self.resid_dropout = rnn.Dropout(cfg.resid_pdrop)

    def forward(self, x):
        return self.resid_dropout(self.c_proj(self.core(x)))

    def named_parameters(self, *a, **k):    # expose c_attn once (under the minGPT name)
        return super().named_parameters(*a, **k)


class MLP(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.c_fc = rnn.Linear(cfg.n_embd, 4 * cfg.n_embd, activation='gelu_tanh')
        self.c_proj = rnn.Linear(4 * cfg.n_embd, cfg.n_embd)
        self.dropout = rnn.Dropout(cfg.resid_pdrop)

    def forward(self, x):
        return self.dropout(self.c_proj(self.c_fc(x)))


class Block(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.ln_1 = rnn.LayerNorm(cfg.n_embd)
        self.attn = CausalSelfAttention(cfg)
        self.ln_2 = rnn.LayerNorm(cfg.n_embd)
        self.mlp = MLP(cfg)

    def forward(self, x):
        x = x + self.attn(self.ln_1(x))
        return x + self.mlp(self.ln_2(x))


class _Transformer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.emb = _Positions(cfg)
        self.wte, self.wpe = self.emb.wte, self.emb.wpe
        self.drop = rnn.Dropout(cfg.embd_pdrop)
        self.h = nn.ModuleList([Block(cfg) for _ in range(cfg.n_layer)])
        self.ln_f = rnn.LayerNorm(cfg.n_embd)

    def forward(self, idx):
        x = self.drop(self.emb(idx))
        for block in self.h:
            x = block(x)
        return self.ln_f(x)


class GPT(nn.Module):
    def __init__(self, config: GPTConfig = None):
        super().__init__()
        self.config = (config or GPTConfig()).resolve()
        self.block_size = self.config.block_size
        self.transformer = _Transformer(self.config)
        self.lm_head = rnn.Linear(self.config.n_embd, self.config.vocab_size, bias=False)
        self.apply(self._init)
        for name, p in self.named_parameters():
            if name.endswith('c_proj.weight'):
                nn.init.normal_(p, mean=0.0, std=0.02 / (2 * self.config.n_layer) ** 0.5)

    @staticmethod
    def _init(m):
        if isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Embedding):
            nn.init.normal_(m.weight, std=0.02)
        elif isinstance(m, nn.LayerNorm):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)

    @staticmethod
    def get_default_config():
        return GPTConfig()

    def forward(self, idx):
        return self.lm_head(self.transformer(idx))

    @torch.no_grad()
    def generate(self, idx, max_new_tokens, do_sample=False, temperature=1.0):
        """Greedy / sampled decoding (reference mingpt generate; used by examples/sorter_inference)."""
        for _ in range(max_new_tokens):
            cond = idx if idx.size(1) <= self.block_size else idx[:, -self.block_size:]
            logits = self(cond)[:, -1, :] / temperature
            if do_sample:
                nxt = torch.multinomial(torch.softmax(logits, dim=-1), 1)
            else:
                nxt = logits.argmax(dim=-1, keepdim=True)
            idx = torch.cat((idx, nxt), dim=1)
        return idx
