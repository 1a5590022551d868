"""BERT-base for pre-training (MLM + sentence-order/NSP heads), FX-traceable, built on ravnest_b200.nn.

Architecture parity: HF ``BertForPreTraining(BertConfig())`` used by the reference
(``cluster_formation.py:49-66``, ``examples/bert/provider.py``): 12 layers, hidden 768, 12 heads,
FFN 3072, vocab 30522, 512 positions, 2 token types, LayerNorm eps 1e-12, dropout 0.1, erf-GELU,
tied MLM decoder — 110 106 428 parameters (tests/test_models.py asserts the count).  Parameter
names follow HF (``bert.encoder.layer.3.attention.self.query.weight`` ...) so HF state dicts load.

Forward signature matches how the reference drives it: ``forward(input_ids, attention_mask,
token_type_ids)`` -> ``(prediction_logits [B,S,V], seq_relationship_logits [B,2])``
(``return_dict=False``, cluster_formation.py:50).
"""
from dataclasses import dataclass

import torch
from torch import nn

from .. import nn as rnn


@dataclass
class BertConfig:
    vocab_size: int = 30522
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    max_position_embeddings: int = 512
    type_vocab_size: int = 2
    layer_norm_eps: float = 1e-12
    hidden_dropout_prob: float = 0.1
    attention_probs_dropout_prob: float = 0.1
    tie_word_embeddings: bool = True


class BertEmbeddings(nn.Module):
    _ravnest_leaf = True      # one FX node: three gathers + add + LayerNorm + dropout

    def __init__(self, cfg: BertConfig):
        super().__init__()
        self.word_embeddings = rnn.Embedding(cfg.vocab_size, cfg.hidden_size, padding_idx=getattr(cfg, 'pad_token_id', 0))
        self.position_embeddings = rnn.Embedding(cfg.max_position_embeddings, cfg.hidden_size)
        self.token_type_embeddings = rnn.Embedding(cfg.type_vocab_size, cfg.hidden_size)
        self.LayerNorm = rnn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)
        self.dropout = rnn.Dropout(cfg.hidden_dropout_prob)

    def forward(self, input_ids, token_type_ids):
        pos = torch.arange(input_ids.shape[1], device=input_ids.device)
        x = self.word_embeddings(input_ids) + self.token_type_embeddings(token_type_ids)
        x = self.LayerNorm(x, self.position_embeddings(pos).unsqueeze(0).expand_as(x))
        return self.dropout(x)


class BertSelfOutput(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.dense = rnn.Linear(cfg.hidden_size, cfg.hidden_size)
        self.LayerNorm = rnn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)
        self.dropout = rnn.Dropout(cfg.hidden_dropout_prob)


class BertAttention(nn.Module):
    """Attention half of an encoder block — one planner leaf, four fused kernels forward:
    QKV GEMM(+bias) -> FMHA(+mask, +dropout) -> proj GEMM(+bias) -> LN(dropout(.) + residual)."""
    _ravnest_leaf = True

    def __init__(self, cfg):
        super().__init__()
        self.self = rnn.SelfAttention(cfg.hidden_size, cfg.num_attention_heads, cfg.attention_probs_dropout_prob)
        self.output = BertSelfOutput(cfg)
        self.hidden_dropout = cfg.hidden_dropout_prob

    def forward(self, x, mask):
        from .. import ops
        p = self.hidden_dropout if self.training else 0.0
        out = self.output
        a = out.dense(self.self(x, mask))
        return ops.layer_norm(a, out.LayerNorm.weight, out.LayerNorm.bias, out.LayerNorm.eps, x, p)


class BertIntermediate(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.dense = rnn.Linear(cfg.hidden_size, cfg.intermediate_size, activation='gelu')


class BertOutput(nn.Module):
    """Feed-forward half of an encoder block — the second planner leaf: FFN1 GEMM(+bias+GELU, act' side output) / FFN2
    GEMM(+bias) as one fused MLP, then LN(dropout(.) + residual).  The FFN1 parameters keep their HF home
    (``intermediate.dense``) and arrive as arguments, so state-dict keys are unchanged."""
    _ravnest_leaf = True

    def __init__(self, cfg):
        super().__init__()
        self.dense = rnn.Linear(cfg.intermediate_size, cfg.hidden_size)
        self.LayerNorm = rnn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)
        self.dropout = rnn.Dropout(cfg.hidden_dropout_prob)
        self.hidden_dropout = cfg.hidden_dropout_prob
        self.activation = 'gelu'

    def forward(self, a, w1, b1):
        from .. import ops
        p = self.hidden_dropout if self.training else 0.0
        h = ops.mlp(a, w1, b1, self.dense.weight, self.dense.bias, self.activation)
        return ops.layer_norm(h, self.LayerNorm.weight, self.LayerNorm.bias, self.LayerNorm.eps, a, p)


class BertLayer(nn.Module):
    """One encoder block = two planner leaves (attention half, feed-forward half), so a pipeline cut can fall in the middle of
    a block: with 12 blocks + embeddings + MLM head on 2 stages, whole-block granularity left the stages 3.6-4.1 vs 4.5-5.1 ms
    busy (profiles/RESULTS_r1.md).  Sub-module / parameter names follow HF (attention.self / attention.output /
    intermediate / output)."""

    def __init__(self, cfg):
        super().__init__()
        self.attention = BertAttention(cfg)
        self.intermediate = BertIntermediate(cfg)
        self.output = BertOutput(cfg)

    def forward(self, x, mask):
        a = self.attention(x, mask)
        return self.output(a, self.intermediate.dense.weight, self.intermediate.dense.bias)


class BertEncoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.layer = nn.ModuleList([BertLayer(cfg) for _ in range(cfg.num_hidden_layers)])
        self.mask_bias = rnn.KeyPaddingBias()

    def forward(self, x, mask):
        mask = self.mask_bias(mask)                 # [B, S] 0 / -inf once, not once per layer
        for layer in self.layer:
            x = layer(x, mask)
        return x


class BertPooler(nn.Module):
    _ravnest_leaf = True

    def __init__(self, cfg):
        super().__init__()
        self.dense = rnn.Linear(cfg.hidden_size, cfg.hidden_size)

    def forward(self, x):
        return torch.tanh(self.dense(x[:, 0]))


class BertModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.embeddings = BertEmbeddings(cfg)
        self.encoder = BertEncoder(cfg)
        self.pooler = BertPooler(cfg)

    def forward(self, input_ids, attention_mask, token_type_ids):
        x = self.embeddings(input_ids, token_type_ids)
        x = self.encoder(x, attention_mask)
        return x, self.pooler(x)


class BertPredictionHeadTransform(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.dense = rnn.Linear(cfg.hidden_size, cfg.hidden_size, activation='gelu')
        self.LayerNorm = rnn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)

    def forward(self, x):
        return self.LayerNorm(self.dense(x))


class BertLMPredictionHead(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.transform = BertPredictionHeadTransform(cfg)
        self.decoder = rnn.Linear(cfg.hidden_size, cfg.vocab_size, bias=True)
        self.bias = self.decoder.bias

    def forward(self, x):
        return self.decoder(self.transform(x))


class BertPreTrainingHeads(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.predictions = BertLMPredictionHead(cfg)
        self.seq_relationship = rnn.Linear(cfg.hidden_size, 2)

    def forward(self, sequence_output, pooled_output):
        return self.predictions(sequence_output), self.seq_relationship(pooled_output)


class BertForPreTraining(nn.Module):
    def __init__(self, config: BertConfig = None):
        super().__init__()
        self.config = config or BertConfig()
        self.bert = BertModel(self.config)
        self.cls = BertPreTrainingHeads(self.config)
        if self.config.tie_word_embeddings:
            self.cls.predictions.decoder.weight = self.bert.embeddings.word_embeddings.weight
        self.apply(self._init)

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

    def forward(self, input_ids, attention_mask, token_type_ids):
        seq, pooled = self.bert(input_ids, attention_mask, token_type_ids)
        prediction_logits, seq_relationship_logits = self.cls(seq, pooled)
        return prediction_logits, seq_relationship_logits


def to_hf_state_dict(model: BertForPreTraining):
    """State dict under HuggingFace's exact key set (``transformers.BertForPreTraining.load_state_dict`` accepts it strictly):
    the fused ``attention.self.qkv`` parameter is split back into ``query`` / ``key`` / ``value`` and the tied decoder bias is
    listed under both of HF's names.  The inverse direction needs no helper: ``load_state_dict`` merges q/k/v on the fly."""
    out = {}
    for k, v in model.state_dict().items():
        if '.attention.self.qkv.' in k:
            prefix, kind = k.split('.attention.self.qkv.')
            H = v.shape[0] // 3
            for i, n in enumerate(('query', 'key', 'value')):
                out['{}.attention.self.{}.{}'.format(prefix, n, kind)] = v[i * H:(i + 1) * H]
        else:
            out[k] = v
    if 'cls.predictions.decoder.bias' in out:
        out.setdefault('cls.predictions.bias', out['cls.predictions.decoder.bias'])
    if 'cls.predictions.bias' in out:
        out.setdefault('cls.predictions.decoder.bias', out['cls.predictions.bias'])
    if 'cls.predictions.decoder.weight' not in out and 'bert.embeddings.word_embeddings.weight' in out:
        out['cls.predictions.decoder.weight'] = out['bert.embeddings.word_embeddings.weight']
    return out


def bert_pretraining_loss(outputs, labels, sentence_order_label, ignore_index=-100):
    """MLM + SOP loss of the reference's BERT provider (examples/bert/provider.py:30-40)."""
    from .. import ops
    logits, sop = outputs
    mlm = ops.cross_entropy(logits.view(-1, logits.shape[-1]), labels.view(-1), ignore_index)
    nsp = ops.cross_entropy(sop.view(-1, 2), sentence_order_label.view(-1), ignore_index)
    return mlm + nsp
