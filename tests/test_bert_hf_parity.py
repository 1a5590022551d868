"""models/bert.py vs HuggingFace ``BertForPreTraining`` (what the reference trains: examples/bert/provider.py:48-49,
cluster_formation.py:49-66): same state-dict keys, and — with HF's weights copied in — same logits, loss and gradients in fp32."""
import pytest
import torch

import ravnest_b200 as rv
from ravnest_b200 import models, ops

transformers = pytest.importorskip('transformers')


def _pair(layers=2, hidden=64, heads=4, inter=128, vocab=512, maxpos=64):
    hf_cfg = transformers.BertConfig(vocab_size=vocab, hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads,
                                     intermediate_size=inter, max_position_embeddings=maxpos, hidden_dropout_prob=0.0,
                                     attention_probs_dropout_prob=0.0)
    hf_cfg.return_dict = False
    hf = transformers.BertForPreTraining(hf_cfg).eval()
    cfg = models.BertConfig(vocab_size=vocab, hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads,
                            intermediate_size=inter, max_position_embeddings=maxpos, hidden_dropout_prob=0.0,
                            attention_probs_dropout_prob=0.0)
    ours = models.BertForPreTraining(cfg).eval()
    return hf, ours


def test_state_dict_keys_match_hf():
    """Our only structural difference is the fused [3H, H] QKV parameter; ``to_hf_state_dict`` undoes it, and HF loads the result
    strictly.  The other direction (HF -> ours) goes through ``load_state_dict`` directly (next test)."""
    hf, ours = _pair()
    exported = models.to_hf_state_dict(ours)
    hf_keys = {k for k in hf.state_dict() if not k.endswith('position_ids') and not k.endswith('token_type_ids')}
    assert set(exported) == hf_keys, (sorted(set(exported) ^ hf_keys))
    hf.load_state_dict(exported, strict=False)
    for k, v in hf.state_dict().items():
        if k in exported:
            assert torch.equal(v, exported[k]), k


def test_default_config_matches_hf_bert_base():
    hf_cfg = transformers.BertConfig()
    cfg = models.BertConfig()
    for k in ('vocab_size', 'hidden_size', 'num_hidden_layers', 'num_attention_heads', 'intermediate_size', 'max_position_embeddings',
              'type_vocab_size', 'layer_norm_eps', 'hidden_dropout_prob', 'attention_probs_dropout_prob'):
        assert getattr(cfg, k) == getattr(hf_cfg, k), k
    assert hf_cfg.hidden_act == 'gelu'


def test_logits_loss_and_grads_match_hf_fp32():
    torch.manual_seed(0)
    ops.set_backend('torch')
    hf, ours = _pair()
    missing, unexpected = ours.load_state_dict({k: v for k, v in hf.state_dict().items()
                                                if not k.endswith('position_ids') and not k.endswith('token_type_ids')}, strict=False)
    assert set(missing) <= {'cls.predictions.decoder.bias', 'cls.predictions.bias'}
    assert set(unexpected) <= {'cls.predictions.decoder.bias', 'cls.predictions.bias'}
    B, S = 3, 32
    ids = torch.randint(0, 512, (B, S))
    mask = torch.ones(B, S, dtype=torch.long)
    mask[1, 20:] = 0
    mask[2, 9:] = 0
    tt = torch.randint(0, 2, (B, S))
    labels = torch.where(torch.rand(B, S) < 0.3, ids, torch.full_like(ids, -100))
    sop = torch.randint(0, 2, (B,))

    a_logits, a_sop = hf(input_ids=ids, attention_mask=mask, token_type_ids=tt)
    b_logits, b_sop = ours(ids, mask, tt)
    assert torch.allclose(a_logits, b_logits, atol=2e-4, rtol=1e-4), (a_logits - b_logits).abs().max()
    assert torch.allclose(a_sop, b_sop, atol=1e-4, rtol=1e-4)

    ce = torch.nn.CrossEntropyLoss()
    loss_hf = ce(a_logits.view(-1, 512), labels.view(-1)) + ce(a_sop.view(-1, 2), sop.view(-1))     # reference provider.py:30-40
    loss_ours = models.bert_pretraining_loss((b_logits, b_sop), labels, sop)
    assert torch.allclose(loss_hf, loss_ours, atol=1e-5, rtol=1e-5)
    loss_hf.backward()
    loss_ours.backward()
    hf_grads = {k: p.grad for k, p in hf.named_parameters()}
    checked = 0
    for name, p in ours.named_parameters():
        key = name if name in hf_grads else {'cls.predictions.decoder.bias': 'cls.predictions.bias',
                                             'cls.predictions.bias': 'cls.predictions.decoder.bias'}.get(name)
        g = hf_grads.get(key)
        if g is None or p.grad is None:
            continue
        assert torch.allclose(g, p.grad, atol=3e-5, rtol=1e-3), (name, (g - p.grad).abs().max())
        checked += 1
    assert checked >= 30
    ops.set_backend('native' if torch.cuda.is_available() else 'torch')


def test_parameter_count_is_hf_bert_base():
    assert sum(p.numel() for p in models.BertForPreTraining().parameters()) == 110_106_428
    assert rv is not None
