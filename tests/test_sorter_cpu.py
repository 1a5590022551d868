"""GPT-Sorter plumbing on CPU: 2-node gRPC pipeline with plain-tensor labels (examples/sorter/provider.py), save,
model_fusion and greedy decoding through the fused model (examples/sorter/sorter_inference.py)."""
import json
import os
import time

import torch
from torch.utils.data import DataLoader

import ravnest_b200 as rv
from ravnest_b200 import models


def _sort_data(n=256, seed=0):
    g = torch.Generator().manual_seed(seed)
    inp = torch.randint(0, 3, (n, 6), generator=g)
    sol = torch.sort(inp, dim=1)[0]
    cat = torch.cat([inp, sol], dim=1)
    x, y = cat[:, :-1].clone(), cat[:, 1:].clone()
    y[:, :5] = -1
    return x, y


def _criterion(outputs, targets):
    return torch.nn.functional.cross_entropy(outputs.reshape(-1, outputs.size(-1)), targets[1].reshape(-1), ignore_index=-1)


def test_sorter_two_node_pipeline_trains_and_decodes(workdir, free_ports):
    rv.set_seed(42)
    ports = free_ports(2)
    cfg = models.GPTConfig(model_type='gpt-nano', vocab_size=3, block_size=11, embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0)
    model = models.GPT(cfg)
    rv.clusterize(model, example_args=(torch.randint(0, 3, (64, 11)),), topology={'stages': 2, 'replicas': 1},
                  node_configs={str(i): {'IP': '127.0.0.1:{}'.format(p), 'benchmarks': {'ram': 1, 'bandwidth': 10}} for i, p in enumerate(ports)})
    x, y = _sort_data()
    loader = DataLoader(list(zip(x, y)), shuffle=False, batch_size=64)
    nodes = [rv.Node(name='node_{}'.format(i), optimizer=torch.optim.Adam, optimizer_params={'lr': 3e-3}, criterion=_criterion,
                     labels=y, label_batch_size=64, device=torch.device('cpu'), transport='grpc', reset=(i == 0)) for i in range(2)]
    assert [n.node_type for n in nodes] == ['root', 'leaf']
    trainer = rv.Trainer(node=nodes[0], train_loader=loader, epochs=30, save=True)
    trainer.train()
    deadline = time.time() + 20
    while not nodes[1]._saved.is_set() and time.time() < deadline:
        time.sleep(0.05)
    nodes[1].compute_session.flush_losses()
    losses = [float(l) for l in open('losses.txt')]
    assert len(losses) == 30 * 4 and losses[-1] < 0.6 * losses[0]
    fused = torch.load(rv.model_fusion(0))
    full = models.GPT(cfg)
    missing = full.load_state_dict(fused, strict=False)
    # minGPT-style aliases (transformer.wte / wpe point at the same Parameters as transformer.emb.*) are the only keys not in the fused dict
    assert set(missing.missing_keys) <= {'transformer.wte.weight', 'transformer.wpe.weight'}
    assert torch.equal(full.transformer.wte.weight, fused['transformer.emb.wte.weight'])
    full.eval()
    # the saved stage modules chained in-process (examples/sorter/sorter_inference.py) compute what the fused model computes
    chained = rv.ClusterPipeline(cluster_id=0).eval()
    with torch.no_grad():
        assert torch.allclose(chained(x[:8]), full(x[:8]), atol=1e-5)
    probe = torch.tensor([[2, 0, 1, 1, 0, 2]])
    out = full.generate(probe, 6)
    assert out.shape == (1, 12)
    for n in nodes:
        n.stop()
