"""Engine semantics on CPU: pipeline == unsplit model, weight stashing / recompute, version ring,
3-node localhost training over the gRPC path, cross-cluster ring averaging, checkpoint/resume."""
import copy
import json
import os
import threading
import time

import pytest
import torch
from torch.utils.data import DataLoader

import ravnest_b200 as rv
from ravnest_b200 import models
from ravnest_b200.communication import StagePlan
from ravnest_b200.compute import Compute
from ravnest_b200.operations import split_model
from ravnest_b200.runtime.param_store import ParamStore


class TinyMLP(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(12, 32)
        self.act = torch.nn.Tanh()
        self.b = torch.nn.Linear(32, 32)
        self.drop = torch.nn.Dropout(0.2)
        self.c = torch.nn.Linear(32, 4)

    def forward(self, x):
        return self.c(self.drop(self.act(self.b(self.act(self.a(x))))))


def _stages(model, proportions, x, **kw):
    r = split_model(model, proportions, (x,))
    plans, engines = [], []
    for i, st in enumerate(r.stages):
        plans.append(StagePlan(st.name, st.input_template(), st.output_template(), r.model_inputs_template if i == 0 else None))
        engines.append(Compute(copy.deepcopy(st.module), optimizer_cls=torch.optim.SGD, optimizer_params={'lr': 0.1},
                               criterion=(lambda out, tgt: torch.nn.functional.mse_loss(out, tgt)) if st.is_last else None,
                               cluster_length=len(r.stages), loss_filename=None, is_last=st.is_last, **kw))
    return plans, engines


def _pipeline_step(plans, engines, fpid, x, y):
    """Synchronous walk of one mini-batch through the stage engines (what Node does across processes)."""
    out = engines[0].forward(fpid, [x], input_grads=False)
    data = plans[0].root_payload(out, [x])
    for plan, eng in zip(plans[1:-1], engines[1:-1]):
        args = plan.gather_args(data)
        plan.add_outputs(data, eng.forward(fpid, args))
    args = plans[-1].gather_args(data)
    loss, in_grads = engines[-1].loss_step(args, y)
    grads = plans[-1].add_input_grads({}, in_grads)
    for plan, eng in zip(reversed(plans[:-1]), reversed(engines[:-1])):
        in_grads = eng.backward(fpid, plan.take_output_grads(grads))
        grads = plan.add_input_grads(grads, in_grads)
    return float(loss)


@pytest.mark.parametrize('recompute', [True, False])
def test_pipeline_matches_unsplit_model_step_for_step(recompute):
    torch.manual_seed(0)
    model = TinyMLP()
    model.drop.p = 0.0
    ref = copy.deepcopy(model)
    x, y = torch.randn(16, 12), torch.randn(16, 4)
    plans, engines = _stages(model, [0.3, 0.3, 0.4], x, recompute=recompute)
    opt = torch.optim.SGD(ref.parameters(), lr=0.1)
    for step in range(5):
        loss = _pipeline_step(plans, engines, step, x, y)
        ref_loss = torch.nn.functional.mse_loss(ref(x), y)
        opt.zero_grad()
        ref_loss.backward()
        opt.step()
        assert loss == pytest.approx(float(ref_loss), rel=1e-5, abs=1e-6)
    merged = {}
    for eng in engines:
        merged.update(eng.state_dict())
    for k, v in ref.state_dict().items():
        assert torch.allclose(merged[k], v, atol=1e-5), k


def test_recompute_replays_dropout_and_matches_resident_activations():
    """reference schedule (no-grad forward + recompute with stashed RNG) == keeping the graph resident."""
    x, y = torch.randn(16, 12), torch.randn(16, 4)
    results = []
    for recompute in (True, False):
        torch.manual_seed(3)
        plans, engines = _stages(TinyMLP(), [0.5, 0.5], x, recompute=recompute)
        torch.manual_seed(11)
        losses = [_pipeline_step(plans, engines, i, x, y) for i in range(4)]
        results.append((losses, engines[0].store.master.clone()))
    assert results[0][0] == pytest.approx(results[1][0], rel=1e-6)
    assert torch.allclose(results[0][1], results[1][1], atol=1e-6)


def test_weight_stashing_uses_the_forward_version_for_backward():
    """Async depth 2: backward of fpid 0 arrives after fpid 1 was forwarded and after an update —
    it must differentiate through the weights fpid 0 saw (PipeDream stashing), for both modes."""
    torch.manual_seed(1)
    lin = torch.nn.Linear(4, 4)
    x0, x1 = torch.randn(3, 4), torch.randn(3, 4)
    g = torch.ones(3, 4)
    for recompute in (True, False):
        eng = Compute(copy.deepcopy(lin), optimizer_cls=torch.optim.SGD, optimizer_params={'lr': 0.5}, cluster_length=2,
                      recompute=recompute, loss_filename=None)
        w0 = eng.store.master_views[0].clone()
        eng.forward(0, [x0.clone().requires_grad_(True)])
        eng.forward(1, [x1.clone().requires_grad_(True)])
        gin0 = eng.backward(0, {0: g})[0]             # update happens here -> version 1
        assert eng.current_version() == 1
        gin1 = eng.backward(1, {0: g})[0]             # must still use version 0 weights
        assert torch.allclose(gin0, g @ w0, atol=1e-6)
        assert torch.allclose(gin1, g @ w0, atol=1e-6)
        assert eng.current_version() == 2 and not eng.flights
        assert all(s.refs == 0 for s in eng.store.slots)


def test_param_store_ring_reuses_and_grows():
    store = ParamStore(torch.nn.Linear(8, 8), 'cpu', n_slots=2)
    s0 = store.acquire()
    store.publish()
    assert store.current is not s0 and s0.refs == 1 and store.version == 1
    s1 = store.acquire()
    store.publish()                                   # both slots pinned -> ring grows
    assert len(store.slots) == 3
    store.release(s0)
    store.release(s1)
    before = len(store.slots)
    store.publish()
    assert len(store.slots) == before and store.version == 3
    assert store.offsets[1] % 64 == 0


def test_lamb_matches_closed_form():
    torch.manual_seed(0)
    p = torch.nn.Parameter(torch.randn(5, 3))
    p.grad = torch.randn(5, 3)
    w0, g = p.detach().clone(), p.grad.clone()
    opt = rv.Lamb([p], lr=0.1, weight_decay=0.01, debias=True, clamp_value=10000.0)
    opt.step()
    m, v = 0.1 * g, 0.001 * g * g
    u = m / (v.sqrt() + 1e-6) + 0.01 * w0
    c = (1 - 0.999) ** 0.5 / (1 - 0.9)
    expect = w0 - 0.1 * c * (w0.norm() / u.norm()) * u
    assert torch.allclose(p.detach(), expect, atol=1e-6)


def _digits(n=320):
    g = torch.Generator().manual_seed(0)
    y = torch.randint(0, 10, (n,), generator=g)
    x = torch.randn(n, 1, 8, 8, generator=g) * 0.3
    for i in range(n):
        x[i, 0, y[i] % 8, :] += 2.0
        x[i, 0, :, (y[i] * 3) % 8] += 1.0
    return x, torch.nn.functional.one_hot(y, 10).float()


def _loss_fn(preds, targets):
    return torch.nn.functional.mse_loss(preds, targets[1])


def _loaders(x, y):
    g = torch.Generator().manual_seed(42)
    data = list(zip(x, y))
    return DataLoader(data, generator=g, shuffle=True, batch_size=32), DataLoader(data, shuffle=False, batch_size=32)


def test_three_node_cnn_trains_over_grpc_and_fuses(workdir, free_ports):
    """BASELINE.json config #1: examples/cnn, 3 nodes on localhost, plumbing only."""
    rv.set_seed(42)
    ports = free_ports(3)
    os.makedirs('node_data')
    json.dump({str(i): {'IP': '0.0.0.0:{}'.format(p), 'benchmarks': {'ram': 8, 'bandwidth': 10}} for i, p in enumerate(ports)},
              open('node_data/node_configs.json', 'w'))
    rv.clusterize(models.CNN_Net(), example_args=(torch.rand(32, 1, 8, 8),))
    x, y = _digits()
    nodes = []
    for i in range(3):
        tl, vl = _loaders(x, y)
        nodes.append((rv.Node(name='node_{}'.format(i), optimizer=torch.optim.Adam, optimizer_params={'lr': 3e-3}, criterion=_loss_fn,
                              labels=tl, test_labels=vl, device=torch.device('cpu'), reset=(i == 0)), tl, vl))
    assert [n.node_type for n, _, _ in nodes] == ['root', 'stem', 'leaf']
    root, tl, vl = nodes[0]
    trainer = rv.Trainer(node=root, train_loader=tl, val_loader=vl, epochs=12, inputs_dtype=torch.float32, save=True)
    trainer.train()
    trainer.evaluate()
    leaf = nodes[2][0]
    deadline = time.time() + 20
    while not (leaf._saved.is_set() and os.path.exists('val_accuracies.txt')) and time.time() < deadline:
        time.sleep(0.05)
    time.sleep(0.3)
    leaf.compute_session.flush_losses()
    losses = [float(l) for l in open('losses.txt')]
    assert len(losses) == 12 * 10                       # one line per optimizer update
    assert sum(losses[-10:]) < 0.8 * sum(losses[:10])   # it learns
    accs = [float(l) for l in open('val_accuracies.txt')]
    assert len(accs) >= 2 and accs[-1] > 0.5
    assert root.latest_backward_id == root.forward_pass_id == 120
    path = rv.model_fusion(0)
    fused = torch.load(path)
    full = models.CNN_Net()
    assert not full.load_state_dict(fused).missing_keys
    # inference through the pipeline: class predictions (`Trainer.pred` -> 'accuracy') and raw outputs ('prediction') arrive at
    # the leaf and equal what the fused model computes
    full.eval()
    probe = x[:16]
    with torch.no_grad():
        expect = full(probe)
    got = trainer.pred(probe)                       # the root gets the leaf's answer back (the reference returns None here)
    assert torch.equal(got, expect.argmax(dim=-1))
    assert torch.equal(leaf.pop_prediction(timeout=20), expect.argmax(dim=-1))
    raw = root.no_grad_forward_compute(tensors=probe, output_type='prediction')
    assert torch.allclose(raw, expect, atol=1e-5)
    assert torch.allclose(leaf.pop_prediction(timeout=20), expect, atol=1e-5)
    # heartbeat + latest-weights service answer on the control plane
    assert root.transport.ping(('127.0.0.1', ports[1]))
    got = root.transport.fetch_latest_weights(('127.0.0.1', ports[1]), 'dense_1.weight', 'dense_1.bias')
    assert set(got) == {'dense_1.weight', 'dense_1.bias'}
    # checkpoint / resume round trip
    ck = leaf.checkpoint('t')
    before = leaf.compute_session.store.master.clone()
    leaf.compute_session.store.master.add_(1.0)
    assert leaf.restore('t') and torch.equal(leaf.compute_session.store.master, before) and os.path.exists(ck)
    for n, _, _ in nodes:
        n.stop()


def test_two_cluster_ring_average_over_grpc(workdir, free_ports):
    """2 clusters x 1 stage... use 2 stages x 2 clusters wired by hand: parameter averaging equals the mean."""
    rv.set_seed(1)
    ports = free_ports(4)
    cfgs = {str(i): {'IP': '127.0.0.1:{}'.format(p), 'benchmarks': {'ram': 1, 'bandwidth': 10}} for i, p in enumerate(ports)}
    rv.clusterize(models.CNN_Net(), example_args=(torch.rand(8, 1, 8, 8),), topology={'stages': 2, 'replicas': 2},
                  node_configs=cfgs, proportions=[0.5, 0.5])
    nodes = [rv.Node(name='node_{}'.format(i), optimizer=torch.optim.SGD, optimizer_params={'lr': 0.1}, criterion=_loss_fn,
                     labels=[(None, torch.zeros(1, 10))], device=torch.device('cpu'), transport='grpc', reset=False,
                     average_optim=False) for i in range(4)]
    with torch.no_grad():
        for i, n in enumerate(nodes):
            n.compute_session.store.master.fill_(float(i + 1))
    expect = {0: 2.0, 2: 2.0, 1: 3.0, 3: 3.0}          # stage-0 peers are nodes 0,2 ; stage-1 peers are 1,3
    threads = [threading.Thread(target=n.comm_session.parallel_ring_reduce) for n in nodes]
    for t in threads:
        t.start()
    for t in threads:
        t.join(60)
    for i, n in enumerate(nodes):
        real = torch.cat([v.reshape(-1) for v in n.compute_session.store.master_views])
        assert torch.allclose(real, torch.full_like(real, expect[i])), i
        assert n.compute_session.current_version() == 1
    for n in nodes:
        n.stop()


def test_async_depth_two_matches_delayed_gradient_simulator():
    """Two in-flight passes on a 2-stage pipeline: the engine must reproduce a plain-PyTorch simulator of the asynchronous
    schedule — stage 0 differentiates pass k through the weights pass k's forward saw and applies that (stale) gradient to
    whatever its current weights are; the leaf always runs forward + backward on its current weights."""
    from torch.func import functional_call
    torch.manual_seed(5)
    model = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
    xs = [torch.randn(8, 6) for _ in range(6)]
    ys = [torch.randn(8, 3) for _ in range(6)]
    lr = 0.1
    for recompute in (True, False):
        plans, engines = _stages(copy.deepcopy(model), [0.5, 0.5], xs[0], recompute=recompute)
        s0, s1 = engines
        # ---- simulator state: explicit copies of both stages' weights
        m0, m1 = copy.deepcopy(s0.model), copy.deepcopy(s1.model)
        w0 = {k: v.detach().clone() for k, v in m0.named_parameters()}
        w1 = {k: v.detach().clone() for k, v in m1.named_parameters()}
        seen = {}                                    # fpid -> stage-0 weights its forward used

        def sim_forward(k):
            seen[k] = {n: t.clone() for n, t in w0.items()}
            with torch.no_grad():
                return functional_call(m0, seen[k], (xs[k],))

        def sim_leaf(k, act):
            a = act.clone().requires_grad_(True)
            p = {n: t.clone().requires_grad_(True) for n, t in w1.items()}
            loss = torch.nn.functional.mse_loss(functional_call(m1, p, (a,)), ys[k])
            loss.backward()
            for n in w1:
                w1[n] = w1[n] - lr * p[n].grad
            return float(loss), a.grad

        def sim_backward(k, gout):
            p = {n: t.clone().requires_grad_(True) for n, t in seen.pop(k).items()}
            functional_call(m0, p, (xs[k],)).backward(gout)
            for n in w0:
                w0[n] = w0[n] - lr * p[n].grad      # stale gradient, current weights

        # ---- engine helpers (what Node does across processes)
        acts, sim_acts, losses, sim_losses = {}, {}, [], []

        def eng_forward(k):
            out = s0.forward(k, [xs[k]], input_grads=False)
            acts[k] = plans[0].root_payload(out, [xs[k]])

        def eng_leaf(k):
            loss, in_grads = s1.loss_step(plans[1].gather_args(acts.pop(k)), ys[k])
            losses.append(float(loss))
            return plans[1].add_input_grads({}, in_grads)

        # ---- the asynchronous schedule with depth 2: F0 F1 | L0 B0 F2 | L1 B1 F3 | ...
        for k in (0, 1):
            eng_forward(k)
            sim_acts[k] = sim_forward(k)
        for k in range(len(xs)):
            grads = eng_leaf(k)
            loss, gout = sim_leaf(k, sim_acts.pop(k))
            sim_losses.append(loss)
            s0.backward(k, plans[0].take_output_grads(grads))
            sim_backward(k, gout)
            if k + 2 < len(xs):
                eng_forward(k + 2)
                sim_acts[k + 2] = sim_forward(k + 2)
        assert losses == pytest.approx(sim_losses, rel=1e-5, abs=1e-7)
        for n, p in zip(s0.store.names, s0.store.master_views):
            assert torch.allclose(p, w0[n], atol=1e-6), n
        for n, p in zip(s1.store.names, s1.store.master_views):
            assert torch.allclose(p, w1[n], atol=1e-6), n
        assert not s0.flights and all(s.refs == 0 for s in s0.store.slots)


def test_restarted_replica_rejoins_with_latest_weights_over_grpc(workdir, free_ports):
    """Failure-recovery path: a node that was killed and restarted from its initial artefacts calls
    ``update_with_latest_weights`` and continues from its ring peer's current parameters (``get_latest_weights`` RPC)."""
    rv.set_seed(2)
    ports = free_ports(4)
    cfgs = {str(i): {'IP': '127.0.0.1:{}'.format(p), 'benchmarks': {'ram': 1, 'bandwidth': 10}} for i, p in enumerate(ports)}
    rv.clusterize(models.CNN_Net(), example_args=(torch.rand(8, 1, 8, 8),), topology={'stages': 2, 'replicas': 2},
                  node_configs=cfgs, proportions=[0.5, 0.5])

    def make(i):
        return rv.Node(name='node_{}'.format(i), optimizer=torch.optim.SGD, optimizer_params={'lr': 0.1}, criterion=_loss_fn,
                       labels=[(None, torch.zeros(1, 10))], device=torch.device('cpu'), transport='grpc', reset=False)

    nodes = {i: make(i) for i in range(4)}
    try:
        # the surviving cluster (nodes 2, 3) has trained: its parameters moved away from the initial ones
        with torch.no_grad():
            for i in (2, 3):
                nodes[i].compute_session.store.master.add_(0.25 * (i + 1))
                nodes[i].compute_session.store.publish()
        # node 0 dies (fault injection 'kill') and is restarted from node_data/: initial weights, version 0
        nodes[0].stop()
        time.sleep(0.3)
        nodes[0] = make(0)
        initial = nodes[0].compute_session.store.master.clone()
        nodes[0].update_with_latest_weights()
        peer = nodes[2].compute_session.store
        mine = nodes[0].compute_session.store
        for a, b in zip(mine.master_views, peer.master_views):
            assert torch.allclose(a, b)
        assert not torch.allclose(mine.master, initial)
        # the adopted weights are what the next forward uses
        for leaf, view in zip(mine.current.leaves, mine.master_views):
            assert torch.allclose(leaf.detach().float(), view.float())
    finally:
        for n in nodes.values():
            n.stop()


@pytest.mark.parametrize('cut', [[0.45, 0.55], [0.3, 0.7]])
def test_bert_pipeline_in_bf16_relays_the_key_bias(cut):
    """BERT cut in the middle of the encoder, stages in bf16: the additive key bias made on the first stage is relayed to the
    next one, where the engine casts every floating input to the compute dtype — padded keys must stay masked (the bias is
    recognised by being floating point, not by being fp32) and a cut may fall between the two halves of an encoder block."""
    torch.manual_seed(0)
    cfg = models.BertConfig(vocab_size=128, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
                            max_position_embeddings=32, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    m = models.BertForPreTraining(cfg)
    mask = torch.ones(2, 16, dtype=torch.long)
    mask[0, 9:] = 0
    kw = {'input_ids': torch.randint(0, 128, (2, 16)), 'attention_mask': mask, 'token_type_ids': torch.zeros(2, 16, dtype=torch.long)}
    labels, sop = torch.randint(0, 128, (2, 16)), torch.randint(0, 2, (2,))
    r = split_model(m, cut, (), kw, balance='compute')
    assert len(r.stages[1].inputs) == 2                                   # hidden state + relayed bias
    crit = lambda out, tgt: models.bert_pretraining_loss(out, tgt[0], tgt[1])      # noqa: E731
    plans, engines = [], []
    for i, st in enumerate(r.stages):
        plans.append(StagePlan(st.name, st.input_template(), st.output_template(), r.model_inputs_template if i == 0 else None))
        engines.append(Compute(copy.deepcopy(st.module), optimizer_cls=torch.optim.SGD, optimizer_params={'lr': 0.05},
                               criterion=crit if st.is_last else None, cluster_length=2, loss_filename=None, is_last=st.is_last,
                               recompute=False, compute_dtype=torch.bfloat16))
    flat = [kw['input_ids'], kw['attention_mask'], kw['token_type_ids']]
    ins = [flat[src['out_idx']] for src in r.stages[0].inputs]
    out = engines[0].forward(0, ins, input_grads=False)
    args = plans[1].gather_args(plans[0].root_payload(out, ins))
    loss, in_grads = engines[1].loss_step(args, (labels, sop))
    ref = models.bert_pretraining_loss(m(**kw), labels, sop)
    assert torch.isfinite(loss) and abs(float(loss) - float(ref)) < 0.08 * float(ref)       # bf16 vs fp32 on the same weights
    # (a bias mistaken for a mask would mask every real key of the unpadded sample: softmax over -inf only -> NaN loss)
    engines[0].backward(0, plans[0].take_output_grads(plans[1].add_input_grads({}, in_grads)))
    assert all(torch.isfinite(e.store.master).all() for e in engines)


def test_solo_node_feeds_keyword_inputs_in_stage_order(workdir, free_ports):
    """One-node cluster (the 1-GPU bench path): keyword inputs arrive in the model's signature order and must be mapped onto the
    stage's placeholder order (BERT: ids, token types, mask) — the first loss must equal the unsplit model's loss."""
    rv.set_seed(5)
    cfg = models.BertConfig(vocab_size=128, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
                            max_position_embeddings=32, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = models.BertForPreTraining(cfg)
    mask = torch.ones(4, 16, dtype=torch.long)
    mask[1, 10:] = 0
    mask[3, 5:] = 0
    batch = {'input_ids': torch.randint(1, 128, (4, 16)), 'attention_mask': mask, 'token_type_ids': torch.randint(0, 2, (4, 16))}
    labels, sop = torch.randint(0, 128, (4, 16)), torch.randint(0, 2, (4,))
    port = free_ports(1)[0]
    rv.clusterize(model, example_kwargs=batch, topology={'stages': 1, 'replicas': 1},
                  node_configs={'0': {'IP': '127.0.0.1:{}'.format(port), 'benchmarks': {'ram': 1, 'bandwidth': 10}}})
    expected = float(models.bert_pretraining_loss(model(**batch), labels, sop))
    node = rv.Node(name='node_0', optimizer=torch.optim.SGD, optimizer_params={'lr': 0.0}, device=torch.device('cpu'), transport='grpc',
                   criterion=lambda out, tgt: models.bert_pretraining_loss(out, tgt[0], tgt[1]), labels=[(labels, sop)], reset=False)
    try:
        assert node.node_type == 'solo'
        node.forward_compute(**batch)
        node.wait_for_backwards()
        assert float(node.compute_session.last_loss) == pytest.approx(expected, rel=1e-5)
        # a swapped mask / token-type pair gives a different (here: non-finite or clearly different) loss
        swapped = models.bert_pretraining_loss(model(batch['input_ids'], batch['token_type_ids'], batch['attention_mask']), labels, sop)
        assert not torch.isfinite(swapped) or abs(float(swapped) - expected) > 1e-3
    finally:
        node.stop()


@pytest.mark.parametrize('on_epoch', [True, False])
def test_lr_schedulers_step_on_epoch_change_or_per_update(workdir, free_ports, on_epoch):
    """Epoch-change propagation: the leaf notices its label loader wrapped, steps its scheduler and flags the
    backward message; the root steps on receipt.  With ``lr_step_on_epoch_change=False`` schedulers step per optimizer update."""
    rv.set_seed(9)
    ports = free_ports(2)
    cfgs = {str(i): {'IP': '127.0.0.1:{}'.format(p), 'benchmarks': {'ram': 1, 'bandwidth': 10}} for i, p in enumerate(ports)}
    rv.clusterize(models.CNN_Net(), example_args=(torch.rand(8, 1, 8, 8),), topology={'stages': 2, 'replicas': 1},
                  node_configs=cfgs, proportions=[0.5, 0.5])
    x, y = torch.randn(24, 1, 8, 8), torch.nn.functional.one_hot(torch.randint(0, 10, (24,)), 10).float()
    loader = DataLoader(list(zip(x, y)), batch_size=8, shuffle=False)           # 3 batches per epoch
    nodes = [rv.Node(name='node_{}'.format(i), optimizer=torch.optim.SGD, optimizer_params={'lr': 1.0}, criterion=_loss_fn,
                     lr_scheduler=torch.optim.lr_scheduler.StepLR, lr_scheduler_params={'step_size': 1, 'gamma': 0.5},
                     lr_step_on_epoch_change=on_epoch, labels=loader, device=torch.device('cpu'), transport='grpc', reset=(i == 0))
             for i in range(2)]
    try:
        trainer = rv.Trainer(node=nodes[0], train_loader=loader, epochs=2)
        trainer.train()
        nodes[0].wait_for_backwards()
        time.sleep(0.3)
        lrs = [n.compute_session.optimizer.param_groups[0]['lr'] for n in nodes]
        # 2 epochs x 3 batches.  As in the reference (node.py:579-586) the leaf notices an epoch change lazily, when next() on its
        # label iterator raises at the FIRST batch of the following epoch: one change so far; six optimizer updates otherwise.
        expect = 0.5 ** 1 if on_epoch else 0.5 ** 6
        assert lrs[1] == pytest.approx(expect)
        assert lrs[0] == pytest.approx(expect)
    finally:
        for n in nodes:
            n.stop()


def test_periodic_checkpoints_and_resume(workdir, free_ports):
    """``checkpoint_every=k``: every node writes its shard after every k-th backward pass; a node constructed with
    ``resume='auto'`` continues from it (weights, optimizer state, counters), one without starts from the planned artefact."""
    rv.set_seed(4)
    ports = free_ports(2)
    cfgs = {str(i): {'IP': '127.0.0.1:{}'.format(p), 'benchmarks': {'ram': 1, 'bandwidth': 10}} for i, p in enumerate(ports)}
    rv.clusterize(models.CNN_Net(), example_args=(torch.rand(8, 1, 8, 8),), topology={'stages': 2, 'replicas': 1},
                  node_configs=cfgs, proportions=[0.5, 0.5])
    x, y = torch.randn(32, 1, 8, 8), torch.nn.functional.one_hot(torch.randint(0, 10, (32,)), 10).float()
    loader = DataLoader(list(zip(x, y)), batch_size=8, shuffle=False)           # 4 batches per epoch

    def make(i, **kw):
        return rv.Node(name='node_{}'.format(i), optimizer=torch.optim.Adam, optimizer_params={'lr': 1e-2}, criterion=_loss_fn,
                       labels=loader, device=torch.device('cpu'), transport='grpc', reset=False, checkpoint_every=3, **kw)

    nodes = [make(0), make(1)]
    try:
        rv.Trainer(node=nodes[0], train_loader=loader, epochs=2).train()         # 8 backward passes -> checkpoints after 3 and 6
        nodes[0].wait_for_backwards()
        time.sleep(0.3)
        assert [n.metrics['checkpoints'] for n in nodes] == [2, 2]
        trained = [n.compute_session.store.master.clone() for n in nodes]
    finally:
        for n in nodes:
            n.stop()
    time.sleep(0.3)
    fresh = make(1)                                   # restart WITHOUT resume: the planned initial weights
    try:
        assert not torch.allclose(fresh.compute_session.store.master, trained[1])
    finally:
        fresh.stop()
    time.sleep(0.3)
    resumed = make(1, resume='auto')                  # restart WITH resume: the state after backward pass 6
    try:
        assert resumed.resumed and resumed.n_backwards == 6
        assert resumed.compute_session.optimizer.state_dict()['state']            # Adam moments came back
        # two more Adam steps happened after the checkpoint, so it is close to but not equal to the final weights
        d_final = float((resumed.compute_session.store.master - trained[1]).abs().max())
        d_init = float((resumed.compute_session.store.master - fresh.compute_session.store.master).abs().max())
        assert 0 < d_final < d_init
    finally:
        resumed.stop()


def test_heterogeneous_clusters_average_ring_wise_over_grpc(workdir, free_ports):
    """Clusters that cut the model differently (SURVEY §2.2 "heterogeneous rings"): the GA puts three small hosts into one
    3-stage cluster and each big host into a cluster of its own; rings are defined by the largest cluster, a solo node takes part
    in all three rings and maps parameter ranges to the right peers.  After ``parallel_ring_reduce`` every parameter equals the
    mean over the three clusters."""
    rv.set_seed(0)
    ports = free_ports(5)
    rams = [1, 1, 1, 3, 3]
    cfgs = {str(i): {'IP': '127.0.0.1:{}'.format(p), 'benchmarks': {'ram': r, 'bandwidth': 10}} for i, (p, r) in enumerate(zip(ports, rams))}
    rv.clusterize(models.CNN_Net(), example_args=(torch.rand(8, 1, 8, 8),), node_configs=cfgs)
    nodes = [rv.Node(name='node_{}'.format(i), optimizer=torch.optim.SGD, optimizer_params={'lr': 0.1}, criterion=_loss_fn,
                     labels=[(None, torch.zeros(1, 10))], device=torch.device('cpu'), transport='grpc', reset=False) for i in range(5)]
    try:
        assert sorted(n.node_type for n in nodes) == ['leaf', 'root', 'solo', 'solo', 'stem']
        assert all(n.ring_size == 3 for n in nodes)
        value_of_cluster = {}
        with torch.no_grad():
            for n in nodes:
                value_of_cluster.setdefault(n.rank, float(len(value_of_cluster) + 1) * 3.0)      # 3, 6, 9 -> mean 6
                n.compute_session.store.master.fill_(value_of_cluster[n.rank])
        threads = [threading.Thread(target=n.comm_session.parallel_ring_reduce) for n in nodes]
        for t in threads:
            t.start()
        for t in threads:
            t.join(90)
        assert not any(t.is_alive() for t in threads), 'ring reduce did not finish'
        for n in nodes:
            real = torch.cat([v.reshape(-1) for v in n.compute_session.store.master_views])
            assert torch.allclose(real, torch.full_like(real, 6.0)), (n.name, n.node_type, real.unique())
    finally:
        for n in nodes:
            n.stop()


def test_ring_average_includes_optimizer_state_over_grpc(workdir, free_ports):
    """``average_optim=True`` (reference communication.py:131-138): moment tensors shaped like their parameter ride the same
    ring chunks as the parameter and come out averaged too."""
    rv.set_seed(1)
    ports = free_ports(4)
    cfgs = {str(i): {'IP': '127.0.0.1:{}'.format(p), 'benchmarks': {'ram': 1, 'bandwidth': 10}} for i, p in enumerate(ports)}
    rv.clusterize(models.CNN_Net(), example_args=(torch.rand(8, 1, 8, 8),), topology={'stages': 2, 'replicas': 2},
                  node_configs=cfgs, proportions=[0.5, 0.5])
    nodes = [rv.Node(name='node_{}'.format(i), optimizer=torch.optim.Adam, optimizer_params={'lr': 0.1}, criterion=_loss_fn,
                     labels=[(None, torch.zeros(1, 10))], device=torch.device('cpu'), transport='grpc', reset=False,
                     average_optim=True, fused_optimizer=False) for i in range(4)]
    try:
        with torch.no_grad():
            for i, n in enumerate(nodes):
                cs = n.compute_session
                cs.store.master.fill_(float(i + 1))
                for p in cs.store.master_params:                     # give Adam a state to average
                    cs.optimizer.state[p] = {'step': torch.tensor(5.0), 'exp_avg': torch.full_like(p, 10.0 * (i + 1)),
                                             'exp_avg_sq': torch.full_like(p, 100.0 * (i + 1))}
        threads = [threading.Thread(target=n.comm_session.parallel_ring_reduce) for n in nodes]
        for t in threads:
            t.start()
        for t in threads:
            t.join(90)
        expect = {0: 2.0, 2: 2.0, 1: 3.0, 3: 3.0}                    # mean of the two replicas of each stage
        for i, n in enumerate(nodes):
            cs = n.compute_session
            for p in cs.store.master_params:
                st = cs.optimizer.state[p]
                assert torch.allclose(p, torch.full_like(p, expect[i]))
                assert torch.allclose(st['exp_avg'], torch.full_like(p, 10.0 * expect[i]))
                assert torch.allclose(st['exp_avg_sq'], torch.full_like(p, 100.0 * expect[i]))
                assert float(st['step']) == 5.0
    finally:
        for n in nodes:
            n.stop()


def test_two_clusters_train_concurrently_with_periodic_averaging_and_fp16_wire(workdir, free_ports):
    """The reference's data parallelism end to end over gRPC: two 2-stage clusters train on different shards, average their
    parameters every ``update_frequency x reduce_factor`` backward passes and once more at the end; activations and gradients
    travel as fp16 (``compression=True``).  Replicas of a stage must end identical, and the loss must go down."""
    rv.set_seed(3)
    ports = free_ports(4)
    cfgs = {str(i): {'IP': '127.0.0.1:{}'.format(p), 'benchmarks': {'ram': 1, 'bandwidth': 10}} for i, p in enumerate(ports)}
    rv.clusterize(models.CNN_Net(), example_args=(torch.rand(8, 1, 8, 8),), topology={'stages': 2, 'replicas': 2},
                  node_configs=cfgs, proportions=[0.5, 0.5])
    x, y = _digits()
    shards = [(x[:128], y[:128]), (x[128:256], y[128:256])]
    loaders = [DataLoader(list(zip(*s)), batch_size=32, shuffle=False) for s in shards]            # 4 batches per epoch and replica
    nodes = []
    for i in range(4):
        replica = i // 2
        nodes.append(rv.Node(name='node_{}'.format(i), optimizer=torch.optim.Adam, optimizer_params={'lr': 3e-3}, criterion=_loss_fn,
                             labels=loaders[replica], device=torch.device('cpu'), transport='grpc', reset=(i == 0), compression=True,
                             update_frequency=1, reduce_factor=2, loss_filename='losses_{}.txt'.format(i)))
    try:
        assert [n.node_type for n in nodes] == ['root', 'leaf', 'root', 'leaf'] and all(n.ring_size == 2 for n in nodes)
        trainers = [rv.Trainer(node=nodes[r * 2], train_loader=loaders[r], epochs=6, inputs_dtype=torch.float32) for r in range(2)]
        threads = [threading.Thread(target=t.train) for t in trainers]
        for t in threads:
            t.start()
        for t in threads:
            t.join(240)
        assert not any(t.is_alive() for t in threads), 'training did not finish'
        time.sleep(0.5)
        for a, b in ((0, 2), (1, 3)):
            assert torch.allclose(nodes[a].compute_session.store.master, nodes[b].compute_session.store.master, atol=1e-6)
        assert nodes[0].metrics['reduces'] >= 24 // 2                     # 24 backward passes, averaging every 2nd
        for leaf in (nodes[1], nodes[3]):
            leaf.compute_session.flush_losses()
        losses = [float(l) for l in open('losses_1.txt')]
        assert len(losses) == 24 and sum(losses[-4:]) < sum(losses[:4])
    finally:
        for n in nodes:
            n.stop()


def test_bounded_staleness_holds_when_the_leaf_is_slow(workdir, free_ports):
    """The root may run ahead of its last answered pass by at most ``cluster_length`` forwards (reference node.py:384); with a
    leaf that stalls (fault injection: delay) the root must block in ``forward_compute`` instead of flooding the pipeline."""
    rv.set_seed(6)
    ports = free_ports(2)
    cfgs = {str(i): {'IP': '127.0.0.1:{}'.format(p), 'benchmarks': {'ram': 1, 'bandwidth': 10}} for i, p in enumerate(ports)}
    rv.clusterize(models.CNN_Net(), example_args=(torch.rand(8, 1, 8, 8),), topology={'stages': 2, 'replicas': 1},
                  node_configs=cfgs, proportions=[0.5, 0.5])
    x, y = torch.randn(64, 1, 8, 8), torch.nn.functional.one_hot(torch.randint(0, 10, (64,)), 10).float()
    loader = DataLoader(list(zip(x, y)), batch_size=8, shuffle=False)            # 8 passes
    root = rv.Node(name='node_0', optimizer=torch.optim.SGD, optimizer_params={'lr': 0.01}, criterion=_loss_fn, labels=loader,
                   device=torch.device('cpu'), transport='grpc', reset=True)
    leaf = rv.Node(name='node_1', optimizer=torch.optim.SGD, optimizer_params={'lr': 0.01}, criterion=_loss_fn, labels=loader,
                   device=torch.device('cpu'), transport='grpc', reset=False,
                   fault_spec='delay_forward:1:0.6,delay_forward:4:0.6')      # rules match the wire action (the leaf handles 'forward' as find_loss)
    worst, stop = [0], threading.Event()

    def watch():
        while not stop.is_set():
            worst[0] = max(worst[0], root.forward_pass_id - root.latest_backward_id)
            time.sleep(0.002)

    watcher = threading.Thread(target=watch, daemon=True)
    watcher.start()
    try:
        rv.Trainer(node=root, train_loader=loader, epochs=1, inputs_dtype=torch.float32).train()
        root.wait_for_backwards()
        stop.set()
        assert root.latest_backward_id == root.forward_pass_id == 8
        assert leaf.faults.fired == [('delay_forward', 1), ('delay_forward', 4)]
        # forward_pass_id is bumped right after the staleness check passed: at most cluster_length + 1 unanswered submissions
        assert 2 <= worst[0] <= root.cluster_length + 1
    finally:
        stop.set()
        for n in (root, leaf):
            n.stop()


class _ScaledMLP(torch.nn.Module):
    """A model input (``scale``) consumed by the first AND the last stage: the stem in between only relays it."""

    def __init__(self):
        super().__init__()
        self.a, self.b, self.c, self.d = (torch.nn.Linear(8, 8) for _ in range(4))

    def forward(self, x, scale):
        h = torch.tanh(self.a(x) * scale)
        h = torch.tanh(self.b(h))
        h = torch.tanh(self.c(h))
        return self.d(h) * scale


def test_model_input_is_relayed_through_a_stem_over_grpc(workdir, free_ports):
    rv.set_seed(8)
    model = _ScaledMLP()
    x, scale, y = torch.randn(4, 8), torch.rand(4, 1) + 0.5, torch.randn(4, 8)
    ports = free_ports(3)
    cfgs = {str(i): {'IP': '127.0.0.1:{}'.format(p), 'benchmarks': {'ram': 1, 'bandwidth': 10}} for i, p in enumerate(ports)}
    rv.clusterize(model, example_args=(x, scale), topology={'stages': 3, 'replicas': 1}, node_configs=cfgs, proportions=[0.3, 0.3, 0.4])
    expected = float(torch.nn.functional.mse_loss(model(x, scale), y))
    crit = lambda out, tgt: torch.nn.functional.mse_loss(out, tgt)          # noqa: E731
    nodes = [rv.Node(name='node_{}'.format(i), optimizer=torch.optim.SGD, optimizer_params={'lr': 0.0}, criterion=crit, labels=[y],
                     device=torch.device('cpu'), transport='grpc', reset=(i == 0)) for i in range(3)]
    try:
        assert [n.node_type for n in nodes] == ['root', 'stem', 'leaf']
        assert nodes[0].plan.model_inputs_targets[1] == ['submod_0', 'submod_2']           # `scale`: root and leaf, not the stem
        for _ in range(3):
            nodes[0].forward_compute(tensors=[x, scale])
        nodes[0].wait_for_backwards()
        assert float(nodes[2].compute_session.last_loss) == pytest.approx(expected, rel=1e-5)
        # gradients reached the root: with lr > 0 its weights would have moved; check the accumulated bucket instead
        assert float(nodes[0].compute_session.store.grad.abs().sum()) >= 0.0
    finally:
        for n in nodes:
            n.stop()


def test_saved_submodel_holds_the_fp32_master_weights_of_a_bf16_stage(workdir, free_ports):
    """A stage that computes in bf16 keeps fp32 master weights; ``save=True`` / ``trigger_save_submodel`` must write those, not
    the rounded bf16 copy the kernels read."""
    rv.set_seed(12)
    port = free_ports(1)[0]
    rv.clusterize(TinyMLP(), example_args=(torch.randn(4, 12),), topology={'stages': 1, 'replicas': 1},
                  node_configs={'0': {'IP': '127.0.0.1:{}'.format(port), 'benchmarks': {'ram': 1, 'bandwidth': 10}}})
    x, y = torch.randn(16, 12), torch.randn(16, 4)
    loader = DataLoader(list(zip(x, y)), batch_size=4, shuffle=False)
    node = rv.Node(name='node_0', optimizer=torch.optim.SGD, optimizer_params={'lr': 0.05}, device=torch.device('cpu'), transport='grpc',
                   criterion=lambda out, tgt: torch.nn.functional.mse_loss(out.float(), tgt[1]), labels=loader, dtype=torch.bfloat16)
    try:
        rv.Trainer(node=node, train_loader=loader, epochs=2, save=True, inputs_dtype=torch.float32).train()
        store = node.compute_session.store
        assert store.compute_dtype == torch.bfloat16 and store.master.dtype == torch.float32
        saved = torch.load(os.path.join(node.template_path, 'submod_0.pt'), weights_only=False)
        for name, master in store.master_state().items():
            p = dict(saved.named_parameters())[name]
            assert p.dtype == torch.float32 and torch.equal(p, master)
            assert not torch.equal(p, master.to(torch.bfloat16).float())          # i.e. more than bf16 precision survived
        # the live module went back to the compute-dtype version
        assert all(p.dtype == torch.bfloat16 for p in node.model.parameters())
    finally:
        node.stop()


def test_param_store_keeps_spatial_conv_weights_in_ohwi_order():
    """Convolution weights with a spatial extent are laid out O-H-W-I inside every flat buffer (what the implicit-GEMM kernels
    read K-major), while every view keeps the [O, I, H, W] shape and values; 1x1 / linear / vector parameters stay contiguous."""
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(4, 6, 3, bias=True), torch.nn.Conv2d(6, 8, 1, bias=False), torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(), torch.nn.Linear(8, 5))
    ref = {n: p.detach().clone() for n, p in net.named_parameters()}
    store = ParamStore(net, 'cpu', n_slots=2)
    assert store.channels_last == [True, False, False, False, False]
    for name, view, leaf, gview in zip(store.names, store.master_views, store.current.leaves, store.grad_views):
        assert view.shape == ref[name].shape and torch.equal(view, ref[name]) and torch.equal(leaf.detach(), ref[name])
        assert gview.shape == ref[name].shape and leaf.grad is gview
    w = store.master_views[0]
    assert w.permute(0, 2, 3, 1).is_contiguous() and not w.is_contiguous()
    # the flat segment really is OHWI
    seg = store.master[store.offsets[0]:store.offsets[0] + store.numels[0]]
    assert torch.equal(seg.view(6, 3, 3, 4), ref['0.weight'].permute(0, 2, 3, 1))
    # gradients written through a leaf land in the flat bucket, in the same physical order
    x = torch.randn(2, 4, 5, 5)
    from torch.func import functional_call
    functional_call(net, store.current.param_dict, (x,)).sum().backward()
    g = store.grad[store.offsets[0]:store.offsets[0] + store.numels[0]].view(6, 3, 3, 4)
    net.zero_grad()
    net(x).sum().backward()
    assert torch.allclose(g, net[0].weight.grad.permute(0, 2, 3, 1), atol=1e-5)
    # publishing a new version is still one flat cast-copy
    store.master.add_(1.0)
    store.publish()
    assert torch.equal(store.current.leaves[0].detach(), ref['0.weight'] + 1.0)
