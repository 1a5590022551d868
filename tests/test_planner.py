"""Planner: GA determinism, split boundaries vs the reference policy, templates, JSON schema."""
import json
import os

import pytest
import torch

import ravnest_b200 as rv
from ravnest_b200 import models
from ravnest_b200.operations import (PlannerNode, calculate_fitness, equal_proportions, genetic_algorithm, round_percentages,
                                     split_model)


def _n(m):
    return sum(p.numel() for p in m.parameters())


def test_param_counts_match_reference_models():
    assert _n(models.CNN_Net()) == 16426
    assert _n(models.resnet50(num_classes=200)) == 23917832
    assert _n(models.Inception3()) == 21638954
    assert _n(models.GPT()) == 85728
    assert _n(models.BertForPreTraining()) == 110106428


def test_equal_proportions_policy():
    assert equal_proportions(3) == [0.3, 0.3, pytest.approx(0.4)]
    assert equal_proportions(4) == [0.2, 0.2, 0.2, pytest.approx(0.4)]
    assert equal_proportions(2) == [0.5, 0.5]
    assert round_percentages([33.3, 33.3, 33.4]) == [33, 33, 34]


def test_cnn_split_matches_reference_boundaries():
    r = split_model(models.CNN_Net(), [0.3, 0.3, 0.4], (torch.rand(64, 1, 8, 8),))
    assert [_n(s.module) for s in r.stages] == [4896, 8448, 3082]
    assert r.stages[1].input_template() == {0: {'submod_0': 'placeholder:tensor'}}
    assert r.stages[0].output_template() == {0: {'target': ['submod_1']}}
    assert r.model_inputs_template == {0: {'target': ['submod_0']}}
    assert 'dense_1.weight' in r.stages[1].module.state_dict()          # original parameter names survive
    assert 'bn_3.weight' in r.stages[2].module.state_dict()


def test_resnet50_split_matches_survey_table():
    """Split policy parity on the reference's own model (torchvision resnet50, cluster_formation.py:23-25)."""
    import torchvision
    m = torchvision.models.resnet50(num_classes=200).eval()
    x = torch.rand(2, 3, 224, 224)
    r = split_model(m, [0.2, 0.2, 0.2, 0.4], (x,))
    assert [_n(s.module) for s in r.stages] == [4337216, 4731392, 3412992, 11436232]
    # cuts land mid-bottleneck: multi-tensor boundaries, one tensor skips a stage
    assert r.stages[1].input_template() == {0: {'submod_0': 0}, 1: {'submod_0': 1}}
    assert r.stages[1].output_template() == {0: {'target': ['submod_2']}, 1: {'target': ['submod_3']}}
    for s in r.stages:
        s.module.eval()
    assert torch.allclose(r.run_local(x), m(x), atol=1e-5)


def test_fused_resnet50_splits_and_matches_torchvision():
    """Our ResNet-50 (fused BN+ReLU+add blocks) keeps torchvision's parameter names and numerics and splits 4 ways."""
    import torchvision
    m = models.resnet50(num_classes=200).eval()
    t = torchvision.models.resnet50(num_classes=200).eval()
    assert not t.load_state_dict(m.state_dict()).missing_keys
    x = torch.rand(2, 3, 96, 96)
    assert torch.allclose(m(x), t(x), atol=1e-5)
    r = split_model(m, [0.2, 0.2, 0.2, 0.4], (x,))
    sizes = [_n(s.module) for s in r.stages]
    assert len(sizes) == 4 and sum(sizes) == 23917832 and sizes[0] == 4337216
    for s in r.stages:
        s.module.eval()
    assert torch.allclose(r.run_local(x), m(x), atol=1e-5)


def test_tuple_valued_nodes_never_cross_a_cut():
    """Identity bottlenecks open with ``conv1(x) -> (y, x)`` (nn.Conv2d(fork=True)): a cut that lands on one of the tuple's getitem
    nodes is moved behind them, so stages exchange tensors only."""
    import operator
    from ravnest_b200.operations.splitter import split_graph, trace_model, propagate_meta
    m = models.resnet50(num_classes=10).eval()
    assert m.layer1[1].conv1.fork and not m.layer1[0].conv1.fork
    x = torch.rand(2, 3, 64, 64)
    gm, names = trace_model(m, (x,))[:2]
    propagate_meta(gm, [x])
    items = [n for n in gm.graph.nodes if n.op == 'call_function' and n.target is operator.getitem]
    assert len(items) == 2 * 12
    for cut in (items[0], items[1], items[7]):
        r = split_graph(gm, [cut], names)
        for st in r.stages:
            st.module.eval()
        assert all(isinstance(o['meta'], dict) or o['meta'] is None or not isinstance(o['meta'], (tuple, list))
                   for st in r.stages for o in st.outputs)
        assert torch.allclose(r.run_local(x), m(x), atol=1e-5)
        assert len(r.stages[0].outputs) == 2          # y and the skip tensor, unpacked on the producing side


def test_inception_and_gpt_splits_run():
    m = models.Inception3().eval()
    x = torch.rand(2, 3, 32, 32)
    r = split_model(m, [0.3, 0.3, 0.4], (x,))
    assert [_n(s.module) for s in r.stages] == [6293536, 5992320, 9353098]
    g = models.GPT().eval()
    idx = torch.randint(0, 3, (4, 11))
    r8 = split_model(g, [1 / 8] * 8, (idx,))
    assert len(r8.stages) == 8 and sum(_n(s.module) for s in r8.stages) == 85728
    for s in r8.stages:
        s.module.eval()
    assert torch.allclose(r8.run_local(idx), g(idx), atol=1e-5)


def test_bert_kwargs_split_routes_mask_to_both_stages():
    cfg = models.BertConfig(vocab_size=512, hidden_size=64, num_hidden_layers=4, num_attention_heads=4, intermediate_size=128,
                            max_position_embeddings=32)
    m = models.BertForPreTraining(cfg).eval()
    kw = {'input_ids': torch.randint(0, 512, (2, 16)), 'attention_mask': torch.ones(2, 16, dtype=torch.long),
          'token_type_ids': torch.zeros(2, 16, dtype=torch.long)}
    r = split_model(m, [0.5, 0.5], (), kw)
    assert r.model_input_names == ['input_ids', 'attention_mask', 'token_type_ids']
    # the integer mask is turned into the additive key bias once, in the stage that holds the first encoder layer; what
    # travels to the leaf is that fp32 bias (an extra output of submod_0), not the mask itself
    assert r.model_inputs_template[1]['target'] == ['submod_0']
    assert len(r.stages[0].outputs) == 2 and len(r.stages[1].inputs) == 2
    for s in r.stages:
        s.module.eval()
    a, b = r.run_local(kw['input_ids'], kw['attention_mask'], kw['token_type_ids'])
    ra, rb = m(**kw)
    assert torch.allclose(a, ra, atol=1e-5) and torch.allclose(b, rb, atol=1e-5)


def test_model_input_used_by_two_stages_is_relayed():
    class TwoUse(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b, self.c = torch.nn.Linear(8, 8), torch.nn.Linear(8, 8), torch.nn.Linear(8, 4)

        def forward(self, x, scale):
            h = self.a(x) * scale
            h = torch.relu(self.b(h))
            return self.c(h) * scale

    m = TwoUse().eval()
    x, scale = torch.randn(3, 8), torch.rand(3, 1)
    r = split_model(m, [0.5, 0.5], (x, scale))
    assert r.model_input_names == ['x', 'scale']
    assert r.model_inputs_template[1]['target'] == ['submod_0', 'submod_1']        # `scale` goes to the root AND the leaf
    for s in r.stages:
        s.module.eval()
    assert torch.allclose(r.run_local(x, scale), m(x, scale), atol=1e-6)


def test_genetic_algorithm_is_deterministic_and_feasible():
    nodes = [PlannerNode(i, '10.0.0.{}:80'.format(i), {'ram': r * 1024, 'bandwidth': b})
             for i, (r, b) in enumerate([(8, 10), (8, 10), (16, 100), (4, 10), (32, 100), (8, 50)])]
    a = genetic_algorithm(nodes, 6000, seed=3)
    b = genetic_algorithm(nodes, 6000, seed=3)
    assert {k: [n.node_id for n in v] for k, v in a.items()} == {k: [n.node_id for n in v] for k, v in b.items()}
    for members in a.values():
        assert sum(n.benchmarks['ram'] for n in members) > 6000
    sol = [0] * len(nodes)
    assert calculate_fitness(sol, nodes, 6000) == 0


def test_clusterize_writes_reference_layout(workdir):
    os.makedirs('node_data')
    cfg = {str(i): {'IP': '0.0.0.0:{}'.format(8080 + i), 'benchmarks': {'ram': 8, 'bandwidth': 10}} for i in range(3)}
    json.dump(cfg, open('node_data/node_configs.json', 'w'))
    rv.clusterize(model=models.CNN_Net(), example_args=(torch.rand(64, 1, 8, 8),))
    base = 'node_data/cluster_0/0.0.0.0:8080'
    assert sorted(os.listdir(base)) == ['model_inputs.pkl', 'plan.json', 'submod.pt', 'submod_0_input.pkl', 'submod_0_output.pkl']
    meta = json.load(open('node_data/nodes/node_1.json'))
    for key in ('node_id', 'local_host', 'local_port', 'template_path', 'rank', 'ring_size', 'cluster_length', 'param_addresses',
                'ring_ids', 'forward_target_host', 'forward_target_port', 'backward_target_host', 'backward_target_port', 'node_type'):
        assert key in meta
    assert meta['node_type'] == 'stem' and meta['cluster_length'] == 3 and meta['ring_size'] == 1
    assert meta['ring_ids'] == {'1': 'dense_1.weight'}
    assert json.load(open('node_data/nodes/node_0.json'))['node_type'] == 'root'
    assert json.load(open('node_data/nodes/node_2.json'))['node_type'] == 'leaf'
    loaded = rv.load_node_json_configs('node_2')
    assert loaded['submod_file'] == 'submod_2' and loaded['ring_ids'] == {2: 'bn_3.weight'}
    rv.reset_node_data()
    assert os.listdir('node_data') == ['node_configs.json']


def test_grid_topology_plan(workdir):
    rv.clusterize(models.CNN_Net(), example_args=(torch.rand(8, 1, 8, 8),), topology={'stages': 2, 'replicas': 2},
                  node_configs=rv.default_gpu_configs(4), proportions=[0.5, 0.5])
    metas = [json.load(open('node_data/nodes/node_{}.json'.format(i))) for i in range(4)]
    assert [m['node_type'] for m in metas] == ['root', 'leaf', 'root', 'leaf']
    assert [m['rank'] for m in metas] == [0, 0, 1, 1] and all(m['ring_size'] == 2 for m in metas)
    assert metas[0]['dp_peer_node_ids'] == [0, 2] and metas[1]['dp_peer_node_ids'] == [1, 3]
    assert list(metas[0]['param_addresses'][0]) == ['127.0.0.1:8082']          # same stage of the next replica


def test_compute_balanced_split_equalises_estimated_step_time():
    """B200 grids balance stages by estimated step time: ResNet-50's parameter-share split puts ~95 % of the work in stage 0."""
    import torchvision
    from ravnest_b200.operations.splitter import split_model
    torch.manual_seed(0)
    model = torchvision.models.resnet50(num_classes=10)
    x = torch.randn(2, 3, 64, 64)
    by_cost = split_model(model, [0.25] * 4, (x,), balance='compute')
    costs = by_cost.stage_costs
    assert len(costs) == 4 and max(costs) / min(costs) < 1.35
    params = [sum(p.numel() for p in st.module.parameters()) for st in by_cost.stages]
    assert params[0] < params[-1] / 4                      # early stages: few parameters, most of the activation work
    by_params = split_model(model, [0.2, 0.2, 0.2, 0.4], (x,), balance='params')
    assert by_params.stage_costs is None
    # both plans compute the same function
    model.eval()
    with torch.no_grad():
        ref = model(x)
        for plan in (by_cost, by_params):
            for st in plan.stages:
                st.module.eval()
            assert torch.allclose(plan.run_local(x), ref, atol=1e-5)


def test_measured_balance_times_forward_and_backward_per_node():
    """Profile-guided planning mechanics on the CPU (wall clock instead of CUDA events): every traced node is run forward and
    differentiated on its own; nodes with in-place updates of their input (torchvision's ``out += identity``) fall back to
    forward-only timing; the resulting stages still compute the model."""
    import torchvision
    from ravnest_b200.operations.splitter import measure_costs, split_model, trace_model
    torch.manual_seed(0)
    model = torchvision.models.resnet18(num_classes=10)
    x = torch.randn(2, 3, 64, 64)
    gm, _ = trace_model(model, (x,))
    costs = measure_costs(gm, [x], 'cpu', iters=1)
    timed = [n for n in gm.graph.nodes if n.op in ('call_module', 'call_function', 'call_method')]
    assert all(costs[n] > 0 for n in timed)
    plan = split_model(model, [0.5, 0.5], (x,), balance='measured', device='cpu')
    assert len(plan.stage_costs) == 2 and max(plan.stage_costs) / min(plan.stage_costs) < 2.0
    model.eval()
    for st in plan.stages:
        st.module.eval()
    with torch.no_grad():
        assert torch.allclose(plan.run_local(x), model(x), atol=1e-5)
