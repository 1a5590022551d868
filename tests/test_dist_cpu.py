"""torch.distributed transport on CPU (gloo, world_size 4): 2 stages x 2 replicas, async pipeline,
periodic parameter averaging (reference semantics) and per-step gradient all-reduce (DP 'gradient' mode)."""
import json
import os
import subprocess
import sys

import pytest
import torch

import ravnest_b200 as rv
from ravnest_b200 import models

HERE = os.path.dirname(os.path.abspath(__file__))


def _launch(workdir, port, extra, stages=2, replicas=2, proportions=(0.5, 0.5)):
    rv.set_seed(42)
    rv.clusterize(models.CNN_Net(), example_args=(torch.rand(32, 1, 8, 8),), topology={'stages': stages, 'replicas': replicas},
                  node_configs=rv.default_gpu_configs(4), proportions=list(proportions) if proportions else None, base_dir=str(workdir))
    env = dict(os.environ, RAVNEST_HOME=str(workdir), OMP_NUM_THREADS='1', CUDA_VISIBLE_DEVICES='')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '4', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(HERE, 'dist_worker.py')] + extra
    proc = subprocess.run(cmd, cwd=str(workdir), env=env, capture_output=True, text=True, timeout=420)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    return [json.load(open(os.path.join(str(workdir), 'result.json.{}'.format(r)))) for r in range(4)]


@pytest.mark.slow
@pytest.mark.parametrize('dp_mode,recompute', [('average', 1), ('gradient', 0)])
def test_two_by_two_grid_over_gloo(workdir, free_ports, dp_mode, recompute):
    res = _launch(workdir, 29000 + free_ports(1)[0] % 900, ['--epochs', '3', '--dp-mode', dp_mode, '--recompute', str(recompute)])
    assert [r['node_type'] for r in res] == ['root', 'leaf', 'root', 'leaf']
    assert all(r['n_backwards'] == 12 for r in res)                      # 3 epochs x 4 batches per replica
    # replicas of the same stage end bit-identical (averaging / identical reduced gradients)
    assert res[0]['checksum'] == pytest.approx(res[2]['checksum'], rel=1e-6)
    if dp_mode == 'gradient':
        assert res[1]['checksum'] == pytest.approx(res[3]['checksum'], rel=1e-6)
    losses = [float(l) for l in open(os.path.join(str(workdir), 'losses_1.txt'))]
    assert len(losses) == 12 and losses[-1] < losses[0]


@pytest.mark.slow
def test_four_stage_pipeline_over_gloo(workdir, free_ports):
    """root -> stem -> stem -> leaf on the torch.distributed transport, stages balanced by the planner's step-time estimate."""
    # RAVNEST_HEADER_FRAME=64 forces every message header onto the two-part path (frame + remainder); the other tests in this
    # file use the default 4 KiB frame that holds the whole header
    os.environ['RAVNEST_HEADER_FRAME'] = '64'
    try:
        res = _launch(workdir, 29000 + free_ports(1)[0] % 900, ['--epochs', '2', '--recompute', '0', '--val', '1'], stages=4, replicas=1,
                      proportions=None)
    finally:
        del os.environ['RAVNEST_HEADER_FRAME']
    assert [r['node_type'] for r in res] == ['root', 'stem', 'stem', 'leaf']
    # validation batches travel root -> stems -> leaf as NO_GRAD_FORWARD messages; the leaf scores them once per epoch
    accs = [float(l) for l in open(os.path.join(str(workdir), 'val_accuracies.txt'))]
    assert len(accs) == 2 and all(0.0 <= a <= 1.0 for a in accs)
    assert all(r['n_backwards'] == 16 for r in res[:3])                  # 2 epochs x 8 batches, one replica
    losses = [float(l) for l in open(os.path.join(str(workdir), 'losses_3.txt'))]
    assert len(losses) == 16 and losses[-1] < losses[0]
    # steady state travels on the static wire plan: per lane only the first message of a shape is pickled, the rest carry a fixed
    # binary header (id, action, fpid, flags)
    for r in res:
        hs = r['header_stats']
        assert hs['static'] > 4 * hs['pickled'] > 0, hs


@pytest.mark.slow
def test_bert_three_stage_pipeline_over_gloo(workdir, free_ports):
    """Keyword model inputs, dict targets, ragged padding masks whose additive bias is relayed root -> stem -> leaf, LAMB with
    gradient accumulation: the BERT path of bench.py, on CPU."""
    rv.set_seed(7)
    cfg = models.BertConfig(vocab_size=128, hidden_size=64, num_hidden_layers=3, num_attention_heads=4, intermediate_size=128,
                            max_position_embeddings=32)
    kw = {'input_ids': torch.randint(0, 128, (4, 16)), 'attention_mask': torch.ones(4, 16, dtype=torch.long),
          'token_type_ids': torch.zeros(4, 16, dtype=torch.long)}
    rv.clusterize(models.BertForPreTraining(cfg), example_kwargs=kw, topology={'stages': 3, 'replicas': 1},
                  node_configs=rv.default_gpu_configs(3), base_dir=str(workdir), balance='compute')
    env = dict(os.environ, RAVNEST_HOME=str(workdir), OMP_NUM_THREADS='1', CUDA_VISIBLE_DEVICES='')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '3', '--master-addr', '127.0.0.1',
           '--master-port', str(29000 + free_ports(1)[0] % 900), os.path.join(HERE, 'dist_bert_worker.py')]
    proc = subprocess.run(cmd, cwd=str(workdir), env=env, capture_output=True, text=True, timeout=420)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    res = [json.load(open(os.path.join(str(workdir), 'result.json.{}'.format(r)))) for r in range(3)]
    assert [r['node_type'] for r in res] == ['root', 'stem', 'leaf']
    assert all(r['finite'] for r in res) and all(r['n_backwards'] == 18 for r in res)        # 3 epochs x 6 batches
    losses = [float(l) for l in open(os.path.join(str(workdir), 'losses_2.txt'))]
    assert len(losses) == 9 and all(l == l for l in losses) and losses[-1] < losses[0]       # one line per optimizer update


def _run_multi(workdir, port, world, model_args):
    env = dict(os.environ, RAVNEST_HOME=str(workdir), OMP_NUM_THREADS='1', CUDA_VISIBLE_DEVICES='')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(HERE, 'dist_multi_worker.py')] + model_args
    proc = subprocess.run(cmd, cwd=str(workdir), env=env, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    return [json.load(open(os.path.join(str(workdir), 'result.json.{}'.format(r)))) for r in range(world)]


@pytest.mark.slow
def test_inception_three_stage_pipeline_over_gloo(workdir, free_ports):
    """BASELINE.json config #4 on CPU: Inception-V3 (CIFAR variant) as a 3-stage pipeline, compute-balanced cuts that fall between
    Inception blocks (branch concatenations stay inside a stage)."""
    rv.set_seed(11)
    rv.clusterize(models.Inception3(), example_args=(torch.rand(2, 3, 32, 32),), topology={'stages': 3, 'replicas': 1},
                  node_configs=rv.default_gpu_configs(3), base_dir=str(workdir), balance='compute')
    res = _run_multi(workdir, 29000 + free_ports(1)[0] % 900, 3, ['--model', 'inception', '--epochs', '2'])
    assert [r['node_type'] for r in res] == ['root', 'stem', 'leaf']
    assert all(r['finite'] and r['n_backwards'] == 6 for r in res)
    losses = [float(l) for l in open(os.path.join(str(workdir), 'losses_2.txt'))]
    assert len(losses) == 6 and all(l == l for l in losses)


@pytest.mark.slow
def test_resnet_four_stage_pipeline_over_gloo(workdir, free_ports):
    """BASELINE.json config #2's pipeline shape on CPU: ResNet-50 cut into 4 stages by the reference's parameter shares
    (cluster_formation.py:23-25).  The cuts land inside bottleneck blocks, where an identity block's opening convolution hands out
    ``(conv1(x), x)`` (nn.Conv2d(fork=True)): stages exchange the unpacked tensors, one of them skipping a stage."""
    rv.set_seed(11)
    rv.clusterize(models.resnet50(num_classes=10), example_args=(torch.rand(2, 3, 32, 32),), topology={'stages': 4, 'replicas': 1},
                  proportions=[0.2, 0.2, 0.2, 0.4], balance='params', node_configs=rv.default_gpu_configs(4), base_dir=str(workdir))
    res = _run_multi(workdir, 29000 + free_ports(1)[0] % 900, 4, ['--model', 'resnet', '--epochs', '2'])
    assert [r['node_type'] for r in res] == ['root', 'stem', 'stem', 'leaf']
    assert all(r['finite'] and r['n_backwards'] == 6 for r in res)
    losses = [float(l) for l in open(os.path.join(str(workdir), 'losses_3.txt'))]
    assert len(losses) == 6 and all(l == l for l in losses)


@pytest.mark.slow
def test_sorter_eight_stage_pipeline_over_gloo(workdir, free_ports):
    """BASELINE.json config #5 on CPU: the minGPT sorter as an 8-stage pure pipeline, one node per stage, tensor labels."""
    rv.set_seed(11)
    cfg = models.GPTConfig(model_type='gpt-nano', vocab_size=3, block_size=11, embd_pdrop=0.0, resid_pdrop=0.0, attn_pdrop=0.0)
    rv.clusterize(models.GPT(cfg), example_args=(torch.randint(0, 3, (32, 11)),), topology={'stages': 8, 'replicas': 1},
                  node_configs=rv.default_gpu_configs(8), base_dir=str(workdir), balance='compute')
    res = _run_multi(workdir, 29000 + free_ports(1)[0] % 900, 8, ['--model', 'sorter', '--epochs', '6'])
    assert [r['node_type'] for r in res] == ['root'] + ['stem'] * 6 + ['leaf']
    assert all(r['finite'] and r['n_backwards'] == 24 for r in res)
    losses = [float(l) for l in open(os.path.join(str(workdir), 'losses_7.txt'))]
    assert len(losses) == 24 and losses[-1] < losses[0]


@pytest.mark.slow
def test_gradient_mode_data_parallel_equals_large_batch_sgd(workdir, free_ports):
    """Two solo replicas in ``dp_mode='gradient'`` (all-reduce of the gradient bucket inside every optimizer step), each on half
    of every global batch, must follow exactly the trajectory of one process doing SGD on the full batches."""
    import sys as _sys
    _sys.path.insert(0, HERE)
    from dist_worker import digits_like, loss_fn
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(64, 32), torch.nn.Tanh(), torch.nn.Linear(32, 10), torch.nn.Softmax(dim=-1))
    rv.set_seed(42)
    rv.clusterize(model, example_args=(torch.rand(32, 1, 8, 8),), topology={'stages': 1, 'replicas': 2},
                  node_configs=rv.default_gpu_configs(2), base_dir=str(workdir))
    env = dict(os.environ, RAVNEST_HOME=str(workdir), OMP_NUM_THREADS='1', CUDA_VISIBLE_DEVICES='')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(29000 + free_ports(1)[0] % 900), os.path.join(HERE, 'dist_worker.py'), '--epochs', '2', '--dp-mode', 'gradient',
           '--recompute', '0', '--optimizer', 'sgd', '--dump', '1']
    proc = subprocess.run(cmd, cwd=str(workdir), env=env, capture_output=True, text=True, timeout=420)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    got = [torch.load(os.path.join(str(workdir), 'master.{}.pt'.format(r))) for r in range(2)]
    # reference: same initial weights, global batch k = [replica 0 batch k ; replica 1 batch k], mean loss
    import copy
    ref = copy.deepcopy(model)
    opt = torch.optim.SGD(ref.parameters(), lr=0.05)
    x, y = digits_like()
    for _epoch in range(2):
        for k in range(4):
            xb = torch.cat([x[k * 32:(k + 1) * 32], x[128 + k * 32:128 + (k + 1) * 32]])
            yb = torch.cat([y[k * 32:(k + 1) * 32], y[128 + k * 32:128 + (k + 1) * 32]])
            opt.zero_grad()
            loss_fn(ref(xb), (None, yb)).backward()
            opt.step()
    for name, p in ref.named_parameters():
        assert torch.allclose(got[0][name], got[1][name], atol=1e-7), name             # replicas identical
        assert torch.allclose(got[0][name], p.detach(), atol=1e-5), name               # ... and equal to large-batch SGD


@pytest.mark.slow
def test_replanning_rebuilds_nodes_in_the_same_processes(workdir, free_ports):
    """Stop a 4-stage pipeline, cut the plan again with ``rebalanced_proportions`` and train on new Node objects over the same
    process group / communicators (what ``bench.py --rebalance`` does between its pre-warm and its warm-up)."""
    res = _launch(workdir, 29000 + free_ports(1)[0] % 900, ['--epochs', '2', '--recompute', '0', '--replan', '1'], stages=4, replicas=1,
                  proportions=None)
    assert [r['node_type'] for r in res] == ['root', 'stem', 'stem', 'leaf']
    assert all(r['n_backwards'] == 16 for r in res[:3])
    losses = [float(l) for l in open(os.path.join(str(workdir), 'losses2_3.txt'))]
    assert len(losses) == 16 and losses[-1] < losses[0]
    shares = rv.rebalanced_proportions([0.25] * 4, [1.0, 1.5, 2.0, 2.5])
    assert abs(sum(shares) - 1.0) < 1e-9 and shares[0] > shares[1] > shares[2] > shares[3]
