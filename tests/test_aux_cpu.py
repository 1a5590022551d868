"""Aux subsystems on CPU: fault injection, heartbeat, tracing, compression, GA tie-break, misc helpers."""
import os
import time

import pytest
import torch

import ravnest_b200 as rv
from ravnest_b200.communication import compress_tensor, decompress_tensor, split_along_largest_axis
from ravnest_b200.runtime.mailbox import Mailbox
from ravnest_b200.runtime.watchdog import FaultInjector, Heartbeat
from ravnest_b200.utils.tracing import StageTracer


def test_fault_injector_rules():
    fi = FaultInjector('drop_forward:3,delay_backward:5:0.05,kill:9')
    assert fi.check('forward', 2) is None
    assert fi.check('forward', 3) == 'drop'
    t0 = time.time()
    assert fi.check('backward', 5) is None and time.time() - t0 >= 0.05
    with pytest.raises(SystemExit):
        fi.check('forward', 9)
    assert [k for k, _ in fi.fired] == ['drop_forward', 'delay_backward', 'kill']


def test_heartbeat_detects_dead_peer():
    class T:
        def __init__(self):
            self.alive = True

        def ping(self, peer, timeout=1.0):
            return self.alive

    t, dead = T(), []
    hb = Heartbeat(t, [('127.0.0.1', 1)], period=0.02, misses_to_fail=2, on_failure=dead.append)
    hb.start()
    time.sleep(0.1)
    assert not hb.dead_peers()
    t.alive = False
    time.sleep(0.15)
    hb.stop()
    assert dead == [('127.0.0.1', 1)] and hb.dead_peers() == [('127.0.0.1', 1)]


def test_mailbox_serves_backward_first_and_blocks_when_full():
    mb = Mailbox(capacity=2)
    mb.put('forward', 'f0')
    mb.put('forward', 'f1')
    assert not mb.has_room('forward') and mb.has_room('backward')
    mb.put('backward', 'b0')
    assert mb.get(0.1) == ('backward', 'b0')
    assert mb.get(0.1) == ('forward', 'f0')
    assert mb.get(0.1) == ('forward', 'f1')
    assert mb.get(0.05) is None


def test_tracer_on_cpu_is_a_noop_but_counts_idle(tmp_path):
    tr = StageTracer('node_0', torch.device('cpu'), path=str(tmp_path / 'm.jsonl'))
    with tr.phase('forward', 1):
        pass
    tr.idle(0.25)
    s = tr.summary()
    assert s['mailbox_idle_s'] == 0.25 and s['node'] == 'node_0'
    assert os.path.exists(tr.dump())


def test_fp16_compression_round_trip_clamps():
    x = torch.tensor([1.0, -2.5, 1e6, -1e6])
    c = compress_tensor(x)
    assert c.dtype == torch.float16 and torch.isfinite(c).all()
    assert decompress_tensor(c, torch.float32).dtype == torch.float32
    assert compress_tensor(torch.ones(2, dtype=torch.float64)).dtype == torch.float32
    idx = torch.arange(3)
    assert compress_tensor(idx) is idx


def test_ring_chunks_split_along_largest_axis():
    chunks = split_along_largest_axis(torch.zeros(3, 10, 4), 4)
    assert [c.shape[1] for c in chunks] == [3, 3, 2, 2] and all(c.shape[0] == 3 for c in chunks)
    assert len(split_along_largest_axis(torch.tensor(1.0), 3)) == 3


def test_set_seed_and_memory_estimate():
    rv.set_seed(7)
    a = torch.rand(3)
    rv.set_seed(7)
    assert torch.equal(a, torch.rand(3))
    from ravnest_b200 import models
    assert rv.get_memory_reqs(models.CNN_Net(), input_size=(64, 1, 8, 8)) >= 1


def test_reference_named_param_helpers_roundtrip():
    """The reference's utils.py plumbing (optimizer over cloned parameters) works on plain modules with our helpers."""
    import torch
    from ravnest_b200 import utils
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    clones = list(utils.current_model_params_clone(model))
    opt = torch.optim.SGD(clones, lr=0.5)
    model(torch.randn(5, 4)).sum().backward()
    before = [p.detach().clone() for p in model.parameters()]
    utils.load_grads_into_optimizer(model, opt)
    opt.step()
    utils.load_optim_weights_into_model(model, opt)
    for p, b, c in zip(model.parameters(), before, clones):
        assert torch.allclose(p, b - 0.5 * p.grad) and p.data_ptr() == c.data_ptr()
    assert utils.get_trainable_param_names(model) == ['0.weight', '0.bias', '1.weight', '1.bias']
    stash = {n: p.detach().clone() * 0 for n, p in model.named_parameters()}
    versions = [p._version for p in model.parameters()]
    utils.load_state_dict_conserve_versions(model, stash)
    assert all(float(p.abs().sum()) == 0 for p in model.parameters()) and versions == [p._version for p in model.parameters()]
    chunks = utils.create_chunks({'w': torch.arange(24.).reshape(2, 12)}, 3)
    assert chunks['w']['split_axis'] == 1 and [c.shape[1] for c in chunks['w']['data']] == [4, 4, 4]
    oc = utils.create_chunks_optim({'w': {'step': torch.tensor(3.), 'exp_avg': torch.ones(6, 2)}}, 2)
    assert oc['w']['step']['reshape'] and oc['w']['exp_avg']['split_axis'] == 0
    adam = torch.optim.Adam(model.parameters())
    state = {n: {'step': torch.tensor(7.), 'exp_avg': torch.zeros_like(p), 'exp_avg_sq': torch.ones_like(p)} for n, p in model.named_parameters()}
    utils.load_optim_state(adam, state, model)
    assert float(adam.state_dict()['state'][0]['step']) == 7.0
    x = torch.tensor([1e9, -1e9, 1.5])
    assert utils.compress_tensor_float16(x).dtype == torch.float16 and float(utils.compress_tensor_float16(x)[0]) == 65504.0
    assert utils.extract_tensor_from_compression_float16(utils.compress_tensor_float16(x), torch.float32).dtype == torch.float32


def test_planner_reference_aliases():
    import torch
    from ravnest_b200 import operations as ops_plan, models
    res = ops_plan.split_model_on_proportions(models.CNN_Net(), [0.3, 0.3, 0.4], (torch.rand(4, 1, 8, 8),))
    assert len(res.stages) == 3
    assert ops_plan.remake_proportions([0.2, 0.6, 0.2]) == [0.2, 0.5, 0.3]
    assert ops_plan.CustomTracer is ops_plan.StageTracer


def test_transfer_watchdog_flags_stalled_and_failed_transfers():
    """The dist transport's NCCL watchdog: a transfer pending beyond the timeout, or a work object carrying an exception, becomes an
    ``on_failure`` callback (the Node turns it into an exception on the training thread)."""
    import time
    from ravnest_b200.runtime.watchdog import TransferWatchdog

    class Work:
        def __init__(self, done, exc=None):
            self.done, self.exc = done, exc

        def is_completed(self):
            return self.done

        def exception(self):
            return self.exc

    class Transport:
        pass

    tr = Transport()
    now = time.time()
    tr._inflight_since = {'forward': [(Work(False), now - 100.0), (Work(True), now - 100.0), (Work(False), now - 1.0)],
                          'backward': [(Work(True, RuntimeError('ncclRemoteError')), now)]}
    seen = []
    wd = TransferWatchdog(tr, timeout=30.0, on_failure=lambda lane, why: seen.append((lane, why)))
    found = wd.check_once(now)
    assert sorted(l for l, _ in found) == ['backward', 'forward']
    assert any('pending for 100 s' in why for _, why in found) and any('ncclRemoteError' in why for _, why in found)


def test_conv2d_fusion_hints_are_inert_off_the_native_path():
    """``fork`` / ``bn_stats`` change neither parameters nor numerics; on CPU they fall back to plain PyTorch ops."""
    from ravnest_b200 import nn as rnn, ops
    torch.manual_seed(0)
    plain, hinted = torch.nn.Conv2d(8, 16, 1, bias=False), rnn.Conv2d(8, 16, 1, bias=False, fork=True, bn_stats=True)
    hinted.load_state_dict(plain.state_dict())
    assert list(hinted.state_dict()) == list(plain.state_dict())
    x = torch.randn(2, 8, 5, 5, requires_grad=True)
    y, skip = hinted(x)
    assert skip is x and torch.allclose(y, plain(x))
    (y.sum() + (skip * 2).sum()).backward()
    xr = x.detach().clone().requires_grad_(True)
    (plain(xr).sum() + (xr * 2).sum()).backward()
    assert torch.allclose(x.grad, xr.grad)
    assert getattr(ops.conv2d(x, plain.weight, bn_stats=True), '_rvn_bn_sums', None) is None
    for bad in (dict(kernel_size=3, padding=1), dict(kernel_size=1, stride=2), dict(kernel_size=1, bias=True)):
        with pytest.raises(ValueError):
            rnn.Conv2d(8, 16, fork=True, **{'bias': False, **bad})


def test_wgrad_staging_is_a_cuda_only_feature():
    """On CPU the stage keeps the plain path: no fp32 accumulator, ``flush_wgrad`` is a no-op, leaves still accumulate into the bucket."""
    from ravnest_b200.runtime.param_store import ParamStore
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.Flatten(), torch.nn.Linear(4 * 36, 2))
    store = ParamStore(model, 'cpu', n_slots=2)
    assert store.enable_wgrad_staging() is False and store.wgrad_acc is None
    store.flush_wgrad()
    leaf = store.make_leaf(store.current.flat.clone(), 0)
    assert leaf.requires_grad and leaf.grad.data_ptr() == store.grad_views[0].data_ptr() and not hasattr(leaf, '_rvn_wgrad_acc')
    slot = store.acquire()
    torch.func.functional_call(model, slot.param_dict, (torch.randn(2, 3, 8, 8),)).sum().backward()
    assert float(store.grad.abs().sum()) > 0
    store.zero_grad(discard_staged=True)
    assert float(store.grad.abs().sum()) == 0


def test_clusterize_pass_data_selects_how_the_model_is_sized(tmp_path):
    """reference operations/utils.py:393-412: shape-only sizing by default, the caller's tensor with ``pass_data=True``."""
    from ravnest_b200 import models
    seen = []
    x = torch.full((4, 1, 8, 8), 7.0)
    for flag in (False, True):
        rv.set_seed(3)
        seen.clear()
        model = models.CNN_Net()
        model.register_forward_pre_hook(lambda _m, args: seen.append(float(args[0].flatten()[0])))     # (not seen by the FX trace)
        rv.clusterize(model, example_args=(x,), pass_data=flag, base_dir=str(tmp_path / 'nd_{}'.format(int(flag))),
                      node_configs=rv.default_gpu_configs(3))
        assert (seen[0] == 7.0) is flag          # the first real forward is the sizing pass
