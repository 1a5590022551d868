"""Offline planner: ``clusterize`` and helpers.

Parity targets (behaviour): reference ``ravnest/operations/utils.py`` —
``spawn_node_pool :24``, ``cluster_formation :52``, ``round_percentages :69``,
``calculate_cluster_power :82``, ``calculate_split_percentages :92``,
``split_model_equal :265``, ``get_memory_reqs :357``, ``clusterize :380`` and the
artefact layout of SURVEY.md §5.4 / appendix A.3-A.4.

New for B200:
  * ``topology={'stages': S, 'replicas': R}`` (or ``"gpu"`` keys in
    ``node_configs.json``) plans an S x R grid on one NVSwitch box without the
    GA — every replica cuts identically so each stage has exactly one DP ring;
  * ``proportions=`` overrides the equal-share policy (root/stem stages pay an
    extra recompute forward, so a balanced pipeline gives them fewer params);
  * a ``plan.json`` per node with shapes/dtypes/wiring for the static NCCL plan;
  * no dependency on pippy / torchinfo / TorchScript.
"""
import json
import math
import os
import shutil
from typing import Tuple, Dict, List, Optional, Sequence

import torch

from .cluster import Cluster
from .genetic import genetic_algorithm
from .node import Node
from .splitter import SplitResult, split_model
from ..utils.serialization import dump_json, dump_pickle, save_submodel

NODE_DATA = 'node_data'


def _root(base_dir: Optional[str]) -> str:
    return os.path.join(base_dir or os.environ.get('RAVNEST_HOME', '.'), NODE_DATA)


def delete_all_folders(path: str) -> None:
    """Remove every sub-folder of ``path`` but keep loose files such as node_configs.json
    (reference operations/utils.py:351-355, reset.py:4-12)."""
    if not os.path.isdir(path):
        return
    for entry in os.listdir(path):
        full = os.path.join(path, entry)
        if os.path.isdir(full):
            shutil.rmtree(full)


def spawn_node_pool(num_nodes=None, mode='load_from_configs', ram_variants=None, bandwidth_variants=None,
                    base_dir=None, configs=None, seed=0):
    """Build the planner's node pool.

    ``load_from_configs`` reads ``node_data/node_configs.json`` (schema of the reference's
    node_data/node_configs.json:1-23 plus an optional ``"gpu"`` ordinal); RAM is given in GB and
    converted to MB.  The ``random`` mode — broken in the reference (operations/utils.py:40-49
    references undefined names) — works here and is seeded.
    """
    import random as _random
    pool: List[Node] = []
    total_ram = 0
    if mode == 'load_from_configs':
        if configs is None:
            with open(os.path.join(_root(base_dir), 'node_configs.json')) as fh:
                configs = json.load(fh)
        for nid in range(len(configs)):
            cfg = configs[str(nid)] if str(nid) in configs else configs[nid]
            bench = dict(cfg.get('benchmarks', {}))
            bench['ram'] = int(bench.get('ram', 180) * 1024)
            bench.setdefault('bandwidth', 900)
            total_ram += bench['ram']
            pool.append(Node(node_id=nid, address=cfg['IP'], benchmarks=bench, gpu=cfg.get('gpu')))
    else:
        rng = _random.Random(seed)
        ram_variants = ram_variants or [8, 16, 32]
        bandwidth_variants = bandwidth_variants or [10, 100, 1000]
        for nid in range(int(num_nodes)):
            bench = {'ram': rng.choice(ram_variants) * 1024, 'bandwidth': rng.choice(bandwidth_variants)}
            total_ram += bench['ram']
            addr = '{}:{}'.format('.'.join(str(rng.randint(1, 254)) for _ in range(4)), 8080)
            pool.append(Node(node_id=nid, address=addr, benchmarks=bench))
    return pool, total_ram


def round_percentages(percentages: Sequence[float]) -> List[int]:
    """Largest-remainder rounding so the integer percentages sum to 100."""
    whole = [int(p) for p in percentages]
    frac = [p - w for p, w in zip(percentages, whole)]
    for _ in range(100 - sum(whole)):
        i = frac.index(max(frac))
        whole[i] += 1
        frac[i] = -1.0
    return whole


def calculate_cluster_power(cluster: Cluster) -> None:
    cluster.total_ram = sum(n.benchmarks['ram'] for n in cluster.nodes.values())
    cluster.total_speed = int(sum(n.benchmarks['ram'] / max(1, n.benchmarks['bandwidth']) for n in cluster.nodes.values()))


def calculate_split_percentages(cluster: Cluster, full_model_size: int) -> None:
    rams = cluster.rams
    if sum(rams) < full_model_size:
        raise ValueError('The sum of the cluster rams ({}) does not exceed full_model_size ({}).'.format(
            sum(rams), full_model_size))
    pct = round_percentages([r / sum(rams) * 100 for r in rams])
    splits = [full_model_size * p // 100 for p in pct]
    for i in range(full_model_size - sum(splits)):
        splits[i % len(splits)] += 1
    cluster.splits = splits
    cluster.assign_split_quotas_to_nodes()


def cluster_formation(full_model_size: int, node_pool: List[Node], seed: int = 42) -> List[Cluster]:
    groups = genetic_algorithm(node_pool, full_model_size, seed=seed)
    clusters = []
    for cid, key in enumerate(sorted(groups)):
        cluster = Cluster(cid)
        for node in groups[key]:
            cluster.add_node(node)
        calculate_split_percentages(cluster, full_model_size)
        calculate_cluster_power(cluster)
        clusters.append(cluster)
    return clusters


def grid_formation(node_pool: List[Node], stages: int, replicas: int) -> List[Cluster]:
    """B200 topology mode: replica r owns nodes [r*stages, (r+1)*stages)."""
    if stages * replicas != len(node_pool):
        raise ValueError('topology {}x{} needs {} nodes, pool has {}'.format(stages, replicas, stages * replicas, len(node_pool)))
    clusters = []
    for r in range(replicas):
        cluster = Cluster(r)
        for s in range(stages):
            cluster.add_node(node_pool[r * stages + s])
        calculate_cluster_power(cluster)
        clusters.append(cluster)
    return clusters


def get_memory_reqs(model=None, input_size=None, input_data=None, depth=3, call_args=None, call_kwargs=None) -> int:
    """Peak-memory estimate in MB = inputs + every leaf-module output + parameters.

    Same accounting as the reference's torchinfo-based estimate (operations/utils.py:357-378) but
    measured with forward hooks on one no-grad pass.
    """
    if call_args is not None or call_kwargs is not None:
        input_data = None
    else:
        assert input_size is not None or input_data is not None
        if input_data is None:
            input_data = torch.rand(tuple(input_size))
    if input_data is None:
        args, kwargs = tuple(call_args or ()), dict(call_kwargs or {})
    elif isinstance(input_data, dict):
        args, kwargs = (), input_data
    elif isinstance(input_data, (tuple, list)):
        args, kwargs = tuple(input_data), {}
    else:
        args, kwargs = (input_data,), {}
    out_bytes = [0]

    def hook(_m, _i, out):
        for t in (out if isinstance(out, (tuple, list)) else (out,)):
            if isinstance(t, torch.Tensor):
                out_bytes[0] += t.numel() * t.element_size()

    handles = [m.register_forward_hook(hook) for m in model.modules() if not list(m.children())]
    was = model.training
    model.eval()
    try:
        with torch.no_grad():
            model(*args, **kwargs)
    finally:
        model.train(was)
        for h in handles:
            h.remove()
    in_bytes = sum(t.numel() * t.element_size() for t in list(args) + list(kwargs.values()) if isinstance(t, torch.Tensor))
    param_bytes = sum(p.numel() * p.element_size() for p in model.parameters())
    return int(math.ceil((in_bytes + out_bytes[0] + param_bytes) / 1e6))


def rebalanced_proportions(proportions: Sequence[float], stage_busy_ms: Sequence[float], damping: float = 1.0) -> List[float]:
    """Profile-guided re-planning (new): given the cost shares a plan was cut with and the device time every stage of that plan
    actually spends per step (``StageTracer`` / ``bench.py: stage_busy_ms_per_rank``), return the shares to cut the next plan with so
    that the stages take equal time: a stage that ran ``r`` times longer than the mean gets ``1/r`` of its share.  The planner's
    per-node estimates are only locally right (eager launch gaps, cuDNN heuristics, graph-replay effects); one or two rounds of this
    feedback remove what they miss.  ``clusterize(..., proportions=rebalanced_proportions(...), balance='measured')``."""
    assert len(proportions) == len(stage_busy_ms) and len(proportions) > 0
    mean = sum(stage_busy_ms) / len(stage_busy_ms)
    raw = [p * (mean / max(b, 1e-6)) ** damping for p, b in zip(proportions, stage_busy_ms)]
    total = sum(raw)
    return [r / total for r in raw]


def equal_proportions(n: int) -> List[float]:
    """round(1/n, 1) for all but the last stage, remainder to the last
    (reference operations/utils.py:430-435: n=3 -> [.3,.3,.4], n=4 -> [.2,.2,.2,.4])."""
    if n == 1:
        return [1.0]
    props = [round(1 / n, 1) or round(1 / n, 3) for _ in range(n - 1)]
    props.append(round(1 - sum(props), 6))
    if props[-1] <= 0:
        props = [1.0 / n] * n
    return props


def write_split(result: SplitResult, cluster_path: str, node_paths: Sequence[str]) -> None:
    """Write submod.pt + templates + plan for every stage of one cluster."""
    for st, node_dir in zip(result.stages, node_paths):
        folder = os.path.join(cluster_path, node_dir)
        os.makedirs(folder, exist_ok=True)
        dump_pickle(st.input_template(), os.path.join(folder, '{}_input.pkl'.format(st.name)))
        dump_pickle(st.output_template(), os.path.join(folder, '{}_output.pkl'.format(st.name)))
        if st.index == 0:
            dump_pickle(result.model_inputs_template, os.path.join(folder, 'model_inputs.pkl'))
        plan = st.describe()
        plan['model_input_names'] = result.model_input_names
        plan['model_input_meta'] = result.model_input_meta
        plan['n_stages'] = len(result.stages)
        plan['stage_names'] = [s.name for s in result.stages]
        plan['all_outputs'] = {s.name: [{'targets': o['targets'], 'meta': o['meta']} for o in s.outputs] for s in result.stages}
        plan['all_inputs'] = {s.name: [{'producer': i['producer'], 'out_idx': i['out_idx']} for i in s.inputs] for s in result.stages}
        dump_json(plan, os.path.join(folder, 'plan.json'))
        save_submodel(st.module, os.path.join(folder, 'submod.pt'))


def split_model_equal(model=None, num_splits=None, proportions=None, example_args=(), example_kwargs=None,
                      cluster_path=None, node_paths=None, model_input_node=None, balance='params', cost_scale=1.0, device=None) -> SplitResult:
    """Split + self-check + write artefacts (reference operations/utils.py:265-349)."""
    if proportions is None:
        proportions = equal_proportions(num_splits) if balance == 'params' else [1.0 / num_splits] * num_splits
    result = split_model(model, proportions, example_args, example_kwargs or {}, check=True, balance=balance, cost_scale=cost_scale,
                         device=device)
    write_split(result, cluster_path, node_paths)
    return result


def default_gpu_configs(n: int, base_port: int = 8080, host: str = '127.0.0.1') -> Dict[str, Dict]:
    """node_configs.json content for an n-GPU B200 box."""
    return {str(i): {'IP': '{}:{}'.format(host, base_port + i), 'gpu': i,
                     'benchmarks': {'ram': 180, 'bandwidth': 900 * 8000}} for i in range(n)}


def clusterize(model=None, example_args=(), example_kwargs=None, pass_data=False, *, topology=None,
               proportions=None, balance=None, cost_scale=1.0, plan_device=None, base_dir=None, node_configs=None, seed=42, verbose=False):
    """Plan the whole job: form clusters, split the model per cluster, derive DP rings and write
    ``node_data/cluster_<c>/<ip:port>/{submod.pt, submod_<k>_{input,output}.pkl, model_inputs.pkl, plan.json}``
    plus ``node_data/nodes/node_<id>.json`` (reference operations/utils.py:380-547).
    """
    example_kwargs = dict(example_kwargs or {})
    root = _root(base_dir)
    os.makedirs(root, exist_ok=True)
    delete_all_folders(root)

    # reference operations/utils.py:393-412: with a single positional example and ``pass_data=False`` the model is sized on a random
    # tensor of the example's SHAPE (torchinfo's input_size mode); ``pass_data=True`` — and every kwargs form — runs the given data
    if len(example_args) == 1 and not example_kwargs and not pass_data and isinstance(example_args[0], torch.Tensor) \
            and example_args[0].is_floating_point():
        full_model_size = get_memory_reqs(model, input_size=tuple(example_args[0].shape))
    else:
        full_model_size = get_memory_reqs(model, call_args=tuple(example_args), call_kwargs=example_kwargs)
    if verbose:
        print('Model memory (MB):', full_model_size)

    pool, total_ram = spawn_node_pool(mode='load_from_configs', base_dir=base_dir, configs=node_configs)
    assert total_ram > full_model_size, 'node pool RAM {} MB cannot hold model {} MB'.format(total_ram, full_model_size)
    if topology is not None:
        clusters = grid_formation(pool, int(topology['stages']), int(topology['replicas']))
    else:
        clusters = cluster_formation(full_model_size, pool, seed=seed)

    os.makedirs(os.path.join(root, 'nodes'), exist_ok=True)
    plans: Dict[Tuple, SplitResult] = {}           # identical chains (GPU grids) share one traced / measured plan
    for cluster in clusters:
        chain = cluster.ordered_nodes()
        cluster.wire()
        # reference policy: parameter shares (heterogeneous hosts are sized by RAM).  GPU grids are homogeneous, so unless the
        # caller pins ``proportions`` the stages are balanced by estimated step time instead (balance='compute').
        mode = balance or ('compute' if (topology is not None and proportions is None) else 'params')
        props = list(proportions) if proportions is not None and len(proportions) == len(chain) else None
        key = (len(chain), tuple(props) if props else None, mode)
        if key in plans:
            res = plans[key]
            write_split(res, os.path.join(root, 'cluster_{}'.format(cluster.cid)), [n.address for n in chain])
        else:
            res = plans[key] = split_model_equal(
                model=model, num_splits=len(chain), proportions=props, example_args=example_args, example_kwargs=example_kwargs,
                cluster_path=os.path.join(root, 'cluster_{}'.format(cluster.cid)), node_paths=[n.address for n in chain],
                model_input_node=chain[0].address, balance=mode, cost_scale=cost_scale, device=plan_device)
        if verbose and res.stage_costs:
            print('cluster', cluster.cid, 'estimated stage step times (ms):', [round(1e3 * c, 3) for c in res.stage_costs])

    home = os.path.dirname(root.rstrip('/')) or '.'
    for node in pool:
        node.set_submodel(home)
        node.set_trainable_parameter_keys()
        node.cluster_length = clusters[node.cluster_id].size

    # ---- DP rings: one per stage of the largest cluster (reference :463-497) -----------------
    largest = max(clusters, key=lambda c: c.size)
    rid = 0
    for node in largest.nodes.values():
        if node.trainable_param_keys:
            node.ring_ids[rid] = node.trainable_param_keys[0]
            rid += 1
    for node in largest.nodes.values():
        if not node.ring_ids:   # parameter-free stage: share the previous ring id (it owns no keys anyway)
            node.ring_ids[max(0, rid - 1)] = None
    largest.set_ringwise_params()
    for cluster in clusters:
        if cluster is largest:
            continue
        for node in cluster.nodes.values():
            current = None
            for key in node.trainable_param_keys:
                ring = largest.all_param_to_ring.get(key)
                if ring is not None and ring != current:
                    node.ring_ids[ring] = key
                    current = ring
    ring_size = len(clusters)

    # ---- which node of the NEXT cluster owns each contiguous parameter range (reference :499-516) --
    for ci, cluster in enumerate(clusters):
        nxt = clusters[(ci + 1) % len(clusters)]
        for node in cluster.nodes.values():
            current = None
            for key in node.trainable_param_keys:
                for peer in nxt.nodes.values():
                    if key in peer.trainable_param_keys:
                        if current != peer.address:
                            node.address_to_param[peer.address] = key
                            current = peer.address
                        break

    for node in pool:
        meta = {
            'node_id': node.node_id,
            'local_host': node.address.split(':')[0],
            'local_port': int(node.address.split(':')[1]),
            'template_path': '{}/cluster_{}/{}/'.format(NODE_DATA, node.cluster_id, node.address),
            'rank': node.cluster_id,
            'ring_size': ring_size,
            'cluster_length': node.cluster_length,
            # the reference writes a 1-tuple here by accident (operations/utils.py:529); keep the list
            # wrapper so either loader works
            'param_addresses': [node.address_to_param],
            'ring_ids': {int(k): v for k, v in node.ring_ids.items()},
            'forward_target_host': node.forward_target_host,
            'forward_target_port': int(node.forward_target_port) if node.forward_target_port is not None else None,
            'backward_target_host': node.backward_target_host,
            'backward_target_port': int(node.backward_target_port) if node.backward_target_port is not None else None,
            'node_type': node.role,
            # ---- B200 extensions -------------------------------------------------------------
            'gpu': node.gpu,
            'stage_index': node.stage_index,
            'cluster_node_ids': [n.node_id for n in clusters[node.cluster_id].ordered_nodes()],
            'dp_peer_node_ids': [c.ordered_nodes()[node.stage_index].node_id for c in clusters
                                 if c.size == clusters[node.cluster_id].size] if topology is not None else None,
        }
        dump_json(meta, os.path.join(root, 'nodes', 'node_{}.json'.format(node.node_id)))
    if verbose:
        print('Clusters formed:', clusters)
    return clusters


# ---- reference names (operations/utils.py:200-263, pippy_utils.py:125) ---------------------------------------------------------
def split_model_on_proportions(model, proportions, example_args=(), example_kwargs=None):
    """``split_model`` under the reference's name: trace, cut at the parameter-share thresholds, return the per-stage modules."""
    return split_model(model, proportions, example_args, example_kwargs or {}, check=False)


def remake_proportions(proportions):
    """The reference's retry policy when a split does not run: move 0.1 from the largest share to its neighbour."""
    from .splitter import nudge_proportions
    return nudge_proportions(list(proportions))


def split_on_proportions(gm, proportions):
    """Cut points for a traced graph module (the reference delegates this to pippy's size-threshold pass)."""
    from .splitter import choose_split_points
    return choose_split_points(gm, proportions)


from .splitter import StageTracer as CustomTracer  # noqa: E402  (reference: operations/utils.py:200, unused by its current path)
