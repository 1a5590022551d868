"""FX-based pipeline-stage splitter (own implementation; no pippy).

What it reproduces from the reference (behaviour, not code):
  * split policy — cumulative parameter-size thresholds derived from a list of
    proportions (``ravnest/operations/pippy_utils.py:43-155``);
  * the three template artefacts written per stage
    (``ravnest/operations/utils.py:280-342``):
      ``submod_<k>_input.pkl``  {arg_pos: {producer: 'placeholder:tensor' | out_idx}}
                                 or {arg_pos: {'model_inputs': input_idx}}
      ``submod_<k>_output.pkl`` {out_idx: {'target': [consumer submods]}}
      ``model_inputs.pkl``      {input_idx: {'target': [consumer submods]}}

What is new:
  * stages are carved directly out of a ``torch.fx`` graph and re-rooted on the
    original module, so every stage keeps the ORIGINAL qualified parameter
    names (``layer3.2.conv2.weight``) — ``model_fusion`` becomes a dict union;
  * a static ``plan`` (shapes / dtypes / producer->consumer wiring) is emitted
    next to the templates so the NCCL transport can move tensors producer ->
    consumer directly (no relay hops, no metadata on the wire);
  * if the threshold policy yields fewer stages than requested (tiny models,
    e.g. gpt-nano over 8 GPUs) the largest stage is bisected until it fits.
"""
import inspect
import operator
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch
import torch.fx as fx
from torch import nn

_LEAF_PREFIXES = ('ravnest_b200.nn',)


class StageTracer(fx.Tracer):
    """Symbolic tracer whose leaves are torch.nn builtins, ravnest_b200.nn modules
    and any user module flagged with ``_ravnest_leaf = True``."""

    def is_leaf_module(self, m: nn.Module, module_qualified_name: str) -> bool:
        if getattr(m, '_ravnest_leaf', False):
            return True
        if m.__class__.__module__.startswith(_LEAF_PREFIXES):
            return True
        return super().is_leaf_module(m, module_qualified_name)


def _strip_none_placeholders(gm: fx.GraphModule, dropped: Sequence[str]) -> None:
    """Remove placeholders that were specialised to a constant default (``concrete_args``)."""
    for node in list(gm.graph.nodes):
        if node.op != 'placeholder' or node.target not in dropped:
            continue
        default = node.args[0] if node.args else None
        for user in list(node.users):
            if user.op == 'call_function' and getattr(user.target, '__name__', '') in ('_assert_is_none', '_assert'):
                gm.graph.erase_node(user)
            else:
                user.replace_input_with(node, default) if isinstance(default, fx.Node) else _replace_const(user, node, default)
        gm.graph.erase_node(node)
    gm.graph.lint()
    gm.recompile()


def _replace_const(user: fx.Node, old: fx.Node, const: Any) -> None:
    user.args = fx.node.map_arg(user.args, lambda n: const if n is old else n)
    user.kwargs = fx.node.map_arg(user.kwargs, lambda n: const if n is old else n)


def trace_model(model: nn.Module, example_args: Sequence = (), example_kwargs: Optional[Dict] = None
                ) -> Tuple[fx.GraphModule, List[str]]:
    """Trace ``model`` and return (graph module, ordered list of model-input names)."""
    example_kwargs = dict(example_kwargs or {})
    sig = inspect.signature(model.forward)
    names = [p for p in sig.parameters if sig.parameters[p].kind in
             (inspect.Parameter.POSITIONAL_OR_KEYWORD, inspect.Parameter.KEYWORD_ONLY)]
    provided = list(names[:len(example_args)]) + [k for k in names[len(example_args):] if k in example_kwargs]
    unknown = [k for k in example_kwargs if k not in names]
    if unknown:
        raise ValueError('example_kwargs {} do not appear in forward{}'.format(unknown, tuple(names)))
    concrete = {}
    for name in names:
        if name not in provided:
            default = sig.parameters[name].default
            concrete[name] = None if default is inspect.Parameter.empty else default
    tracer = StageTracer()
    graph = tracer.trace(model, concrete_args=concrete or None)
    gm = fx.GraphModule(tracer.root, graph, model.__class__.__name__)
    if concrete:
        _strip_none_placeholders(gm, list(concrete))
    placeholders = [n.target for n in gm.graph.nodes if n.op == 'placeholder']
    return gm, placeholders


class _MetaProp(fx.Interpreter):
    """Run the graph once on example inputs recording (shape, dtype) for every node."""

    def run_node(self, n: fx.Node) -> Any:
        out = super().run_node(n)
        n.meta['ravnest_val'] = _describe(out)
        return out


def _describe(v: Any) -> Any:
    if isinstance(v, torch.Tensor):
        return {'shape': list(v.shape), 'dtype': str(v.dtype).replace('torch.', ''),
                'requires_grad': bool(v.is_floating_point() or v.is_complex())}
    if isinstance(v, (tuple, list)):
        return [_describe(x) for x in v]
    if isinstance(v, dict):
        return {k: _describe(x) for k, x in v.items()}
    return {'const': type(v).__name__}


def propagate_meta(gm: fx.GraphModule, inputs: Sequence[Any]) -> Any:
    was_training = gm.training
    gm.eval()
    try:
        with torch.no_grad():
            return _MetaProp(gm).run(*inputs)
    finally:
        gm.train(was_training)


def _node_param_sizes(gm: fx.GraphModule) -> Dict[fx.Node, Dict[str, int]]:
    state = dict(gm.named_parameters())
    state.update(dict(gm.named_buffers()))
    sizes: Dict[fx.Node, Dict[str, int]] = {}
    for node in gm.graph.nodes:
        if node.op == 'get_attr' and node.target in state:
            for user in node.users:
                sizes.setdefault(user, {}).setdefault(node.target, state[node.target].numel())
        elif node.op == 'call_module':
            sub = gm.get_submodule(node.target)
            own = {node.target + '.' + k: p.numel() for k, p in sub.named_parameters()}
            if own:
                sizes.setdefault(node, {}).update(own)
    return sizes


def choose_split_points(gm: fx.GraphModule, proportions: Sequence[float]) -> List[fx.Node]:
    """Return the first node of stage 1..n-1 (so ``len(result) == len(proportions) - 1``)."""
    n_stages = len(proportions)
    if n_stages <= 1:
        return []
    total = sum(p.numel() for p in gm.parameters()) + sum(b.numel() for b in gm.buffers())
    thresholds = [int(p * total) for p in proportions]
    sizes = _node_param_sizes(gm)
    boundaries: List[fx.Node] = []
    seen: Dict[str, int] = {}
    acc = 0
    ti = 0
    for node in gm.graph.nodes:
        if node not in sizes:
            continue
        thr = thresholds[min(ti, n_stages - 1)]
        new = sum(sz for name, sz in sizes[node].items() if name not in seen)
        repeated = {name: sz for name, sz in sizes[node].items() if name in seen}
        if acc + new <= thr:
            acc += new
            seen.update(sizes[node])
        elif acc == 0 and new > thr:
            nxt = node.next
            if nxt.op != 'output':
                boundaries.append(nxt)
            ti += 1
            seen = {}
        else:
            boundaries.append(node)
            seen = dict(repeated)
            seen.update(sizes[node])
            acc = sum(seen.values())
            ti += 1
        if len(boundaries) >= n_stages - 1:
            break
    boundaries = boundaries[:n_stages - 1]
    if len(boundaries) < n_stages - 1:
        boundaries = _bisect_until(gm, boundaries, n_stages)
    return boundaries


# ------------------------------------------------------------------------------------------------
# compute-balanced planning (B200 grids: every stage owns an identical GPU, so what must be equal is
# the stage's STEP TIME, not its parameter bytes — ResNet-50's first half of the parameters is ~95 % of
# its work).  The reference only knows the parameter-share policy above (operations/utils.py:430-435).
# ------------------------------------------------------------------------------------------------
# Calibrated on the measured single-GPU steps (profiles/RESULTS_r1.md): ResNet-50 batch 100 = 16.6 ms for 0.82 TFLOP forward and
# 7.5 GB of produced activations; BERT-base 8 x 512 = 8.3 ms for 0.24 TFLOP and 2.5 GB.
_PEAK_FLOPS = 6.0e14        # dense bf16 as sustained by the mid-sized GEMM / implicit-GEMM tiles of these models (fwd + bwd = 3x fwd flops)
_HBM_BPS = 6.4e12           # measured copy bandwidth (MEASURED_PEAKS.json)
_ACT_PASSES = 9.0           # HBM passes per produced activation byte over forward + backward (norm / activation / residual traffic)
_LAUNCH_S = 1.0e-6          # per-kernel floor inside a CUDA graph


class _ByteCounter(torch.utils._python_dispatch.TorchDispatchMode):
    """Counts aten calls and the bytes of the tensors they produce (as 2-byte elements: the engine computes in bf16)."""

    def __init__(self):
        super().__init__()
        self.ops = 0
        self.bytes = 0

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        self.ops += 1
        for t in (out if isinstance(out, (tuple, list)) else (out,)):
            if isinstance(t, torch.Tensor) and t.is_floating_point():
                self.bytes += t.numel() * 2
        return out


class _CostProp(fx.Interpreter):
    def __init__(self, gm, scale: float = 1.0):
        super().__init__(gm)
        self.scale = scale

    def run_node(self, n: fx.Node) -> Any:
        from torch.utils.flop_counter import FlopCounterMode
        if n.op not in ('call_module', 'call_function', 'call_method'):
            return super().run_node(n)
        with FlopCounterMode(display=False) as fc, _ByteCounter() as bc:
            out = super().run_node(n)
        flops = float(fc.get_total_flops())
        # forward + backward: 3x the GEMM work, ~3x the elementwise traffic (read x, write y, read dy, write dx ...)
        n.meta['ravnest_cost'] = self.scale * (3.0 * flops / _PEAK_FLOPS + _ACT_PASSES * bc.bytes / _HBM_BPS) + 2.0 * bc.ops * _LAUNCH_S
        n.meta['ravnest_flops'] = flops
        return out


def estimate_costs(gm: fx.GraphModule, inputs: Sequence[Any], scale: float = 1.0) -> Dict[fx.Node, float]:
    """Per-node step-time estimate (seconds, forward + backward on one B200) from one instrumented run on the example inputs.
    ``scale`` = training batch / example batch when the plan is made on a smaller example than the job will run."""
    was_training = gm.training
    gm.eval()
    try:
        with torch.no_grad():
            _CostProp(gm, scale).run(*inputs)
    finally:
        gm.train(was_training)
    return {n: float(n.meta.get('ravnest_cost', 0.0)) for n in gm.graph.nodes}


def measure_costs(gm: fx.GraphModule, inputs: Sequence[Any], device, dtype=torch.bfloat16, batch_scale: int = 1,
                  iters: int = 3) -> Dict[fx.Node, float]:
    """Profile-guided planning: run a copy of the traced model on ``device`` in the compute dtype with the REAL per-replica batch
    (the example tiled ``batch_scale`` times along dim 0) and time every node's forward AND backward on the kernels the job
    will run (CUDA events; wall clock on CPU).  Each node is differentiated on its own — inputs detached, gradient of ones —
    so a node's cost includes its data and weight gradients; nodes that cannot be differentiated in isolation (in-place
    updates of their input) fall back to 3 x forward."""
    import copy
    import time
    dev = torch.device(device)
    cuda = dev.type == 'cuda'
    twin = copy.deepcopy(gm).to(dev)
    for p in list(twin.parameters()) + list(twin.buffers()):
        if p.is_floating_point() and cuda:
            p.data = p.data.to(dtype)
    twin.train()

    def stage(t):
        if not isinstance(t, torch.Tensor):
            return t
        if batch_scale > 1 and t.dim() >= 1:
            t = t.repeat(batch_scale, *([1] * (t.dim() - 1)))
        t = t.to(dev)
        if t.is_floating_point() and cuda:
            t = t.to(dtype)
            if t.dim() == 4:
                t = t.contiguous(memory_format=torch.channels_last)
        return t

    def mark():
        if cuda:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e
        return time.perf_counter()

    def span(a, b):
        return a.elapsed_time(b) * 1e-3 if cuda else b - a

    def leafify(v):
        if isinstance(v, torch.Tensor):
            v = v.detach()
            return v.requires_grad_(True) if v.is_floating_point() else v
        if isinstance(v, (tuple, list)):
            return type(v)(leafify(x) for x in v)
        if isinstance(v, dict):
            return {k: leafify(x) for k, x in v.items()}
        return v

    def detach_tree(v):
        if isinstance(v, torch.Tensor):
            return v.detach()
        if isinstance(v, (tuple, list)):
            return type(v)(detach_tree(x) for x in v)
        if isinstance(v, dict):
            return {k: detach_tree(x) for k, x in v.items()}
        return v

    def grad_roots(v, out):
        if isinstance(v, torch.Tensor):
            if v.is_floating_point() and v.requires_grad:
                out.append(v)
        elif isinstance(v, (tuple, list)):
            for x in v:
                grad_roots(x, out)
        elif isinstance(v, dict):
            for x in v.values():
                grad_roots(x, out)
        return out

    args = [stage(t) for t in inputs]
    samples: Dict[str, List[float]] = {}

    class _Timed(fx.Interpreter):
        def run_node(self, n: fx.Node) -> Any:
            if n.op not in ('call_module', 'call_function', 'call_method'):
                return super().run_node(n)
            a, kw = self.fetch_args_kwargs_from_env(n)
            try:
                la, lkw = leafify(a), leafify(kw)
                with torch.enable_grad():
                    t0 = mark()
                    out = getattr(self, n.op)(n.target, la, lkw)
                    t1 = mark()
                    roots = grad_roots(out, [])
                    seeds = [torch.ones_like(r) for r in roots]
                    t2 = mark()
                    if roots:
                        torch.autograd.backward(roots, seeds)
                    t3 = mark()
                pending.append((n.name, (t0, t1), (t2, t3)))
                return detach_tree(out)
            except RuntimeError:                              # e.g. an in-place update of a (now leaf) input: forward only
                with torch.no_grad():
                    t0 = mark()
                    out = getattr(self, n.op)(n.target, a, kw)
                    t1 = mark()
                pending.append((n.name, (t0, t1), None))
                return out

    for it in range(iters + 2):
        pending: List[Tuple[str, Any, Any]] = []
        if cuda and it >= 2:
            # Give the host a head start: eager launches (~5-10 us each) are slower than most of the kernels they launch, so an event
            # pair around a node would measure the LAUNCH rate of small kernels, not their device time — and the job itself replays
            # CUDA graphs, where launches are free.  A spin kernel keeps the GPU busy while the whole pass is enqueued behind it; the
            # events then bracket back-to-back device execution.  (Without this the late, small-kernel layers of ResNet-50 looked
            # 40 % more expensive than they are and the 2-stage cut left the stages 10.3 vs 7.3 ms busy.)
            torch.cuda._sleep(int(0.35 * 1.9e9))
        with torch.no_grad():
            _Timed(twin).run(*args)
        if cuda:
            torch.cuda.synchronize(dev)
        if it >= 2:                                          # two warm-up passes (descriptor caches, cuDNN heuristics, allocator)
            for name, fwd, bwd in pending:
                f = span(*fwd)
                samples.setdefault(name, []).append(f + span(*bwd) if bwd is not None else 3.0 * f)
        twin.zero_grad(set_to_none=True)
    del twin
    by_name = {k: sorted(v)[len(v) // 2] for k, v in samples.items()}
    return {n: by_name.get(n.name, 0.0) for n in gm.graph.nodes}


def choose_split_points_by_cost(gm: fx.GraphModule, proportions: Sequence[float], costs: Dict[fx.Node, float]) -> List[fx.Node]:
    """First node of stage 1..n-1 such that stage i's estimated step time is ``proportions[i]`` of the total."""
    n_stages = len(proportions)
    if n_stages <= 1:
        return []
    nodes = _compute_nodes(gm)
    if len(nodes) < n_stages:
        raise ValueError('model has too few operations ({}) for {} stages'.format(len(nodes), n_stages))
    total = sum(costs.get(n, 0.0) for n in nodes) or 1.0
    share = sum(proportions)
    targets, acc = [], 0.0
    for p in proportions[:-1]:
        acc += p / share
        targets.append(acc * total)
    prefix = [0.0]
    for n in nodes:
        prefix.append(prefix[-1] + costs.get(n, 0.0))
    cuts: List[int] = []
    for t in targets:
        lo = (cuts[-1] + 1) if cuts else 1
        hi = len(nodes) - (n_stages - 1 - len(cuts))          # leave one node for every later stage
        best = min(range(lo, hi + 1), key=lambda i: abs(prefix[i] - t))
        cuts.append(best)
    return [nodes[c] for c in cuts]


def _compute_nodes(gm: fx.GraphModule) -> List[fx.Node]:
    return [n for n in gm.graph.nodes if n.op in ('call_module', 'call_function', 'call_method')]


def _bisect_until(gm: fx.GraphModule, boundaries: List[fx.Node], n_stages: int) -> List[fx.Node]:
    """Fewer boundaries than stages: repeatedly cut the longest stage in half (by compute nodes)."""
    nodes = _compute_nodes(gm)
    index = {n: i for i, n in enumerate(nodes)}
    cuts = sorted({index[b] for b in boundaries if b in index})
    while len(cuts) < n_stages - 1:
        edges = [0] + cuts + [len(nodes)]
        spans = [(edges[i + 1] - edges[i], i) for i in range(len(edges) - 1)]
        length, which = max(spans)
        if length < 2:
            raise ValueError('model has too few operations ({}) for {} stages'.format(len(nodes), n_stages))
        cuts.append(edges[which] + length // 2)
        cuts = sorted(set(cuts))
    return [nodes[c] for c in cuts]


class StagePlan:
    """Everything the runtime needs to know about one stage."""

    def __init__(self, index: int, name: str):
        self.index = index
        self.name = name                      # 'submod_<k>'
        self.module: Optional[fx.GraphModule] = None
        self.inputs: List[Dict[str, Any]] = []    # [{'producer','out_idx','name','meta'}] in arg order
        self.outputs: List[Dict[str, Any]] = []   # [{'name','targets':[stage names],'meta'}]
        self.is_last = False

    def input_template(self) -> Dict[int, Dict[str, Any]]:
        tmpl = {}
        for pos, src in enumerate(self.inputs):
            if src['producer'] == 'model_inputs':
                tmpl[pos] = {'model_inputs': src['out_idx']}
            elif src['single']:
                tmpl[pos] = {src['producer']: 'placeholder:tensor'}
            else:
                tmpl[pos] = {src['producer']: src['out_idx']}
        return tmpl

    def output_template(self) -> Dict[int, Dict[str, Any]]:
        return {i: {'target': list(o['targets'])} for i, o in enumerate(self.outputs) if o['targets']}

    def describe(self) -> Dict[str, Any]:
        return {'index': self.index, 'name': self.name, 'is_last': self.is_last,
                'inputs': [{k: v for k, v in s.items()} for s in self.inputs],
                'outputs': [{k: v for k, v in o.items()} for o in self.outputs]}


class SplitResult:
    def __init__(self, stages: List[StagePlan], model_input_names: List[str],
                 model_inputs_template: Dict[int, Dict[str, Any]], model_input_meta: List[Any]):
        self.stages = stages
        self.model_input_names = model_input_names
        self.model_inputs_template = model_inputs_template
        self.model_input_meta = model_input_meta
        self.stage_costs: Optional[List[float]] = None       # balance='compute': estimated seconds per stage step

    def run_local(self, *inputs):
        """Chain the stages in-process (used for the planner's self-check and tests)."""
        produced: Dict[Tuple[str, int], Any] = {('model_inputs', i): v for i, v in enumerate(inputs)}
        out = None
        for st in self.stages:
            args = [produced[(s['producer'], s['out_idx'])] for s in st.inputs]
            out = st.module(*args)
            if st.is_last:
                return out
            vals = (out,) if len(st.outputs) == 1 else tuple(out)
            for i, v in enumerate(vals):
                produced[(st.name, i)] = v
        return out


def split_graph(gm: fx.GraphModule, boundaries: Sequence[fx.Node], input_names: Sequence[str]) -> SplitResult:
    """Carve ``gm`` into ``len(boundaries) + 1`` stage GraphModules."""
    n_stages = len(boundaries) + 1
    boundary_set = {b: i + 1 for i, b in enumerate(boundaries)}
    stage_of: Dict[fx.Node, int] = {}
    current = 0
    output_node = None
    for node in gm.graph.nodes:
        if node in boundary_set:
            current = max(current, boundary_set[node])
        if node.op == 'placeholder':
            continue
        if node.op == 'output':
            output_node = node
            stage_of[node] = n_stages - 1
            continue
        stage_of[node] = current
        if node.op == 'call_function' and node.target is operator.getitem and isinstance(node.args[0], fx.Node) \
                and node.args[0] in stage_of:
            stage_of[node] = stage_of[node.args[0]]      # a tuple never crosses a cut: its elements are unpacked where it is made
    # parameters / constants referenced with get_attr live with their first consumer
    for node in gm.graph.nodes:
        if node.op == 'get_attr' and node.users:
            stage_of[node] = min(stage_of[u] for u in node.users)
    placeholders = [n for n in gm.graph.nodes if n.op == 'placeholder']
    ph_index = {n: i for i, n in enumerate(placeholders)}

    stages = [StagePlan(k, 'submod_{}'.format(k)) for k in range(n_stages)]
    stages[-1].is_last = True
    # which values leave each stage (in first-use order) ---------------------------------
    exported: List[List[fx.Node]] = [[] for _ in range(n_stages)]
    for node in gm.graph.nodes:
        if node.op in ('placeholder',):
            continue
        k = stage_of[node]
        for src in node.all_input_nodes:
            if src.op == 'placeholder':
                continue
            sk = stage_of[src]
            if sk != k and src not in exported[sk]:
                if sk > k:
                    raise RuntimeError('split produced a backward edge {} -> {}'.format(src, node))
                exported[sk].append(src)

    model_inputs_template: Dict[int, Dict[str, Any]] = {i: {'target': []} for i in range(len(placeholders))}

    for k, st in enumerate(stages):
        g = fx.Graph()
        env: Dict[fx.Node, fx.Node] = {}
        ext_inputs: List[fx.Node] = []
        members = [n for n in gm.graph.nodes if n.op not in ('placeholder',) and stage_of.get(n) == k]
        for node in members:
            for src in node.all_input_nodes:
                if (src.op == 'placeholder' or stage_of[src] != k) and src not in ext_inputs:
                    ext_inputs.append(src)
        for src in ext_inputs:
            env[src] = g.placeholder(src.name)
            if src.op == 'placeholder':
                idx = ph_index[src]
                st.inputs.append({'producer': 'model_inputs', 'out_idx': idx, 'single': False,
                                  'name': src.name, 'meta': src.meta.get('ravnest_val')})
                if st.name not in model_inputs_template[idx]['target']:
                    model_inputs_template[idx]['target'].append(st.name)
            else:
                sk = stage_of[src]
                st.inputs.append({'producer': stages[sk].name, 'out_idx': exported[sk].index(src),
                                  'single': len(exported[sk]) == 1, 'name': src.name,
                                  'meta': src.meta.get('ravnest_val')})
        for node in members:
            if node.op == 'output':
                continue
            env[node] = g.node_copy(node, lambda n: env[n])
        if st.is_last:
            g.output(fx.node.map_arg(output_node.args[0], lambda n: env[n]))
            st.outputs = [{'name': 'output', 'targets': [], 'meta': None}]
        else:
            outs = exported[k]
            if not outs:
                raise RuntimeError('stage {} exports nothing; choose different proportions'.format(k))
            g.output(env[outs[0]] if len(outs) == 1 else tuple(env[o] for o in outs))
            for o in outs:
                targets = []
                for user in o.users:
                    name = stages[stage_of[user]].name
                    if stage_of[user] != k and name not in targets:
                        targets.append(name)
                st.outputs.append({'name': o.name, 'targets': sorted(targets, key=lambda s: int(s.split('_')[-1])),
                                   'meta': o.meta.get('ravnest_val')})
        g.lint()
        st.module = fx.GraphModule(gm, g, 'Submod{}'.format(k))
    input_meta = [p.meta.get('ravnest_val') for p in placeholders]
    return SplitResult(stages, list(input_names), model_inputs_template, input_meta)


def nudge_proportions(proportions: List[float]) -> List[float]:
    """Reference fallback when a split does not run (operations/utils.py:256-263): move 0.1 from the
    largest share to its right neighbour (or to the first share if the largest is last)."""
    props = list(proportions)
    big = props.index(max(props))
    dst = 0 if big == len(props) - 1 else big + 1
    props[dst] = round(props[dst] + 0.1, 6)
    props[big] = round(props[big] - 0.1, 6)
    return props


def split_model(model: nn.Module, proportions: Sequence[float], example_args: Sequence = (),
                example_kwargs: Optional[Dict] = None, check: bool = True, balance: str = 'params',
                cost_scale: float = 1.0, device=None, dtype=torch.bfloat16) -> SplitResult:
    """Trace + split policy + carve.  ``balance='params'`` is the reference's parameter-share threshold policy;
    ``balance='compute'`` makes ``proportions`` shares of the estimated step time (``estimate_costs``); ``balance='measured'``
    times every node's forward + backward on ``device`` with the native kernels at the real batch (``measure_costs``; falls
    back to the estimate when no GPU is visible and ``device`` is not explicitly the CPU).  ``check`` runs the
    chained stages once and compares against the traced module on the example inputs (the reference only checks "does not raise")."""
    assert balance in ('params', 'compute', 'measured')
    example_kwargs = dict(example_kwargs or {})
    gm, input_names = trace_model(model, example_args, example_kwargs)
    flat_inputs = list(example_args) + [example_kwargs[k] for k in input_names[len(example_args):]]
    propagate_meta(gm, flat_inputs)
    costs = None
    if len(proportions) > 1 and balance in ('compute', 'measured'):          # a single stage needs no balancing
        if balance == 'measured' and (torch.cuda.is_available() or (device is not None and torch.device(device).type == 'cpu')):
            try:
                costs = measure_costs(gm, flat_inputs, device or torch.device('cuda', torch.cuda.current_device()), dtype,
                                      batch_scale=max(1, int(round(cost_scale))))
            except Exception as err:      # planning must not take the job down: fall back to the analytic estimate
                import warnings
                warnings.warn('measured stage balancing failed ({!r}); using the estimated costs'.format(err))
                if torch.cuda.is_available():
                    torch.cuda.synchronize()
        if costs is None:
            costs = estimate_costs(gm, flat_inputs, cost_scale)
    attempt = list(proportions)
    last_err = None
    for _ in range(2):
        try:
            points = choose_split_points(gm, attempt) if costs is None else choose_split_points_by_cost(gm, attempt, costs)
            result = split_graph(gm, points, input_names)
            if costs is not None:       # estimated step time per stage (seconds), for the plan file / logs
                order = _compute_nodes(gm)
                edges = [0] + [order.index(b) for b in points] + [len(order)]
                result.stage_costs = [sum(costs.get(n, 0.0) for n in order[a:b]) for a, b in zip(edges[:-1], edges[1:])]
            if check:
                was = gm.training
                gm.eval()
                for st in result.stages:
                    st.module.eval()
                with torch.no_grad():
                    result.run_local(*flat_inputs)
                gm.train(was)
                for st in result.stages:
                    st.module.train(was)
            return result
        except Exception as err:  # mirror the reference's single retry with nudged proportions
            last_err = err
            attempt = nudge_proportions(attempt)
    raise last_err
