"""``Pipe`` — trace with dynamo's torch-IR export, split at ``pipe_split`` markers, own every parameter in the
stage that uses it.  Mirrors what torchpippy 0.2.0 produced on torch 2.2 closely enough for the reference's planner
and runtime: ``call_module`` nodes named ``L__self___<path>`` (so state-dict keys carry the ``L__self___`` prefix
the reference strips in ``model_fusion``, ``ravnest/utils.py:247``), keyword inputs become placeholders
``l_<name>_`` (the reference re-keys ``forward_compute(**kwargs)`` that way, ``ravnest/node.py:437-442``), and
``split_gm`` has children ``submod_0 … submod_{n-1}`` connected by plain tensors / ``getitem``s.
"""
import copy
import inspect
import operator
import re

import torch
import torch.fx as fx
from torch.fx.passes.split_module import split_module


def pipe_split():
    """Marker: a new pipeline stage starts here."""
    return None


fx.wrap('pipe_split')


def _rename_placeholders(gm: fx.GraphModule, mod, n_args):
    """dynamo (torch 2.11) names inputs ``l_args_0_`` / ``l_kwargs_<k>_``; torch 2.2 named them after the forward
    parameter (``l_<k>_``) — the reference's root relies on the latter."""
    params = [p for p in inspect.signature(mod.forward).parameters]
    for node in gm.graph.nodes:
        if node.op != 'placeholder':
            continue
        m = re.fullmatch(r'l_kwargs_(.+)_', node.name)
        if m:
            node.name = 'l_{}_'.format(m.group(1))
            continue
        m = re.fullmatch(r'l_args_(\d+)_', node.name)
        if m and int(m.group(1)) < len(params):
            node.name = 'l_{}_'.format(params[int(m.group(1))])


class Pipe(torch.nn.Module):
    def __init__(self, split_gm: fx.GraphModule, num_stages: int):
        super().__init__()
        self.split_gm = split_gm
        self.num_stages = num_stages

    def forward(self, *args, **kwargs):
        flat = list(args) + list(kwargs.values())
        return self.split_gm(*flat)

    # ------------------------------------------------------------------ tracing
    @staticmethod
    def _trace_with_export(mod, example_args=(), example_kwargs=None, constraints=None) -> fx.GraphModule:
        example_kwargs = example_kwargs or {}
        import torch._dynamo
        torch._dynamo.reset()
        with torch._dynamo.config.patch(inline_inbuilt_nn_modules=False):
            exported = torch._dynamo.export(mod, aten_graph=False, assume_static_by_default=True)(*example_args, **example_kwargs)
        gm = exported.graph_module
        # plain positional calling convention (drop the pytree in/out spec dynamo installs)
        gm.graph._codegen = fx.graph.CodeGen()
        for node in list(gm.graph.nodes):
            if node.op == 'output':
                outs = node.args[0]
                if isinstance(outs, (list, tuple)) and len(outs) == 1:
                    node.args = (outs[0],)
                elif isinstance(outs, list):
                    node.args = (tuple(outs),)
        _rename_placeholders(gm, mod, len(example_args))
        gm.graph.eliminate_dead_code()
        gm.recompile()
        return gm

    # ------------------------------------------------------------------ splitting
    @staticmethod
    def _from_traced(mod, traced: fx.GraphModule, multi_use_param_spec=None, output_loss_value_spec=None,
                     split_policy=None) -> 'Pipe':
        if split_policy is not None:
            traced = split_policy(traced)
        part = 0

        def stage_of(node):
            nonlocal part
            if node.op == 'call_function' and node.target is pipe_split:
                part += 1
            return part

        split = split_module(traced, mod, stage_of, keep_original_order=True)
        # pipe_split markers carry no data
        for sub in split.children():
            if isinstance(sub, fx.GraphModule):
                for node in list(sub.graph.nodes):
                    if node.op == 'call_function' and node.target is pipe_split:
                        sub.graph.erase_node(node)
                sub.recompile()
        Pipe._lift_attributes(split)
        split.graph.eliminate_dead_code()
        split.recompile()
        return Pipe(split, part + 1)

    @staticmethod
    def _lift_attributes(split: fx.GraphModule):
        """``split_module`` leaves ``get_attr`` (parameters / buffers used directly) in the root graph and passes
        them as stage arguments; move each into the stage(s) that use it (a tensor used by several stages is
        replicated — every provider owns what it needs, nothing but activations crosses a stage boundary)."""
        for node in list(split.graph.nodes):
            if node.op != 'get_attr':
                continue
            value = split
            for atom in node.target.split('.'):
                value = getattr(value, atom)
            if not isinstance(value, torch.Tensor):
                continue
            is_buffer = not isinstance(value, torch.nn.Parameter)
            new_name = 'moved_' + node.target.replace('.', '_')
            for user in list(node.users):
                if user.op != 'call_module':
                    continue
                callee = split.get_submodule(user.target)
                idxs = [i for i, a in enumerate(user.args) if a is node]
                if not hasattr(callee, new_name):
                    held = value if len(node.users) == 1 else copy.deepcopy(value)
                    if is_buffer:
                        callee.register_buffer(new_name, held)
                    else:
                        setattr(callee, new_name, held)
                phs = [n for n in callee.graph.nodes if n.op == 'placeholder']
                for i in idxs:
                    with callee.graph.inserting_before(phs[i]):
                        attr = callee.graph.get_attr(new_name)
                    phs[i].replace_all_uses_with(attr)
                    callee.graph.erase_node(phs[i])
                callee.recompile()
                user.args = tuple(a for i, a in enumerate(user.args) if i not in idxs)
            if not node.users:
                split.graph.erase_node(node)


def annotate_split_points(mod, spec):      # pragma: no cover - API placeholder, unused by the reference
    raise NotImplementedError('annotate_split_points is not part of the shim')
