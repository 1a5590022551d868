"""Names the reference pulls in through ``from pippy.ModelSplit import *`` (``pippy_utils.py:2``): it uses
``logging``, ``torch``, ``pipe_split`` and the typing aliases from that star import."""
import logging  # noqa: F401
from typing import Callable, Dict, List, Tuple  # noqa: F401

import torch  # noqa: F401
import torch.fx as fx  # noqa: F401

from .IR import pipe_split  # noqa: F401


def split_into_equal_size(nstages: int = 1):
    """Equal parameter-size split policy (imported, never called, by the reference)."""
    def policy(gm: fx.GraphModule) -> fx.GraphModule:
        total = sum(p.numel() for p in gm.parameters()) + sum(b.numel() for b in gm.buffers())
        per_stage = total // max(1, nstages)
        seen, done = 0, 1
        for node in list(gm.graph.nodes):
            if node.op == 'call_module':
                size = sum(p.numel() for p in gm.get_submodule(node.target).parameters())
                if seen + size > per_stage and seen > 0 and done < nstages:
                    with gm.graph.inserting_before(node):
                        gm.graph.call_function(pipe_split, (), {})
                    done += 1
                    seen = 0
                seen += size
        gm.recompile()
        return gm
    return policy
