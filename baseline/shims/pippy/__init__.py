"""Stand-in for ``torchpippy==0.2.0`` (pinned by the reference, not installable offline, incompatible with torch 2.11).

Only the surface the reference's planner touches is provided (reference ``ravnest/operations/utils.py:7-9,:243-248``,
``pippy_utils.py:2``): ``pippy.IR.Pipe._trace_with_export`` / ``Pipe._from_traced`` / ``pipe.split_gm`` /
``pipe.forward``, ``pippy.IR.pipe_split``, ``pippy.split_into_equal_size`` and the names ``pippy.ModelSplit`` star-exports.
Implemented on stock torch (``torch._dynamo.export`` + ``torch.fx.passes.split_module``) — no ravnest_b200 code.
"""
from .ModelSplit import split_into_equal_size  # noqa: F401
from .IR import Pipe, pipe_split  # noqa: F401

__version__ = '0.2.0+rvn.shim'
