"""ravnest_b200 — a Blackwell-native asynchronous pipeline+data-parallel training engine with the
capabilities and user API of ravenprotocol/ravnest (``Node`` / ``Trainer`` / ``clusterize`` /
``model_fusion`` / ``set_seed``).  Like the reference (``ravnest/__init__.py:1-8`` star-imports its node, trainer, utils and
operations modules) every public helper is reachable from the package root."""
from .strings import NodeTypes, ActionTypes, BufferStatus, NodeStatus
from . import utils as _utils
from . import operations as _operations
from .utils import *            # noqa: F401,F403  set_seed, model_fusion, load_node_json_configs, parameter / chunk helpers ...
from .operations import *       # noqa: F401,F403  clusterize, split_model_equal, spawn_node_pool, cluster_formation ...
from .compute import Compute
from .communication import Communication, compress_tensor, decompress_tensor
from .endpoints import GrpcService
from .node import Node
from .trainer import Trainer
from .optim import Lamb

__version__ = '0.1.0'
__all__ = (['NodeTypes', 'ActionTypes', 'BufferStatus', 'NodeStatus', 'Compute', 'Communication', 'compress_tensor',
            'decompress_tensor', 'GrpcService', 'Node', 'Trainer', 'Lamb'] + list(_utils.__all__) + list(_operations.__all__))
