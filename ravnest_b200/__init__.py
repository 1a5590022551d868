"""ravnest_b200 — a Blackwell-native asynchronous pipeline+data-parallel training engine with the
capabilities and user API of ravenprotocol/ravnest (``Node`` / ``Trainer`` / ``clusterize`` /
``model_fusion`` / ``set_seed``).  Everything is exported from the package root like the reference
(``ravnest/__init__.py:1-8``)."""
from .strings import NodeTypes, ActionTypes, BufferStatus, NodeStatus
from .utils import set_seed, model_fusion, load_node_json_configs, reset_node_data, check_gpu_usage
from .operations import clusterize, split_model, get_memory_reqs, default_gpu_configs
from .compute import Compute
from .communication import Communication
from .endpoints import GrpcService
from .node import Node
from .trainer import Trainer
from .optim import Lamb

__version__ = '0.1.0'
__all__ = ['NodeTypes', 'ActionTypes', 'BufferStatus', 'NodeStatus', 'set_seed', 'model_fusion',
           'load_node_json_configs', 'reset_node_data', 'check_gpu_usage', 'clusterize', 'split_model',
           'get_memory_reqs', 'default_gpu_configs', 'Compute', 'Communication', 'GrpcService', 'Node', 'Trainer', 'Lamb']
