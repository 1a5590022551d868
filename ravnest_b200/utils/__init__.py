"""Runtime utilities (reference ``ravnest/utils.py``)."""
from .serialization import save_submodel, load_submodel, dump_pickle, load_pickle, dump_json, load_json
from .config import load_node_json_configs
from .misc import set_seed, model_fusion, find_files_with_extension, reset_node_data, check_gpu_usage

__all__ = ['save_submodel', 'load_submodel', 'dump_pickle', 'load_pickle', 'dump_json', 'load_json',
           'load_node_json_configs', 'set_seed', 'model_fusion', 'find_files_with_extension', 'reset_node_data',
           'check_gpu_usage']
