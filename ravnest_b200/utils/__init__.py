"""Runtime utilities (reference ``ravnest/utils.py``)."""
from .serialization import save_submodel, load_submodel, dump_pickle, load_pickle, dump_json, load_json
from .config import load_node_json_configs
from .misc import (set_seed, model_fusion, find_files_with_extension, reset_node_data, check_gpu_usage, compare_models,
                   ClusterPipeline)
from .streams import generate_stream, generate_weights_stream, generate_data_stream
from .params import (current_model_params_clone, optimizer_params, load_grads_into_optimizer, load_optim_weights_into_model,
                     load_model_weights_into_optim, get_trainable_param_names, get_trainable_parameters,
                     load_state_dict_conserve_versions, load_optim_state, create_chunks, create_chunks_optim,
                     compress_tensor_float16, extract_tensor_from_compression_float16)

__all__ = ['save_submodel', 'load_submodel', 'dump_pickle', 'load_pickle', 'dump_json', 'load_json',
           'load_node_json_configs', 'set_seed', 'model_fusion', 'find_files_with_extension', 'reset_node_data',
           'check_gpu_usage', 'compare_models', 'ClusterPipeline', 'generate_stream', 'generate_weights_stream', 'generate_data_stream',
           'current_model_params_clone', 'optimizer_params', 'load_grads_into_optimizer', 'load_optim_weights_into_model',
           'load_model_weights_into_optim', 'get_trainable_param_names', 'get_trainable_parameters',
           'load_state_dict_conserve_versions', 'load_optim_state', 'create_chunks', 'create_chunks_optim',
           'compress_tensor_float16', 'extract_tensor_from_compression_float16']
