"""Parameter / optimizer plumbing helpers with the reference's names (``ravnest/utils.py:84-194``).

Inside the engine these jobs are done by ``runtime.param_store.ParamStore`` (flat master / slot / gradient buffers: "clone the
parameters for the optimizer", "load grads into the optimizer", "copy the optimizer's weights back into the model" are pointer
identities there, not per-tensor loops).  User scripts written against the reference import the helpers directly
(``from ravnest.utils import ...``), so they are provided here on plain modules / optimizers with the same call signatures.
"""
from typing import Dict, Iterator, List

import torch

FP16_MAX = float(torch.finfo(torch.float16).max)
FP32_MAX = float(torch.finfo(torch.float32).max)


@torch.no_grad()
def current_model_params_clone(model: torch.nn.Module) -> Iterator[torch.Tensor]:
    """Detached copies of every parameter, in ``model.parameters()`` order (what root / stem optimizers are built over)."""
    return (p.detach().clone() for p in model.parameters())


def optimizer_params(optimizer: torch.optim.Optimizer) -> Iterator[torch.Tensor]:
    """Every tensor the optimizer updates, group after group."""
    for group in optimizer.param_groups:
        yield from group['params']


@torch.no_grad()
def load_grads_into_optimizer(model: torch.nn.Module, optimizer: torch.optim.Optimizer) -> None:
    """Hand the model's gradients to the optimizer's own (cloned) parameters — by reference, no copy."""
    for src, dst in zip(model.parameters(), optimizer_params(optimizer)):
        dst.grad = src.grad


@torch.no_grad()
def load_optim_weights_into_model(model: torch.nn.Module, optimizer: torch.optim.Optimizer) -> None:
    """Point the model's parameters at the optimizer's freshly updated values."""
    for dst, src in zip(model.parameters(), optimizer_params(optimizer)):
        dst.data = src.data


@torch.no_grad()
def load_model_weights_into_optim(model: torch.nn.Module, optimizer: torch.optim.Optimizer) -> None:
    """The reverse direction (after parameter averaging wrote into the model)."""
    for src, dst in zip(model.parameters(), optimizer_params(optimizer)):
        dst.data = src.data


def get_trainable_param_names(model: torch.nn.Module) -> List[str]:
    return [name for name, p in model.named_parameters() if p.requires_grad]


def get_trainable_parameters(model: torch.nn.Module) -> Iterator[torch.nn.Parameter]:
    """``operations/misc.py:4`` of the reference."""
    return (p for p in model.parameters() if p.requires_grad)


@torch.no_grad()
def load_state_dict_conserve_versions(model: torch.nn.Module, state_dict: Dict[str, torch.Tensor]) -> None:
    """Swap parameter storage without touching autograd version counters (weight-stash restore around a recompute)."""
    for name, p in model.named_parameters():
        p.data = state_dict[name].data


@torch.no_grad()
def load_optim_state(optimizer: torch.optim.Optimizer, state: Dict[str, Dict], model: torch.nn.Module) -> None:
    """Install per-parameter optimizer state keyed by parameter NAME (the wire format of optimizer-state averaging)."""
    packed = optimizer.state_dict()
    packed['state'] = {i: state[name] for i, (name, _) in enumerate(model.named_parameters()) if name in state}
    optimizer.load_state_dict(packed)


def _split(t: torch.Tensor, parts: int):
    axis = max(range(t.dim()), key=lambda d: t.shape[d]) if t.dim() else 0
    return list(torch.tensor_split(t.detach().to('cpu'), parts, axis)), axis


def create_chunks(data: Dict[str, torch.Tensor], size: int) -> Dict[str, Dict]:
    """Per-parameter ring chunks: every tensor cut into ``size`` pieces along its longest axis."""
    out = {}
    for key, t in data.items():
        pieces, axis = _split(t, size)
        out[key] = {'data': pieces, 'split_axis': axis}
    return out


def create_chunks_optim(data: Dict[str, Dict[str, torch.Tensor]], size: int) -> Dict[str, Dict]:
    """The same for optimizer state (``{param: {exp_avg: ..., step: ...}}``); 0-d entries are sent as 1-element vectors."""
    out = {}
    for key, entry in data.items():
        packed = {}
        for name, t in entry.items():
            scalar = t.dim() == 0
            pieces, axis = _split(t.reshape(1) if scalar else t, size)
            packed[name] = {'data': pieces, 'split_axis': axis, 'reshape': scalar}
        out[key] = packed
    return out


def compress_tensor_float16(tensor: torch.Tensor) -> torch.Tensor:
    """fp32 -> fp16 (clamped), fp64 -> fp32: the reference's activation "compression"."""
    from ..communication import compress_tensor
    if tensor.dtype == torch.float64:
        return tensor.clamp(-FP32_MAX, FP32_MAX).to(torch.float32)
    return compress_tensor(tensor)


def extract_tensor_from_compression_float16(tensor: torch.Tensor, original_dtype: torch.dtype) -> torch.Tensor:
    return tensor.to(original_dtype)
