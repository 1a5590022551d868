"""Stage artefact (``submod.pt``) and template I/O.

The reference stores stages as TorchScript (``torch.jit.script(...).save``,
operations/utils.py:345-349) which is deprecated in torch 2.11.  We keep the
file name and folder layout but store a pickled
``torch.fx.GraphModule`` whose state-dict keys are the ORIGINAL parameter names.
"""
import json
import os
import pickle
from typing import Any

import torch


def save_submodel(module: torch.nn.Module, path: str) -> None:
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    tmp = path + '.tmp'
    torch.save(module, tmp)
    os.replace(tmp, path)


def load_submodel(path: str, map_location: Any = 'cpu') -> torch.nn.Module:
    return torch.load(path, map_location=map_location, weights_only=False)


def dump_pickle(obj: Any, path: str) -> None:
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    with open(path, 'wb') as fh:
        pickle.dump(obj, fh)


def load_pickle(path: str) -> Any:
    with open(path, 'rb') as fh:
        return pickle.load(fh)


def dump_json(obj: Any, path: str) -> None:
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    tmp = path + '.tmp'
    with open(tmp, 'w') as fh:
        json.dump(obj, fh, indent=1)
    os.replace(tmp, path)


def load_json(path: str) -> Any:
    with open(path) as fh:
        return json.load(fh)
