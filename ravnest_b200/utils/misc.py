"""Seeding, fusion, housekeeping helpers (reference ``ravnest/utils.py:196-255``, ``reset.py``)."""
import os
import random
import shutil

import numpy as np
import torch

from .serialization import load_submodel


def set_seed(seed=42):
    """Seed python / numpy / torch (CPU + every CUDA device), reset the Philox state of the native dropout kernels and ask for
    deterministic cuDNN (reference utils.py:196-209)."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    os.environ['PYTHONHASHSEED'] = str(seed)
    try:                                   # the dropout kernels' Philox {seed, step} is re-derived from the new torch seed
        from ..ops import native
        for state in native._RNG.values():         # in place: captured CUDA graphs keep pointing at these tensors
            state.copy_(torch.tensor([int(seed) & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64))
    except Exception:
        pass


def find_files_with_extension(root_folder, extension):
    hits = []
    for folder, _dirs, files in os.walk(root_folder):
        for f in sorted(files):
            if f.endswith('.' + extension.lstrip('.')):
                hits.append(os.path.join(folder, f))
    return sorted(hits)


def model_fusion(cluster_id=0, base_dir=None):
    """Merge the saved stage modules of one cluster into ``trained/trained_state_dict.pt``
    (reference utils.py:232-255).  Stage state dicts already use the original parameter names; a
    legacy ``L__self___`` prefix is stripped if present."""
    base_dir = base_dir or os.environ.get('RAVNEST_HOME', '.')
    folder = os.path.join(base_dir, 'node_data', 'cluster_{}'.format(cluster_id))
    if not os.path.isdir(folder):
        print('Submodels not found!')
        return None
    combined = {}
    for path in find_files_with_extension(folder, 'pt'):
        if os.path.basename(path) == 'submod.pt' and any(
                f.startswith('submod_') and f.endswith('.pt') for f in os.listdir(os.path.dirname(path))):
            continue                      # prefer the trained submod_<k>.pt over the initial stage
        state = load_submodel(path).state_dict()
        combined.update({k.replace('L__self___', ''): v.detach().to('cpu') for k, v in state.items()})
    if not combined:
        print('{} path has no submodels'.format(folder))
        return None
    out_dir = os.path.join(base_dir, 'trained')
    os.makedirs(out_dir, exist_ok=True)
    save_path = os.path.join(out_dir, 'trained_state_dict.pt')
    torch.save(combined, save_path)
    return save_path


def reset_node_data(base_dir=None):
    """``reset.py`` of the reference: drop every sub-folder of node_data/ (keeps node_configs.json)."""
    base_dir = base_dir or os.environ.get('RAVNEST_HOME', '.')
    root = os.path.join(base_dir, 'node_data')
    if os.path.isdir(root):
        for entry in os.listdir(root):
            full = os.path.join(root, entry)
            if os.path.isdir(full):
                shutil.rmtree(full)


def check_gpu_usage(device=None):
    """Memory in use on ``device`` in MB.  The reference calls nvmlInit/Shutdown around every
    forward/backward (utils.py:211-221); this reads the caching allocator's counter instead."""
    if not torch.cuda.is_available():
        return 0.0
    return torch.cuda.memory_allocated(device) / 1e6


def compare_models(a, b, atol: float = 0.0, rtol: float = 0.0):
    """Debug helper (reference compute.py:542-555): names of entries that differ between two modules / state dicts, with the
    largest absolute difference of each.  Returns ``[]`` when they match."""
    sa = a.state_dict() if hasattr(a, 'state_dict') else dict(a)
    sb = b.state_dict() if hasattr(b, 'state_dict') else dict(b)
    diffs = []
    for key in sorted(set(sa) | set(sb)):
        if key not in sa or key not in sb:
            diffs.append((key, None))
            continue
        x, y = sa[key].detach().to('cpu'), sb[key].detach().to('cpu')
        if x.shape != y.shape:
            diffs.append((key, None))
        elif not torch.allclose(x.double(), y.double(), atol=atol, rtol=rtol):
            diffs.append((key, float((x.double() - y.double()).abs().max())))
    return diffs


class ClusterPipeline(torch.nn.Module):
    """The saved stages of one cluster chained in-process (no nodes, no transport): what the reference's
    ``examples/sorter/sorter_inference.py`` does with its TorchScript sub-models.  Stages are loaded from
    ``node_data/cluster_<c>/<host:port>/submod_<k>.pt`` (``submod.pt`` before training was saved) and wired with each stage's
    ``plan.json``; call it with the model's inputs in signature order."""

    def __init__(self, cluster_id: int = 0, base_dir=None, trained: bool = True):
        super().__init__()
        from .serialization import load_json, load_submodel
        root = os.path.join(base_dir or os.environ.get('RAVNEST_HOME', '.'), 'node_data', 'cluster_{}'.format(cluster_id))
        plans = []
        for folder in sorted(os.listdir(root)):
            plan_path = os.path.join(root, folder, 'plan.json')
            if os.path.exists(plan_path):
                plans.append((load_json(plan_path), os.path.join(root, folder)))
        if not plans:
            raise FileNotFoundError('no stage plans under {}'.format(root))
        plans.sort(key=lambda item: item[0]['index'])
        self.plans = [p for p, _ in plans]
        modules = []
        for plan, folder in plans:
            saved = os.path.join(folder, '{}.pt'.format(plan['name']))
            modules.append(load_submodel(saved if trained and os.path.exists(saved) else os.path.join(folder, 'submod.pt')))
        self.stages = torch.nn.ModuleList(modules)

    def forward(self, *model_inputs):
        produced = {('model_inputs', i): v for i, v in enumerate(model_inputs)}
        out = None
        for plan, stage in zip(self.plans, self.stages):
            out = stage(*[produced[(src['producer'], src['out_idx'])] for src in plan['inputs']])
            if plan['is_last']:
                return out
            values = (out,) if len(plan['outputs']) == 1 else tuple(out)
            for i, v in enumerate(values):
                produced[(plan['name'], i)] = v
        return out
