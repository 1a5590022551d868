"""Locate and import the in-tree CUDA extension."""
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, 'rvn_C.so')
_mod = None


def load():
    global _mod
    if _mod is None:
        if not os.path.exists(_PATH):
            raise RuntimeError('ravnest_b200 CUDA extension not built: run `python -m ravnest_b200.ops.build` '
                               '(or __graft_entry__.build()).  Refusing to fall back silently on a GPU box.')
        import torch  # noqa: F401  (libtorch symbols must be loaded first)
        spec = importlib.util.spec_from_file_location('rvn_C', _PATH)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def available() -> bool:
    return os.path.exists(_PATH)
