"""Data-parallel machinery over NVLink peer memory: symmetric buffers + fused all-reduce/optimizer."""
from .symmetric import SymmetricBuffer, symmetric_available

__all__ = ['SymmetricBuffer', 'symmetric_available']
