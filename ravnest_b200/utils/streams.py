"""Chunked gRPC stream builders (2 MiB blocks) and their inverse.

Parity: reference ``ravnest/utils.py:31-83`` (``generate_stream`` / ``generate_weights_stream`` /
``generate_data_stream``).  The size fields are ``int32`` on the wire (tensor.proto:16), so a payload
larger than 2 GiB is refused here with a clear error instead of silently overflowing.
"""
import pickle
from typing import Any, Iterable, Iterator

from ..protos import pb

CHUNK_BYTES = 2 * 1024 * 1024
_INT32_MAX = 2 ** 31 - 1


def _blocks(blob: bytes) -> Iterator[bytes]:
    if len(blob) > _INT32_MAX:
        raise ValueError('payload of {} bytes exceeds the int32 size field of the wire schema; '
                         'use the NCCL transport for tensors this large'.format(len(blob)))
    view = memoryview(blob)
    for start in range(0, max(len(blob), 1), CHUNK_BYTES):
        yield bytes(view[start:start + CHUNK_BYTES])


def generate_stream(data: Any, type: str = None) -> Iterator:
    """Payload dict -> stream of ``SendTensor`` (for ``send_buffer``)."""
    blob = pickle.dumps(data, protocol=pickle.HIGHEST_PROTOCOL)
    size = len(blob)
    for block in _blocks(blob):
        yield pb.SendTensor(tensor_chunk=pb.TensorChunk(buffer=block, type='$tensor', tensor_size=size), type=type or '')


def generate_weights_stream(data: Any) -> Iterator:
    """State dict -> stream of ``WeightsChunk`` (for ``get_latest_weights``)."""
    blob = pickle.dumps(data, protocol=pickle.HIGHEST_PROTOCOL)
    size = len(blob)
    for block in _blocks(blob):
        yield pb.WeightsChunk(tensor_chunk=pb.TensorChunk(buffer=block, type='$tensor', tensor_size=size))


def generate_data_stream(data: Any, ring_id: int = 0, type: str = None) -> Iterator:
    """Ring chunk dict -> stream of ``ReduceChunk`` (the reference sends ReduceChunk objects to
    ``gather_chunk`` too; GatherChunk is wire-identical, server.proto:25-33)."""
    blob = pickle.dumps(data, protocol=pickle.HIGHEST_PROTOCOL)
    size = len(blob)
    for block in _blocks(blob):
        yield pb.ReduceChunk(ring_id=int(ring_id), data_chunk=pb.DataChunk(buffer=block, type=type or '$tensor', data_size=size))


def join_chunks(parts: Iterable[bytes]) -> bytes:
    return b''.join(parts)
