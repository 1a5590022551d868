"""gRPC servicer: the 8 ``CommServer`` endpoints.

Parity: reference ``ravnest/endpoints.py:12-157`` (``GrpcService``).  Same RPC surface and flow
control contract (``buffer_status`` answers ``send_buffer`` / ``wait``; chunked streams are
reassembled by concatenation and un-pickled), but the servicer runs on threads INSIDE the node
process and drops payloads straight into the scheduler's ``Mailbox`` / the collective's condition-
variable buffers — no Manager proxies, no second pickle hop, no polling of shared lists.

On a B200 box these endpoints are the control plane (rendezvous, Ping heartbeat, introspection) and
the CPU/localhost compatibility data path; GPU tensors use ``runtime.dist_transport``.
"""
import pickle
import threading
from collections import defaultdict
from typing import Dict

from .protos import pb
from .runtime.mailbox import BACKWARD, FORWARD, Mailbox
from .strings import BufferStatus
from .utils.streams import generate_weights_stream, join_chunks


class RingBuffers:
    """Per-ring landing zone for reduce / gather chunks, keyed by iteration."""

    def __init__(self):
        self.cv = threading.Condition()
        self.reduce: Dict[int, Dict[int, Dict[str, object]]] = defaultdict(lambda: defaultdict(dict))
        self.gather: Dict[int, Dict[int, Dict[str, object]]] = defaultdict(lambda: defaultdict(dict))
        self.reduce_iteration: Dict[int, int] = defaultdict(int)
        self.gather_iteration: Dict[int, int] = defaultdict(int)

    def deposit(self, phase: str, ring_id: int, iteration: int, chunks: Dict[str, object]) -> None:
        with self.cv:
            getattr(self, phase)[ring_id][iteration].update(chunks)
            self.cv.notify_all()

    def collect(self, phase: str, ring_id: int, iteration: int, keys, timeout=None) -> Dict[str, object]:
        keys = set(keys)
        with self.cv:
            store = getattr(self, phase)[ring_id]
            ok = self.cv.wait_for(lambda: keys.issubset(store[iteration].keys()), timeout=timeout)
            if not ok:
                raise TimeoutError('ring {} {} iteration {}: missing {}'.format(
                    ring_id, phase, iteration, sorted(keys - set(store[iteration].keys()))[:4]))
            return store.pop(iteration)

    def advance(self, phase: str, ring_id: int) -> None:
        with self.cv:
            counter = self.reduce_iteration if phase == 'reduce' else self.gather_iteration
            counter[ring_id] += 1
            self.cv.notify_all()

    def reset(self, ring_id: int) -> None:
        with self.cv:
            self.reduce_iteration[ring_id] = 0
            self.gather_iteration[ring_id] = 0


class GrpcService(object):
    """Servicer object handed to ``pb.add_CommServerServicer_to_server``."""

    def __init__(self, mailbox: Mailbox, rings: RingBuffers, weights_provider=None, name: str = ''):
        self.mailbox = mailbox
        self.rings = rings
        self.weights_provider = weights_provider   # callable(first, last) -> {name: tensor}
        self.name = name
        self.stats = {'send_buffer': 0, 'reduce_chunk': 0, 'gather_chunk': 0, 'ping': 0, 'bytes_in': 0}

    # ---- pipeline payloads -------------------------------------------------------------------
    def send_buffer(self, request_iterator, context):
        lane, parts = None, []
        for req in request_iterator:
            lane = req.type or lane
            parts.append(req.tensor_chunk.buffer)
        blob = join_chunks(parts)
        self.stats['send_buffer'] += 1
        self.stats['bytes_in'] += len(blob)
        message = pickle.loads(blob)
        self.mailbox.put(BACKWARD if lane == BACKWARD else FORWARD, message)
        return pb.SendTensorReply(reply=True)

    def buffer_status(self, request, context):
        lane = BACKWARD if request.type == BACKWARD else FORWARD
        ok = self.mailbox.has_room(lane)
        return pb.BufferStatusReply(status=BufferStatus.SEND_BUFFER if ok else BufferStatus.WAIT)

    # ---- ring all-reduce ---------------------------------------------------------------------
    def _ring_chunk(self, phase, request_iterator):
        ring_id, parts = 0, []
        for req in request_iterator:
            ring_id = req.ring_id
            parts.append(req.data_chunk.buffer)
        payload = pickle.loads(join_chunks(parts))
        self.stats[phase + '_chunk'] += 1
        self.rings.deposit(phase, ring_id, payload['iteration'], payload['chunks'])
        return pb.ReceivedChunk(reply=True)

    def reduce_chunk(self, request_iterator, context):
        return self._ring_chunk('reduce', request_iterator)

    def gather_chunk(self, request_iterator, context):
        return self._ring_chunk('gather', request_iterator)

    def reduce_iteration(self, request, context):
        return pb.ReduceIterationReply(iteration=self.rings.reduce_iteration[request.ring_id])

    def gather_iteration(self, request, context):
        return pb.GatherIterationReply(iteration=self.rings.gather_iteration[request.ring_id])

    # ---- latest weights / liveness -----------------------------------------------------------
    def get_latest_weights(self, request, context):
        first, last = pickle.loads(request.param_names)
        weights = self.weights_provider(first, last) if self.weights_provider else {}
        for chunk in generate_weights_stream(weights):
            yield chunk

    def Ping(self, request, context):
        self.stats['ping'] += 1
        return pb.PingResponse(data='Pong')
