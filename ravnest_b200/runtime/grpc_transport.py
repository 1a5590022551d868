"""gRPC transport: control plane + CPU/localhost compatibility data path.

Client-side parity: reference ``ravnest/communication.py:67-77`` (``trigger_send``: poll
``buffer_status`` then stream the pickled payload), ``:292-330`` (ring chunk / latest-weights
senders).  Server side: ``node.py:259-328`` — but the server lives on threads in this process.
Differences: channels are cached per peer instead of re-dialled per message; sends are ordered per
lane by a dedicated sender thread (the reference spawns a Thread per message, node.py:486); waits
back off instead of spinning; every RPC carries a deadline and ``Ping`` is actually used (heartbeat).
"""
import pickle
import queue
import threading
import time
from concurrent import futures
from typing import Any, Dict, Optional, Tuple

import grpc
import torch

from ..endpoints import GrpcService, RingBuffers
from ..protos import pb
from ..strings import ActionTypes, BufferStatus
from ..utils.streams import generate_data_stream, generate_stream, join_chunks
from .mailbox import BACKWARD, FORWARD, Mailbox

_OPTIONS = [('grpc.max_send_message_length', 64 << 20), ('grpc.max_receive_message_length', 64 << 20)]


def _cpu_tree(obj: Any) -> Any:
    if isinstance(obj, torch.Tensor):
        return obj.detach().to('cpu')
    if isinstance(obj, dict):
        return {k: _cpu_tree(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_cpu_tree(v) for v in obj)
    return obj


class GrpcTransport:
    kind = 'grpc'

    def __init__(self, name: str, local_host: str, local_port: int, mailbox: Mailbox,
                 forward_target: Optional[Tuple[str, int]] = None, backward_target: Optional[Tuple[str, int]] = None,
                 weights_provider=None, rpc_timeout: float = 600.0):
        self.name = name
        self.local = (local_host, int(local_port))
        self.mailbox = mailbox
        self.targets = {FORWARD: forward_target, BACKWARD: backward_target}
        self.rings = RingBuffers()
        self.service = GrpcService(mailbox, self.rings, weights_provider, name=name)
        self.rpc_timeout = rpc_timeout
        self._server = None
        self._channels: Dict[Tuple[str, int], grpc.Channel] = {}
        self._stubs: Dict[Tuple[str, int], pb.CommServerStub] = {}
        self._queues = {FORWARD: queue.Queue(), BACKWARD: queue.Queue()}
        self._senders = {}
        self._pending = 0
        self._pending_cv = threading.Condition()
        self._stop = threading.Event()
        self.bytes_out = 0

    # ------------------------------------------------------------------ lifecycle
    def start(self) -> None:
        self._server = grpc.server(futures.ThreadPoolExecutor(max_workers=8), options=_OPTIONS)
        pb.add_CommServerServicer_to_server(self.service, self._server)
        bound = self._server.add_insecure_port('{}:{}'.format(*self.local))
        if bound == 0:
            raise RuntimeError('{}: cannot bind {}:{}'.format(self.name, *self.local))
        self._server.start()
        for lane in (FORWARD, BACKWARD):
            t = threading.Thread(target=self._sender_loop, args=(lane,), name='{}-send-{}'.format(self.name, lane), daemon=True)
            t.start()
            self._senders[lane] = t

    def _announce_shutdown(self, deadline_s: float = 5.0) -> None:
        """Best effort: tell the next stage that this node is leaving, so a stopping root takes its whole pipeline down (the
        stem / leaf provider scripts return from ``Trainer.train()``).  Never blocks on a peer that is already gone."""
        target = self.targets.get(FORWARD)
        if target is None or not self.ping(target, timeout=0.5):
            return
        stub, probe = self._stub(target), pb.CheckBufferStatus(name=self.name, type=FORWARD)
        message = {'action': ActionTypes.SHUTDOWN, 'forward_pass_id': -1, 'data': {}}
        end = time.time() + deadline_s
        try:
            while time.time() < end:
                if stub.buffer_status(probe, timeout=1.0).status == BufferStatus.SEND_BUFFER:
                    stub.send_buffer(generate_stream(message, type=FORWARD), timeout=5.0)
                    return
                time.sleep(0.01)
        except grpc.RpcError:
            pass

    def stop(self) -> None:
        if self._stop.is_set():
            return
        self.flush(timeout=5.0)                # payloads already queued go out first
        self._announce_shutdown()
        self._stop.set()
        for q in self._queues.values():
            q.put(None)
        if self._server is not None:
            self._server.stop(grace=0.2)
        for ch in self._channels.values():
            ch.close()

    # ------------------------------------------------------------------ client helpers
    def _stub(self, target: Tuple[str, int]) -> pb.CommServerStub:
        target = (target[0], int(target[1]))
        stub = self._stubs.get(target)
        if stub is None:
            host = '127.0.0.1' if target[0] in ('0.0.0.0', '') else target[0]
            ch = grpc.insecure_channel('{}:{}'.format(host, target[1]), options=_OPTIONS)
            self._channels[target] = ch
            stub = self._stubs[target] = pb.CommServerStub(ch)
        return stub

    def ping(self, target: Tuple[str, int], timeout: float = 2.0) -> bool:
        try:
            return self._stub(target).Ping(pb.PingRequest(data='Ping'), timeout=timeout).data == 'Pong'
        except grpc.RpcError:
            return False

    def wait_for_peer(self, target: Tuple[str, int], timeout: float = 120.0) -> None:
        deadline = time.time() + timeout
        while not self.ping(target, timeout=1.0):
            if time.time() > deadline:
                raise TimeoutError('{}: peer {}:{} did not answer Ping within {}s'.format(self.name, target[0], target[1], timeout))
            time.sleep(0.05)

    # ------------------------------------------------------------------ pipeline payloads
    def send(self, lane: str, message: Dict[str, Any]) -> None:
        """Queue ``message`` for the next (lane='forward') / previous (lane='backward') stage."""
        with self._pending_cv:
            self._pending += 1
        self._queues[lane].put(_cpu_tree(message))

    def _sender_loop(self, lane: str) -> None:
        q = self._queues[lane]
        while not self._stop.is_set():
            message = q.get()
            if message is None:
                return
            try:
                self.trigger_send(message, lane, self.targets[lane])
            finally:
                with self._pending_cv:
                    self._pending -= 1
                    self._pending_cv.notify_all()

    def trigger_send(self, message: Dict[str, Any], lane: str, target: Tuple[str, int]) -> None:
        stub = self._stub(target)
        probe = pb.CheckBufferStatus(name=self.name, type=lane)
        delay = 0.0002
        first = True
        while True:
            try:
                status = stub.buffer_status(probe, timeout=10.0).status
            except grpc.RpcError:
                if first:
                    self.wait_for_peer(target)
                    first = False
                    continue
                raise
            if status == BufferStatus.SEND_BUFFER:
                break
            time.sleep(delay)
            delay = min(delay * 2, 0.01)
        stub.send_buffer(generate_stream(message, type=lane), timeout=self.rpc_timeout)

    def flush(self, timeout: Optional[float] = None) -> None:
        with self._pending_cv:
            self._pending_cv.wait_for(lambda: self._pending == 0, timeout=timeout)

    # ------------------------------------------------------------------ ring collectives / weights
    def send_ring_chunk(self, phase: str, target: Tuple[str, int], ring_id: int, iteration: int, chunks: Dict[str, torch.Tensor]) -> None:
        stub = self._stub(target)
        payload = {'iteration': iteration, 'chunks': _cpu_tree(chunks)}
        rpc = stub.reduce_chunk if phase == 'reduce' else stub.gather_chunk
        try:
            rpc(generate_data_stream(payload, ring_id=ring_id), timeout=self.rpc_timeout)
        except grpc.RpcError:
            self.wait_for_peer(target)
            rpc(generate_data_stream(payload, ring_id=ring_id), timeout=self.rpc_timeout)

    def peer_iteration(self, phase: str, target: Tuple[str, int], ring_id: int) -> int:
        stub = self._stub(target)
        if phase == 'reduce':
            return stub.reduce_iteration(pb.CheckReduceIteration(ring_id=ring_id), timeout=10.0).iteration
        return stub.gather_iteration(pb.CheckGatherIteration(ring_id=ring_id), timeout=10.0).iteration

    def fetch_latest_weights(self, target: Tuple[str, int], first: str, last: str) -> Dict[str, torch.Tensor]:
        stub = self._stub(target)
        request = pb.SendLatestWeights(param_names=pickle.dumps((first, last)))
        parts = [chunk.tensor_chunk.buffer for chunk in stub.get_latest_weights(request, timeout=self.rpc_timeout)]
        return pickle.loads(join_chunks(parts))
