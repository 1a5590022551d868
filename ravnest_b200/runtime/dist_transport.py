"""torch.distributed transport: the B200 data plane for pipeline payloads.

Replaces the reference's ``GPU -> .to('cpu') -> pickle -> 2 MiB gRPC chunks -> unpickle -> Manager
proxy -> .to(device)`` hop (``ravnest/communication.py:67-77,:108``; ``endpoints.py:36-69``;
SURVEY.md §2.3) with:

  * tensors: NCCL p2p (``isend``/``irecv``) straight between HBMs over NVLink/NVSwitch, on dedicated
    send / recv CUDA streams coupled to the compute stream by events (no host sync on the hot path);
  * headers: a tiny pickled dict (action, fpid, flags, tensor shapes/dtypes, non-tensor values) over a
    gloo pair-group — the only thing the CPU ever touches;
  * one communicator per (stage pair, direction): forward traffic s -> s+1 and backward traffic
    s+1 -> s never share a communicator, so an in-flight activation send can never deadlock against
    an in-flight gradient send (NCCL executes a communicator's ops in issue order);
  * DP groups (one per stage across replicas) for the collectives in ``parallel/``.

With the ``gloo`` default backend the same code runs on CPU (tests run it with world_size > 1).
rank == node_id (``node_<id>``), one process per GPU.
"""
import os
import pickle
import queue
import threading
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from ..strings import ActionTypes
from .mailbox import BACKWARD, FORWARD

_TAG = {FORWARD: 11, BACKWARD: 13}


def init_distributed(backend: Optional[str] = None, device: Optional[torch.device] = None) -> Tuple[int, int, torch.device]:
    """Join the job described by RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT."""
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', rank))
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29531')
    use_cuda = torch.cuda.is_available() and (device is None or torch.device(device).type == 'cuda')
    if use_cuda:
        dev = torch.device('cuda', local % torch.cuda.device_count())
        torch.cuda.set_device(dev)
    else:
        dev = torch.device('cpu')
    if not dist.is_initialized():
        if backend is None:
            backend = 'cuda:nccl,cpu:gloo' if use_cuda else 'gloo'
        # NOT bound to a device (no ``device_id=``): on a device-bound default group every later ``new_group`` — even a gloo-only
        # one — makes its NON-members run a blocking nocolor ``ncclCommSplit`` that the members of a gloo group never join; with
        # more than two stages per chain that desynchronises the split sequence and dead-locks group creation (seen on 4 x B200).
        # Communicators are created lazily instead and warmed up in one global order by ``_Groups._warm_up``.
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, dev


def _milestone(text: str) -> None:
    """Start-up progress on stderr when RAVNEST_TRACE_STARTUP=1 (multi-GPU hangs are expensive to reproduce)."""
    if os.environ.get('RAVNEST_TRACE_STARTUP', '0') == '1':
        import sys
        import time
        print('[rank {} t={:.1f}] {}'.format(os.environ.get('RANK', '0'), time.time() % 10000, text), file=sys.stderr, flush=True)


def _load_chains(base_dir: str) -> Dict[int, List[int]]:
    import json
    folder = os.path.join(base_dir, 'node_data', 'nodes')
    chains: Dict[int, List[int]] = {}
    for fname in sorted(os.listdir(folder)):
        if fname.startswith('node_') and fname.endswith('.json'):
            with open(os.path.join(folder, fname)) as fh:
                cfg = json.load(fh)
            ids = cfg.get('cluster_node_ids')
            if ids:
                chains[int(cfg['rank'])] = [int(i) for i in ids]
    return chains


_FRAME = max(64, int(os.environ.get('RAVNEST_HEADER_FRAME', '4096')))    # bytes of the fixed header frame (8-byte length + pickled skeleton)


class _Groups:
    """Every rank builds the same groups in the same order (``new_group`` is collective)."""
    _cache: Dict[str, '_Groups'] = {}

    def __init__(self, chains: Dict[int, List[int]]):
        self.pair: Dict[Tuple[int, int, str], Any] = {}       # (src, dst, lane) -> tensor group
        self.header: Dict[Tuple[int, int, str], Any] = {}     # (src, dst, lane) -> gloo group
        self.dp: Dict[int, Any] = {}                          # stage index -> group
        self.dp_ranks: Dict[int, List[int]] = {}
        need_gloo = dist.get_backend() != 'gloo'
        for cid in sorted(chains):
            chain = chains[cid]
            for a, b in zip(chain[:-1], chain[1:]):
                for lane, (src, dst) in ((FORWARD, (a, b)), (BACKWARD, (b, a))):
                    # tensors: one NCCL communicator per (pair, direction) so the two lanes never serialise on one comm stream;
                    # headers: a gloo pair group (explicit backends: the default 'cuda:nccl,cpu:gloo' would build both per group)
                    g = dist.new_group(ranks=sorted([a, b]), backend='nccl' if need_gloo else None)
                    self.pair[(src, dst, lane)] = g
                    self.header[(src, dst, lane)] = dist.new_group(ranks=sorted([a, b]), backend='gloo') if need_gloo else g
        lengths = {len(c) for c in chains.values()}
        if len(chains) > 1 and len(lengths) == 1:
            for s in range(lengths.pop()):
                ranks = [chains[cid][s] for cid in sorted(chains)]
                self.dp[s] = dist.new_group(ranks=ranks)
                self.dp_ranks[s] = ranks
        _milestone('groups created: {} pair lanes, {} dp'.format(len(self.pair), len(self.dp)))
        self._warm_up()
        _milestone('communicators warm')

    def _warm_up(self) -> None:
        """NCCL builds a communicator on first use (for unbatched p2p on a 2-rank group: a dedicated one), from whichever thread
        gets there first — and that creation synchronises the device, which invalidates a CUDA-graph capture in progress on
        another thread.  So every communicator the transport will use is created HERE, before any node thread exists, one at a
        time and in the same global order on every rank: a real send/recv per lane, an all-reduce per DP group."""
        if not torch.cuda.is_available() or 'nccl' not in str(dist.get_backend()).lower():
            return
        me = dist.get_rank()
        dev = torch.device('cuda', torch.cuda.current_device())
        for (src, dst, lane), g in self.pair.items():
            if me == src:
                dist.send(torch.ones(8, device=dev), dst=dst, group=g)
            elif me == dst:
                dist.recv(torch.empty(8, device=dev), src=src, group=g)
            if me in (src, dst):
                torch.cuda.synchronize(dev)
        for s, g in self.dp.items():
            if me in self.dp_ranks[s]:
                dist.all_reduce(torch.zeros(1, device=dev), group=g)
                torch.cuda.synchronize(dev)

    @classmethod
    def get(cls, chains: Dict[int, List[int]]) -> '_Groups':
        key = repr(sorted(chains.items()))
        if key not in cls._cache:
            cls._cache[key] = cls(chains)
        return cls._cache[key]


def _split_tensors(obj: Any, bag: List[torch.Tensor]) -> Any:
    """Replace every tensor in a nested message by a placeholder; tensors go to ``bag``."""
    if isinstance(obj, torch.Tensor):
        bag.append(obj)
        return ('__tensor__', len(bag) - 1, tuple(obj.shape), obj.dtype)
    if isinstance(obj, dict):
        return {k: _split_tensors(v, bag) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_split_tensors(v, bag) for v in obj]
    if isinstance(obj, tuple) and not (len(obj) == 4 and obj and obj[0] == '__tensor__'):
        return tuple(_split_tensors(v, bag) for v in obj)
    return obj


def _join_tensors(obj: Any, bag: List[torch.Tensor]) -> Any:
    if isinstance(obj, tuple) and len(obj) == 4 and obj[0] == '__tensor__':
        return bag[obj[1]]
    if isinstance(obj, dict):
        return {k: _join_tensors(v, bag) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_join_tensors(v, bag) for v in obj]
    if isinstance(obj, tuple):
        return tuple(_join_tensors(v, bag) for v in obj)
    return obj


def _tensor_specs(skeleton: Any, out: List[Tuple[Tuple[int, ...], torch.dtype]]) -> None:
    if isinstance(skeleton, tuple) and len(skeleton) == 4 and skeleton[0] == '__tensor__':
        while len(out) <= skeleton[1]:
            out.append(None)
        out[skeleton[1]] = (skeleton[2], skeleton[3])
    elif isinstance(skeleton, dict):
        for v in skeleton.values():
            _tensor_specs(v, out)
    elif isinstance(skeleton, (list, tuple)):
        for v in skeleton:
            _tensor_specs(v, out)


_VOLATILE = ('forward_pass_id', 'epoch_change', 'action')
_FAST_MAGIC = 0x52564E46          # 'RVNF'


def _signature(obj: Any) -> Any:
    """Hashable structural signature of a tensor-free skeleton (keys, placeholders with shape / dtype, scalar leaves)."""
    if isinstance(obj, dict):
        return ('d',) + tuple((k, _signature(v)) for k, v in obj.items())
    if isinstance(obj, (list, tuple)):
        return ('l' if isinstance(obj, list) else 't',) + tuple(_signature(v) for v in obj)
    return obj if isinstance(obj, (int, float, str, bool, bytes, type(None), torch.dtype)) else repr(obj)


def _strip_volatile(message: Dict[str, Any]) -> Tuple[Dict[str, Any], Tuple[Any, ...]]:
    """Split a message into its per-step scalars and the rest (whose tensor-free skeleton is identical step after step)."""
    rest = {k: v for k, v in message.items() if k not in _VOLATILE}
    return rest, (message.get('action'), int(message.get('forward_pass_id', -1)), bool(message.get('epoch_change', False)))


class DistTransport:
    kind = 'dist'

    def __init__(self, node, chains: Optional[Dict[int, List[int]]] = None, **_unused):
        self.node = node
        self.rank = dist.get_rank()
        self.device = node.device
        self.mailbox = node.mailbox
        self.cuda = self.device.type == 'cuda'
        cfg = node.config
        chains = chains or _load_chains(node.base_dir)
        self.chain = chains[int(cfg['rank'])]
        self.stage = self.chain.index(self.rank)
        self.groups = _Groups.get(chains)
        self.next_rank = self.chain[self.stage + 1] if self.stage + 1 < len(self.chain) else None
        self.prev_rank = self.chain[self.stage - 1] if self.stage > 0 else None
        self.dp_group = self.groups.dp.get(self.stage)
        self.dp_ranks = self.groups.dp_ranks.get(self.stage)
        self._queues = {FORWARD: queue.Queue(), BACKWARD: queue.Queue()}
        self._threads: List[threading.Thread] = []
        self._stop = threading.Event()
        self._pending = 0
        self._pending_cv = threading.Condition()
        self._inflight: Dict[str, List[Any]] = {FORWARD: [], BACKWARD: []}
        self._inflight_since: Dict[str, List[Any]] = {FORWARD: [], BACKWARD: []}     # (work, enqueue time) for the TransferWatchdog
        self.watchdog = None
        self.send_stream = {FORWARD: None, BACKWARD: None}
        self.recv_stream = {FORWARD: None, BACKWARD: None}
        self.bytes_out = 0
        self.rings = None
        self.just_in_time = os.environ.get('RAVNEST_P2P_JIT', '1') != '0'
        # static wire plan, learned: the first message of a given shape on a lane travels with its pickled skeleton and gets an id;
        # every later message with the same skeleton is announced by a fixed 64-byte binary header (id, action code, fpid, flags) —
        # no pickle / unpickle, no shape metadata on the wire in steady state (the cut tensors of a stage never change shape)
        self.static_headers = os.environ.get('RAVNEST_STATIC_HEADERS', '1') != '0'
        self._sent_skeletons: Dict[str, Dict[bytes, int]] = {FORWARD: {}, BACKWARD: {}}
        self._known_skeletons: Dict[str, Dict[int, Any]] = {FORWARD: {}, BACKWARD: {}}
        self._known_volatile: Dict[Tuple[str, int], Tuple[str, ...]] = {}
        self.header_stats = {'pickled': 0, 'static': 0}

    # ------------------------------------------------------------------ lifecycle
    def start(self) -> None:
        if self.cuda:
            for lane in (FORWARD, BACKWARD):
                self.send_stream[lane] = torch.cuda.Stream(self.device)
                self.recv_stream[lane] = torch.cuda.Stream(self.device)
        plans = []
        if self.next_rank is not None:
            plans.append((self._sender_loop, FORWARD, self.next_rank))
            plans.append((self._receiver_loop, BACKWARD, self.next_rank))
        if self.prev_rank is not None:
            plans.append((self._sender_loop, BACKWARD, self.prev_rank))
            plans.append((self._receiver_loop, FORWARD, self.prev_rank))
        for fn, lane, peer in plans:
            t = threading.Thread(target=fn, args=(lane, peer), daemon=True,
                                 name='{}-{}-{}'.format(self.node.name, fn.__name__, lane))
            t.start()
            self._threads.append(t)
        timeout = float(os.environ.get('RAVNEST_TRANSFER_TIMEOUT', '0') or 0)
        if self.cuda and timeout > 0:
            from .watchdog import TransferWatchdog

            def failed(lane, reason):
                with self.node._cv:
                    self.node._errors.append(RuntimeError('{} lane of {}: {}'.format(lane, self.node.name, reason)))
                    self.node._cv.notify_all()
            self.watchdog = TransferWatchdog(self, timeout=timeout, on_failure=failed)
            self.watchdog.start()

    def stop(self) -> None:
        if self._stop.is_set():
            return
        # tell both neighbours so their receiver threads leave their blocking recv
        for lane, peer in ((FORWARD, self.next_rank), (BACKWARD, self.prev_rank)):
            if peer is not None:
                self.send(lane, {'action': ActionTypes.SHUTDOWN, 'forward_pass_id': -1, 'data': {}})
        self.flush(timeout=10)
        if self.watchdog is not None:
            self.watchdog.stop()
        self._stop.set()
        for q in self._queues.values():
            q.put(None)

    # ------------------------------------------------------------------ send side
    def send(self, lane: str, message: Dict[str, Any]) -> None:
        ev = None
        if self.cuda:
            # blocking-sync event: the lane's sender thread SLEEPS on it until the producing kernels have finished (see
            # _sender_loop: NCCL kernels are only launched once the payload exists)
            ev = torch.cuda.Event(blocking=True)
            ev.record(torch.cuda.current_stream(self.device))
        with self._pending_cv:
            self._pending += 1
        self._queues[lane].put((message, ev))

    def flush(self, timeout: Optional[float] = None) -> None:
        with self._pending_cv:
            self._pending_cv.wait_for(lambda: self._pending == 0, timeout=timeout)

    def _sender_loop(self, lane: str, peer: int) -> None:
        if self.cuda:
            torch.cuda.set_device(self.device)
        group = self.groups.pair[(self.rank, peer, lane)]
        hgroup = self.groups.header[(self.rank, peer, lane)]
        q = self._queues[lane]
        while True:
            item = q.get()
            if item is None:
                return
            message, ev = item
            try:
                bag: List[torch.Tensor] = []
                skeleton = _split_tensors(message, bag)
                if bag and ev is not None and self.just_in_time:
                    # Just-in-time rendezvous.  The host runs far ahead of the GPU (a whole stage step is one graph launch), so a
                    # header sent at enqueue time makes the receiver launch its ncclRecv kernel milliseconds before the data exists:
                    # that kernel then sits on the peer's SMs spinning (measured: 3.6 - 4.6 ms of resident ncclDevKernel_SendRecv per
                    # 5 ms BERT step, 10 ms per 12.7 ms ResNet step, and every concurrent GEMM / BN kernel 20 - 28 % slower).  Waiting
                    # for the producer's event HERE (this thread sleeps, the scheduler thread is not involved) keeps both NCCL
                    # kernels down to the transfer itself.
                    ev.synchronize()
                frame = torch.zeros(_FRAME, dtype=torch.uint8)
                fast_id = None
                if self.static_headers and bag:
                    from ..strings import ACTION_CODES
                    rest_skel, (action, fpid, epoch_change) = _strip_volatile(skeleton)
                    code = ACTION_CODES.get(action)
                    if code is not None:
                        key = _signature(rest_skel)
                        known = self._sent_skeletons[lane]
                        fast_id = known.get(key)
                        if fast_id is None:
                            known[key] = len(known) + 1              # announced by THIS (pickled) message, usable from the next one
                            skeleton = dict(skeleton, _skeleton_id=known[key])
                if fast_id is not None:
                    import struct
                    frame[:8] = torch.frombuffer(bytearray((0).to_bytes(8, 'little')), dtype=torch.uint8)       # length 0 = static header
                    frame[8:40] = torch.frombuffer(bytearray(struct.pack('<IIqqII', _FAST_MAGIC, fast_id, fpid, code, int(epoch_change), 0)),
                                                   dtype=torch.uint8)
                    dist.send(frame, dst=peer, group=hgroup, tag=_TAG[lane])
                    self.header_stats['static'] += 1
                else:
                    raw = pickle.dumps(skeleton, protocol=pickle.HIGHEST_PROTOCOL)
                    # ONE fixed-size frame carries the length and (almost always) the whole pickled skeleton; only a header that
                    # does not fit is followed by a second message with the rest (was: a length message + a blob message per hop)
                    frame[:8] = torch.frombuffer(bytearray(len(raw).to_bytes(8, 'little')), dtype=torch.uint8)
                    head = min(len(raw), _FRAME - 8)
                    frame[8:8 + head] = torch.frombuffer(bytearray(raw[:head]), dtype=torch.uint8)
                    dist.send(frame, dst=peer, group=hgroup, tag=_TAG[lane])
                    if len(raw) > head:
                        dist.send(torch.frombuffer(bytearray(raw[head:]), dtype=torch.uint8), dst=peer, group=hgroup, tag=_TAG[lane])
                    self.header_stats['pickled'] += 1
                if bag:
                    if self.cuda:
                        stream = self.send_stream[lane]
                        stream.wait_event(ev)
                        with torch.cuda.stream(stream):
                            works = []
                            for t in bag:
                                t = t if t.is_contiguous() else t.contiguous()
                                if t.device != self.device:
                                    t = t.to(self.device, non_blocking=True)
                                works.append((dist.isend(t, dst=peer, group=group), t))
                                self.bytes_out += t.numel() * t.element_size()
                        keep = self._inflight[lane]
                        keep.extend(works)
                        self._inflight[lane] = [(w, t) for w, t in keep if not w.is_completed()][-64:]
                        import time as _time
                        now = _time.time()
                        self._inflight_since[lane] = [(w, s0) for w, s0 in self._inflight_since[lane] if not w.is_completed()][-64:] + \
                            [(w, now) for w, _t in works]
                    else:
                        for t in bag:
                            dist.send(t.contiguous(), dst=peer, group=group, tag=_TAG[lane] + 1)
                            self.bytes_out += t.numel() * t.element_size()
            finally:
                with self._pending_cv:
                    self._pending -= 1
                    self._pending_cv.notify_all()

    # ------------------------------------------------------------------ receive side
    def _receiver_loop(self, lane: str, peer: int) -> None:
        if self.cuda:
            torch.cuda.set_device(self.device)
        group = self.groups.pair[(peer, self.rank, lane)]
        hgroup = self.groups.header[(peer, self.rank, lane)]
        while not self._stop.is_set():
            frame = torch.zeros(_FRAME, dtype=torch.uint8)
            try:
                dist.recv(frame, src=peer, group=hgroup, tag=_TAG[lane])
                total = int.from_bytes(frame[:8].numpy().tobytes(), 'little')
                head = min(total, _FRAME - 8)
                raw = frame[8:8 + head].numpy().tobytes()
                if total > head:
                    rest = torch.empty(total - head, dtype=torch.uint8)
                    dist.recv(rest, src=peer, group=hgroup, tag=_TAG[lane])
                    raw += rest.numpy().tobytes()
            except RuntimeError:
                if self._stop.is_set():
                    return
                raise
            if total == 0:
                # static header: the skeleton (structure, shapes, dtypes) was announced by an earlier message of this lane
                import struct
                from ..strings import CODE_ACTIONS
                magic, skel_id, fpid, code, epoch_change, _pad = struct.unpack('<IIqqII', frame[8:40].numpy().tobytes())
                assert magic == _FAST_MAGIC, 'corrupt static header'
                skeleton = dict(self._known_skeletons[lane][skel_id])
                skeleton['action'] = CODE_ACTIONS[int(code)]
                skeleton['forward_pass_id'] = int(fpid)
                if epoch_change or 'epoch_change' in self._known_volatile.get((lane, skel_id), ()):
                    skeleton['epoch_change'] = bool(epoch_change)
            else:
                skeleton = pickle.loads(raw)
                skel_id = skeleton.pop('_skeleton_id', None)
                if skel_id is not None:
                    rest_skel, _ = _strip_volatile(skeleton)
                    self._known_skeletons[lane][skel_id] = rest_skel
                    self._known_volatile[(lane, skel_id)] = tuple(k for k in _VOLATILE if k in skeleton)
            specs: List[Tuple[Tuple[int, ...], torch.dtype]] = []
            _tensor_specs(skeleton, specs)
            bag: List[torch.Tensor] = []
            ev = None
            if specs:
                if self.cuda:
                    stream = self.recv_stream[lane]
                    with torch.cuda.stream(stream):
                        works = []
                        for shape, dtype in specs:
                            t = torch.empty(shape, dtype=dtype, device=self.device)
                            works.append(dist.irecv(t, src=peer, group=group))
                            bag.append(t)
                        for w in works:
                            w.wait()
                        ev = torch.cuda.Event()
                        ev.record(stream)
                else:
                    for shape, dtype in specs:
                        t = torch.empty(shape, dtype=dtype)
                        dist.recv(t, src=peer, group=group, tag=_TAG[lane] + 1)
                        bag.append(t)
            message = _join_tensors(skeleton, bag)
            if message.get('action') == ActionTypes.SHUTDOWN:
                # a neighbour is leaving: stop serving this lane; the scheduler decides about the node
                if lane == FORWARD:
                    self.mailbox.put(FORWARD, message, block=False)
                return
            if ev is not None:
                message['_ready_event'] = ev
                message['_tensors'] = bag
            self.mailbox.put(lane, message)

    # ------------------------------------------------------------------ helpers for the scheduler
    @staticmethod
    def adopt(message: Dict[str, Any], device: torch.device, tracer=None) -> None:
        """Make the scheduler's stream wait for the payload and own its memory; the stall (if any) is
        accounted as exposed communication."""
        ev = message.pop('_ready_event', None)
        tensors = message.pop('_tensors', None)
        if ev is not None:
            stream = torch.cuda.current_stream(device)
            if tracer is not None:
                tracer.exposed_wait(ev)
            else:
                stream.wait_event(ev)
            for t in tensors or ():
                t.record_stream(stream)
