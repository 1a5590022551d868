"""Failure detection + fault injection (new; SURVEY.md §5.3: the reference defines ``Ping`` but never calls it and
every wait is an unbounded spin).

* ``Heartbeat``: a daemon thread pings the pipeline neighbours over the gRPC control plane every ``period`` seconds
  and records (alive, last_seen, consecutive_misses); ``on_failure(peer)`` fires once per outage.  With the NCCL
  transport the async-error watchdog of ProcessGroupNCCL stays enabled (``TORCH_NCCL_ASYNC_ERROR_HANDLING``) and the
  heartbeat is what tells a healthy stage *which* neighbour died.
* ``TransferWatchdog``: watches the NCCL p2p transfers of the dist transport and surfaces a stalled / failed one as an exception on the
  training thread long before ProcessGroupNCCL's own process-killing timeout.
* ``FaultInjector``: deterministic hooks (``RAVNEST_FAULT='drop_forward:7,delay_backward:3:0.5,kill:40'``) consulted by
  the node scheduler — used by tests to exercise timeouts and the rejoin path (``update_with_latest_weights``).
"""
import os
import threading
import time
from typing import Callable, Dict, List, Optional, Tuple


class Heartbeat(threading.Thread):
    def __init__(self, transport, peers: List[Tuple[str, int]], period: float = 1.0, misses_to_fail: int = 3,
                 on_failure: Optional[Callable[[Tuple[str, int]], None]] = None):
        super().__init__(daemon=True, name='heartbeat')
        self.transport, self.peers, self.period = transport, list(peers), period
        self.misses_to_fail = misses_to_fail
        self.on_failure = on_failure
        self.state: Dict[Tuple[str, int], Dict[str, float]] = {p: {'alive': True, 'last_seen': 0.0, 'misses': 0} for p in self.peers}
        self._halt = threading.Event()

    def run(self) -> None:
        while not self._halt.wait(self.period):
            for peer in self.peers:
                st = self.state[peer]
                if self.transport.ping(peer, timeout=min(1.0, self.period)):
                    st.update(alive=True, last_seen=time.time(), misses=0)
                else:
                    st['misses'] += 1
                    if st['misses'] == self.misses_to_fail:
                        st['alive'] = False
                        if self.on_failure is not None:
                            self.on_failure(peer)

    def stop(self) -> None:
        self._halt.set()

    def dead_peers(self) -> List[Tuple[str, int]]:
        return [p for p, st in self.state.items() if not st['alive']]


class TransferWatchdog(threading.Thread):
    """NCCL-side failure detection for the dist transport (new).  ProcessGroupNCCL's own async-error handling tears the whole process
    down after ``TORCH_NCCL_TIMEOUT`` (minutes); a pipeline stage wants to know much earlier — and as a Python exception on the thread
    that drives training — that a neighbour stopped answering.  The watchdog looks at the transport's in-flight ``isend`` work objects
    (kept per lane together with their enqueue time): a transfer that has not completed after ``timeout`` seconds, or a work object
    that reports an exception, is turned into ``on_failure(lane, reason)`` (the Node records it; the next ``forward_compute`` /
    ``wait_for_backwards`` raises).  Polling never touches a CUDA stream: ``is_completed()`` is a host-side query."""

    def __init__(self, transport, timeout: float = 60.0, period: float = 1.0,
                 on_failure: Optional[Callable[[str, str], None]] = None):
        super().__init__(daemon=True, name='transfer-watchdog')
        self.transport, self.timeout, self.period, self.on_failure = transport, float(timeout), float(period), on_failure
        self._halt = threading.Event()
        self.fired: List[Tuple[str, str]] = []

    def check_once(self, now: Optional[float] = None) -> List[Tuple[str, str]]:
        now = time.time() if now is None else now
        found = []
        for lane, works in list(getattr(self.transport, '_inflight_since', {}).items()):
            for work, since in list(works):
                reason = None
                try:
                    if work.is_completed():
                        exc = work.exception() if hasattr(work, 'exception') else None
                        if exc is not None:
                            reason = 'transfer failed: {!r}'.format(exc)
                    elif now - since > self.timeout:
                        reason = 'transfer pending for {:.0f} s (peer of the {} lane not receiving)'.format(now - since, lane)
                except RuntimeError as err:          # NCCL reports async errors through the query itself
                    reason = 'NCCL error: {}'.format(err)
                if reason is not None:
                    found.append((lane, reason))
        return found

    def run(self) -> None:
        while not self._halt.wait(self.period):
            for lane, reason in self.check_once():
                if (lane, reason[:24]) in [(l, r[:24]) for l, r in self.fired]:
                    continue
                self.fired.append((lane, reason))
                if self.on_failure is not None:
                    self.on_failure(lane, reason)

    def stop(self) -> None:
        self._halt.set()


class FaultInjector:
    """Parse ``RAVNEST_FAULT`` into per-fpid actions."""

    def __init__(self, spec: Optional[str] = None):
        self.rules = []
        spec = spec if spec is not None else os.environ.get('RAVNEST_FAULT', '')
        for item in filter(None, (s.strip() for s in spec.split(','))):
            parts = item.split(':')
            self.rules.append((parts[0], int(parts[1]) if len(parts) > 1 else -1, float(parts[2]) if len(parts) > 2 else 0.0))
        self.fired = []

    def check(self, action: str, fpid: int) -> Optional[str]:
        """Returns 'drop' to swallow the message, sleeps for 'delay', raises SystemExit for 'kill'."""
        for kind, at, arg in self.rules:
            if at != fpid:
                continue
            if kind == 'kill':
                self.fired.append((kind, fpid))
                raise SystemExit('fault injection: kill at fpid {}'.format(fpid))
            if kind == 'drop_' + action:
                self.fired.append((kind, fpid))
                return 'drop'
            if kind == 'delay_' + action:
                self.fired.append((kind, fpid))
                time.sleep(arg)
        return None
