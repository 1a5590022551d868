"""Tracing / metrics (new; the reference has only ``print`` and an NVML init per step — SURVEY.md §5.1, §5.5).

* NVTX ranges per (stage, fpid, phase) when ``RAVNEST_NVTX=1`` (visible to nsys / ncu range filters);
* device-timed phase accounting with CUDA events (forward / backward / update), resolved lazily so the hot
  path never synchronises;
* exposed-communication accounting: the time the compute stream actually stalls on an incoming payload's
  ready-event (``record A -> wait_event(recv) -> record B``; ``elapsed(A, B)`` is ~0 when the transfer was
  fully overlapped) — the "exposed (non-overlapped) comm time per step" metric of BASELINE.json;
* a JSON-lines metrics file (``<name>_metrics.jsonl``) instead of NodeStatus fields nobody reads.
"""
import contextlib
import json
import os
import time
from collections import defaultdict
from typing import Dict, List, Optional, Tuple

import torch

NVTX = os.environ.get('RAVNEST_NVTX', '0') == '1'


@contextlib.contextmanager
def nvtx_range(name: str):
    if NVTX and torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield


class StageTracer:
    def __init__(self, name: str, device: torch.device, enabled: bool = True, path: Optional[str] = None):
        self.name, self.device = name, device
        self.enabled = enabled and device.type == 'cuda'
        self.path = path
        self._pending: List[Tuple[str, torch.cuda.Event, torch.cuda.Event]] = []
        self.totals_ms: Dict[str, float] = defaultdict(float)
        self.counts: Dict[str, int] = defaultdict(int)
        self.wall_wait_s = 0.0
        self._t0 = time.time()

    @contextlib.contextmanager
    def phase(self, kind: str, fpid: int = -1):
        """Device-time a phase on the current stream (and open an NVTX range)."""
        if not self.enabled:
            with nvtx_range('{}:{}:{}'.format(self.name, kind, fpid)):
                yield
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with nvtx_range('{}:{}:{}'.format(self.name, kind, fpid)):
            a.record()
            try:
                yield
            finally:
                b.record()
                self._pending.append((kind, a, b))
                if len(self._pending) > 4096:
                    self.resolve(block=False)

    def exposed_wait(self, event: torch.cuda.Event) -> None:
        """Make the current stream wait for ``event`` and account the stall as exposed communication."""
        stream = torch.cuda.current_stream(self.device)
        if not self.enabled:
            stream.wait_event(event)
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        stream.wait_event(event)
        b.record(stream)
        self._pending.append(('exposed_comm', a, b))

    def idle(self, seconds: float) -> None:
        self.wall_wait_s += seconds

    def resolve(self, block: bool = True) -> None:
        keep = []
        for kind, a, b in self._pending:
            if not b.query():
                if not block:
                    keep.append((kind, a, b))
                    continue
                b.synchronize()
            self.totals_ms[kind] += a.elapsed_time(b)
            self.counts[kind] += 1
        self._pending = keep

    def summary(self) -> Dict[str, float]:
        self.resolve(block=True)
        out = {'node': self.name, 'wall_s': time.time() - self._t0, 'mailbox_idle_s': self.wall_wait_s}
        for k, v in self.totals_ms.items():
            out[k + '_ms'] = v
            out[k + '_n'] = self.counts[k]
        return out

    def dump(self) -> Optional[str]:
        if self.path is None:
            return None
        with open(self.path, 'a') as fh:
            fh.write(json.dumps(self.summary()) + '\n')
        return self.path
