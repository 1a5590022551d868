"""Two-lane mailbox feeding a Node's scheduler thread.

Replaces the reference's ``multiprocessing.Manager`` list proxies + busy-wait polling
(``ravnest/node.py:92-98``, ``:330-367``; ``endpoints.py:71-89``): one process, condition variables,
no pickling between the receiver and the scheduler.  Scheduling policy is the reference's: pending
backward messages are always served before forward messages (``node.py:338-362``).
"""
import threading
from collections import deque
from typing import Any, Optional, Tuple

FORWARD = 'forward'
BACKWARD = 'backward'


class Mailbox:
    def __init__(self, capacity: int = 4):
        self._lanes = {FORWARD: deque(), BACKWARD: deque()}
        self._cv = threading.Condition()
        self._capacity = max(1, int(capacity))
        self._closed = False
        self.received = {FORWARD: 0, BACKWARD: 0}

    def put(self, lane: str, message: Any, block: bool = True) -> None:
        with self._cv:
            if block:
                while len(self._lanes[lane]) >= self._capacity and not self._closed:
                    self._cv.wait(0.5)
            self._lanes[lane].append(message)
            self.received[lane] += 1
            self._cv.notify_all()

    def has_room(self, lane: str) -> bool:
        with self._cv:
            return len(self._lanes[lane]) < self._capacity

    def depth(self, lane: str) -> int:
        with self._cv:
            return len(self._lanes[lane])

    def get(self, timeout: Optional[float] = None) -> Optional[Tuple[str, Any]]:
        """Next message, backward lane first.  Returns ``None`` on timeout/close."""
        with self._cv:
            while True:
                if self._lanes[BACKWARD]:
                    msg = self._lanes[BACKWARD].popleft()
                    self._cv.notify_all()
                    return BACKWARD, msg
                if self._lanes[FORWARD]:
                    msg = self._lanes[FORWARD].popleft()
                    self._cv.notify_all()
                    return FORWARD, msg
                if self._closed:
                    return None
                if not self._cv.wait(timeout):
                    return None

    def close(self) -> None:
        with self._cv:
            self._closed = True
            self._cv.notify_all()

    def empty(self) -> bool:
        with self._cv:
            return not self._lanes[FORWARD] and not self._lanes[BACKWARD]
