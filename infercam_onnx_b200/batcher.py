"""Stream batcher — SURVEY.md §8f row N1, the first component "next" to the hot path.

The reference infers one frame at a time in one task (`Inferer::run`, /root/reference/infer_server/src/inferer.rs:29-50):
`recv_ref().await` -> decode -> `self.model.run(&image)` -> draw/encode -> `sender.send(..)`. Frames reach it through
`INFER_IMAGES_CHANNEL`, a bounded (capacity 10, lib.rs:37) *lossy* queue: the router uses `try_send_ref` and simply drops
the frame when the queue is full (router.rs:64-72). A GPU wants batches, so this module replaces that loop:

* `try_submit(stream, frame, on_result)` has the router's semantics: it never blocks and returns False (frame dropped)
  when `capacity` frames are already waiting;
* worker threads drain up to `max_batch` frames (or whatever has arrived when `max_delay` expires, so a lone webcam is
  not held back), call `model.run_batch` — several batches are in flight at once, which is what the lanes of
  `libultraface_b200` are for (the H2D copies of one batch overlap the kernels of another) — and fan the detections
  back to each frame's callback, the analogue of the per-frame `BroadcastSender` (inferer.rs:41-46);
* results of one stream are delivered in submission order even when batches finish out of order.

Host-side plumbing only: no arithmetic happens here.
"""
from __future__ import annotations

import collections
import threading
import time
from typing import Any, Callable, Deque, List, Optional, Tuple


class StreamBatcher:
    def __init__(self, model: Any, max_batch: int = 256, max_delay: float = 0.002, capacity: int = 1024,
                 workers: int = 2, cap: int = 256):
        """model: anything with `run_batch(frames, cap) -> (list of [n,5] arrays, counts)` (UltrafaceModel)."""
        if max_batch < 1 or capacity < 1 or workers < 1:
            raise ValueError("max_batch, capacity and workers must be positive")
        self.model, self.max_batch, self.max_delay, self.capacity, self.cap = model, max_batch, max_delay, capacity, cap
        self._q: Deque[Tuple[Any, Any, Callable]] = collections.deque()
        self._cv = threading.Condition()
        self._closed = False
        self._next_batch = 0      # sequence number handed to the next batch that is formed
        self._next_deliver = 0    # sequence number allowed to deliver
        self._deliver_cv = threading.Condition()
        self.dropped = 0
        self.batches = 0
        self.frames = 0
        self._threads = [threading.Thread(target=self._worker, daemon=True) for _ in range(workers)]
        for t in self._threads:
            t.start()

    # -- producer side (router.rs:64-72)
    def try_submit(self, stream: Any, frame: Any, on_result: Callable[[Any, list], None]) -> bool:
        with self._cv:
            if self._closed:
                raise RuntimeError("batcher is closed")
            if len(self._q) >= self.capacity:
                self.dropped += 1
                return False
            self._q.append((stream, frame, on_result))
            self._cv.notify()
            return True

    # -- consumer side (inferer.rs:29-50, batched)
    def _take_batch(self) -> Optional[Tuple[int, List[Tuple[Any, Any, Callable]]]]:
        with self._cv:
            while not self._q and not self._closed:
                self._cv.wait()
            if not self._q:
                return None
            deadline = time.monotonic() + self.max_delay
            while len(self._q) < self.max_batch and not self._closed:
                left = deadline - time.monotonic()
                if left <= 0:
                    break
                self._cv.wait(left)
            n = min(len(self._q), self.max_batch)
            items = [self._q.popleft() for _ in range(n)]
            seq = self._next_batch
            self._next_batch += 1
            return seq, items

    def _worker(self) -> None:
        while True:
            got = self._take_batch()
            if got is None:
                return
            seq, items = got
            try:
                dets, counts = self.model.run_batch([it[1] for it in items], self.cap)
                results: List[Any] = [[((float(d[0]), float(d[1]), float(d[2]), float(d[3])), float(d[4])) for d in det]
                                      for det in dets]
                error = None
            except Exception as e:  # a failed batch skips its frames, like `if let Ok(..)` in inferer.rs:37
                results, error = [None] * len(items), e
            with self._deliver_cv:  # deliver batches in the order they were formed => per-stream order is kept
                while self._next_deliver != seq:
                    self._deliver_cv.wait()
                try:
                    for (stream, _, cb), res in zip(items, results):
                        if res is not None:
                            cb(stream, res)
                    self.batches += 1
                    self.frames += len(items)
                    self.last_error = error
                finally:
                    self._next_deliver += 1
                    self._deliver_cv.notify_all()

    def close(self) -> None:
        """Stop accepting frames, finish what is queued, join the workers."""
        with self._cv:
            self._closed = True
            self._cv.notify_all()
        for t in self._threads:
            t.join()

    last_error: Optional[Exception] = None
