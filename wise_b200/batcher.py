"""Request micro-batcher: coalesce concurrent single-query searches into one
`index.search(n > 1)` call so that the tensor-core path (K2, 5+ queries) is reachable from the product.

WISE's REST handlers are `async def` with a blocking body (/root/reference/api/routes.py:1212,1257): every request
is a separate `index.search(features, end)` with n = 1 on the event-loop thread.  With this wrapper the handler does

    dist, ids = await batcher.search(features, end)

requests that arrive within `max_wait_ms` (or until `max_batch` are queued) share one GPU call; the blocking call runs
on a worker thread, off the event loop.  Pure host logic: the index can be any object with a faiss-style `search`.
"""
from __future__ import annotations

import asyncio
import threading
import time
from concurrent.futures import Future

import numpy as np


class SearchBatcher:
    def __init__(self, index, max_batch: int = 64, max_wait_ms: float = 1.0):
        self.index = index
        self.max_batch = int(max_batch)
        self.max_wait = float(max_wait_ms) / 1e3
        self._lock = threading.Condition()
        self._queue: list[tuple[np.ndarray, int, Future]] = []
        self._stop = False
        self.batches = 0
        self.requests = 0
        self._thread = threading.Thread(target=self._run, name="wise-b200-batcher", daemon=True)
        self._thread.start()

    # -- client side ------------------------------------------------------------------------------------
    def submit(self, x: np.ndarray, k: int) -> Future:
        """x: float32 (1, d) (or (m, d)); returns a Future of (D, I) with the same row count."""
        x = np.ascontiguousarray(x, np.float32)
        if x.ndim != 2:
            raise AssertionError("x must be (m, d)")
        fut: Future = Future()
        with self._lock:
            if self._stop:
                raise RuntimeError("batcher is closed")
            self._queue.append((x, int(k), fut))
            self._lock.notify()
        return fut

    def search_blocking(self, x: np.ndarray, k: int):
        return self.submit(x, k).result()

    async def search(self, x: np.ndarray, k: int):
        return await asyncio.wrap_future(self.submit(x, k))

    def close(self):
        with self._lock:
            self._stop = True
            self._lock.notify()
        self._thread.join(timeout=5)

    # -- worker -----------------------------------------------------------------------------------------
    def _take(self):
        with self._lock:
            while not self._queue and not self._stop:
                self._lock.wait()
            if self._stop and not self._queue:
                return None
            deadline = time.monotonic() + self.max_wait
            while sum(r[0].shape[0] for r in self._queue) < self.max_batch and not self._stop:
                left = deadline - time.monotonic()
                if left <= 0:
                    break
                self._lock.wait(left)
            batch, rows = [], 0
            while self._queue and (not batch or rows + self._queue[0][0].shape[0] <= self.max_batch):
                r = self._queue.pop(0)
                batch.append(r)
                rows += r[0].shape[0]
            return batch

    def _run(self):
        while True:
            batch = self._take()
            if batch is None:
                return
            try:
                kmax = max(k for _, k, _ in batch)  # one call at the largest k; each request keeps its prefix
                xs = np.concatenate([x for x, _, _ in batch], axis=0)
                D, I = self.index.search(xs, kmax)
                self.batches += 1
                self.requests += len(batch)
                o = 0
                for x, k, fut in batch:
                    m = x.shape[0]
                    fut.set_result((D[o:o + m, :k].copy(), I[o:o + m, :k].copy()))
                    o += m
            except Exception as e:  # propagate to every waiter
                for _, _, fut in batch:
                    if not fut.done():
                        fut.set_exception(e)
